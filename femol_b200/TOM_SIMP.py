"""SIMP_COMP — compliance topology optimisation (solid and coating formulations) on the device.

Mirror of the reference's FEMOL/TOM_SIMP.py: same constructor (`Problem, volfrac, penal, rmin`), same
`solve(converge, max_iter, plot, save)` loop and attributes (`X, U, c, dc, change, loop, f, p, rmin,
FEM_solver, objective_function`, optional `solid_domain` / `void_domain`).  One iteration is
(TOM_SIMP.py:48-60)

    U = FEM_solver(X)            femol_assemble (x^p fused) + femol_pcg          (was assemble + spsolve)
    c, dc = objective(X)         femol_sens_compliance                           (was a Python loop, :126-226)
    filter                       femol_filter_grid / femol_filter_list           (was O(N^2), :228-277)
    X = OC update                femol_oc_update (device-side bisection)         (was 30 numpy passes, :279-313)
    change = max|X - X_old|      fused in femol_oc_update

The densities stay on the device between iterations; one scalar pair (change, objective) is read back per
iteration.  Reference quirks are reproduced (SURVEY §8 a-quirks): the filter excludes the centre element,
OC exponent 0.3 / move 0.5 / x_min 0.001, `any(dc > 0)` sign switch, double penalisation of the objective on
unstructured meshes (FEM.py:630-631).
"""
import numpy as np
import torch

from . import _lib
from .utils import unique_time_string

__all__ = ['SIMP_COMP']


def build_filter(mesh, rmin, device):
    """Filter weights (TOM_SIMP.py:237-267).  Structured meshes: a (2rx+1) x (2ry+1) stencil evaluated with
    the reference's own formula on the cell centres; otherwise a neighbour list with weights."""
    h = mesh.element_size()
    ctype = mesh.contains[0]
    if mesh.structured and mesh.ElementClasses[ctype].N_nodes in (4, 8):
        centers = mesh.all_cell_centers
        n = centers.shape[0]
        # cells are numbered e = i * n_fast + j (mesh.py:852-860 / meshzoo ordering): find n_fast
        y0 = centers[0, 1]
        same_col = np.isclose(centers[:, 0], centers[0, 0])
        n_fast = int(np.count_nonzero(same_col))
        n_slow = n // n_fast
        if n_fast * n_slow == n and (n_fast == 1 or centers[1, 1] > y0):
            dx = centers[n_fast, 0] - centers[0, 0] if n_slow > 1 else h
            dy = centers[1, 1] - centers[0, 1] if n_fast > 1 else h
            rx = int(np.floor(rmin * h / dx + 1e-9))
            ry = int(np.floor(rmin * h / dy + 1e-9))
            di = np.arange(-rx, rx + 1)[:, None] * dx
            dj = np.arange(-ry, ry + 1)[None, :] * dy
            dist = np.sqrt(di ** 2 + dj ** 2)
            w = np.maximum(0.0, rmin - dist / h)
            w[dist > rmin * h * (1 + 1e-12)] = 0.0
            w[rx, ry] = 0.0  # the centre element is excluded (TOM_SIMP.py:255)
            return {'kind': 'grid', 'n_slow': n_slow, 'n_fast': n_fast, 'rx': rx, 'ry': ry,
                    'w': torch.from_numpy(np.ascontiguousarray(w)).to(device)}
    D = neighbour_distances(mesh.all_cell_centers, rmin * h * (1 + 1e-9))
    w = np.maximum(0.0, rmin - D.data / h)
    return {'kind': 'list', 'ptr': torch.from_numpy(D.indptr.astype(np.int32)).to(device),
            'idx': torch.from_numpy(D.indices.astype(np.int32)).to(device),
            'w': torch.from_numpy(np.ascontiguousarray(w)).to(device)}


def neighbour_distances(centers, radius):
    """CSR matrix of the centre distances d_ij <= radius, i != j, sorted columns, zero distances dropped (what
    `tree.sparse_distance_matrix(tree, radius)` minus its diagonal gives, one third faster at 4 M elements: unordered
    pairs from `query_pairs`, distances in numpy)."""
    import scipy.sparse
    from scipy.spatial import cKDTree
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    n = centers.shape[0]
    pairs = cKDTree(centers).query_pairs(radius, output_type='ndarray')
    d = np.sqrt(((centers[pairs[:, 0]] - centers[pairs[:, 1]]) ** 2).sum(axis=1))
    keep = d > 0
    pairs, d = pairs[keep], d[keep]
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    D = scipy.sparse.csr_matrix((np.concatenate([d, d]), (i, j)), shape=(n, n))
    D.sort_indices()
    return D


def apply_filter(flt, x, dc, out):
    if flt['kind'] == 'grid':
        _lib.call('femol_filter_grid', flt['n_slow'], flt['n_fast'], flt['rx'], flt['ry'], _lib.ptr(flt['w']),
                  _lib.ptr(x), _lib.ptr(dc), _lib.ptr(out), _lib.stream())
    else:
        _lib.call('femol_filter_list', x.shape[0], _lib.ptr(flt['ptr']), _lib.ptr(flt['idx']), _lib.ptr(flt['w']),
                  _lib.ptr(x), _lib.ptr(dc), _lib.ptr(out), _lib.stream())


def domain_mask(mesh, solid_domain, void_domain, device):
    """Per-element mask 1 = solid, 2 = void: an element is constrained when ALL its nodes lie in the domain
    (TOM_SIMP.py:315-331); void is applied after solid like the reference."""
    if solid_domain is None and void_domain is None:
        return None
    cells = mesh.cells[mesh.contains[0]]
    mask = np.zeros(cells.shape[0], dtype=np.uint8)
    for dom, val in ((solid_domain, 1), (void_domain, 2)):
        if dom is None:
            continue
        x, y = mesh.points[:, 0], mesh.points[:, 1]
        if getattr(dom, 'vectorized', False):
            inside = np.asarray(dom(x, y)).astype(bool)
        else:
            inside = np.array([bool(dom(*c[:2])) for c in mesh.points])
        mask[inside[cells].all(axis=1)] = val
    return torch.from_numpy(mask).to(device)


class SIMP_COMP(object):
    """Solid Isotropic Material with Penalization, minimum compliance (TOM_SIMP.py:8-32)."""

    def __init__(self, Problem, volfrac=0.5, penal=3, rmin=1.5):
        self.FEM = Problem
        self.mesh = Problem.mesh
        self.mesh.compute_element_centers()
        self.FEM_solver = self._SIMP_displacement_solver
        kind = 'coating' * self.FEM.coating + 'solid' * (not self.FEM.coating)
        self.objective_function = {'solid': self._solid_compliance_objective_function,
                                   'coating': self._coating_compliance_objective_function}[kind]
        self.f = volfrac
        self.rmin = rmin
        self.p = penal
        self.X = {key: np.ones(self.mesh.cells[key].shape[0]) * volfrac for key in self.mesh.contains}
        self.archive = True          # keep X{k} / d{k} histories in the mesh like the reference (:62-69)
        self.iteration_log = []
        # SIMP iterations whose linear solve stopped short of its tolerance (each one also raises a
        # FemolConvergenceWarning, or a FemolError with Problem.on_nonconvergence = 'raise')
        self.unconverged_solves = 0
        d = self.FEM.device
        self._x = torch.full((d.n_elem,), float(volfrac), dtype=torch.float64, device=d.dev)
        self._dc = torch.empty(d.n_elem, dtype=torch.float64, device=d.dev)
        self._dcf = torch.empty(d.n_elem, dtype=torch.float64, device=d.dev)
        self._xn = torch.empty(d.n_elem, dtype=torch.float64, device=d.dev)
        self._scal = torch.zeros(8, dtype=torch.float64, device=d.dev)   # [c, -, -, -, change, sum, lmid, nbis]
        self._filter = None
        self._mask = None
        self._mask_key = None

    # ---- host <-> device density -----------------------------------------------------------------------
    def set_density(self, X):
        """Upload densities (dict / array); used by the end-to-end path with host buffers."""
        if isinstance(X, dict):
            X = X[self.mesh.contains[0]]
        self._x.copy_(torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)))

    def density(self):
        return self._x.cpu().numpy()

    @property
    def x_device(self):
        return self._x

    # ---- reference-compatible pieces ---------------------------------------------------------------------
    def _SIMP_displacement_solver(self, X=None):
        """TOM_SIMP.py:118-124: assemble K(x^p) and solve; returns the device displacement."""
        self.FEM.assemble('K', X=self._x if X is None else X, p=self.p)
        return self.FEM.solve_device()

    def _objective(self, X, coating):
        d = self.FEM.device
        x = self._x if X is None else self.FEM._x_device(X)
        Eb, pe_b = self.FEM._element_matrices('K', False)
        Ec, pe_c = self.FEM._element_matrices('K', True) if coating else (None, 0)
        structured = self.mesh.structured
        pc = self.p if structured else 2 * self.p          # FEM.py:630-631 double penalisation quirk
        pd = pc - 1
        _lib.call('femol_sens_compliance', d.plan, _lib.ptr(Eb), pe_b, _lib.ptr(Ec), pe_c, _lib.ptr(self._U),
                  _lib.ptr(x), float(self.p), float(pc), float(pd), _lib.ptr(self._dc), _lib.ptr(self._scal),
                  _lib.ptr(d.red_ws), _lib.stream())
        return self._scal[0], self._dc

    def _solid_compliance_objective_function(self, X=None):
        """TOM_SIMP.py:126-174 -> (c, dc) as device tensors"""
        return self._objective(X, False)

    def _coating_compliance_objective_function(self, X=None):
        """TOM_SIMP.py:176-226"""
        return self._objective(X, True)

    def _filter_sensibility(self, X=None):
        """TOM_SIMP.py:228-277"""
        if self._filter is None:
            self._filter = build_filter(self.mesh, self.rmin, self.FEM.device.dev)
        apply_filter(self._filter, self._x, self._dc, self._dcf)
        self._dc, self._dcf = self._dcf, self._dc

    def _get_new_x(self, X=None):
        """TOM_SIMP.py:279-313 on the device; writes self._xn and the scalars (change, volume, lmid)."""
        key = (getattr(self, 'solid_domain', None), getattr(self, 'void_domain', None))
        if key != self._mask_key:
            self._mask = domain_mask(self.mesh, key[0], key[1], self.FEM.device.dev)
            self._mask_key = key
        d = self.FEM.device
        _lib.call('femol_oc_update', d.n_elem, _lib.ptr(self._x), _lib.ptr(self._dc), 0, float(self.f), None,
                  _lib.ptr(self._mask), _lib.ptr(self._xn), self._scal.data_ptr() + 4 * 8, _lib.ptr(d.red_ws),
                  _lib.stream())
        return self._xn

    def iterate(self):
        """One SIMP iteration entirely on the device; returns (change, c) after a single 2-scalar read-back."""
        self._U = self.FEM_solver(None)
        self.objective_function(None)
        self._filter_sensibility(None)
        self._get_new_x(None)
        self._x, self._xn = self._xn, self._x
        s = self._scal.cpu().numpy()
        self.c, self.change = float(s[0]), float(s[4])
        return self.change, self.c

    def solve(self, converge=0.01, max_iter=100, plot=True, save=True):
        """TOM_SIMP.py:34-81.  `plot` (matplotlib) is ignored.  `save`: the reference rewrites ONE VTK file
        (named after the start time) on every iteration; the same file is written once, after the last iteration."""
        start_time = unique_time_string()
        self.change = 1
        self.loop = 0
        ctype = self.mesh.contains[0]
        self.set_density(self.X)
        while (self.change > converge) & (self.loop < max_iter):
            self.loop += 1
            self.iterate()
            info = dict(self.FEM.solver_info)
            info.update(loop=self.loop, change=self.change, c=self.c)
            self.iteration_log.append(info)
            self.unconverged_solves += 0 if info.get('converged', True) else 1
            if self.archive:
                self.X = {ctype: self.density()}
                self.U = self._U.cpu().numpy()
                self.dc = self._dc.cpu().numpy()
                X_key = 'X{}'.format(self.loop - 1)
                if 'X' in self.mesh.cell_data.keys():
                    self.mesh.cell_data[X_key] = self.mesh.cell_data['X']
                self.mesh.cell_data['X'] = self.X
                self.mesh.add_mode('d{}'.format(self.loop), self.U, self.FEM.N_dof)
            if not plot:
                print(f"Iteration : {self.loop}, variation: {self.change}, objective : {self.c}")
        if not self.archive:
            self.X = {ctype: self.density()}
            self.U = self._U.cpu().numpy()
            self.dc = self._dc.cpu().numpy()
            self.mesh.cell_data['X'] = self.X
        self.mesh.U = self.U
        if save:
            self._save_TOM_result(start_time)
        return self.mesh

    def _save_TOM_result(self, timestring):
        """TOM_SIMP.py:108-116: Results/_topopt_cache/TOM_<time>.vtk, or the working directory if that folder
        does not exist."""
        try:
            self.mesh.save('Results/_topopt_cache/TOM_' + timestring)
        except FileNotFoundError:
            self.mesh.save('TOM_' + unique_time_string())
