"""SIMP_VIBE — eigenfrequency topology optimisation with MAC mode tracking, on the device.

Mirror of the reference's FEMOL/TOM.py (`SIMP_VIBE(Problem, volfrac, p, q, rmin, FEM_solver_type)`,
`solve(v_ref, converge, min_iter, max_iter, ...)`, attributes `lmbd, dlmbd, lmbds, v, X, change, loop`).
One iteration (TOM.py:110-127):

    K(x^p), M(x^q)               femol_assemble x2
    lowest modes                  LOBPCG over the block kernels (was dense eigh of ALL modes, FEM.py:277)
    mode pick by MAC vs v_ref     femol_mac + argmax                      (TOM.py:383-386)
    dl_e                          femol_sens_eig, lambda in Hz            (TOM.py:408-495, FEM.py:256)
    filter / OC (move 0.3)        femol_filter_* / femol_oc_update mode 1 (TOM.py:305-354, 251-303)

Differences from the reference, by construction: only the `n_modes` lowest PHYSICAL modes are computed
(the reference computes all n, including one spurious lambda = 1 pair per fixed dof); the tracked mode must
be among them.  Save / plot / core-height post-processing are outside the accelerated path.
"""
import numpy as np
import torch

from . import _lib
from .TOM_SIMP import apply_filter, build_filter, domain_mask

__all__ = ['SIMP_VIBE']


class SIMP_VIBE(object):

    def __init__(self, Problem, volfrac=0.5, p=3, q=1, rmin=1.5, FEM_solver_type='fast', saving_root=None):
        self.FEM = Problem
        self.mesh = Problem.mesh
        self.mesh.compute_element_centers()
        if FEM_solver_type != 'fast':
            raise NotImplementedError("only FEM_solver_type='fast' is on the accelerated path")
        self.FEM_solver = self._SIMP_fast_modal_solver
        kind = 'coating' * self.FEM.coating + 'solid' * (not self.FEM.coating)
        self.objective_function = {'solid': self._solid_max_eigs_objective_function,
                                   'coating': self._coating_max_eigs_objective_function}[kind]
        self.element_areas = np.hstack([self.mesh.areas[c] for c in self.mesh.contains])
        self.mesh_area = np.sum(self.element_areas)
        self.f, self.rmin, self.p, self.q = volfrac, rmin, p, q
        self.X = {key: np.ones(self.mesh.cells[key].shape[0]) * volfrac for key in self.mesh.contains}
        self.lmbds, self.all_lmbds, self.eigen_vectors, self.macs = [], [], [], []
        self.v = 0
        self.lmbd = 0
        self.dlmbd = 0
        self.loop = 0
        self.change = None
        self.save_root = 'Results/_topopt_cache/' if saving_root is None else saving_root   # TOM.py:84-86
        self.archive = True
        d = self.FEM.device
        self._x = torch.full((d.n_elem,), float(volfrac), dtype=torch.float64, device=d.dev)
        self._dl = torch.empty(d.n_elem, dtype=torch.float64, device=d.dev)
        self._dlf = torch.empty(d.n_elem, dtype=torch.float64, device=d.dev)
        self._xn = torch.empty(d.n_elem, dtype=torch.float64, device=d.dev)
        self._scal = torch.zeros(8, dtype=torch.float64, device=d.dev)
        self._lam = torch.zeros(1, dtype=torch.float64, device=d.dev)
        self._areas = None if self.mesh.structured else torch.from_numpy(np.ascontiguousarray(self.element_areas)).to(d.dev)
        self._filter = None
        self._mask = None
        self._mask_key = None
        self.unconverged_solves = 0   # SIMP iterations whose eigen-solve stopped short of its tolerance

    def set_density(self, X):
        if isinstance(X, dict):
            X = X[self.mesh.contains[0]]
        self._x.copy_(torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)))

    def density(self):
        return self._x.cpu().numpy()

    def _SIMP_fast_modal_solver(self, X=None, v_ref=None, verbose=False):
        """TOM.py:374-386: assemble K(x^p), M(x^q), solve, pick the mode with the best MAC to v_ref.
        Returns (f [Hz], v (device tensor))."""
        d = self.FEM.device
        x = self._x if X is None else X
        self.FEM.assemble('K', X=x, p=self.p)
        self.FEM.assemble('M', X=x, q=self.q)
        lam, V = self.FEM.modal_device()
        w = np.sqrt(np.maximum(lam, 0.0)) / (2 * np.pi)
        self.all_lmbds.append(w)
        vr = v_ref if isinstance(v_ref, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v_ref, dtype=np.float64)).to(d.dev)
        mac = torch.empty(V.shape[0], dtype=torch.float64, device=d.dev)
        _lib.call('femol_mac', d.n, V.shape[0], _lib.ptr(V), _lib.ptr(vr), _lib.ptr(mac), _lib.ptr(d.red_ws), _lib.stream())
        mac_h = mac.cpu().numpy()
        i = int(np.argmax(mac_h))
        self.macs.append(float(mac_h[i]))
        if verbose:
            print('Best mac match (LOBPCG) :', mac_h[i])
        return w[i], V[i].contiguous()

    def _objective(self, coating):
        d = self.FEM.device
        Ke, pe_k = self.FEM._element_matrices('K', coating)
        Me, pe_m = self.FEM._element_matrices('M', coating)
        structured = self.mesh.structured
        # unstructured SOLID problems reuse the x^p / x^q scaled element matrices (FEM.py:630, 779); the
        # unstructured coating K_coat is scaled too (FEM.py:673) but M_coat is not (FEM.py:845-849)
        ek = self.p - 1 if structured else 2 * self.p - 1
        em = self.q - 1 if (structured or coating) else 2 * self.q - 1
        self._lam.fill_(float(self.lmbd))
        _lib.call('femol_sens_eig', d.plan, _lib.ptr(Ke), pe_k, _lib.ptr(Me), pe_m, _lib.ptr(self._v), _lib.ptr(self._x),
                  float(self.p), float(self.q), float(ek), float(em), _lib.ptr(self._lam), _lib.ptr(self._dl),
                  _lib.stream())
        return self.lmbd, self._dl

    def _solid_max_eigs_objective_function(self, X=None):
        """TOM.py:408-451"""
        return self._objective(False)

    def _coating_max_eigs_objective_function(self, X=None):
        """TOM.py:453-495"""
        return self._objective(True)

    def _filter_sensibility(self, X=None):
        """TOM.py:305-354"""
        if self._filter is None:
            self._filter = build_filter(self.mesh, self.rmin, self.FEM.device.dev)
        apply_filter(self._filter, self._x, self._dl, self._dlf)
        self._dl, self._dlf = self._dlf, self._dl

    def _get_new_x(self, X=None):
        """TOM.py:251-303 (move 0.3, negative sensitivities clamped, rescale when sum < 1)."""
        key = (getattr(self, 'solid_domain', None), getattr(self, 'void_domain', None))
        if key != self._mask_key:
            self._mask = domain_mask(self.mesh, key[0], key[1], self.FEM.device.dev)
            self._mask_key = key
        d = self.FEM.device
        _lib.call('femol_oc_update', d.n_elem, _lib.ptr(self._x), _lib.ptr(self._dl), 1, float(self.f),
                  _lib.ptr(self._areas), _lib.ptr(self._mask), _lib.ptr(self._xn), self._scal.data_ptr() + 4 * 8,
                  _lib.ptr(d.red_ws), _lib.stream())
        return self._xn

    def iterate(self, v_ref):
        self.lmbd, self._v = self.FEM_solver(None, v_ref=v_ref, verbose=False)
        self.lmbds.append(self.lmbd)
        self.objective_function(None)
        self._filter_sensibility(None)
        self._get_new_x(None)
        self._x, self._xn = self._xn, self._x
        self.change = float(self._scal[4].item())
        return self.change, self.lmbd

    def check_convergence(self, converge, min_iter, max_iter, convergence_criteria):
        """TOM.py:155-173"""
        solved = False
        if convergence_criteria == 'change':
            if (self.change < converge) and (self.loop > min_iter):
                solved = True
        elif convergence_criteria == 'objective':
            if np.abs(np.mean(self.lmbds[-3:-1]) - self.lmbds[-1]) < converge:
                solved = True
        if self.loop == max_iter:
            solved = True
        return solved

    def solve(self, v_ref, converge=0.01, min_iter=1, max_iter=100, plot=True, save=True, verbose=True,
              convergence_criteria='change', mesh_filename=None, eigvals_filename=None, eigvecs_filename=None):
        """TOM.py:87-132.  `plot` (matplotlib) is ignored.  `save`: the reference rewrites the same result files on
        every iteration (TOM.py:134-152); they are written once, after the last iteration."""
        files = self._get_saving_filenames(mesh_filename, eigvals_filename, eigvecs_filename)
        self.change = 1
        self.loop = 0
        ctype = self.mesh.contains[0]
        self.set_density(self.X)
        solved = False
        while not solved:
            self.loop += 1
            self.iterate(v_ref)
            self.unconverged_solves += 0 if self.FEM.solver_info.get('converged', True) else 1
            if self.archive:
                self.X = {ctype: self.density()}
                self.v = self._v.cpu().numpy()
                self.eigen_vectors.append(self.v)
                self.dlmbd = self._dl.cpu().numpy()
                X_key = 'X{}'.format(self.loop - 1)
                if 'X' in self.mesh.cell_data.keys():
                    self.mesh.cell_data[X_key] = self.mesh.cell_data['X']
                self.mesh.cell_data['X'] = self.X
                self.mesh.add_mode('m{}'.format(self.loop), self.v, self.FEM.N_dof)
            if verbose:
                print(f'Iteration : {self.loop}, Variation : {self.change}, EigenVal : {self.lmbd}')
            solved = self.check_convergence(converge, min_iter, max_iter, convergence_criteria)
        if not self.archive:
            self.X = {ctype: self.density()}
            self.v = self._v.cpu().numpy()
            self.dlmbd = self._dl.cpu().numpy()
            self.mesh.cell_data['X'] = self.X
        if save:
            self.save_TOM_iteration(*files)
        return self.mesh

    def _get_saving_filenames(self, mesh_filename, eigvals_filename, eigvecs_filename):
        """TOM.py:211-223"""
        from .utils import unique_time_string
        stamp = unique_time_string()
        root = self.save_root
        return [root + (name if name is not None else base + stamp)
                for name, base in zip((mesh_filename, eigvals_filename, eigvecs_filename), ('TOM_', 'eigvals_', 'eigvecs_'))]

    def save_TOM_iteration(self, msh_file, eig_file, vec_file):
        """TOM.py:134-152: core heights (two-laminate problems), penalised densities `X_real`, the mesh as VTK and the
        tracked eigenvalue history as .npy."""
        import os
        if len(self.FEM.layups) == 2:                 # the reference assumes a base + coating laminate here
            self.mesh = self.density_to_core_height()
        ctype = self.mesh.contains[0]
        self.mesh.cell_data['X_real'] = {ctype: np.asarray(self.mesh.cell_data['X'][ctype]) ** self.p}
        folder = os.path.dirname(msh_file)
        if folder and not os.path.isdir(folder):
            os.makedirs(folder, exist_ok=True)
        self.mesh.save(msh_file)
        np.save(eig_file, np.array(self.lmbds))

    def density_to_core_height(self):
        """TOM.py:497-540: map the optimised densities to a core height per element (`cell_data['zc']`), the input
        of `FEM_Problem.solve(solve_from_zc=True)`.  The bending stiffness of the two shifted laminates is sampled
        at 100 core heights, normalised to [0, 1], its p-th root taken as the density-height law and inverted by
        linear interpolation.  Host-side post-processing (100 small laminates), as in the reference."""
        from .laminate import Layup
        lb, lc = self.FEM.layups[0], self.FEM.layups[1]
        hc = -lb.zc - lb.hA / 2 + lc.zc - lc.hA / 2
        h_values = np.linspace(0, hc, 100)
        D1, D2 = [], []
        for hi in h_values:
            b = Layup(plies=lb.angles, material=lb.mtr, symetric=False, z_core=-hi / 2 - lb.hA / 2)
            c = Layup(plies=lc.angles, material=lc.mtr, symetric=False, z_core=hi / 2 + lc.hA / 2)
            D1.append(b.D_mat[0, 0] + c.D_mat[0, 0])
            D2.append(b.D_mat[1, 1] + c.D_mat[1, 1])
        base = Layup(plies=lb.angles, material=lb.mtr, symetric=False)
        Db1, Db2 = base.D_mat[0, 0], base.D_mat[1, 1]
        X1 = (np.array(D1) - Db1) / (D1[-1] - Db1)
        X2 = (np.array(D2) - Db2) / (D2[-1] - Db2)
        X_interp = np.mean([X1, X2], axis=0) ** (1 / self.p)
        X_interp = np.append(np.append(0, X_interp), 1)
        h_values = np.append(np.append(0, h_values), hc)
        zcore = {}
        for key in self.mesh.cell_data['X']:
            z = np.interp(self.mesh.cell_data['X'][key], X_interp, h_values)   # scipy interp1d (linear) in the reference
            z[z < 0] = np.min(h_values)
            zcore[key] = z
        self.mesh.cell_data['zc'] = zcore
        return self.mesh
