"""FEM_Problem — B200-native mirror of the reference's FEMOL/FEM.py.

Same public surface (`FEM_Problem(physics, model, mesh)`, `define_materials`, `define_tensors`,
`add_fixed_domain`, `add_forces`, `assemble(which, X, p, q)`, `solve(...)`, attributes `K`, `M`, `F`,
`tensors`, `coat_tensors`, `ho`, `coating`, `N_dof`) and the same conventions, but the per-iteration work
runs in hand-written sm_100a kernels behind the C-ABI of include/femol_b200.h:

  reference (FEM.py)                                   here
  _K/_M_*_mesh_data_* :550-871 + BC scan :902-930     femol_ke_batch / femol_me_batch + femol_assemble
  csr_matrix + sum_duplicates :498-501                 symbolic plan (femol_plan_create) + femol_assemble
  spsolve :258-265                                     femol_pcg (Jacobi / block-Jacobi PCG, warm start)
  dense eigh :267-295 / eigsh :234                     LOBPCG over femol_spmm / femol_gram / ... (lobpcg.py)

K / M live on the device as CSR values in the plan's order; `problem.K` / `problem.M` materialise scipy CSR
matrices on demand (bit-identical pattern: sorted int32 indices, explicit zeros kept).  There is no CPU
fallback: without the CUDA library or a GPU every compute call raises.
"""
import ctypes
import warnings

import numpy as np
import scipy.sparse
import torch

from . import _lib
from . import elements
from . import lobpcg as _lobpcg
from .laminate import Layup
from .mesh import Mesh

__all__ = ['FEM_Problem', 'Mesh', 'Layup']


class _TrackedVector(np.ndarray):
    """`problem.F` as handed back after an assembly: an ordinary float64 array whose in-place edits
    (`F[i] = v`, `F[a:b] += w`, writes through views) raise a shared dirty flag, so the next assembly re-applies
    the boundary conditions and re-uploads it - the reference re-reads `self.F` on every assembly (FEM.py:536-544);
    checksumming 4 M entries per SIMP iteration instead would cost 2 - 3 ms."""

    def __new__(cls, data):
        obj = np.asarray(data, dtype=np.float64).view(cls)
        obj._flag = [False]
        return obj

    def __array_finalize__(self, obj):
        self._flag = getattr(obj, '_flag', None) or [False]

    def __setitem__(self, key, value):
        self._flag[0] = True
        super().__setitem__(key, value)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        plain = [np.asarray(a) if isinstance(a, _TrackedVector) else a for a in inputs]
        if out is not None:
            for o in out:
                if isinstance(o, _TrackedVector):
                    o._flag[0] = True
            kwargs['out'] = tuple(np.asarray(o) if isinstance(o, _TrackedVector) else o for o in out)
        res = getattr(ufunc, method)(*plain, **kwargs)
        if out is not None and len(out) == 1 and isinstance(out[0], _TrackedVector) and res is kwargs['out'][0]:
            return out[0]              # `F += w` must rebind the name to the tracked array, not to a plain view
        return res

    @property
    def dirty(self):
        return self._flag[0]

_PRECOND = {'none': _lib.PRECOND_NONE, 'jacobi': _lib.PRECOND_JACOBI, 'block_jacobi': _lib.PRECOND_BLOCK_JACOBI}


class _Device(object):
    """Device-resident mesh data, the symbolic plan and workspaces of one FEM_Problem."""

    def __init__(self, mesh, N_dof):
        _lib.require_gpu()
        if len(mesh.contains) != 1:
            raise NotImplementedError('femol_b200 handles meshes with a single cell type '
                                      '(mixed triangle/quad meshes are outside the accelerated path)')
        self.ctype = mesh.contains[0]
        cls = mesh.ElementClasses[self.ctype]
        self.etype = cls.ETYPE
        self.nn = cls.N_nodes
        self.ndof = N_dof
        self.dev = torch.device('cuda', torch.cuda.current_device())
        self.n_nodes = mesh.points.shape[0]
        self.n_elem = mesh.cells[self.ctype].shape[0]
        self.n = self.n_nodes * N_dof
        self.xy = torch.from_numpy(np.ascontiguousarray(mesh.points[:, :2], dtype=np.float64)).to(self.dev)
        self.conn = torch.from_numpy(np.ascontiguousarray(mesh.cells[self.ctype], dtype=np.int32)).to(self.dev)
        handle = ctypes.c_void_p()
        _lib.call('femol_plan_create', self.n_elem, self.nn, N_dof, self.n_nodes, _lib.ptr(self.conn), _lib.stream(),
                  ctypes.byref(handle))
        self.plan = handle
        info = (ctypes.c_int64 * 8)()
        _lib.call('femol_plan_info', self.plan, info)
        self.nnz_node, self.nnz = int(info[0]), int(info[1])
        self.max_valence, self.max_row_nodes = int(info[2]), int(info[3])
        self.xpow = torch.empty(self.n_elem, dtype=torch.float64, device=self.dev)
        self.xpow_m = torch.empty(self.n_elem, dtype=torch.float64, device=self.dev)
        self.red_ws = torch.zeros(int(_lib.load().femol_reduce_workspace_doubles()), dtype=torch.float64, device=self.dev)
        self._csr = None
        self._pcg_ws = None

    def __del__(self):
        try:
            if getattr(self, 'plan', None):
                _lib.load().femol_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def csr_pattern(self):
        if self._csr is None:
            indptr = torch.empty(self.n + 1, dtype=torch.int32, device=self.dev)
            indices = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
            _lib.call('femol_plan_export_csr', self.plan, _lib.ptr(indptr), _lib.ptr(indices), _lib.stream())
            self._csr = (indptr, indices)
        return self._csr

    def scatter_map(self):
        out = torch.empty((self.n_elem, self.nn, self.nn), dtype=torch.int32, device=self.dev)
        _lib.call('femol_plan_export_scatter_map', self.plan, _lib.ptr(out), _lib.stream())
        return out

    def pcg_ws(self):
        if self._pcg_ws is None:
            nd = int(_lib.load().femol_pcg_workspace_doubles(self.plan))
            self._pcg_ws = torch.empty(nd, dtype=torch.float64, device=self.dev)
        return self._pcg_ws

    def to_scipy(self, vals):
        indptr, indices = self.csr_pattern()
        return scipy.sparse.csr_matrix((vals.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()),
                                       shape=(self.n, self.n))


class FEM_Problem(object):
    """General structural FEM problem: displacement (K U = F) or modal (K v = lambda M v), plane stress
    (2 dof / node) or Reissner-Mindlin plate (6 dof / node).  Reference: FEM.py:21-59."""
    dof_dict = {'plane': 2, 'plate': 6}

    def __init__(self, physics, model, mesh):
        self.mesh = mesh
        self.physics = physics
        self.model = model
        self.N_dof = self.dof_dict[model]
        self.coating = False
        # reference FEM.py:49 builds the COO index arrays here; the device plan replaces them and
        # mesh.rows / mesh.cols stay available lazily (mesh.global_matrix_indexes)
        ctype = mesh.contains[0]
        if mesh.structured:
            mesh.element = mesh.ElementClasses[ctype](mesh.points[mesh.cells[ctype][0]], N_dof=self.N_dof)
        self.fixed_nodes = np.array([])
        self.fixed_domains = []
        self.fixed_ddls = []
        self.forces = []
        self.force_domains = []
        self.layups = []
        self.F = np.zeros(self.N_dof * self.mesh.N_nodes)
        self.V = None
        # solver controls (no reference counterpart: spsolve is direct)
        self.solver_rtol = 1e-10
        self.solver_max_iter = 200000
        # a solve that stops without reaching its tolerance: 'warn' (FemolConvergenceWarning, results are still
        # returned), 'raise' (FemolError) or 'ignore'.  PCG status 3 (recurrence converged, true residual at its
        # attainable floor ~ eps * cond(K)) counts as converged up to `stagnation_bound`
        self.on_nonconvergence = 'warn'
        self.stagnation_bound = 1e-6
        # 'auto': geometric multigrid on rectangle_Q4-ordered structured meshes, block-Jacobi otherwise;
        # 'gmg' | 'block_jacobi' | 'jacobi' | 'none' force one
        self.preconditioner = 'auto'
        self.gmg_smoothing_steps = 1     # pre = post sweeps of damped block-Jacobi (tuned on cfg 2, tools/tune_gmg.py)
        self.gmg_omega = 0.7
        # Q4 Mindlin plates: second multigrid cycle on the sign-modulated operator S K S, which removes the hourglass
        # family of the selectively integrated element from the CG spectrum (femol_gmg_attach_modulated).  'auto': plate
        # models whose boundary is not fully supported in w (with all edges supported the family is restrained) and
        # whose thickness is at least 8 element sizes: the cycle costs as much as the first one and, measured on a
        # cantilever (n = 200): t/h = 20: 262 -> 67 iterations, 10: 270 -> 119, 4: 386 -> 279, 2: 540 -> 558,
        # 0.5: 2264 -> 2978; far below one element size it can even prevent convergence
        self.gmg_hourglass = 'auto'
        self.warm_start = True
        # structured rectangle_Q4 grids: PCG / multigrid apply K matrix-free (csrc/stencil.cu); K itself is
        # still assembled (it is what .K returns and what the block eigen-solver multiplies with)
        self.matrix_free = True
        self.n_modes = 10
        self.modal_tol = 1e-8
        self.modal_max_iter = 500
        self.modal_inner_iters = 200     # block-CG iterations on K per LOBPCG preconditioner application
        self.modal_inner_tol = 1e-3
        self.modal_gmg_iters = 8         # structured grids: multigrid-CG iterations per column instead of the block-CG
        # Reissner-Mindlin drilling rotations: K_tz is the diagonal set to 1 per element (elements.py:1071)
        # and the consistent mass couples tz only with tz, so the pencil has one exactly decoupled pure-tz
        # mode per free node (a dense cluster of spurious low frequencies).  They are projected out of the
        # LOBPCG iteration, like the lambda = 1 pairs of the fixed dofs.
        self.deflate_drilling = True
        self.solver_info = {}
        self._dev = None
        self._E = {}          # cached element matrices on the device
        self._E_host = {}     # host copies of the shared (structured-mesh) element matrices
        self._Kv = self._Mv = None
        self._K_host = self._M_host = None
        self._U_dev = None
        self._U_mask_key = None
        self._X_modal = None  # last eigenvector block (warm start)
        self._epoch = 0
        self._mask_key = None
        self._mask_dev = None
        self._mask_host = None
        self._fixed_dofs = None
        self._fixed_dofs_key = None
        self._F_key = None
        self._F_ref = None
        self._grid_shape = None
        self._lattice = None
        self._gmg = None
        self._gmg2 = None           # hierarchy of S K S (plates), on a view of the plan
        self._plan_view = None
        self._gmg2_refs = None
        self._gmg_mask_key = None
        self._gmg_is_lattice = False
        self._vc_refs = None        # variable-core laminates: (base matrices, coating matrices, 0/1 indicator) on the device
        self._gmg_stale = True
        self._gmg_ws = None
        self._K_scaled = False
        self._mf_refs = None  # keeps the tensors the plan's matrix-free descriptor borrows alive
        self._variable_core = False   # K / M came from assemble_variable_core (per-element laminates)

    def __del__(self):
        try:
            self._destroy_gmg()
        except Exception:
            pass

    # ---- set-up (host) ----------------------------------------------------------------------------
    def define_materials(self, *materials):
        """FEM.py:105-111"""
        self.materials = materials
        if len(materials) == 2:
            self.coating = True
        self._E.clear()
        self._E_host.clear()

    def define_tensors(self, *tensors):
        """FEM.py:113-154: thickness (isotropic) or Layup (orthotropic), base [+ coating]."""
        C_A, C_B, C_D, C_G, ho = [], [], [], [], []
        for tensor, material in zip(tensors, self.materials):
            if material.kind == 'orthotropic':
                layup = tensor
                ho.append(layup.hA)
                C_A.append(layup.get_A_FEM()); C_B.append(layup.B_mat)
                C_D.append(layup.D_mat); C_G.append(layup.G_mat)
                self.layups.append(layup)
            elif material.kind == 'isotropic':
                t = tensor
                ho.append(t)
                C_A.append(material.plane_tensor(t)); C_B.append(material.coupled_tensor())
                C_D.append(material.bending_tensor(t)); C_G.append(material.shear_tensor(t))
        self.tensors = [C_A[0], C_D[0], C_G[0], C_B[0]]
        self.ho = ho[0]
        if len(C_A) == 2:
            self.coat_tensors = [C_A[1], C_D[1], C_G[1], C_B[1]]
            self.coat_ho = ho[1]
            self.coating = True
        self._E.clear()
        self._E_host.clear()

    def add_fixed_domain(self, domain, ddls=None):
        """FEM.py:424-438"""
        if ddls is None:
            ddls = np.arange(self.N_dof)
        self.fixed_domains.append(domain)
        self.fixed_ddls.append(ddls)

    def add_forces(self, force, domain):
        """FEM.py:440-464: the force vector is split equally over the nodes of the domain."""
        F = np.zeros(self.N_dof * self.mesh.N_nodes)
        self.forces.append(force)
        self.force_domains.append(domain)
        nodes = self.mesh.domain_nodes(domain)
        if len(nodes) > 0:
            f = np.array(force, dtype=float) / len(nodes)
            for ddl in range(self.N_dof):
                F[self.N_dof * nodes + ddl] = f[ddl]
        else:
            print('invalid force domain')
        self.F = self.F + F

    # ---- device plumbing --------------------------------------------------------------------------
    @property
    def device(self):
        if self._dev is None:
            self._dev = _Device(self.mesh, self.N_dof)
            self.mesh._femol_device = self._dev      # Mesh.stress_from_displacement reuses the plan
            shape = self.structured_grid_shape()
            if shape:                                # assembly derives the scatter pattern from grid coordinates
                _lib.call('femol_plan_set_grid', self._dev.plan, shape[0], shape[1])
        return self._dev

    def _bc_key(self):
        """Identity of the Dirichlet set: the domain callables themselves (replacing one in place invalidates the
        cached mask / fixed dofs / multigrid hierarchy / F), their dof lists and the invalidate() epoch."""
        return (self._epoch, tuple(id(d) for d in self.fixed_domains),
                tuple(tuple(np.asarray(d).tolist()) for d in self.fixed_ddls))

    def invalidate(self):
        """Drop every cache derived from the mesh geometry, the domains and the loads (call after editing
        `mesh.points` in place; replaced domains / forces / `F` edits are detected without it)."""
        self._epoch += 1
        self._E.clear()
        self._E_host.clear()
        self._F_key = None
        self._grid_shape = None
        self._lattice = None
        self._gmg_stale = True
        if self._dev is not None:
            self._dev.xy.copy_(torch.from_numpy(np.ascontiguousarray(self.mesh.points[:, :2], dtype=np.float64)))

    def _report_nonconvergence(self, message):
        if self.on_nonconvergence == 'raise':
            raise _lib.FemolError(message)
        if self.on_nonconvergence != 'ignore':
            warnings.warn(message, _lib.FemolConvergenceWarning, stacklevel=3)

    def fixed_dofs(self):
        """Sorted unique global dof numbers fixed by the registered domains (cached: the reference
        re-evaluates every domain predicate over all nodes on each assembly, mesh.py:281)."""
        key = self._bc_key()
        if self._fixed_dofs_key != key:
            out = []
            for domain, ddls in zip(self.fixed_domains, self.fixed_ddls):
                nodes = self.mesh.domain_nodes(domain)
                out.append((nodes[:, None] * self.N_dof + np.asarray(ddls)[None]).reshape(-1))
            self._fixed_dofs = np.unique(np.concatenate(out)) if out else np.zeros(0, dtype=np.int64)
            self._fixed_dofs_key = key
        return self._fixed_dofs

    def _fixed_mask(self):
        """uint8 per node, bit d = dof d fixed (device), or None."""
        key = self._bc_key()
        if key != self._mask_key:
            mask = np.zeros(self.mesh.N_nodes, dtype=np.uint8)
            nodes_all = []
            for domain, ddls in zip(self.fixed_domains, self.fixed_ddls):
                nodes = self.mesh.domain_nodes(domain)
                nodes_all.append(nodes)
                bits = np.uint8(sum(1 << int(d) for d in ddls))
                mask[nodes] |= bits
            self.fixed_nodes = np.concatenate(nodes_all) if nodes_all else np.array([])
            self._mask_dev = torch.from_numpy(mask).to(self.device.dev) if self.fixed_ddls else None
            self._mask_host = mask
            self._mask_key = key
        return self._mask_dev

    def _x_device(self, X):
        if X is None:
            return None
        if isinstance(X, dict):
            X = X[self.mesh.contains[0]]
        if isinstance(X, torch.Tensor):
            return X.to(self.device.dev, torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(self.device.dev)

    def _element_matrices(self, which, coat):
        """Cached device element matrices: (tensor [1 or n_elem, s, s], per_elem flag).
        Structured meshes use the single matrix of cell 0 (FEM.py:571, 757), others one per element."""
        key = (which, coat)
        if key not in self._E:
            d = self.device
            conn = d.conn[:1].contiguous() if self.mesh.structured else d.conn
            if which == 'K':
                T = self.coat_tensors if coat else self.tensors
                E = elements.ke_batch_device(d.etype, self.N_dof, d.xy, conn, T)
            else:
                if self.V is not None and not coat:
                    V = elements.mass_tensor(d.etype, self.N_dof, V=self.V)
                else:
                    mat = self.materials[1] if coat else self.materials[0]
                    V = elements.mass_tensor(d.etype, self.N_dof, mat, self.coat_ho if coat else self.ho)
                E = elements.me_batch_device(d.etype, self.N_dof, d.xy, conn, V)
            self._E[key] = (E, 0 if self.mesh.structured else 1)
        return self._E[key]

    def _element_matrix_host(self, E):
        """Host copy of a shared element matrix (one device read per cached matrix, not per assembly)."""
        if E is None:
            return None
        key = (E.data_ptr(), E._version)
        hit = self._E_host.get(key)
        if hit is None:
            # the tensor is kept with its copy so the address cannot be recycled while the entry lives
            hit = self._E_host[key] = (E, np.ascontiguousarray(E[0].cpu().numpy().reshape(-1)))
        return hit[1]

    def _assemble_values(self, which, X, expo):
        d = self.device
        Eb, pe_b = self._element_matrices(which, False)
        Ec, pe_c = self._element_matrices(which, True) if self.coating else (None, 0)
        x = self._x_device(X)
        vals = torch.empty(d.nnz, dtype=torch.float64, device=d.dev)
        xpow = d.xpow if which == 'K' else d.xpow_m      # x^p of K is reused by the multigrid set-up
        _lib.call('femol_assemble', d.plan, _lib.ptr(Eb), pe_b, _lib.ptr(Ec), pe_c, _lib.ptr(x),
                  float(expo) if expo is not None else 1.0, _lib.ptr(self._fixed_mask()), _lib.ptr(xpow),
                  _lib.ptr(vals), _lib.stream())
        if which == 'K':
            self._K_scaled = x is not None
            self._gmg_stale = True
            # structured grids: the solvers apply K from (Ke, x^p, mask) instead of streaming the CSR values
            shape = self.structured_grid_shape() if self.matrix_free else None
            if shape:
                mask = self._fixed_mask()
                hb, hc = self._element_matrix_host(Eb), self._element_matrix_host(Ec)
                _lib.call('femol_plan_set_matrix_free', d.plan, shape[0], shape[1], hb.ctypes.data,
                          hc.ctypes.data if hc is not None else None,
                          _lib.ptr(xpow) if x is not None else None, _lib.ptr(mask), _lib.ptr(vals))
                self._mf_refs = (xpow, mask, vals)
            else:
                _lib.call('femol_plan_set_matrix_free', d.plan, 0, 0, None, None, None, None, None)
        return vals

    # ---- variable-core laminates (FEM.py:697-731, 873-896) -------------------------------------------------
    def _variable_core_tensors(self, zc, h_min):
        """Per-element [A, D, G, B] blocks (n_elem x 31) of the base and of the coating laminate.  For zc > h_min
        the base laminate (full ply list, symetric=False) sits at z = -h/2 + hA0/2 and the coating at
        +h/2 - hA1/2 with h = zc + hA0 + hA1; below, the unshifted base laminate alone.  A shift c of a laminate
        changes its tensors by B = B0 + c A, D = D0 + 2 c B0 + c^2 A (A and G do not move), which vectorises the
        reference's per-element `Layup(...)` construction."""
        from .elements import pack_tensors
        lb, lc = self.layups[0], self.layups[1]
        zc = np.asarray(zc, dtype=np.float64)
        above = zc > h_min
        h = zc + lb.hA + lc.hA
        out = []
        for lay, mat, shift in ((lb, self.materials[0], np.where(above, -h / 2 + lb.hA / 2, 0.0)),
                                (lc, self.materials[1], np.where(above, h / 2 - lc.hA / 2, 0.0))):
            base = Layup(mat, plies=lay.angles, symetric=False)
            A, B0, D0, G = (np.asarray(base.A_mat, float), np.asarray(base.B_mat, float), np.asarray(base.D_mat, float),
                            np.asarray(base.G_mat, float))
            c = shift[:, None, None]
            T = np.zeros((zc.shape[0], 31))
            T[:, 0:9] = np.broadcast_to(A, (zc.shape[0], 3, 3)).reshape(-1, 9)
            T[:, 9:18] = (D0[None] + 2 * c * B0[None] + c * c * A[None]).reshape(-1, 9)
            T[:, 18:22] = np.broadcast_to(G, (zc.shape[0], 2, 2)).reshape(-1, 4)
            T[:, 22:31] = (B0[None] + c * A[None]).reshape(-1, 9)
            out.append(T)
        out[1][~above] = 0.0
        return out[0], out[1], above

    def assemble_variable_core(self, which, h_min=0):
        """K or M from the per-element core heights `mesh.cell_data['zc']` (the data `solve(solve_from_zc=True)`
        assembles, FEM.py:208, 226-227): element = base part + (zc > h_min) * coating part, through the ordinary
        assembly kernel with per-element matrices and the 0/1 indicator as the 'density'."""
        d = self.device
        ctype = self.mesh.contains[0]
        zc = np.asarray(self.mesh.cell_data['zc'][ctype], dtype=np.float64)
        Tb, Tc, above = self._variable_core_tensors(zc, h_min)
        size = d.nn * self.N_dof
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(d.dev)
        if which == 'K':
            E = []
            for T in (Tb, Tc):
                out = torch.empty((d.n_elem, size, size), dtype=torch.float64, device=d.dev)
                t_dev = up(T)            # named: a temporary would be recycled by the caching allocator mid-call
                _lib.call('femol_ke_batch_var', d.etype, self.N_dof, d.n_elem, _lib.ptr(d.xy), _lib.ptr(d.conn), _lib.ptr(t_dev), 0,
                          _lib.ptr(out), _lib.stream())
                E.append(out)
            self.assemble('F')
        else:
            E = [elements.me_batch_device(d.etype, self.N_dof, d.xy, d.conn,
                                          elements.mass_tensor(d.etype, self.N_dof, mat, ho))
                 for mat, ho in ((self.materials[0], self.ho), (self.materials[1], self.coat_ho))]
        vals = torch.empty(d.nnz, dtype=torch.float64, device=d.dev)
        xpow = d.xpow if which == 'K' else d.xpow_m
        indicator, mask = up(above.astype(np.float64)), self._fixed_mask()
        _lib.call('femol_assemble', d.plan, _lib.ptr(E[0]), 1, _lib.ptr(E[1]), 1, _lib.ptr(indicator), 1.0,
                  _lib.ptr(mask), _lib.ptr(xpow), _lib.ptr(vals), _lib.stream())
        self._variable_core = True
        if which == 'K':
            self._vc_refs = (E[0], E[1], xpow)          # what the lattice multigrid coarsens (xpow = the 0/1 indicator)
            self._Kv, self._K_host, self._gmg_stale, self._K_scaled = vals, None, True, False
            _lib.call('femol_plan_set_matrix_free', d.plan, 0, 0, None, None, None, None, None)
        else:
            self._Mv, self._M_host = vals, None

    # ---- assembly -----------------------------------------------------------------------------------
    def assemble(self, which, X=None, p=None, q=None):
        """FEM.py:160-183.  which in 'K' | 'M' | 'F'; X = {cell_type: densities} or None."""
        if which == 'M':
            self._Mv = self._assemble_values('M', X, q)
            self._M_host = None
        elif which == 'K':
            self.assemble('F')
            self._Kv = self._assemble_values('K', X, p)
            self._K_host = None
            self._variable_core = False
        elif which == 'F':
            self._assemble_F()

    def _assemble_F(self):
        """FEM.py:536-544: zero the force at fixed dofs.

        The uploaded vector is cached.  It is refreshed when `problem.F` was rebound (what `add_forces` does,
        FEM.py:464; detected by identity), when it was edited in place (`problem.F` is handed back as a
        `_TrackedVector` whose writes raise a dirty flag) and when the Dirichlet set changed.  (Checksumming the
        4 M-entry vector on every assembly cost 2-3 ms per SIMP iteration on cfg 2.)"""
        key = (self._bc_key(), tuple(id(f) for f in self.forces), tuple(id(d) for d in self.force_domains))
        if self._F_key == key and self._F_ref is self.F and not self.F.dirty:
            return                                   # nothing changed since the last upload
        F = np.array(self.F, dtype=np.float64)       # private copy
        if self.fixed_ddls:
            F[self.fixed_dofs()] = 0
        self._F_dev = torch.from_numpy(F).to(self.device.dev)
        self.F = self._F_ref = _TrackedVector(F)
        self._F_key = key

    @property
    def K(self):
        if self._Kv is None:
            raise AttributeError('K')
        if self._K_host is None:
            self._K_host = self.device.to_scipy(self._Kv)
        return self._K_host

    @property
    def M(self):
        if self._Mv is None:
            raise AttributeError('M')
        if self._M_host is None:
            self._M_host = self.device.to_scipy(self._Mv)
        return self._M_host

    @property
    def K_values(self):
        """CSR values of K on the device (torch, plan order)."""
        return self._Kv

    @property
    def M_values(self):
        return self._Mv

    # ---- solvers --------------------------------------------------------------------------------------
    def solve(self, verbose=True, filtre=0, solve_from_zc=False, h_min=0, modal_solver='lobpcg', sigma=None):
        """FEM.py:189-199.  displacement -> mesh (with mesh.U and point data); modal -> (f[Hz], v[k, ndof])."""
        if solve_from_zc:
            # FEM.py:201-256: the displacement solve re-assembles from `cell_data['zc']` only if K does not exist
            # yet; the modal solve always re-assembles both matrices
            if self.physics == 'modal' or self._Kv is None:
                self.assemble_variable_core('K', h_min)
            if self.physics == 'modal':
                self.assemble_variable_core('M', h_min)
        if self.physics == 'displacement':
            return self._displacement_solve(verbose)
        elif self.physics == 'modal':
            return self._modal_solve(verbose, filtre)

    def solve_device(self):
        """K U = F on the device; returns the device displacement tensor (no host copy)."""
        d = self.device
        if self._Kv is None:
            self.assemble('K')
        # warm start from the previous displacement - unless the Dirichlet set changed since (the old field is not
        # zero on the newly fixed dofs, and an iterative solve would only drive it there to within its tolerance)
        self._fixed_mask()
        if self._U_dev is None or not self.warm_start or self._U_mask_key != self._mask_key:
            self._U_dev = torch.zeros(d.n, dtype=torch.float64, device=d.dev)
            self._U_mask_key = self._mask_key
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        precond = self._effective_preconditioner()
        if precond == 'gmg':
            gmg = self._gmg_ready()
            if self._gmg_ws is None:
                self._gmg_ws = torch.empty(int(_lib.load().femol_pcg_gmg_workspace_doubles(d.plan)), dtype=torch.float64, device=d.dev)
            rc = _lib.call('femol_pcg_gmg', d.plan, gmg, _lib.ptr(self._Kv), _lib.ptr(self._F_dev), _lib.ptr(self._U_dev),
                           float(self.solver_rtol), int(self.solver_max_iter), _lib.ptr(self._gmg_ws),
                           ctypes.byref(iters), ctypes.byref(relres), _lib.stream())
        else:
            rc = _lib.call('femol_pcg', d.plan, _lib.ptr(self._Kv), _lib.ptr(self._F_dev), _lib.ptr(self._U_dev),
                           float(self.solver_rtol), int(self.solver_max_iter), _PRECOND[precond],
                           _lib.ptr(d.pcg_ws()), ctypes.byref(iters), ctypes.byref(relres), _lib.stream())
        # status 0: converged to rtol; 3: recurrence converged and the TRUE residual sits at its attainable
        # floor (~ eps * cond(K), the level a direct solver reaches too); 1: iteration limit
        converged = rc == 0 or (rc == _lib.S_STAGNATED and relres.value <= self.stagnation_bound)
        self.solver_info = {'iterations': iters.value, 'relres': relres.value, 'status': rc,
                            'converged': converged, 'preconditioner': precond}
        if rc == _lib.S_BREAKDOWN:
            raise _lib.FemolError('PCG breakdown (matrix not positive definite?) after %d iterations' % iters.value)
        if not converged:
            why = ('iteration limit %d reached' % self.solver_max_iter if rc == _lib.S_NOT_CONVERGED else
                   'stagnated above stagnation_bound = %.1e' % self.stagnation_bound)
            self._report_nonconvergence('PCG (%s) did not converge: %s after %d iterations, relative residual %.3e '
                                        '(rtol %.1e); the returned displacement is the last iterate'
                                        % (precond, why, iters.value, relres.value, self.solver_rtol))
        return self._U_dev

    # ---- geometric multigrid (structured Q4 meshes) ---------------------------------------------------
    def structured_grid_shape(self):
        """(nx, ny) when the mesh is a rectangle_Q4-ordered structured quad grid (mesh.py:852-860), else None."""
        if self._grid_shape is None:
            shape = False
            m = self.mesh
            if m.structured and m.contains == ['quad'] and m.ElementClasses['quad'].N_nodes == 4:
                nxp = int(np.count_nonzero(m.points[:, 1] == m.points[0, 1]))
                nx = nxp - 1
                if nx >= 1 and m.N_nodes % nxp == 0:
                    ny = m.N_nodes // nxp - 1
                    if ny >= 1 and nx * ny == m.N_ele:
                        a = np.add.outer(np.arange(nx), nxp * np.arange(ny)).reshape(-1)
                        exp = np.stack([a, a + 1, a + nx + 2, a + nx + 1], axis=1)
                        if np.array_equal(np.asarray(m.cells['quad']), exp):
                            shape = (nx, ny)
            self._grid_shape = shape
        return self._grid_shape or None

    def node_lattice(self):
        """Lattice description of a mesh whose nodes occupy every point of a uniform (Nx+1) x (Ny+1) lattice exactly
        once and whose cells (one type: T3, Q4 or T6) each lie inside one cell of the grid over every second lattice
        line - `rectangle_T3`, `rectangle_T6` (corner + mid-side nodes, Nx = 2 nelx), `rectangle_Q4` in any
        numbering - or None.  Feeds femol_gmg_create_lattice: dict(Nx, Ny, lat_node, node_lat, child_ptr, child_idx)."""
        if self._lattice is None:
            self._lattice = self._find_node_lattice() or False
        return self._lattice or None

    def _find_node_lattice(self):
        m = self.mesh
        if len(m.contains) != 1:
            return None
        cells = np.asarray(m.cells[m.contains[0]])
        if cells.shape[1] not in (3, 4, 6) or m.N_nodes >= 2 ** 31:
            return None
        idx = []
        for c in (m.points[:, 0], m.points[:, 1]):
            lo, hi = float(c.min()), float(c.max())
            if not hi > lo:
                return None
            srt = np.sort(c)
            lines = srt[np.concatenate([[True], np.diff(srt) > 1e-9 * (hi - lo)])]
            n = lines.shape[0] - 1
            if n < 1 or not np.allclose(np.diff(lines), (hi - lo) / n, rtol=1e-6, atol=0):
                return None
            k = np.rint((c - lo) / ((hi - lo) / n)).astype(np.int64)
            if np.abs(lo + k * ((hi - lo) / n) - c).max() > 1e-6 * (hi - lo) / n:
                return None
            idx.append((k, n))
        (li, Nx), (lj, Ny) = idx
        if (Nx + 1) * (Ny + 1) != m.N_nodes:
            return None
        lat = lj * (Nx + 1) + li
        if np.bincount(lat, minlength=m.N_nodes).max() != 1:
            return None
        ci, cj = li[cells], lj[cells]
        I, J = ci.min(axis=1) // 2, cj.min(axis=1) // 2
        if (ci.max(axis=1) > 2 * I + 2).any() or (cj.max(axis=1) > 2 * J + 2).any():
            return None                                  # a cell straddles two coarse cells
        nxc, nyc = (Nx + 1) // 2, (Ny + 1) // 2
        parent = I * nyc + J
        order = np.argsort(parent, kind='stable')        # ascending element id inside a parent: deterministic sums
        ptr = np.concatenate([[0], np.cumsum(np.bincount(parent, minlength=nxc * nyc))])
        lat_node = np.empty(m.N_nodes, dtype=np.int32)
        lat_node[lat] = np.arange(m.N_nodes, dtype=np.int32)
        return {'Nx': int(Nx), 'Ny': int(Ny), 'lat_node': lat_node, 'node_lat': lat.astype(np.int32),
                'child_ptr': ptr.astype(np.int32), 'child_idx': order.astype(np.int32)}

    def _effective_preconditioner(self):
        p = self.preconditioner
        # per-element matrices (variable-core laminates, T3 / T6 meshes) have no shared-Ke hierarchy: they take the
        # lattice hierarchy when the nodes form a lattice
        shared = bool(self.structured_grid_shape()) and not self._variable_core
        if p == 'auto':
            ok = self.mesh.N_ele >= 64 and (shared or self.node_lattice() is not None)
            return 'gmg' if ok else 'block_jacobi'
        if p == 'gmg' and not shared and self.node_lattice() is None:
            raise _lib.FemolError("preconditioner='gmg' needs a mesh whose nodes form a lattice (rectangle_Q4 / T3 / T6)")
        return p

    def _gmg_ready(self):
        """Create (once per Dirichlet set) and refresh (once per assembled K) the multigrid hierarchy."""
        d = self.device
        mask = self._fixed_mask()
        lattice = self._variable_core or not self.structured_grid_shape()
        if lattice:
            return self._gmg_ready_lattice(mask)
        hourglass = self._hourglass_enabled()
        if (self._gmg is None or self._gmg_mask_key != self._mask_key or hourglass != (self._gmg2 is not None)
                or self._gmg_is_lattice):
            self._destroy_gmg()
            nx, ny = self.structured_grid_shape()
            handle = ctypes.c_void_p()
            _lib.call('femol_gmg_create', nx, ny, self.N_dof, d.plan, _lib.ptr(mask), int(self.gmg_smoothing_steps),
                      float(self.gmg_omega), _lib.stream(), ctypes.byref(handle))
            self._gmg, self._gmg_mask_key, self._gmg_stale, self._gmg_is_lattice = handle, self._mask_key, True, False
            if hourglass:
                view, h2 = ctypes.c_void_p(), ctypes.c_void_p()
                _lib.call('femol_plan_create_view', d.plan, ctypes.byref(view))
                self._plan_view = view
                _lib.call('femol_gmg_create', nx, ny, self.N_dof, view, _lib.ptr(mask), int(self.gmg_smoothing_steps),
                          float(self.gmg_omega), _lib.stream(), ctypes.byref(h2))
                self._gmg2 = h2
                _lib.call('femol_gmg_attach_modulated', self._gmg, h2)
        if self._gmg_stale:
            Eb, _ = self._element_matrices('K', False)
            Ec = self._element_matrices('K', True)[0] if self.coating else None
            xp = _lib.ptr(d.xpow) if self._K_scaled else None
            _lib.call('femol_gmg_setup', self._gmg, _lib.ptr(Eb), _lib.ptr(Ec), xp, _lib.ptr(self._Kv), _lib.stream())
            if self._gmg2 is not None:
                # S_e Ke S_e with S_e = (+, -, +, -) over the element's nodes (Q4 order (0,0) (1,0) (1,1) (0,1)); the
                # view's matrix-free descriptor is tagged with K's value array: level 0 of the second hierarchy is
                # applied matrix-free, and its block-diagonal inverse is that of K (S is constant within a node)
                sgn = torch.tensor([1.0, -1.0, 1.0, -1.0], dtype=torch.float64, device=d.dev).repeat_interleave(self.N_dof)
                flip = sgn[:, None] * sgn[None, :]
                Ebh = (Eb[:1] * flip).contiguous()
                Ech = (Ec[:1] * flip).contiguous() if Ec is not None else None
                hb = np.ascontiguousarray(Ebh[0].cpu().numpy().reshape(-1))
                hc = np.ascontiguousarray(Ech[0].cpu().numpy().reshape(-1)) if Ech is not None else None
                nx, ny = self.structured_grid_shape()
                _lib.call('femol_plan_set_matrix_free', self._plan_view, nx, ny, hb.ctypes.data,
                          hc.ctypes.data if hc is not None else None, xp, _lib.ptr(mask), _lib.ptr(self._Kv))
                _lib.call('femol_gmg_setup', self._gmg2, _lib.ptr(Ebh), _lib.ptr(Ech), xp, _lib.ptr(self._Kv), _lib.stream())
                self._gmg2_refs = (Ebh, Ech, mask)
            self._gmg_stale = False
        return self._gmg

    def _gmg_ready_lattice(self, mask):
        """Lattice hierarchy (femol_gmg_create_lattice): level 0 = the assembled operator of this mesh, element
        matrices shared or one per element (T3 / T6 meshes, variable-core laminates)."""
        d = self.device
        if self._gmg is None or self._gmg_mask_key != self._mask_key or not self._gmg_is_lattice:
            self._destroy_gmg()
            lat = self.node_lattice()
            dev = {k: torch.from_numpy(lat[k]).to(d.dev) for k in ('lat_node', 'node_lat', 'child_ptr', 'child_idx')}
            handle = ctypes.c_void_p()
            _lib.call('femol_gmg_create_lattice', lat['Nx'], lat['Ny'], self.N_dof, d.plan, _lib.ptr(mask),
                      _lib.ptr(dev['lat_node']), _lib.ptr(dev['node_lat']), _lib.ptr(dev['child_ptr']),
                      _lib.ptr(dev['child_idx']), int(self.gmg_smoothing_steps), float(self.gmg_omega), _lib.stream(),
                      ctypes.byref(handle))
            torch.cuda.current_stream().synchronize()      # the library copies the index arrays in stream order
            self._gmg, self._gmg_mask_key, self._gmg_stale, self._gmg_is_lattice = handle, self._mask_key, True, True
        if self._gmg_stale:
            if self._variable_core:
                Eb, Ec, xp_t = self._vc_refs
                pe_b = pe_c = 1
                xp = _lib.ptr(xp_t)
            else:
                Eb, pe_b = self._element_matrices('K', False)
                Ec, pe_c = self._element_matrices('K', True) if self.coating else (None, 0)
                xp = _lib.ptr(d.xpow) if self._K_scaled else None
            _lib.call('femol_gmg_setup_lattice', self._gmg, _lib.ptr(Eb), pe_b, _lib.ptr(Ec), pe_c, xp, _lib.ptr(self._Kv),
                      _lib.stream())
            self._gmg_stale = False
        return self._gmg

    def _hourglass_enabled(self):
        h = self.gmg_hourglass
        if self.N_dof != 6 or not self.matrix_free or not self.structured_grid_shape() or not h:
            return False
        if h != 'auto':
            return True
        nx, ny = self.structured_grid_shape()
        self._fixed_mask()
        if self._mask_host is None:
            return float(self.ho) >= 8.0 * float(self.mesh.element_size())
        m = self._mask_host.reshape(ny + 1, nx + 1)
        edge = np.concatenate([m[0], m[-1], m[:, 0], m[:, -1]])
        free_edge = bool(((edge >> 2) & 1).min() == 0)      # some boundary node is free to move in w
        ratio = float(self.ho) / float(self.mesh.element_size())
        # fully supported plates: the family is restrained by the boundary, but between ~8 and ~50 element sizes of
        # thickness it still dominates the spectrum (n = 200, t/h = 20: simply supported 90 -> 33 iterations, clamped
        # 105 -> 29; t/h = 100: 15 -> 23 and 35 -> 18; t/h = 4: 177 -> 137 and 147 -> 113 - profiles/r01_plate_hourglass.md)
        return ratio >= 8.0 if free_edge else 8.0 <= ratio <= 50.0

    def _gmg_workspace_views(self, gmg, m):
        """m workspace views of the current hierarchy (femol_gmg_clone_workspace), created once and kept until the
        hierarchy is destroyed: LOBPCG solves the m columns of its preconditioner concurrently, one view per column."""
        if gmg is None or m < 2:
            return None
        views = getattr(self, '_gmg_views', None)
        if views is None:
            views = self._gmg_views = []
        try:
            while len(views) < m:
                h = ctypes.c_void_p()
                _lib.call('femol_gmg_clone_workspace', gmg, ctypes.byref(h))
                views.append(h)
        except _lib.FemolError:
            return None                      # no memory for more work vectors: sequential column solves
        return views[:m]

    def _destroy_gmg(self):
        L = _lib.load()
        for h in getattr(self, '_gmg_views', None) or []:       # the views borrow the hierarchy: they go first
            L.femol_gmg_destroy(h)
        self._gmg_views = []
        if getattr(self, '_gmg', None) is not None:
            L.femol_gmg_destroy(self._gmg)
            self._gmg = None
        if getattr(self, '_gmg2', None) is not None:
            L.femol_gmg_destroy(self._gmg2)
            self._gmg2 = None
        if getattr(self, '_plan_view', None) is not None:
            L.femol_plan_destroy(self._plan_view)
            self._plan_view = None

    def _displacement_solve(self, verbose):
        """FEM.py:201-214"""
        U = self.solve_device().cpu().numpy()
        if verbose:
            print('Solved with femol_pcg: %(iterations)d iterations, relative residual %(relres).2e' % self.solver_info)
        self.mesh.add_displacement(U, self.N_dof)
        return self.mesh

    def modal_device(self, k=None, X0=None):
        """Lowest k physical eigenpairs on the device: (lambda[k] (rad/s)^2 host, V [k][n] device tensor)."""
        d = self.device
        if self._Mv is None:
            self.assemble('M')
        if self._Kv is None:
            self.assemble('K')
        k = k or self.n_modes
        m = _lobpcg.pick_block_size(k)
        mask = self._fixed_mask()
        if self.N_dof == 6 and self.deflate_drilling:
            drill = torch.full((d.n_nodes,), 1 << 5, dtype=torch.uint8, device=d.dev)
            mask = drill if mask is None else (mask | drill)
        if X0 is None and self._X_modal is not None and self._X_modal.shape == (d.n, m) and self.warm_start:
            X0 = self._X_modal
        if X0 is None:
            g = torch.Generator(device=d.dev)
            g.manual_seed(1234)
            X0 = torch.randn((d.n, m), dtype=torch.float64, device=d.dev, generator=g)
        X0 = X0.clone()
        if mask is not None:  # fixed dofs are projected out (the reference keeps them as lambda = 1 pairs)
            bits = (mask.to(torch.int32)[:, None] >> torch.arange(self.N_dof, device=d.dev, dtype=torch.int32)[None]) & 1
            X0[bits.reshape(-1).bool()] = 0.0
        inner = self.preconditioner if self.preconditioner in ('none', 'jacobi', 'block_jacobi') else 'block_jacobi'
        # structured grids: T ~ K^-1 by a few multigrid-preconditioned CG iterations per column; otherwise by
        # `modal_inner_iters` block-CG iterations with the (block-)diagonal inverse
        gmg = self._gmg_ready() if (self.preconditioner in ('auto', 'gmg') and self._effective_preconditioner() == 'gmg') else None
        ops = _lobpcg.DeviceOps(d.plan, self._Kv, self._Mv, mask, m, _PRECOND[inner], X0,
                                inner_iters=self.modal_inner_iters, inner_tol=self.modal_inner_tol, gmg=gmg,
                                gmg_iters=self.modal_gmg_iters, gmg_views=self._gmg_workspace_views(gmg, m))
        lam, rel, its = _lobpcg.lobpcg(ops, m, k, tol=self.modal_tol, max_iter=self.modal_max_iter)
        V = ops.vectors()
        self._X_modal = V.clone()
        # converged to modal_tol, or stopped at the attainable accuracy (LOBPCG's stall rule) below stagnation_bound
        worst = float(np.max(rel[:k])) if k else 0.0
        converged = bool(np.all(np.isfinite(lam[:k]))) and worst <= max(self.modal_tol, self.stagnation_bound)
        self.solver_info = {'iterations': its, 'relres': rel[:k].copy(), 'block': m, 'converged': converged}
        if not converged:
            self._report_nonconvergence('LOBPCG did not converge: %d iterations (limit %d), worst relative residual '
                                        '%.3e over the %d requested pairs (tol %.1e); the returned pairs are the last '
                                        'Ritz pairs' % (its, self.modal_max_iter, worst, k, self.modal_tol))
        return lam[:k], V[:, :k].t().contiguous()

    def _modal_solve(self, verbose, filtre=0):
        """FEM.py:216-256: returns (sqrt(lambda)/2pi [Hz], v) with v rows = M-orthonormal modes.

        The eigen-filters (FEM.py:335-418) are heuristics the reference needs to clean the FULL dense spectrum
        (lambda = 1 pairs of the fixed dofs, drilling / mesh modes).  The block eigen-solver projects those out
        and returns the lowest `n_modes` physical pairs, so filter 0 is the natural choice; 1 and 2 are applied
        to the returned pairs with the reference's own rules for callers that pass them."""
        lam, V = self.modal_device()
        if verbose:
            print('Solved with LOBPCG: %(iterations)d iterations' % self.solver_info)
        w, v = lam, V.cpu().numpy()
        keep = ~np.isnan(w)
        w, v = w[keep], v[keep]
        keep = ~np.isclose(w, 0) & (w > 0)
        w, v = w[keep], v[keep]
        w, v = {0: lambda a, b: (a, b), 1: self._filter_eigenvalues_1, 2: self._filter_eigenvalues_2}[filtre](w, v)
        return np.sqrt(w) / (2 * np.pi), v

    # ---- eigen-filters of the reference (host post-processing of the returned pairs) --------------------
    def _filter_eigenvalues_1(self, w, v):
        """FEM.py:343-351 = low-range duplicates (:405-418) then mesh modes (:392-403)."""
        f = np.sqrt(w) / (2 * np.pi)
        distinct = np.flatnonzero(~np.isclose(f[:-1], f[1:], 10e-5))      # third positional argument = rtol
        start = distinct[1]                                               # IndexError like the reference if < 2
        w, v = w[start:], v[start:]
        if self.N_dof != 6:
            # the reference's _filter_mesh_modes returns None for plane models and fails on unpacking
            raise TypeError('cannot unpack non-iterable NoneType object')
        live = ~np.isclose([np.std(vi[2::6]) for vi in v], 0)
        return w[live], v[live]

    def _filter_eigenvalues_2(self, w, v):
        """FEM.py:353-390: skip runs of frequencies within 32 Hz of the run's first one, keep a run's first
        frequency if it is above 1.1 Hz and its vector is not too peaked (max / mean|.| < 5)."""
        f = np.sqrt(w) / (2 * np.pi)
        n = f.shape[0]
        kept = []
        i = 0
        while i < n - 2:
            j = 1
            while np.isclose(f[i], f[i + j], atol=32):
                j += 1
                if i + j == n:
                    break
            if f[i] > 1.1 and np.abs(v[i].max()) / np.abs(v[i]).mean() < 5:
                kept.append(i)
            i += j
        kept = np.array(kept, dtype=int)
        return (f[kept] * 2 * np.pi) ** 2, v[kept]
