"""CPU stand-in for femol_b200.lobpcg.DeviceOps (TEST INFRASTRUCTURE: lets the LOBPCG driver logic be
exercised without a GPU, the way gloo stands in for NCCL).  Same method contract, numpy/scipy arithmetic."""
import numpy as np


class NumpyOps(object):
    def __init__(self, K, M, fixed_idx, m, X0, precond='jacobi'):
        n = K.shape[0]
        self.K, self.M, self.m, self.n = K.tocsr(), M.tocsr(), m, n
        self.free = np.ones(n, bool); self.free[fixed_idx] = False
        z = lambda: np.zeros((n, 3 * m))
        self.buf = [{'S': z(), 'KS': z(), 'MS': z()}, {'S': z(), 'KS': z(), 'MS': z()}]
        self.cur = 0
        X0 = X0.copy(); X0[~self.free] = 0
        self.buf[0]['S'][:, :m] = X0
        d = self.K.diagonal()
        if precond == 'jacobi':
            dinv = 1.0 / np.where(d != 0, d, 1.0)
            self.apply_T = lambda R: R * dinv[:, None]
        elif precond == 'exact':  # T = K^-1 (sparse LU): the ideal preconditioner, for testing the driver
            import scipy.sparse.linalg as sla
            lu = sla.splu(self.K.tocsc())
            self.apply_T = lambda R: lu.solve(np.ascontiguousarray(R))
        else:
            self.apply_T = lambda R: R

    def _a(self, name, other=False):
        return self.buf[self.cur ^ (1 if other else 0)][name]

    def spmm(self, which, col, m):
        A = self.K if which == 'K' else self.M
        dst = 'KS' if which == 'K' else 'MS'
        self._a(dst)[:, col:col + m] = A @ self._a('S')[:, col:col + m]

    def gram(self, a, b, ca, qa, cb, qb):
        return self._a(a)[:, ca:ca + qa].T @ self._a(b)[:, cb:cb + qb]

    def combine(self, C, src_cols, dst_col, swap):
        c0, q = src_cols[0], C.shape[0]
        for name in ('S', 'KS', 'MS'):
            self._a(name, other=True)[:, dst_col:dst_col + C.shape[1]] = self._a(name)[:, c0:c0 + q] @ C
        if swap:
            self.cur ^= 1

    def axpy_cols(self, G, src_cols, dst_col, arrays):
        c0, q = src_cols[0], G.shape[0]
        for name in arrays:
            self._a(name)[:, dst_col:dst_col + G.shape[1]] -= self._a(name)[:, c0:c0 + q] @ G

    def scale_cols(self, C, col, arrays):
        q = C.shape[0]
        for name in arrays:
            self._a(name)[:, col:col + C.shape[1]] = self._a(name)[:, col:col + q] @ C

    def residual(self, lam):
        m = self.m
        self._a('S')[:, m:2 * m] = self._a('KS')[:, :m] - self._a('MS')[:, :m] * lam[None, :]

    def precondition(self, skip=None):
        m = self.m
        W = self.apply_T(self._a('S')[:, m:2 * m])
        W[~self.free] = 0
        self._a('S')[:, m:2 * m] = W

    def vectors(self):
        return self._a('S')[:, :self.m]


def numpy_lattice_vcycle(K, lat, fixed_dofs, r, omega, nd=2, return_levels=False):
    """V(1,1) cycle with damped block-Jacobi on Galerkin operators; the numpy statement of gmg.cu for a lattice mesh.
    return_levels: (z, levels, cyc) with levels[l] = (A_l, Dinv_l [n_nodes][nd][nd], P_l or None at the coarsest level)
    and cyc(l, r_l) the cycle started at level l (tests/test_dist.py runs the upper levels strip-distributed)."""
    import scipy.sparse as sp

    def p1(n):
        nc = (n + 1) // 2
        P = np.zeros((n + 1, nc + 1))
        for i in range(n + 1):
            if i % 2 == 0: P[i, i // 2] = 1
            elif i == n: P[i, (i + 1) // 2] = 1
            else: P[i, i // 2] = .5; P[i, i // 2 + 1] = .5
        return sp.csr_matrix(P)

    def dinv(A):
        n = A.shape[0] // nd
        B = A.tobsr(blocksize=(nd, nd)); B.sort_indices()
        D = np.zeros((n, nd, nd))
        for i in range(n):
            for k in range(B.indptr[i], B.indptr[i + 1]):
                if B.indices[k] == i: D[i] = B.data[k]
        return np.linalg.inv(D)

    def jac(A, D, rr, e):
        res = rr if e is None else rr - A @ e
        dz = omega * np.einsum('nab,nb->na', D, res.reshape(-1, nd)).reshape(-1)
        return dz if e is None else e + dz

    levels = []
    A, Nx, Ny, perm = K, lat['Nx'], lat['Ny'], lat['node_lat']
    fx = np.zeros(K.shape[0], bool); fx[fixed_dofs] = True
    while True:
        if A.shape[0] <= 100 or (Nx == 1 and Ny == 1):
            levels.append((A, None, None)); break
        P2 = sp.kron(p1(Ny), p1(Nx)).tocsr()
        if perm is not None:
            P2 = P2[perm]
        P = sp.diags((~fx).astype(float)) @ sp.kron(P2, sp.identity(nd)).tocsr()
        nxc, nyc = (Nx + 1) // 2, (Ny + 1) // 2
        I, J = np.meshgrid(np.arange(nxc + 1), np.arange(nyc + 1))
        fl = (np.minimum(2 * J, Ny) * (Nx + 1) + np.minimum(2 * I, Nx)).reshape(-1)
        if perm is not None:
            inv = np.empty_like(perm); inv[perm] = np.arange(perm.shape[0]); fl = inv[fl]
        fc = fx.reshape(-1, nd)[fl].reshape(-1)
        keep = sp.diags((~fc).astype(float))
        P = P @ keep
        Ac = (P.T @ A @ P + sp.diags(fc.astype(float))).tocsr()
        levels.append((A, dinv(A), P))
        A, Nx, Ny, perm, fx = Ac, nxc, nyc, None, fc

    def cyc(l, rr):
        A, D, P = levels[l]
        if P is None:
            # the device inverse decouples dead coarse dofs; none here
            return np.linalg.solve(A.toarray(), rr)
        e = jac(A, D, rr, None)
        e = e + P @ cyc(l + 1, P.T @ (rr - A @ e))
        return jac(A, D, rr, e)
    z = cyc(0, r)
    return (z, levels, cyc) if return_levels else z
