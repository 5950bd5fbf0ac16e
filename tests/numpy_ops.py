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

    def precondition(self):
        m = self.m
        W = self.apply_T(self._a('S')[:, m:2 * m])
        W[~self.free] = 0
        self._a('S')[:, m:2 * m] = W

    def vectors(self):
        return self._a('S')[:, :self.m]
