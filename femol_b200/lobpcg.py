"""LOBPCG outer iteration for the lowest generalised eigenpairs  K v = lambda M v.

Replaces the reference's dense `scipy.linalg.eigh(K.toarray(), M.toarray())` (FEM.py:277) and its
ARPACK option (FEM.py:234).  All O(n) work goes through an `ops` object; the product implementation
(`DeviceOps`) calls the CUDA block kernels of csrc/lobpcg.cu through the C-ABI.  Only the 3m x 3m
Rayleigh-Ritz pencil (m <= 16) is handled on the host, as every GPU LOBPCG does.

Block vectors live in three [n][3m] row-major device arrays  S = [X | W | P],  KS = K S,  MS = M S
(double buffered, so the new X and P are written straight into the other buffer: no copies).
"""
import numpy as np
import scipy.linalg

ALLOWED_M = (1, 2, 4, 8, 12, 16)


def pick_block_size(k):
    for m in ALLOWED_M:
        if m >= k + min(2, k):
            return m
    raise ValueError('at most %d eigenpairs per call' % (ALLOWED_M[-1] - 2))


def _rayleigh_ritz(A, B, m):
    """Lowest m eigenpairs of the small pencil (A, B) with diagonal scaling; raises LinAlgError when B is
    numerically not positive definite."""
    A = 0.5 * (A + A.T)
    B = 0.5 * (B + B.T)
    d = 1.0 / np.sqrt(np.maximum(np.abs(np.diag(B)), 1e-300))
    As, Bs = A * d[:, None] * d[None, :], B * d[:, None] * d[None, :]
    # whiten with the eigen-decomposition of B and drop numerically dependent directions, so a
    # nearly singular basis (W or P almost inside span(X)) cannot break the iteration
    s, V = np.linalg.eigh(Bs)
    keep = s > 1e-11 * s.max()
    if keep.sum() < m:
        raise np.linalg.LinAlgError('LOBPCG basis has rank %d < block size %d' % (keep.sum(), m))
    T = V[:, keep] / np.sqrt(s[keep])[None, :]
    At = T.T @ As @ T
    w, c = np.linalg.eigh(0.5 * (At + At.T))
    return w[:m], (T @ c[:, :m]) * d[:, None]


def lobpcg(ops, m, k, tol=1e-8, max_iter=500, verbose=False, stall_window=5):
    """Run LOBPCG on `ops` (which already holds the initial block in S[:, :m]).

    Returns (lambda[m], rel_resid[m], iterations).  On exit ops.S[:, :m] holds M-orthonormal Ritz vectors,
    ascending.  Convergence: ||K v - l M v||_2 / ||K v||_2 <= tol for the first k pairs.
    """
    q3 = 3 * m
    # M-orthonormalise the start block and do a first Rayleigh-Ritz on span(X)
    ops.spmm('M', 0, m)
    ops.spmm('K', 0, m)
    A = ops.gram('S', 'KS', 0, m, 0, m)
    B = ops.gram('S', 'MS', 0, m, 0, m)
    lam, C = _rayleigh_ritz(A, B, m)
    ops.combine(C, src_cols=(0, m), dst_col=0, swap=True)      # X, KX, MX <- (.) C   (into the other buffer)
    have_p = False
    rel = np.full(m, np.inf)
    it = 0
    lam_prev = lam.copy()
    best, since_best = np.inf, 0
    for it in range(1, max_iter + 1):
        ops.residual(lam)                                       # R -> W slot
        rn = np.sqrt(np.diag(ops.gram('S', 'S', m, m, m, m)))
        kn = np.sqrt(np.diag(ops.gram('KS', 'KS', 0, m, 0, m)))
        rel = rn / np.where(kn > 0, kn, 1.0)
        if verbose:
            print('lobpcg it %d  lambda[:k]=%s  rel=%s' % (it, lam[:k], rel[:k]))
        if np.all(rel[:k] <= tol):
            it -= 1
            break
        # Attainable accuracy: ||K v - l M v|| / ||K v|| cannot drop below ~eps * cond.  Once the Ritz values
        # have stopped moving (1e-11 relative) AND the worst residual has set no new minimum (by 2 %) for
        # `stall_window` iterations, further iterations only add noise (the residuals drift up again).  Slow but
        # steady convergence keeps setting new minima and is not stopped.
        worst = float(np.max(rel[:k]))
        if worst < 0.98 * best:
            best, since_best = worst, 0
        else:
            since_best += 1
        if since_best >= stall_window and np.all(np.abs(lam[:k] - lam_prev[:k]) <= 1e-11 * np.abs(lam[:k])):
            it -= 1
            break
        lam_prev = lam.copy()
        # soft locking: a pair whose residual already meets the tolerance keeps its (tiny) residual as search
        # direction instead of an expensive T ~ K^-1 application; it is re-activated if its residual grows again
        locked = np.zeros(m, dtype=bool)
        locked[:k] = rel[:k] <= tol
        ops.precondition(skip=locked)                           # W <- T R  (fixed dofs zeroed)
        # M-orthogonalise W against X, then M-orthonormalise W (Cholesky)
        ops.spmm('M', m, m)
        G = ops.gram('S', 'MS', 0, m, m, m)                     # X' M W
        ops.axpy_cols(G, src_cols=(0, m), dst_col=m, arrays=('S', 'MS'))   # W -= X G ; MW -= MX G
        Gw = ops.gram('S', 'MS', m, m, m, m)
        Gw = 0.5 * (Gw + Gw.T)
        try:
            L = np.linalg.cholesky(Gw)
            Ci = np.linalg.inv(L).T
        except np.linalg.LinAlgError:
            d = 1.0 / np.sqrt(np.maximum(np.diag(Gw), 1e-300))
            Ci = np.diag(d)
        ops.scale_cols(Ci, col=m, arrays=('S', 'MS'))           # W <- W Ci ; MW <- MW Ci
        ops.spmm('K', m, m)
        q = q3 if have_p else 2 * m
        A = ops.gram('S', 'KS', 0, q, 0, q)
        B = ops.gram('S', 'MS', 0, q, 0, q)
        if not (np.isfinite(A).all() and np.isfinite(B).all()):
            break                                                # keep the last good block
        try:
            lam_new, C = _rayleigh_ritz(A, B, m)
        except (np.linalg.LinAlgError, ValueError):
            if not have_p:
                break
            # ill-conditioned basis: restart without the conjugate directions
            q = 2 * m
            lam_new, C = _rayleigh_ritz(A[:q, :q], B[:q, :q], m)
        lam = lam_new
        # new P = [W P] C[m:], new X = S C ; written into the alternate buffers, then swap
        Cp = C.copy()
        Cp[:m, :] = 0.0
        ops.combine(Cp, src_cols=(0, q), dst_col=2 * m, swap=False)
        ops.combine(C, src_cols=(0, q), dst_col=0, swap=True)
        have_p = True
    return lam, rel, it


class DeviceOps(object):
    """CUDA implementation of the block operations (csrc/lobpcg.cu through the C-ABI)."""

    def __init__(self, plan, k_vals, m_vals, fixed_mask, m, precond_kind, X0, inner_iters=0, inner_tol=1e-2, gmg=None,
                 gmg_iters=8, gmg_views=None):
        """precond_kind: diagonal / block-diagonal inverse used directly (inner_iters = 0) or as the inner
        preconditioner of `inner_iters` block-CG iterations on K (T ~ K^-1, femol_block_pcg).
        gmg: multigrid hierarchy of K (femol_gmg handle, structured grids): T ~ K^-1 by at most `gmg_iters`
        multigrid-preconditioned CG iterations per column (femol_pcg_gmg) instead of the block-CG.
        gmg_views: m workspace views of `gmg` (femol_gmg_clone_workspace, owned by the caller): the columns are then
        solved concurrently, one stream and one host thread per column."""
        import torch
        from . import _lib
        self._lib, self._torch = _lib, torch
        self.plan, self.k_vals, self.m_vals, self.fixed = plan, k_vals, m_vals, fixed_mask
        self.m = m
        self.inner_iters, self.inner_tol = int(inner_iters), float(inner_tol)
        self.gmg, self.gmg_iters = gmg, int(gmg_iters)
        self._given_views = list(gmg_views) if gmg_views else []
        self._bp_ws = None
        self._gmg_ws = None
        self._keep = None
        self._views, self._streams, self._pool = [], [], None
        n = X0.shape[0]
        self.n = n
        dev = X0.device
        z = lambda: torch.zeros((n, 3 * m), dtype=torch.float64, device=dev)
        self.buf = [{'S': z(), 'KS': z(), 'MS': z()}, {'S': z(), 'KS': z(), 'MS': z()}]
        self.cur = 0
        self.buf[0]['S'][:, :m] = X0
        self.small = torch.zeros(48 * 48, dtype=torch.float64, device=dev)
        self.lam_dev = torch.zeros(m, dtype=torch.float64, device=dev)
        self.gram_ws = torch.zeros(int(_lib.load().femol_gram_workspace_doubles(48, 48)), dtype=torch.float64, device=dev)
        self.gram_out = torch.zeros(48 * 48, dtype=torch.float64, device=dev)
        ndof = int(n // self._n_nodes())
        self.dinv = torch.empty(self._n_nodes() * ndof * ndof, dtype=torch.float64, device=dev)
        _lib.call('femol_precond_setup', plan, _lib.ptr(k_vals), precond_kind, _lib.ptr(self.dinv), _lib.stream())

    def _n_nodes(self):
        import ctypes
        info = (ctypes.c_int64 * 8)()
        self._lib.call('femol_plan_info', self.plan, info)
        return int(info[5])

    def _p(self, name, col=0, other=False):
        t = self.buf[self.cur ^ (1 if other else 0)][name]
        return t.data_ptr() + 8 * col

    def spmm(self, which, col, m):
        vals = self.k_vals if which == 'K' else self.m_vals
        dst = 'KS' if which == 'K' else 'MS'
        L = self._lib
        L.call('femol_spmm', self.plan, L.ptr(vals), self._p('S', col), 3 * self.m, self._p(dst, col), 3 * self.m, m, L.stream())

    def gram(self, a, b, ca, qa, cb, qb):
        L = self._lib
        L.call('femol_gram', self.n, qa, qb, self._p(a, ca), 3 * self.m, self._p(b, cb), 3 * self.m,
               L.ptr(self.gram_out), L.ptr(self.gram_ws), L.stream())
        return self.gram_out[:qa * qb].cpu().numpy().reshape(qa, qb)

    def _upload(self, C):
        C = np.ascontiguousarray(C, dtype=np.float64)
        self.small[:C.size].copy_(self._torch.from_numpy(C.reshape(-1)), non_blocking=False)
        return C.shape

    def combine(self, C, src_cols, dst_col, swap):
        """dst[:, dst_col:dst_col+mc] (other buffer) = cur[:, c0:c0+q] C for S, KS, MS."""
        L = self._lib
        q, mc = self._upload(C)
        c0 = src_cols[0]
        for name in ('S', 'KS', 'MS'):
            L.call('femol_block_gemm', self.n, q, mc, self._p(name, c0), 3 * self.m, L.ptr(self.small),
                   self._p(name, dst_col, other=True), 3 * self.m, 1.0, 0.0, L.stream())
        if swap:
            self.cur ^= 1

    def axpy_cols(self, G, src_cols, dst_col, arrays):
        """arr[:, dst] -= arr[:, src] G"""
        L = self._lib
        q, mc = self._upload(G)
        for name in arrays:
            L.call('femol_block_gemm', self.n, q, mc, self._p(name, src_cols[0]), 3 * self.m, L.ptr(self.small),
                   self._p(name, dst_col), 3 * self.m, -1.0, 1.0, L.stream())

    def scale_cols(self, C, col, arrays):
        """arr[:, col:col+m] <- arr[:, col:col+m] C   (in place)"""
        L = self._lib
        q, mc = self._upload(C)
        for name in arrays:
            L.call('femol_block_gemm', self.n, q, mc, self._p(name, col), 3 * self.m, L.ptr(self.small),
                   self._p(name, col), 3 * self.m, 1.0, 0.0, L.stream())

    def residual(self, lam):
        L = self._lib
        self.lam_dev.copy_(self._torch.from_numpy(np.ascontiguousarray(lam, dtype=np.float64)))
        L.call('femol_block_residual', self.n, self.m, self._p('KS', 0), 3 * self.m, self._p('MS', 0), 3 * self.m,
               L.ptr(self.lam_dev), self._p('S', self.m), 3 * self.m, L.stream())

    def _gmg_setup_concurrent(self):
        """One workspace view of the hierarchy (given by the caller), one stream, one PCG workspace and one host thread
        per column: a single multigrid-CG solve is a chain of small latency-bound kernels, m of them side by side fill
        the GPU.  FEMOL_MODAL_CONCURRENT=0 (or no views): one column after the other on the current stream."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        L, torch = self._lib, self._torch
        dev = self.buf[0]['S'].device
        nd = int(L.load().femol_pcg_gmg_workspace_doubles(self.plan))
        self._gmg_ws = [torch.empty(nd, dtype=torch.float64, device=dev)]
        if self.fixed is not None:
            ndof = self.n // self._n_nodes()
            bits = (self.fixed.to(torch.int32)[:, None] >> torch.arange(ndof, device=dev, dtype=torch.int32)[None]) & 1
            self._keep = (1 - bits).reshape(-1).to(torch.float64)
        self._views, self._streams, self._pool = [], [], None
        if os.environ.get('FEMOL_MODAL_CONCURRENT', '1') != '0' and self.m > 1 and len(self._given_views) >= self.m:
            self._views = self._given_views[:self.m]
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.m)]
            self._gmg_ws = [torch.empty(nd, dtype=torch.float64, device=dev) for _ in range(self.m)]
            self._pool = ThreadPoolExecutor(max_workers=self.m)

    def __del__(self):
        try:
            if self._pool is not None:
                self._pool.shutdown(wait=True)
        except Exception:
            pass

    def _solve_column(self, c, gmg, ws, stream):
        """z = (approximately) K^-1 S[:, c] by at most gmg_iters multigrid-CG iterations, written back to S[:, c]."""
        import ctypes
        L, torch = self._lib, self._torch
        S = self.buf[self.cur]['S']
        it, rr = ctypes.c_int(0), ctypes.c_double(0.0)
        r = S[:, c].contiguous()
        z = torch.zeros_like(r)
        L.call('femol_pcg_gmg', self.plan, gmg, L.ptr(self.k_vals), L.ptr(r), L.ptr(z), self.inner_tol, self.gmg_iters, L.ptr(ws),
               ctypes.byref(it), ctypes.byref(rr), stream)
        if self._keep is not None:
            z.mul_(self._keep)
        S[:, c] = z

    def _precondition_gmg(self, skip=None):
        """W <- P K^-1 W column by column (P: zero the masked dofs), K^-1 by a few multigrid-preconditioned CG
        iterations (a variable preconditioner, which LOBPCG tolerates); the columns are solved concurrently, each on its
        own stream with its own workspace view of the hierarchy."""
        L, torch = self._lib, self._torch
        m = self.m
        if self._gmg_ws is None:
            self._gmg_setup_concurrent()
        cols = [c for c in range(m, 2 * m) if not (skip is not None and skip[c - m])]   # locked pairs keep W = R
        if not self._views:
            for c in cols:
                self._solve_column(c, self.gmg, self._gmg_ws[0], L.stream())
            return
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        dev_index = self.buf[0]['S'].device.index

        def work(c):
            k = c - m
            torch.cuda.set_device(dev_index)
            s = self._streams[k]
            s.wait_event(start)
            with torch.cuda.stream(s):
                self._solve_column(c, self._views[k], self._gmg_ws[k], s.cuda_stream)
                done = torch.cuda.Event()
                done.record(s)
            return done

        for done in list(self._pool.map(work, cols)):
            main.wait_event(done)

    def precondition(self, skip=None):
        """W <- T W.  skip[c] (optional, per column): leave column c as it is (soft locking of converged pairs; only the
        column-by-column multigrid path can skip, the block kernels treat all columns together)."""
        L = self._lib
        fixed = L.ptr(self.fixed) if self.fixed is not None else None
        if self.gmg is not None:
            return self._precondition_gmg(skip)
        if self.inner_iters > 0:
            if self._bp_ws is None:
                nd = int(L.load().femol_block_pcg_workspace_doubles(self.plan, self.m))
                self._bp_ws = self._torch.empty(nd, dtype=self._torch.float64, device=self.k_vals.device)
            L.call('femol_block_pcg', self.plan, L.ptr(self.k_vals), L.ptr(self.dinv), fixed, self._p('S', self.m),
                   3 * self.m, self._p('S', self.m), 3 * self.m, self.m, self.inner_iters, self.inner_tol,
                   L.ptr(self._bp_ws), L.stream())
            return
        L.call('femol_precond_apply', self.plan, L.ptr(self.dinv), fixed,
               self._p('S', self.m), 3 * self.m, self._p('S', self.m), 3 * self.m, self.m, L.stream())

    def vectors(self):
        return self.buf[self.cur]['S'][:, :self.m]
