// K6: block kernels for LOBPCG (lowest generalised eigenpairs K v = lambda M v).
//
// Replaces the dense scipy.linalg.eigh(K.toarray(), M.toarray()) of the reference (FEM.py:277, O(n^3)
// time / O(n^2) memory, all modes) and the ARPACK path (FEM.py:234).  The outer LOBPCG iteration
// (Rayleigh-Ritz on a 3m x 3m pencil, m <= 16) is host logic in femol_b200/lobpcg.py; everything that
// touches O(n) data runs here:
//   femol_spmm           Y = A X for a block of m vectors (matrix values read ONCE for all m)
//   femol_gram           tall-skinny G = A^T B   (deterministic two-stage reduction)
//   femol_block_gemm     Y = S C  (n x q times q x mc)
//   femol_block_residual R = KX - MX diag(lambda)
//   femol_precond_setup / femol_precond_apply   Jacobi / block-Jacobi, fixed dofs projected out
// Block vectors are stored row-major [n][ld] ("interleaved"): the m values of one dof are contiguous,
// so the gathers of an SpMM row are m*8-byte segments and all block kernels are coalesced.
#include <math.h>
#include "common.cuh"

namespace {

// ---- SpMM: lane (row r, vector v) --------------------------------------------------------------
template <int ND, int M>
__global__ void __launch_bounds__(256)
spmm_node_kernel(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                 const double* __restrict__ vals, const double* __restrict__ X, int ldx, double* __restrict__ Y, int ldy) {
    const int v = threadIdx.x % M;
    const int64_t n_rows = n_nodes * ND;
    const int64_t row0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / M;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / M;
    for (int64_t r = row0; r < n_rows; r += stride) {
        const int64_t i = r / ND;
        const int da = (int)(r - i * ND);
        const int pb = __ldg(&adj_ptr[i]);
        const int len = __ldg(&adj_ptr[i + 1]) - pb;
        const double* __restrict__ a = vals + (int64_t)ND * ND * pb + (int64_t)da * len * ND;
        double acc = 0.0;
        for (int k = 0; k < len; ++k) {
            const int64_t col0 = (int64_t)__ldg(&adj_idx[pb + k]) * ND;
#pragma unroll
            for (int d = 0; d < ND; ++d) acc += __ldcs(&a[k * ND + d]) * __ldg(&X[(col0 + d) * ldx + v]);
        }
        Y[r * ldy + v] = acc;
    }
}

template <int ND, int M>
int launch_spmm(femol_plan_t p, const double* vals, const double* X, int ldx, double* Y, int ldy, cudaStream_t st) {
    constexpr int TPB = 256 - 256 % M;
    const int64_t n_rows = p->n_nodes * ND;
    int64_t blocks = (n_rows * M + TPB - 1) / TPB;
    const int64_t cap = (int64_t)p->sm_count * 8;
    if (blocks > cap) blocks = cap;
    spmm_node_kernel<ND, M><<<(int)blocks, TPB, 0, st>>>(p->n_nodes, p->adj_ptr, p->adj_idx, vals, X, ldx, Y, ldy);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

// ---- Gram: G[qa][qb] = A^T B over n rows ----------------------------------------------------------
constexpr int GRAM_ROWS = 32;
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_MAXQ = 48;

__global__ void __launch_bounds__(GRAM_THREADS)
gram_partial_kernel(int64_t n, int qa, int qb, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                    double* __restrict__ partial /*[grid][qa*qb]*/) {
    __shared__ double sA[GRAM_ROWS][GRAM_MAXQ + 1];
    __shared__ double sB[GRAM_ROWS][GRAM_MAXQ + 1];
    const int nout = qa * qb;
    constexpr int MAXPT = (GRAM_MAXQ * GRAM_MAXQ + GRAM_THREADS - 1) / GRAM_THREADS;
    double acc[MAXPT];
#pragma unroll
    for (int t = 0; t < MAXPT; ++t) acc[t] = 0.0;
    const int64_t n_tiles = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * GRAM_ROWS;
        const int rows = (int)min((int64_t)GRAM_ROWS, n - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < GRAM_ROWS * qa; t += blockDim.x) {
            const int rr = t / qa, c = t - rr * qa;
            sA[rr][c] = rr < rows ? A[(r0 + rr) * lda + c] : 0.0;
        }
        for (int t = threadIdx.x; t < GRAM_ROWS * qb; t += blockDim.x) {
            const int rr = t / qb, c = t - rr * qb;
            sB[rr][c] = rr < rows ? B[(r0 + rr) * ldb + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < MAXPT; ++t) {
            const int o = threadIdx.x + t * GRAM_THREADS;
            if (o < nout) {
                const int a = o / qb, b = o - a * qb;
                double s = 0.0;
#pragma unroll 8
                for (int rr = 0; rr < GRAM_ROWS; ++rr) s += sA[rr][a] * sB[rr][b];
                acc[t] += s;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < MAXPT; ++t) {
        const int o = threadIdx.x + t * GRAM_THREADS;
        if (o < nout) partial[(int64_t)blockIdx.x * nout + o] = acc[t];
    }
}

__global__ void gram_reduce_kernel(int nblk, int nout, const double* __restrict__ partial, double* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nout) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * nout + o];
    out[o] = s;
}

// ---- Y = alpha S C + beta Y -------------------------------------------------------------------------
// A CTA stages GEMM_ROWS rows of S in shared memory before writing, so Y may alias columns of S
// (in-place orthonormalisation W <- W C) as long as it covers the same rows.
constexpr int GEMM_ROWS = 32;
__global__ void __launch_bounds__(256)
block_gemm_kernel(int64_t n, int q, int mc, const double* S, int lds, const double* __restrict__ C,
                  double* Y, int ldy, double alpha, double beta) {
    extern __shared__ double sm[];  // [q][mc] C | [GEMM_ROWS][q+1] rows of S
    double* sC = sm;
    double* sS = sm + q * mc;
    for (int t = threadIdx.x; t < q * mc; t += blockDim.x) sC[t] = C[t];
    const int64_t n_tiles = (n + GEMM_ROWS - 1) / GEMM_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * GEMM_ROWS;
        const int rows = (int)min((int64_t)GEMM_ROWS, n - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < rows * q; t += blockDim.x) {
            const int rr = t / q, a = t - rr * q;
            sS[rr * (q + 1) + a] = S[(r0 + rr) * lds + a];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < rows * mc; t += blockDim.x) {
            const int rr = t / mc, c = t - rr * mc;
            double acc = 0.0;
            for (int a = 0; a < q; ++a) acc += sS[rr * (q + 1) + a] * sC[a * mc + c];
            double* y = Y + (r0 + rr) * ldy + c;
            *y = beta == 0.0 ? alpha * acc : alpha * acc + beta * (*y);
        }
    }
}

__global__ void __launch_bounds__(256)
block_residual_kernel(int64_t n, int m, const double* __restrict__ KX, int ldk, const double* __restrict__ MX, int ldm,
                      const double* __restrict__ lambda, double* __restrict__ R, int ldr) {
    const int64_t total = n * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / m;
        const int c = (int)(t - r * m);
        R[r * ldr + c] = KX[r * ldk + c] - lambda[c] * MX[r * ldm + c];
    }
}

// ---- preconditioner --------------------------------------------------------------------------------
template <int ND>
__global__ void precond_setup_kernel(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                                     const double* __restrict__ vals, int kind, double* __restrict__ dinv /*[n_nodes][ND][ND]*/) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        const int pb = adj_ptr[i], len = adj_ptr[i + 1] - pb;
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (adj_idx[pb + mid] < (int32_t)i) lo = mid + 1; else hi = mid;
        }
        const int64_t base = (int64_t)ND * ND * pb + (int64_t)lo * ND;
        double A[ND][ND], B[ND][ND];
        for (int r = 0; r < ND; ++r)
            for (int c = 0; c < ND; ++c) {
                const double a = vals[base + (int64_t)r * len * ND + c];
                A[r][c] = (kind == FEMOL_PRECOND_BLOCK_JACOBI || r == c) ? a : 0.0;
                if (kind == FEMOL_PRECOND_NONE) A[r][c] = r == c ? 1.0 : 0.0;
                B[r][c] = r == c ? 1.0 : 0.0;
            }
        bool ok = true;
        for (int k = 0; k < ND && ok; ++k) {
            int piv = k;
            double best = fabs(A[k][k]);
            for (int r = k + 1; r < ND; ++r)
                if (fabs(A[r][k]) > best) { best = fabs(A[r][k]); piv = r; }
            if (best == 0.0) { ok = false; break; }
            if (piv != k)
                for (int c = 0; c < ND; ++c) {
                    double t = A[k][c]; A[k][c] = A[piv][c]; A[piv][c] = t;
                    t = B[k][c]; B[k][c] = B[piv][c]; B[piv][c] = t;
                }
            const double ip = 1.0 / A[k][k];
            for (int c = 0; c < ND; ++c) { A[k][c] *= ip; B[k][c] *= ip; }
            for (int r = 0; r < ND; ++r) {
                if (r == k) continue;
                const double f = A[r][k];
                if (f == 0.0) continue;
                for (int c = 0; c < ND; ++c) { A[r][c] -= f * A[k][c]; B[r][c] -= f * B[k][c]; }
            }
        }
        for (int r = 0; r < ND; ++r)
            for (int c = 0; c < ND; ++c) dinv[(i * ND + r) * ND + c] = ok ? B[r][c] : (r == c ? 1.0 : 0.0);
    }
}

// W[node dofs][v] = Dinv_node * R[node dofs][v], fixed dofs zeroed
template <int ND>
__global__ void __launch_bounds__(256)
precond_apply_kernel(int64_t n_nodes, int m, const double* __restrict__ dinv, const uint8_t* __restrict__ fixed,
                     const double* R, int ldr, double* W, int ldw) {  // W may alias R
    const int64_t total = n_nodes * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / m;
        const int v = (int)(t - i * m);
        const uint32_t fx = fixed ? fixed[i] : 0u;
        double rv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) rv[d] = ((fx >> d) & 1u) ? 0.0 : R[(i * ND + d) * ldr + v];
        const double* B = dinv + i * ND * ND;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) s += __ldg(&B[a * ND + b]) * rv[b];
            W[(i * ND + a) * ldw + v] = ((fx >> a) & 1u) ? 0.0 : s;
        }
    }
}

}  // namespace

extern "C" int femol_spmm(femol_plan_t p, const double* vals, const double* X, int ldx, double* Y, int ldy, int m, void* stream) {
    FEMOL_REQUIRE(p && vals && X && Y && ldx >= m && ldy >= m, "femol_spmm: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
#define SPMM_CASE(ND, M) if (p->ndof == ND && m == M) return launch_spmm<ND, M>(p, vals, X, ldx, Y, ldy, st);
    SPMM_CASE(2, 1) SPMM_CASE(2, 2) SPMM_CASE(2, 4) SPMM_CASE(2, 8) SPMM_CASE(2, 12) SPMM_CASE(2, 16)
    SPMM_CASE(6, 1) SPMM_CASE(6, 2) SPMM_CASE(6, 4) SPMM_CASE(6, 8) SPMM_CASE(6, 12) SPMM_CASE(6, 16)
#undef SPMM_CASE
    femol_set_error("femol_spmm: unsupported (ndof=%d, m=%d); m must be one of 1,2,4,8,12,16", p->ndof, m);
    return FEMOL_E_ARG;
}

extern "C" int64_t femol_gram_workspace_doubles(int qa, int qb) {
    return (int64_t)FEMOL_NUM_SMS * 2 * qa * qb;
}

extern "C" int femol_gram(int64_t n, int qa, int qb, const double* A, int lda, const double* B, int ldb, double* out,
                          double* ws, void* stream) {
    FEMOL_REQUIRE(A && B && out && ws && n > 0, "femol_gram: null argument");
    FEMOL_REQUIRE(qa >= 1 && qb >= 1 && qa <= GRAM_MAXQ && qb <= GRAM_MAXQ, "femol_gram: block width must be in [1, 48]");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t grid = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    if (grid > FEMOL_NUM_SMS * 2) grid = FEMOL_NUM_SMS * 2;
    gram_partial_kernel<<<(int)grid, GRAM_THREADS, 0, st>>>(n, qa, qb, A, lda, B, ldb, ws);
    FEMOL_LAUNCH_CHECK();
    const int nout = qa * qb;
    gram_reduce_kernel<<<(nout + 127) / 128, 128, 0, st>>>((int)grid, nout, ws, out);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_block_gemm(int64_t n, int q, int mc, const double* S, int lds, const double* C, double* Y, int ldy,
                                double alpha, double beta, void* stream) {
    FEMOL_REQUIRE(S && C && Y && n > 0 && q >= 1 && mc >= 1 && q <= GRAM_MAXQ && mc <= GRAM_MAXQ, "femol_block_gemm: bad argument");
    int64_t grid = (n + GEMM_ROWS - 1) / GEMM_ROWS;
    const int64_t cap = (int64_t)femol_sm_count() * 8;
    if (grid > cap) grid = cap;
    const size_t smem = ((size_t)q * mc + (size_t)GEMM_ROWS * (q + 1)) * sizeof(double);
    block_gemm_kernel<<<(int)grid, 256, smem, (cudaStream_t)stream>>>(n, q, mc, S, lds, C, Y, ldy, alpha, beta);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_block_residual(int64_t n, int m, const double* KX, int ldk, const double* MX, int ldm,
                                    const double* lambda, double* R, int ldr, void* stream) {
    FEMOL_REQUIRE(KX && MX && lambda && R && n > 0 && m >= 1, "femol_block_residual: bad argument");
    const int grid = femol_grid_for(n * m, 256, femol_sm_count(), 16);
    block_residual_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, KX, ldk, MX, ldm, lambda, R, ldr);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_precond_setup(femol_plan_t p, const double* vals, int kind, double* dinv, void* stream) {
    FEMOL_REQUIRE(p && vals && dinv, "femol_precond_setup: null argument");
    const int grid = femol_grid_for(p->n_nodes, 128, p->sm_count, 16);
    if (p->ndof == 2) precond_setup_kernel<2><<<grid, 128, 0, (cudaStream_t)stream>>>(p->n_nodes, p->adj_ptr, p->adj_idx, vals, kind, dinv);
    else if (p->ndof == 6) precond_setup_kernel<6><<<grid, 128, 0, (cudaStream_t)stream>>>(p->n_nodes, p->adj_ptr, p->adj_idx, vals, kind, dinv);
    else { femol_set_error("femol_precond_setup: unsupported ndof %d", p->ndof); return FEMOL_E_ARG; }
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_precond_apply(femol_plan_t p, const double* dinv, const uint8_t* fixed, const double* R, int ldr,
                                   double* W, int ldw, int m, void* stream) {
    FEMOL_REQUIRE(p && dinv && R && W && m >= 1, "femol_precond_apply: bad argument");
    const int grid = femol_grid_for(p->n_nodes * m, 256, p->sm_count, 16);
    if (p->ndof == 2) precond_apply_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(p->n_nodes, m, dinv, fixed, R, ldr, W, ldw);
    else if (p->ndof == 6) precond_apply_kernel<6><<<grid, 256, 0, (cudaStream_t)stream>>>(p->n_nodes, m, dinv, fixed, R, ldr, W, ldw);
    else { femol_set_error("femol_precond_apply: unsupported ndof %d", p->ndof); return FEMOL_E_ARG; }
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

// =====================================================================================================
// Block PCG: W ~= K^-1 R for m right-hand sides at once (the LOBPCG preconditioner "T ~ K^-1").
// Every column runs its own CG recurrence (alpha_c, beta_c) on the shared SpMM, preconditioned with the
// block-Jacobi inverse.  All scalars live on the device; a fixed number of iterations is enqueued with no
// host synchronisation; a column freezes (alpha = 0) once its preconditioned residual dropped by rel_tol.
// Six launches per iteration: spmm, dots, scalars, update(+dots), scalars, direction.
// =====================================================================================================
namespace {

constexpr int BP_MAXM = 16;
constexpr int BP_BLOCKS = 592;  // 4 CTAs per SM on 148 SMs

// partial[blockIdx][c] = sum over this block's rows of A[r][c] * B[r][c]
__global__ void __launch_bounds__(256)
bp_dots_kernel(int64_t n, int m, const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ partial) {
    __shared__ double sh[256];
    const int c = threadIdx.x % m;
    const int lanes = (256 / m) * m;          // threads that own a column
    double acc = 0.0;
    if ((int)threadIdx.x < lanes) {
        const int64_t r0 = (int64_t)blockIdx.x * (lanes / m) + threadIdx.x / m;
        const int64_t stride = (int64_t)gridDim.x * (lanes / m);
        for (int64_t r = r0; r < n; r += stride) acc += A[r * m + c] * B[r * m + c];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < m) {
        double s = 0.0;
        for (int t = threadIdx.x; t < lanes; t += m) s += sh[t];
        partial[(int64_t)blockIdx.x * m + threadIdx.x] = s;
    }
}

// sc layout (per column c): rz[c] | pq[c] | alpha[c] | beta[c] | rz0[c] | active[c]
enum { BP_RZ = 0, BP_PQ = 1, BP_ALPHA = 2, BP_BETA = 3, BP_RZ0 = 4, BP_ACTIVE = 5, BP_NS = 6 };

// phase 0: rz0 = rz = sum ; phase 1: pq = sum, alpha = rz/pq ; phase 2: rz_new = sum, beta = rz_new/rz, rz = rz_new
__global__ void bp_scalars_kernel(int nblk, int m, int phase, double tol2, const double* __restrict__ partial,
                                  double* __restrict__ sc) {
    const int c = threadIdx.x;
    if (c >= m) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * m + c];
    if (phase == 0) {
        sc[BP_RZ * BP_MAXM + c] = s;
        sc[BP_RZ0 * BP_MAXM + c] = s;
        sc[BP_ACTIVE * BP_MAXM + c] = (s > 0.0) ? 1.0 : 0.0;
        sc[BP_ALPHA * BP_MAXM + c] = 0.0;
        sc[BP_BETA * BP_MAXM + c] = 0.0;
    } else if (phase == 1) {
        sc[BP_PQ * BP_MAXM + c] = s;
        const bool act = sc[BP_ACTIVE * BP_MAXM + c] != 0.0 && s > 0.0;
        if (!act) sc[BP_ACTIVE * BP_MAXM + c] = 0.0;
        sc[BP_ALPHA * BP_MAXM + c] = act ? sc[BP_RZ * BP_MAXM + c] / s : 0.0;
    } else {
        const double rz = sc[BP_RZ * BP_MAXM + c];
        bool act = sc[BP_ACTIVE * BP_MAXM + c] != 0.0;
        if (act && !(s > tol2 * sc[BP_RZ0 * BP_MAXM + c])) act = false;   // converged (or NaN): freeze
        sc[BP_BETA * BP_MAXM + c] = (act && rz > 0.0) ? s / rz : 0.0;
        sc[BP_RZ * BP_MAXM + c] = s;
        sc[BP_ACTIVE * BP_MAXM + c] = act ? 1.0 : 0.0;
    }
}

// init: r = R (fixed dofs zeroed), z = T r, p = z, Z = 0
template <int ND>
__global__ void __launch_bounds__(256)
bp_init_kernel(int64_t n_nodes, int m, const double* __restrict__ dinv, const uint8_t* __restrict__ fixed,
               const double* __restrict__ R, int ldr, double* __restrict__ Z, double* __restrict__ r,
               double* __restrict__ p, double* __restrict__ z) {
    const int64_t total = n_nodes * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / m;
        const int v = (int)(t - i * m);
        const uint32_t fx = fixed ? fixed[i] : 0u;
        double rv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) rv[d] = ((fx >> d) & 1u) ? 0.0 : R[(i * ND + d) * ldr + v];
        const double* B = dinv + i * ND * ND;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) s += __ldg(&B[a * ND + b]) * rv[b];
            if ((fx >> a) & 1u) s = 0.0;
            const int64_t o = (i * ND + a) * m + v;
            Z[o] = 0.0; r[o] = rv[a]; z[o] = s; p[o] = s;
        }
    }
}

// Z += alpha p ; r -= alpha q ; z = T r
template <int ND>
__global__ void __launch_bounds__(256)
bp_update_kernel(int64_t n_nodes, int m, const double* __restrict__ dinv, const uint8_t* __restrict__ fixed,
                 const double* __restrict__ sc, const double* __restrict__ p, const double* __restrict__ q,
                 double* __restrict__ Z, double* __restrict__ r, double* __restrict__ z) {
    const int64_t total = n_nodes * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / m;
        const int v = (int)(t - i * m);
        const double alpha = sc[BP_ALPHA * BP_MAXM + v];
        const uint32_t fx = fixed ? fixed[i] : 0u;
        double rv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const int64_t o = (i * ND + d) * m + v;
            Z[o] += alpha * p[o];
            rv[d] = r[o] - alpha * q[o];
            if ((fx >> d) & 1u) rv[d] = 0.0;
            r[o] = rv[d];
        }
        const double* B = dinv + i * ND * ND;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) s += __ldg(&B[a * ND + b]) * rv[b];
            if ((fx >> a) & 1u) s = 0.0;
            z[(i * ND + a) * m + v] = s;
        }
    }
}

// p = z + beta p   (final = 1: also W[:, :m] = Z)
__global__ void __launch_bounds__(256)
bp_direction_kernel(int64_t n, int m, const double* __restrict__ sc, const double* __restrict__ z, double* __restrict__ p) {
    const int64_t total = n * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t % m);
        p[t] = z[t] + sc[BP_BETA * BP_MAXM + v] * p[t];
    }
}

__global__ void __launch_bounds__(256)
bp_store_kernel(int64_t n, int m, const double* __restrict__ Z, double* __restrict__ W, int ldw) {
    const int64_t total = n * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / m;
        W[r * ldw + (t - r * m)] = Z[t];
    }
}

template <int ND>
int block_pcg_run(femol_plan_t plan, const double* vals, const double* dinv, const uint8_t* fixed, const double* R, int ldr,
                  double* W, int ldw, int m, int iters, double rel_tol, double* ws, cudaStream_t st) {
    const int64_t n = plan->n_nodes * ND, nm = n * m;
    double* Z = ws;
    double* r = Z + nm;
    double* p = r + nm;
    double* q = p + nm;
    double* z = q + nm;
    double* partial = z + nm;                       // [BP_BLOCKS][m]
    double* sc = partial + (int64_t)BP_BLOCKS * BP_MAXM;  // [BP_NS][BP_MAXM]
    const int vgrid = femol_grid_for(plan->n_nodes * m, 256, plan->sm_count, 16);
    int dgrid = (int)((n + 15) / 16);
    if (dgrid > BP_BLOCKS) dgrid = BP_BLOCKS;
    const double tol2 = rel_tol * rel_tol;
    bp_init_kernel<ND><<<vgrid, 256, 0, st>>>(plan->n_nodes, m, dinv, fixed, R, ldr, Z, r, p, z);
    FEMOL_LAUNCH_CHECK();
    bp_dots_kernel<<<dgrid, 256, 0, st>>>(n, m, r, z, partial);
    FEMOL_LAUNCH_CHECK();
    bp_scalars_kernel<<<1, 32, 0, st>>>(dgrid, m, 0, tol2, partial, sc);
    FEMOL_LAUNCH_CHECK();
    for (int it = 0; it < iters; ++it) {
        int rc = femol_spmm(plan, vals, p, m, q, m, m, st);
        if (rc != FEMOL_OK) return rc;
        bp_dots_kernel<<<dgrid, 256, 0, st>>>(n, m, p, q, partial);
        FEMOL_LAUNCH_CHECK();
        bp_scalars_kernel<<<1, 32, 0, st>>>(dgrid, m, 1, tol2, partial, sc);
        FEMOL_LAUNCH_CHECK();
        bp_update_kernel<ND><<<vgrid, 256, 0, st>>>(plan->n_nodes, m, dinv, fixed, sc, p, q, Z, r, z);
        FEMOL_LAUNCH_CHECK();
        bp_dots_kernel<<<dgrid, 256, 0, st>>>(n, m, r, z, partial);
        FEMOL_LAUNCH_CHECK();
        bp_scalars_kernel<<<1, 32, 0, st>>>(dgrid, m, 2, tol2, partial, sc);
        FEMOL_LAUNCH_CHECK();
        bp_direction_kernel<<<vgrid, 256, 0, st>>>(n, m, sc, z, p);
        FEMOL_LAUNCH_CHECK();
    }
    bp_store_kernel<<<vgrid, 256, 0, st>>>(n, m, Z, W, ldw);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

}  // namespace

extern "C" int64_t femol_block_pcg_workspace_doubles(femol_plan_t p, int m) {
    if (!p) return 0;
    return 5 * p->n_nodes * p->ndof * (int64_t)m + (int64_t)BP_BLOCKS * BP_MAXM + BP_NS * BP_MAXM;
}

extern "C" int femol_block_pcg(femol_plan_t p, const double* vals, const double* dinv, const uint8_t* fixed,
                               const double* R, int ldr, double* W, int ldw, int m, int iters, double rel_tol,
                               double* ws, void* stream) {
    FEMOL_REQUIRE(p && vals && dinv && R && W && ws, "femol_block_pcg: null argument");
    FEMOL_REQUIRE(m >= 1 && m <= BP_MAXM && iters >= 0 && rel_tol >= 0.0, "femol_block_pcg: bad block size / iteration count");
    cudaStream_t st = (cudaStream_t)stream;
    if (p->ndof == 2) return block_pcg_run<2>(p, vals, dinv, fixed, R, ldr, W, ldw, m, iters, rel_tol, ws, st);
    if (p->ndof == 6) return block_pcg_run<6>(p, vals, dinv, fixed, R, ldr, W, ldw, m, iters, rel_tol, ws, st);
    femol_set_error("femol_block_pcg: unsupported ndof %d", p->ndof);
    return FEMOL_E_ARG;
}
