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
                     const double* __restrict__ R, int ldr, double* __restrict__ W, int ldw) {
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
