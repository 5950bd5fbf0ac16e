// K4: sparse matrix-vector products.
//
// femol_spmv: y = A x for a matrix whose values are stored in the plan's CSR order.  Column indices are
// NOT read from the CSR `indices` array: every node pair owns a dense ndof x ndof block, so the kernel
// reads the node adjacency (4 B per block instead of 4 B per entry) and derives the columns.  HBM
// traffic per product: 8*nnz (values) + 4*nnz/ndof^2 (adjacency) + 16*n (x once, y once) — 25 % less
// than plain CSR at ndof = 2 and 33 % less at ndof = 6.
// femol_spmv_csr: plain CSR product for arbitrary matrices (parity checks against scipy matrices).
#include <stdlib.h>
#include "common.cuh"
#include "spmv.cuh"

namespace {

template <int G>
__global__ void __launch_bounds__(256)
spmv_csr_kernel(int64_t n_rows, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
    const int g = threadIdx.x % G;
    const int64_t row0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / G;
    for (int64_t r = row0; r < n_rows + (G - 1); r += stride) {  // keep whole groups converged
        double acc = 0.0;
        if (r < n_rows) {
            const int b = indptr[r], e = indptr[r + 1];
            for (int c = b + g; c < e; c += G) acc += vals[c] * __ldg(&x[indices[c]]);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (g == 0 && r < n_rows) y[r] = acc;
    }
}

}  // namespace

size_t femol_l2_persist_bytes() {
    static long long cached = -1;
    if (cached >= 0) return (size_t)cached;
    cached = 0;
    const char* env = getenv("FEMOL_L2_PERSIST_MB");
    const long mb = env ? atol(env) : 32;   // default: 32 MB of the 126 MB L2 (measured: SpMV 152 -> 147 us; 96 MB: 235 us)
    if (mb <= 0) return 0;
    int dev = 0, max_persist = 0, max_window = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    long long want = (long long)mb << 20;
    if (want > max_persist) want = max_persist;
    if (want <= 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want) != cudaSuccess) { cudaGetLastError(); return 0; }
    cached = want < max_window ? want : max_window;
    return (size_t)cached;
}

int femol_spmv_launch(femol_plan_t p, const double* vals, const double* x, double* y, const double* dot_with,
                      double* dot_out, double* red_ws, cudaStream_t st) {
    int rc = FEMOL_E_ARG;
    if (p->ndof == 2) rc = spmv_node_launch<2>(p, vals, x, y, dot_with, dot_out, red_ws, st);
    else if (p->ndof == 6) rc = spmv_node_launch<6>(p, vals, x, y, dot_with, dot_out, red_ws, st);
    else femol_set_error("femol_spmv: unsupported ndof %d", p->ndof);
    return rc;
}

int femol_spmv_f32_epilogue_launch(femol_plan_t p, const float* vals32, const double* x, double* y, const femol_epilogue& ep,
                                   cudaStream_t st) {
    if (ep.mode == FEMOL_EPI_RESIDUAL) {
        if (p->ndof == 2) return spmv_node_launch<2, 1, float>(p, vals32, x, y, nullptr, nullptr, nullptr, st, ep.rhs);
        if (p->ndof == 6) return spmv_node_launch<6, 1, float>(p, vals32, x, y, nullptr, nullptr, nullptr, st, ep.rhs);
    } else if (ep.mode == FEMOL_EPI_JACOBI && p->ndof == 2) {
        return spmv_node_launch<2, 2, float>(p, vals32, x, y, nullptr, ep.dot_out, ep.red_ws, st, ep.rhs, ep.dinv, ep.omega);
    } else if (ep.mode == FEMOL_EPI_NONE) {
        if (p->ndof == 2) return spmv_node_launch<2, 0, float>(p, vals32, x, y, nullptr, nullptr, nullptr, st);
        if (p->ndof == 6) return spmv_node_launch<6, 0, float>(p, vals32, x, y, nullptr, nullptr, nullptr, st);
    }
    femol_set_error("femol_spmv_f32: mode %d not available for ndof %d", ep.mode, p->ndof);
    return FEMOL_E_ARG;
}

extern "C" int femol_spmv(femol_plan_t p, const double* vals, const double* x, double* y, void* stream) {
    FEMOL_REQUIRE(p && vals && x && y, "femol_spmv: null argument");
    return femol_spmv_launch(p, vals, x, y, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int femol_spmv_csr(int64_t n_rows, const int32_t* indptr, const int32_t* indices, const double* vals,
                              const double* x, double* y, void* stream) {
    FEMOL_REQUIRE(indptr && indices && vals && x && y && n_rows >= 0, "femol_spmv_csr: bad argument");
    if (n_rows == 0) return FEMOL_OK;
    constexpr int G = 8;
    int64_t blocks = (n_rows * G + 255) / 256;
    const int64_t cap = (int64_t)femol_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    spmv_csr_kernel<G><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n_rows, indptr, indices, vals, x, y);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}
