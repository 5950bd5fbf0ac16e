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

// ---- single-precision copies of assembled operators (mixed-precision multigrid levels) ---------------------------
// ndof = 2: NODE-BLOCKED layout, one float4 (a00 a01 a10 a11) per node pair in adjacency order, so a lane fetches a
// whole 2 x 2 block with one 16-byte request and uses the gathered x for both rows (the row-major float2 stream
// needed twice the requests per byte and ran at 3.2 TB/s).  ndof = 6: CSR order, converted element-wise.
namespace {

__global__ void __launch_bounds__(256)
to_f32_blocked2(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const double* __restrict__ vals, float4* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        const int pb = adj_ptr[i], len = adj_ptr[i + 1] - pb;
        const double* r0 = vals + 4 * (int64_t)pb;
        const double* r1 = r0 + 2 * len;
        for (int k = 0; k < len; ++k)
            out[pb + k] = make_float4((float)r0[2 * k], (float)r0[2 * k + 1], (float)r1[2 * k], (float)r1[2 * k + 1]);
    }
}

__global__ void __launch_bounds__(256) to_f32_plain(int64_t n, const double* __restrict__ a, float* __restrict__ b) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = (float)a[i];
}

// G lanes per NODE (ndof = 2); MODE as in spmv_node_kernel
template <int G, int MODE>
__global__ void __launch_bounds__(256)
spmv_block2_f32_kernel(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                       const float4* __restrict__ blk, const double* __restrict__ x, double* __restrict__ y,
                       const double* __restrict__ rhs, const double* __restrict__ dinv, double omega) {
    FEMOL_PDL_SYNC();
    const int g = threadIdx.x % G;
    const int64_t node0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / G;
    const int64_t n_iter = (n_nodes + stride - 1) / stride;
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(x);
    // Two nodes per thread in flight and every load level (row pointers -> column indices + blocks -> gathered x)
    // issued for both before the first use: the kernel is a chain of dependent L2 / DRAM round trips, not a stream
    constexpr int U = 2, Q = 3;      // Q slots per lane cover rows of up to Q * G = 12 node blocks; longer rows: tail loop
    for (int64_t it = 0; it < n_iter; it += U) {
        int64_t i[U]; int pb[U], len[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            i[u] = node0 + (it + u) * stride;
            const bool ok = it + u < n_iter && i[u] < n_nodes;
            pb[u] = ok ? __ldg(&adj_ptr[i[u]]) : 0;
            len[u] = ok ? __ldg(&adj_ptr[i[u] + 1]) - pb[u] : 0;
        }
        int32_t j[U][Q]; float4 a[U][Q];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int k = g + q * G;
                const bool ok = k < len[u];
                j[u][q] = ok ? __ldg(&adj_idx[pb[u] + k]) : -1;
                a[u][q] = ok ? __ldcs(&blk[pb[u] + k]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        double2 bx[U][Q];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < Q; ++q) bx[u][q] = j[u][q] >= 0 ? __ldg(&x2[j[u][q]]) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                a0 += (double)a[u][q].x * bx[u][q].x; a0 += (double)a[u][q].y * bx[u][q].y;
                a1 += (double)a[u][q].z * bx[u][q].x; a1 += (double)a[u][q].w * bx[u][q].y;
            }
            for (int k = g + Q * G; k < len[u]; k += G) {      // rows longer than Q * G node blocks
                const int32_t jj = __ldg(&adj_idx[pb[u] + k]);
                const float4 aa = __ldcs(&blk[pb[u] + k]);
                const double2 b = __ldg(&x2[jj]);
                a0 += (double)aa.x * b.x; a0 += (double)aa.y * b.y;
                a1 += (double)aa.z * b.x; a1 += (double)aa.w * b.y;
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            }
            if (g == 0 && len[u] > 0) {
                const int64_t ii = i[u];
                double2 out = make_double2(a0, a1);
                if (MODE >= 1) {
                    const double2 r = __ldg(reinterpret_cast<const double2*>(rhs) + ii);
                    out.x = r.x - a0; out.y = r.y - a1;
                }
                if (MODE == 2) {
                    const double2 d0 = __ldg(reinterpret_cast<const double2*>(dinv) + 2 * ii);
                    const double2 d1 = __ldg(reinterpret_cast<const double2*>(dinv) + 2 * ii + 1);
                    const double2 xi = __ldg(&x2[ii]);
                    out = make_double2(xi.x + omega * (d0.x * out.x + d0.y * out.y), xi.y + omega * (d1.x * out.x + d1.y * out.y));
                }
                reinterpret_cast<double2*>(y)[ii] = out;
            }
        }
    }
}

template <int MODE>
int launch_block2(femol_plan_t p, const float* vals32, const double* x, double* y, const femol_epilogue& ep, cudaStream_t st) {
    constexpr int G = 4;
    int64_t blocks = (p->n_nodes * G + 255) / 256;
    const int64_t cap = (int64_t)p->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    FEMOL_CUDA_CHECK(femol_launch_pdl(spmv_block2_f32_kernel<G, MODE>, dim3((unsigned)blocks), dim3(256), 0, st, p->n_nodes,
                                      (const int32_t*)p->adj_ptr, (const int32_t*)p->adj_idx, reinterpret_cast<const float4*>(vals32),
                                      x, y, ep.rhs, ep.dinv, ep.omega));
    femol_count_launch();
    return FEMOL_OK;
}

}  // namespace

int femol_to_f32_launch(femol_plan_t p, const double* vals, float* vals32, cudaStream_t st) {
    if (p->ndof == 2) {
        const int grid = femol_grid_for(p->n_nodes, 256, p->sm_count, 8);
        to_f32_blocked2<<<grid, 256, 0, st>>>(p->n_nodes, p->adj_ptr, vals, reinterpret_cast<float4*>(vals32));
    } else {
        const int grid = femol_grid_for(p->nnz, 256, p->sm_count, 8);
        to_f32_plain<<<grid, 256, 0, st>>>(p->nnz, vals, vals32);
    }
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

int femol_spmv_f32_epilogue_launch(femol_plan_t p, const float* vals32, const double* x, double* y, const femol_epilogue& ep,
                                   cudaStream_t st) {
    if (p->ndof == 2 && !ep.dot_out) {
        if (ep.mode == FEMOL_EPI_NONE) return launch_block2<0>(p, vals32, x, y, ep, st);
        if (ep.mode == FEMOL_EPI_RESIDUAL) return launch_block2<1>(p, vals32, x, y, ep, st);
        if (ep.mode == FEMOL_EPI_JACOBI) return launch_block2<2>(p, vals32, x, y, ep, st);
    } else if (p->ndof == 6) {
        if (ep.mode == FEMOL_EPI_RESIDUAL) return spmv_node_launch<6, 1, float>(p, vals32, x, y, nullptr, nullptr, nullptr, st, ep.rhs);
        if (ep.mode == FEMOL_EPI_NONE) return spmv_node_launch<6, 0, float>(p, vals32, x, y, nullptr, nullptr, nullptr, st);
    }
    femol_set_error("femol_spmv_f32: mode %d not available for ndof %d", ep.mode, p->ndof);
    return FEMOL_E_ARG;
}

// Single-precision copy of an assembled operator and the product from it (what the multigrid levels use): exposed for the
// parity tests and the bench.  vals32: nnz floats (ndof = 2: node-blocked float4 layout).  mode: FEMOL_EPI_* (0 y = K x,
// 1 y = rhs - K x, 2 y = x + omega Dinv (rhs - K x)).
extern "C" int femol_to_f32(femol_plan_t p, const double* vals, float* vals32, void* stream) {
    FEMOL_REQUIRE(p && vals && vals32, "femol_to_f32: null argument");
    return femol_to_f32_launch(p, vals, vals32, (cudaStream_t)stream);
}

extern "C" int femol_apply_f32(femol_plan_t p, const float* vals32, const double* x, double* y, int mode, const double* rhs,
                               const double* dinv, double omega, void* stream) {
    FEMOL_REQUIRE(p && vals32 && x && y && x != y, "femol_apply_f32: null or aliased argument");
    FEMOL_REQUIRE(mode >= 0 && mode <= 2 && (mode == 0 || rhs) && (mode != 2 || dinv), "femol_apply_f32: bad mode / missing operand");
    const femol_epilogue ep = {mode, rhs, dinv, omega, nullptr, nullptr};
    return femol_spmv_f32_epilogue_launch(p, vals32, x, y, ep, (cudaStream_t)stream);
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
