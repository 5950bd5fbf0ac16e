// Node-block SpMV kernel shared by the PCG and LOBPCG drivers (see spmv.cu for the traffic model).
#pragma once
#include "common.cuh"

// bytes of L2 set aside for persisting accesses (0 = feature off); configured once per process
size_t femol_l2_persist_bytes();

// One group of G lanes per CSR row; the row's values are read as a contiguous double2 stream with
// streaming (evict-first) loads so that the L2 stays available for the gathered x vector; x is read
// through the read-only path as double2 (two dofs of one neighbour node per load).
// Optional fused dot product: dot_out[0] = sum_r y[r] * dot_with[r] (deterministic grid reduction).
// MODE (femol_epilogue): 0  y = K x;  1  y = rhs - K x;  2  y = x + omega Dinv (rhs - K x)  (ND = 2 only: the two
// rows of a node sit in neighbouring lane groups of one warp, so the 2 x 2 block product is one shuffle away).
// VT: storage type of the matrix values (double; float for the mixed-precision multigrid levels - the arithmetic
// stays in double).
template <typename VT> struct femol_pair;
template <> struct femol_pair<double> { typedef double2 type; };
template <> struct femol_pair<float> { typedef float2 type; };

template <int ND, int G, int MODE = 0, typename VT = double>
__global__ void __launch_bounds__(256)
spmv_node_kernel(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                 const VT* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                 const double* __restrict__ dot_with, double* __restrict__ dot_out, double* __restrict__ red_ws,
                 const double* __restrict__ rhs = nullptr, const double* __restrict__ dinv = nullptr, double omega = 0.0) {
    static_assert(MODE != 2 || (ND == 2 && 32 % (2 * G) == 0), "fused Jacobi needs both rows of a node in one warp");
    static_assert(ND % 2 == 0, "ndof must be even for the double2 path");
    constexpr int H = ND / 2;  // double2 pairs per neighbour node
    __shared__ double s_red[32];
    const int g = threadIdx.x % G;
    const int64_t n_rows = n_nodes * ND;
    const int64_t row0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / G;
    const int64_t n_iter = (n_rows + stride - 1) / stride;
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(x);
    double dot[1] = {0.0};
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t r = row0 + it * stride;
        double acc = 0.0;
        if (r < n_rows) {
            const int64_t i = r / ND;
            const int da = (int)(r - i * ND);
            const int pb = __ldg(&adj_ptr[i]);
            const int len = __ldg(&adj_ptr[i + 1]) - pb;
            const int64_t base2 = ((int64_t)ND * ND * pb + (int64_t)da * len * ND) >> 1;
            typedef typename femol_pair<VT>::type VT2;
            const VT2* __restrict__ v2 = reinterpret_cast<const VT2*>(vals) + base2;
            const int n2 = len * H;
            // four entries per lane in flight: the column index, the value and then the gathered x of all four are
            // requested before the first product (the plain loop serialised index -> x round trips)
            for (int c0 = g; c0 < n2; c0 += 4 * G) {
                int32_t j[4]; VT2 a[4]; int hh[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = c0 + q * G;
                    const bool ok = c < n2;
                    const int k = c / H;
                    hh[q] = c - k * H;
                    j[q] = ok ? __ldg(&adj_idx[pb + k]) : -1;
                    if (ok) a[q] = __ldcs(&v2[c]); else { a[q].x = 0; a[q].y = 0; }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (j[q] >= 0) {
                        const double2 b = __ldg(&x2[(int64_t)j[q] * H + hh[q]]);
                        acc += (double)a[q].x * b.x;
                        acc += (double)a[q].y * b.y;
                    }
                }
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (MODE == 2) {
            const double other = __shfl_xor_sync(0xffffffffu, acc, G);   // K x of the node's other dof
            if (g == 0 && r < n_rows) {
                const int64_t i = r >> 1;
                const int da = (int)(r & 1);
                const double mine = __ldg(&rhs[r]) - acc, theirs = __ldg(&rhs[r ^ 1]) - other;
                const double2 d = __ldg(reinterpret_cast<const double2*>(dinv) + i * 2 + da);   // row da of the 2 x 2 block
                const double out = __ldg(&x[r]) + omega * (da == 0 ? d.x * mine + d.y * theirs : d.x * theirs + d.y * mine);
                y[r] = out;
                if (dot_out) dot[0] += out * __ldg(&rhs[r]);      // fused y . rhs (the r.z of preconditioned CG)
            }
        } else if (g == 0 && r < n_rows) {
            const double out = MODE == 1 ? __ldg(&rhs[r]) - acc : acc;
            y[r] = out;
            if (dot_with) dot[0] += out * dot_with[r];
        }
    }
    if (dot_out) {
        unsigned long long* ticket = reinterpret_cast<unsigned long long*>(red_ws + FEMOL_RED_CHANNELS * FEMOL_RED_MAX_BLOCKS);
        grid_sum_finish<1>(dot, red_ws, ticket, dot_out, s_red);
    }
}

template <int ND, int MODE = 0, typename VT = double>
int spmv_node_launch(femol_plan_t p, const VT* vals, const double* x, double* y, const double* dot_with,
                     double* dot_out, double* red_ws, cudaStream_t st, const double* rhs = nullptr,
                     const double* dinv = nullptr, double omega = 0.0) {
    constexpr int G = ND == 2 ? 4 : 8;
    const int64_t n_rows = p->n_nodes * ND;
    int64_t blocks = (n_rows * G + 255) / 256;
    int64_t cap = (int64_t)p->sm_count * 8;
    if (cap > FEMOL_RED_MAX_BLOCKS) cap = FEMOL_RED_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t xbytes = (size_t)n_rows * sizeof(double);
    const size_t window = femol_l2_persist_bytes();
    if (window && xbytes >= (4u << 20)) {
        // Large problems: the gathered vector x (re-used ~18x per entry, but evicted by the 576 MB value stream:
        // the first ncu capture showed ~150 MB of x traffic for a 32 MB vector) is pinned in the persisting
        // part of L2 for this launch; everything else in the kernel streams.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<double*>(x);
        attr[0].val.accessPolicyWindow.num_bytes = xbytes < window ? xbytes : window;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const int64_t nn = p->n_nodes;
        FEMOL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, spmv_node_kernel<ND, G, MODE, VT>, nn, (const int32_t*)p->adj_ptr,
                                            (const int32_t*)p->adj_idx, vals, x, y, dot_with, dot_out, red_ws, rhs, dinv,
                                            omega));
        femol_count_launch();
        return FEMOL_OK;
    }
    spmv_node_kernel<ND, G, MODE, VT><<<(int)blocks, 256, 0, st>>>(p->n_nodes, p->adj_ptr, p->adj_idx, vals, x, y, dot_with,
                                                                dot_out, red_ws, rhs, dinv, omega);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

int femol_spmv_launch(femol_plan_t p, const double* vals, const double* x, double* y, const double* dot_with,
                      double* dot_out, double* red_ws, cudaStream_t st);
// solver-side K x: matrix-free stencil when available for `vals` (stencil.cu), else the CSR kernel above
int femol_apply_launch(femol_plan_t p, const double* vals, const double* x, double* y, const double* dot_with,
                       double* dot_out, double* red_ws, cudaStream_t st);
// the same product with a fused epilogue (multigrid smoother / residual); y must not alias x
struct femol_epilogue {
    int mode;            // FEMOL_EPI_*
    const double* rhs;   // r in modes 1, 2
    const double* dinv;  // [n_nodes][ND][ND] block-diagonal inverse, mode 2
    double omega;        // damping, mode 2
    double* dot_out;     // mode 2, optional: dot_out[0] = y . rhs (deterministic grid sum, needs red_ws)
    double* red_ws;
};
#define FEMOL_EPI_NONE 0      // y = K x
#define FEMOL_EPI_RESIDUAL 1  // y = rhs - K x
#define FEMOL_EPI_JACOBI 2    // y = x + omega Dinv (rhs - K x)
// whether mode 2 is available for this (plan, vals): every ND on the stencil path, ND = 2 on the CSR path
bool femol_apply_has_jacobi(femol_plan_t p, const double* vals);
int femol_apply_epilogue_launch(femol_plan_t p, const double* vals, const double* x, double* y, const femol_epilogue& ep,
                                cudaStream_t st);
// single-precision copy of an assembled operator (ndof = 2: node-blocked float4 layout, ndof = 6: CSR order) and the
// product from it (mixed-precision multigrid levels): modes 0-2 for ndof = 2, modes 0-1 for ndof = 6
int femol_to_f32_launch(femol_plan_t p, const double* vals, float* vals32, cudaStream_t st);
int femol_spmv_f32_epilogue_launch(femol_plan_t p, const float* vals32, const double* x, double* y, const femol_epilogue& ep,
                                   cudaStream_t st);
