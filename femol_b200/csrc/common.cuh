// Shared device/host helpers for the femol_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/femol_b200.h"

#define FEMOL_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void femol_set_error(const char* fmt, ...);

#define FEMOL_CUDA_CHECK(expr)                                                             \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            femol_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return FEMOL_E_CUDA;                                                           \
        }                                                                                  \
    } while (0)

// every kernel launch goes through this: counts the launch and surfaces launch errors
void femol_count_launch();
void femol_count_launches(long long k);  // kernels replayed through a CUDA graph
extern "C" int64_t femol_launch_count(int reset);
#define FEMOL_LAUNCH_CHECK()                \
    do {                                    \
        femol_count_launch();               \
        FEMOL_CUDA_CHECK(cudaGetLastError()); \
    } while (0)

// cudaFuncSetAttribute is per device (context): remember it per device, not per process (a process may drive
// several GPUs, and two host threads may race on the first call; setting it twice is harmless)
#define FEMOL_FUNC_ATTR_PER_DEVICE(kern, attr, value)                                        \
    do {                                                                                     \
        static std::atomic<unsigned long long> _done{0ULL};                                  \
        int _dev = 0;                                                                        \
        FEMOL_CUDA_CHECK(cudaGetDevice(&_dev));                                              \
        const unsigned long long _bit = 1ULL << (_dev & 63);                                 \
        if (!(_done.load(std::memory_order_acquire) & _bit)) {                               \
            FEMOL_CUDA_CHECK(cudaFuncSetAttribute(kern, attr, value));                       \
            _done.fetch_or(_bit, std::memory_order_release);                                 \
        }                                                                                    \
    } while (0)

#define FEMOL_REQUIRE(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            femol_set_error("%s:%d: %s", __FILE__, __LINE__, msg); \
            return FEMOL_E_ARG;                                    \
        }                                                          \
    } while (0)

// Matrix-free description of K on a structured rectangle_Q4 mesh (stencil.cu); device pointers are borrowed,
// the element matrices are host copies (they are passed to the kernel as parameters).
struct femol_matrix_free {
    int64_t nx, ny;
    double e_base[576], e_coat[576];
    bool has_coat;
    const double* xpow;
    const uint8_t* fixed;
    const double* vals;  // the assembled value array this description is equivalent to (identity tag)
};

// The symbolic plan (plan.cu).  All pointers are device memory owned by the plan.
struct femol_plan {
    int64_t n_elem, n_nodes, nnz_node, nnz;
    int nn, ndof;
    int max_valence;    // max elements incident to a node
    int max_row_nodes;  // max nodes adjacent to a node (incl. itself)
    int32_t* conn;      // [n_elem][nn]
    int32_t* n2e_ptr;   // [n_nodes+1]
    int32_t* n2e_idx;   // [n_elem*nn]  packed (elem << 3 | local node), ascending per node
    int32_t* adj_ptr;   // [n_nodes+1]
    int32_t* adj_idx;   // [nnz_node (+pad)] sorted per node
    int32_t* src_ptr;   // [nnz_node+1] contributions of each node block (transposed scatter map)
    uint32_t* src;      // [n_elem*nn*nn] packed (elem << 6 | local row node << 3 | local col node), ascending
    uint32_t* src_ell;  // [nnz_node][ell_width] the same lists at fixed width (0xFFFFFFFF = empty), or NULL
    int ell_width;      // 4, 8 or 0 (a node pair with more than 8 contributing elements: lists only)
    int sm_count;
    femol_matrix_free* mf;  // host struct, optional (femol_plan_set_matrix_free)
    int64_t grid_nx, grid_ny;  // > 0: the mesh is an nx x ny rectangle_Q4 grid (femol_plan_set_grid)
    bool is_view;              // shares the device arrays of another plan (femol_plan_create_view): owns only `mf`
};

// Deterministic grid reduction scratch: FEMOL_RED_SLOTS partial slots + a ticket counter.
#define FEMOL_RED_MAX_BLOCKS 2048
#define FEMOL_RED_CHANNELS 4
// layout (doubles): [CHANNELS][MAX_BLOCKS] partials | [8] tickets (as unsigned long long)
#define FEMOL_RED_WS_DOUBLES (FEMOL_RED_CHANNELS * FEMOL_RED_MAX_BLOCKS + 8)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NCH channels; result valid in thread 0.  Requires blockDim.x % 32 == 0, <= 1024.
template <int NCH>
__device__ __forceinline__ void block_sum(double (&v)[NCH], double* smem /* >= 32*NCH */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = warp_sum(v[c]);
    __syncthreads();  // protect smem reuse across calls
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) smem[c * 32 + wid] = v[c];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            double t = lane < nw ? smem[c * 32 + lane] : 0.0;
            v[c] = warp_sum(t);
        }
    }
}

// Deterministic grid sum: every block deposits its partials; the last block to arrive (ticket) adds
// them in block order.  Returns true for thread 0 of that last block only, with the totals in v[];
// the caller publishes them.  `ticket` must be zero on entry and is reset before returning.
template <int NCH>
__device__ __forceinline__ bool grid_sum_last(double (&v)[NCH], double* partials /*[NCH][MAX_BLOCKS]*/,
                                              unsigned long long* ticket, double* smem) {
    block_sum<NCH>(v, smem);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) __stcg(&partials[c * FEMOL_RED_MAX_BLOCKS + blockIdx.x], v[c]);
        __threadfence();
        unsigned long long t = atomicAdd(ticket, 1ULL);
        is_last = (t == (unsigned long long)gridDim.x - 1ULL);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = 0.0;
    // fixed order: thread t sums blocks t, t+blockDim, ... then a fixed tree
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) v[c] += __ldcg(&partials[c * FEMOL_RED_MAX_BLOCKS + b]);
    }
    block_sum<NCH>(v, smem);
    if (threadIdx.x == 0) {
        *ticket = 0ULL;
        return true;
    }
    return false;
}

template <int NCH>
__device__ __forceinline__ void grid_sum_finish(double (&v)[NCH], double* partials, unsigned long long* ticket,
                                                double* out, double* smem) {
    if (grid_sum_last<NCH>(v, partials, ticket, smem)) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) out[c] = v[c];
    }
}

__device__ __forceinline__ unsigned long long* red_ticket(double* red_ws) {
    return reinterpret_cast<unsigned long long*>(red_ws + FEMOL_RED_CHANNELS * FEMOL_RED_MAX_BLOCKS);
}

static inline int femol_grid_for(int64_t n, int block, int sm_count, int waves = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * waves;
    if (need < 1) need = 1;
    if (need > cap) need = cap;
    if (need > FEMOL_RED_MAX_BLOCKS) need = FEMOL_RED_MAX_BLOCKS;
    return (int)need;
}

int femol_sm_count();

// Side stream for work that must not run on the legacy default stream (CUDA-graph capture): one library-owned
// non-blocking stream per device.  fork: *side waits for everything enqueued on `user` so far; join: `user` waits for
// everything enqueued on `side`.
bool femol_stream_is_default(cudaStream_t st);
int femol_stream_fork(cudaStream_t user, cudaStream_t* side);
int femol_stream_join(cudaStream_t side, cudaStream_t user);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// A CG iteration is ~26 short kernels back to back in one stream (a CUDA graph); the ~1.8 us between the end of one
// and the start of the next was 9 % of a SIMP step.  Kernels launched through femol_launch_pdl may be SCHEDULED
// while their predecessor is still running; they execute FEMOL_PDL_SYNC() as their first statement, which lets
// their own successor be scheduled and then waits until every predecessor grid has completed and its writes are
// visible, so no global memory is touched early.  Without the launch attribute both instructions are no-ops.
// FEMOL_PDL=0 disables the attribute.
#define FEMOL_PDL_SYNC() asm volatile("griddepcontrol.launch_dependents;\n griddepcontrol.wait;\n" ::: "memory")

bool femol_pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t femol_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = femol_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
