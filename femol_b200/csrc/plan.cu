// Symbolic plan: sparsity pattern of K / M and the element <-> CSR maps, built on the device.
//
// Replaces Mesh.global_matrix_indexes (reference mesh.py:284-308, a Python loop emitting N_e*size^2
// COO rows/cols) and the pattern half of scipy's csr_matrix(...).sum_duplicates() (FEM.py:498-501).
// The pattern is value independent: mesh connectivity x ndof^2, sorted int32 column indices, explicit
// zeros kept (SURVEY §7 "hard parts").  Instead of sorting N_e*size^2 COO pairs we build the NODE
// adjacency (node i ~ node j when they share an element) through the node->element incidence; the dof
// level CSR follows arithmetically because every node pair owns a dense ndof x ndof block.
#include <cub/device/device_scan.cuh>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

static thread_local char g_err[512] = "";
void femol_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* femol_last_error(void) { return g_err; }
#include <atomic>
static std::atomic<long long> g_launches{0};
void femol_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void femol_count_launches(long long k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
extern "C" int64_t femol_launch_count(int reset) {
    const long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return (int64_t)v;
}
extern "C" const char* femol_version(void) { return "femol_b200 0.1 (sm_100a)"; }
extern "C" int femol_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
bool femol_pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* env = getenv("FEMOL_PDL");
        cached = (env && atoi(env) == 0) ? 0 : 1;
    }
    return cached == 1;
}

int femol_sm_count() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
        cached = n;
    else
        cached = FEMOL_NUM_SMS;
    return cached;
}

bool femol_stream_is_default(cudaStream_t st) {
    return st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread;
}

namespace {
struct SideStream { cudaStream_t st; cudaEvent_t fork_ev, join_ev; bool ready; };
SideStream g_side[64];
std::atomic_flag g_side_lock = ATOMIC_FLAG_INIT;
}  // namespace

static int side_stream_get(SideStream** out) {
    int dev = 0;
    FEMOL_CUDA_CHECK(cudaGetDevice(&dev));
    SideStream& s = g_side[dev & 63];
    while (g_side_lock.test_and_set(std::memory_order_acquire)) {}
    cudaError_t e = cudaSuccess;
    if (!s.ready) {
        e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.fork_ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.join_ev, cudaEventDisableTiming);
        s.ready = e == cudaSuccess;
    }
    g_side_lock.clear(std::memory_order_release);
    FEMOL_CUDA_CHECK(e);
    *out = &s;
    return FEMOL_OK;
}

int femol_stream_fork(cudaStream_t user, cudaStream_t* side) {
    SideStream* s = nullptr;
    int rc = side_stream_get(&s);
    if (rc != FEMOL_OK) return rc;
    FEMOL_CUDA_CHECK(cudaEventRecord(s->fork_ev, user));
    FEMOL_CUDA_CHECK(cudaStreamWaitEvent(s->st, s->fork_ev, 0));
    *side = s->st;
    return FEMOL_OK;
}

int femol_stream_join(cudaStream_t side, cudaStream_t user) {
    SideStream* s = nullptr;
    int rc = side_stream_get(&s);
    if (rc != FEMOL_OK) return rc;
    FEMOL_CUDA_CHECK(cudaEventRecord(s->join_ev, side));
    FEMOL_CUDA_CHECK(cudaStreamWaitEvent(user, s->join_ev, 0));
    return FEMOL_OK;
}

namespace {

constexpr int MAX_CAND = 128;  // max distinct nodes adjacent to one node

__global__ void count_valence(int64_t total, const int32_t* __restrict__ conn, int32_t* __restrict__ deg) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&deg[conn[i]], 1);
}

__global__ void fill_n2e(int64_t n_elem, int nn, const int32_t* __restrict__ conn, const int32_t* __restrict__ ptr,
                         int32_t* __restrict__ cursor, int32_t* __restrict__ idx) {
    const int64_t total = n_elem * nn;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t node = conn[i];
        const int32_t slot = atomicAdd(&cursor[node], 1);
        const int64_t e = i / nn;
        const int a = (int)(i % nn);
        idx[ptr[node] + slot] = (int32_t)((e << 3) | a);
    }
}

// ascending order per node -> deterministic summation order in the assembly
__global__ void sort_n2e(int64_t n_nodes, const int32_t* __restrict__ ptr, int32_t* __restrict__ idx, int* max_val) {
    int local_max = 0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
        const int b = ptr[n], e = ptr[n + 1];
        for (int i = b + 1; i < e; ++i) {
            const int32_t v = idx[i];
            int j = i - 1;
            while (j >= b && idx[j] > v) { idx[j + 1] = idx[j]; --j; }
            idx[j + 1] = v;
        }
        local_max = max(local_max, e - b);
    }
    if (local_max) atomicMax(max_val, local_max);
}

// Node adjacency: pass 0 counts distinct neighbours, pass 1 writes them sorted.
template <int PASS>
__global__ void build_adjacency(int64_t n_nodes, int nn, const int32_t* __restrict__ conn,
                                const int32_t* __restrict__ n2e_ptr, const int32_t* __restrict__ n2e_idx,
                                int32_t* __restrict__ adj_cnt_or_ptr, int32_t* __restrict__ adj_idx,
                                int* __restrict__ flags /*[0]=max row nodes,[1]=overflow*/) {
    int local_max = 0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
        int32_t cand[MAX_CAND];
        int len = 0;
        bool overflow = false;
        for (int s = n2e_ptr[n]; s < n2e_ptr[n + 1]; ++s) {
            const int64_t e = n2e_idx[s] >> 3;
            for (int b = 0; b < nn; ++b) {
                const int32_t v = conn[e * nn + b];
                int lo = 0, hi = len;  // lower bound in the sorted candidate list
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cand[mid] < v) lo = mid + 1; else hi = mid;
                }
                if (lo < len && cand[lo] == v) continue;
                if (len == MAX_CAND) { overflow = true; continue; }
                for (int j = len; j > lo; --j) cand[j] = cand[j - 1];
                cand[lo] = v;
                ++len;
            }
        }
        if (overflow) atomicExch(&flags[1], 1);
        if (PASS == 0) {
            adj_cnt_or_ptr[n] = len;
            local_max = max(local_max, len);
        } else {
            const int base = adj_cnt_or_ptr[n];
            for (int j = 0; j < len; ++j) adj_idx[base + j] = cand[j];
        }
    }
    if (PASS == 0 && local_max) atomicMax(&flags[0], local_max);
}

__global__ void export_indptr(int64_t n_nodes, int nd, const int32_t* __restrict__ adj_ptr, int32_t* __restrict__ indptr) {
    const int64_t rows = n_nodes * nd;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x) {
        if (r == rows) { indptr[r] = (int32_t)((int64_t)nd * nd * adj_ptr[n_nodes]); continue; }
        const int64_t i = r / nd;
        const int da = (int)(r % nd);
        const int len = adj_ptr[i + 1] - adj_ptr[i];
        indptr[r] = (int32_t)((int64_t)nd * nd * adj_ptr[i] + (int64_t)da * nd * len);
    }
}

__device__ __forceinline__ int64_t find_row(const int32_t* __restrict__ ptr, int64_t n, int64_t pos) {
    int64_t lo = 0, hi = n;  // last i with ptr[i] <= pos
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ptr[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void export_indices(int64_t n_nodes, int nd, const int32_t* __restrict__ adj_ptr,
                               const int32_t* __restrict__ adj_idx, int32_t* __restrict__ indices) {
    const int64_t nblk = adj_ptr[n_nodes];
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < nblk; pos += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = find_row(adj_ptr, n_nodes, pos);
        const int len = adj_ptr[i + 1] - adj_ptr[i];
        const int k = (int)(pos - adj_ptr[i]);
        const int64_t base = (int64_t)nd * nd * adj_ptr[i];
        const int32_t col0 = adj_idx[pos] * nd;
        for (int da = 0; da < nd; ++da)
            for (int db = 0; db < nd; ++db)
                indices[base + (int64_t)da * nd * len + (int64_t)k * nd + db] = col0 + db;
    }
}

__global__ void export_scatter(int64_t n_elem, int nn, const int32_t* __restrict__ conn,
                               const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                               int32_t* __restrict__ map) {
    const int64_t total = n_elem * nn * nn;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nn * nn);
        const int a = (int)((t / nn) % nn), b = (int)(t % nn);
        const int32_t na = conn[e * nn + a], nb = conn[e * nn + b];
        int lo = adj_ptr[na], hi = adj_ptr[na + 1];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (adj_idx[mid] < nb) lo = mid + 1; else hi = mid;
        }
        map[t] = lo;
    }
}

// Transposed scatter map: for node block `pos` the list of (element, a, b) that contribute to it.
template <int PASS>
__global__ void build_sources(int64_t n_elem, int nn, const int32_t* __restrict__ conn,
                              const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                              int32_t* __restrict__ cnt_or_ptr, int32_t* __restrict__ cursor,
                              uint32_t* __restrict__ src) {
    const int64_t total = n_elem * nn * nn;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nn * nn);
        const int a = (int)((t / nn) % nn), b = (int)(t % nn);
        const int32_t na = conn[e * nn + a], nb = conn[e * nn + b];
        int lo = adj_ptr[na], hi = adj_ptr[na + 1];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (adj_idx[mid] < nb) lo = mid + 1; else hi = mid;
        }
        if (PASS == 0) {
            atomicAdd(&cnt_or_ptr[lo], 1);
        } else {
            const int slot = atomicAdd(&cursor[lo], 1);
            src[cnt_or_ptr[lo] + slot] = ((uint32_t)e << 6) | ((uint32_t)a << 3) | (uint32_t)b;
        }
    }
}

__global__ void sort_sources(int64_t nblk, const int32_t* __restrict__ ptr, uint32_t* __restrict__ src) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nblk; n += (int64_t)gridDim.x * blockDim.x) {
        const int b = ptr[n], e = ptr[n + 1];
        for (int i = b + 1; i < e; ++i) {
            const uint32_t v = src[i];
            int j = i - 1;
            while (j >= b && src[j] > v) { src[j + 1] = src[j]; --j; }
            src[j + 1] = v;
        }
    }
}

__global__ void pack_sources_ell(int64_t nblk, int width, const int32_t* __restrict__ ptr, const uint32_t* __restrict__ src,
                                 uint32_t* __restrict__ ell) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nblk; n += (int64_t)gridDim.x * blockDim.x) {
        const int b = ptr[n], cnt = ptr[n + 1] - b;
        for (int s = 0; s < width; ++s) ell[n * width + s] = s < cnt ? src[b + s] : 0xFFFFFFFFu;
    }
}

int grid_of(int64_t n, int block = 256) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = (int64_t)femol_sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int exclusive_scan(int32_t* data, int64_t n, cudaStream_t st) {
    size_t bytes = 0;
    FEMOL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, data, data, (int)n, st));
    void* tmp = nullptr;
    FEMOL_CUDA_CHECK(cudaMallocAsync(&tmp, bytes ? bytes : 16, st));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, data, data, (int)n, st);
    cudaFreeAsync(tmp, st);
    FEMOL_CUDA_CHECK(e);
    return FEMOL_OK;
}

}  // namespace

extern "C" int femol_plan_destroy(femol_plan_t p) {
    if (!p) return FEMOL_OK;
    if (p->is_view) { delete p->mf; delete p; return FEMOL_OK; }
    cudaFree(p->conn); cudaFree(p->n2e_ptr); cudaFree(p->n2e_idx); cudaFree(p->adj_ptr); cudaFree(p->adj_idx);
    cudaFree(p->src_ptr); cudaFree(p->src); cudaFree(p->src_ell);
    delete p->mf;
    delete p;
    return FEMOL_OK;
}

extern "C" int femol_plan_create(int64_t n_elem, int nn, int ndof, int64_t n_nodes, const int32_t* conn,
                                 void* stream, femol_plan_t* plan_out) {
    FEMOL_REQUIRE(plan_out && conn, "femol_plan_create: null argument");
    FEMOL_REQUIRE(n_elem > 0 && n_nodes > 0, "femol_plan_create: empty mesh");
    FEMOL_REQUIRE(nn >= 2 && nn <= 8, "femol_plan_create: nodes per element must be in [2, 8]");
    FEMOL_REQUIRE(ndof >= 1 && ndof <= 6, "femol_plan_create: ndof must be in [1, 6]");
    FEMOL_REQUIRE(n_elem < (1LL << 26), "femol_plan_create: more than 2^26 elements");
    FEMOL_REQUIRE(n_nodes * ndof < 0x7fffffffLL, "femol_plan_create: dof count overflows int32 CSR indices");
    cudaStream_t st = (cudaStream_t)stream;
    femol_plan* p = new femol_plan();
    memset(p, 0, sizeof(*p));
    p->n_elem = n_elem; p->n_nodes = n_nodes; p->nn = nn; p->ndof = ndof;
    p->sm_count = femol_sm_count();
    int* d_flags = nullptr;
    int32_t* d_cursor = nullptr;
    int rc = FEMOL_OK;
    auto fail = [&](int code) {
        if (d_flags) cudaFree(d_flags);
        if (d_cursor) cudaFree(d_cursor);
        femol_plan_destroy(p);
        return code;
    };
#define PLAN_CHECK(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            femol_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return fail(FEMOL_E_CUDA);                                                          \
        }                                                                                       \
    } while (0)
    const int64_t inc = n_elem * nn;
    PLAN_CHECK(cudaMalloc(&p->conn, inc * sizeof(int32_t)));
    PLAN_CHECK(cudaMemcpyAsync(p->conn, conn, inc * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    PLAN_CHECK(cudaMalloc(&p->n2e_ptr, (n_nodes + 1) * sizeof(int32_t)));
    PLAN_CHECK(cudaMalloc(&p->n2e_idx, inc * sizeof(int32_t)));
    PLAN_CHECK(cudaMalloc(&p->adj_ptr, (n_nodes + 1) * sizeof(int32_t)));
    PLAN_CHECK(cudaMalloc(&d_cursor, (n_nodes + 1) * sizeof(int32_t)));
    PLAN_CHECK(cudaMalloc(&d_flags, 4 * sizeof(int)));
    PLAN_CHECK(cudaMemsetAsync(p->n2e_ptr, 0, (n_nodes + 1) * sizeof(int32_t), st));
    PLAN_CHECK(cudaMemsetAsync(d_cursor, 0, (n_nodes + 1) * sizeof(int32_t), st));
    PLAN_CHECK(cudaMemsetAsync(d_flags, 0, 4 * sizeof(int), st));

    count_valence<<<grid_of(inc), 256, 0, st>>>(inc, p->conn, p->n2e_ptr);
    PLAN_CHECK(cudaGetLastError());
    if ((rc = exclusive_scan(p->n2e_ptr, n_nodes + 1, st)) != FEMOL_OK) return fail(rc);
    fill_n2e<<<grid_of(inc), 256, 0, st>>>(n_elem, nn, p->conn, p->n2e_ptr, d_cursor, p->n2e_idx);
    PLAN_CHECK(cudaGetLastError());
    sort_n2e<<<grid_of(n_nodes), 256, 0, st>>>(n_nodes, p->n2e_ptr, p->n2e_idx, &d_flags[2]);
    PLAN_CHECK(cudaGetLastError());

    PLAN_CHECK(cudaMemsetAsync(p->adj_ptr, 0, (n_nodes + 1) * sizeof(int32_t), st));
    build_adjacency<0><<<grid_of(n_nodes, 128), 128, 0, st>>>(n_nodes, nn, p->conn, p->n2e_ptr, p->n2e_idx,
                                                                p->adj_ptr, nullptr, d_flags);
    PLAN_CHECK(cudaGetLastError());
    if ((rc = exclusive_scan(p->adj_ptr, n_nodes + 1, st)) != FEMOL_OK) return fail(rc);
    int32_t h_nnz_node = 0;
    int h_flags[4];
    PLAN_CHECK(cudaMemcpyAsync(&h_nnz_node, p->adj_ptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PLAN_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    PLAN_CHECK(cudaStreamSynchronize(st));
    if (h_flags[1]) {
        femol_set_error("femol_plan_create: a node has more than %d adjacent nodes", MAX_CAND);
        return fail(FEMOL_E_LIMIT);
    }
    p->nnz_node = h_nnz_node;
    p->nnz = (int64_t)h_nnz_node * ndof * ndof;
    p->max_row_nodes = h_flags[0];
    p->max_valence = h_flags[2];
    if (p->nnz >= 0x7fffffffLL) {
        femol_set_error("femol_plan_create: nnz = %lld overflows int32 CSR pointers", (long long)p->nnz);
        return fail(FEMOL_E_LIMIT);
    }
    PLAN_CHECK(cudaMalloc(&p->adj_idx, (p->nnz_node + 8) * sizeof(int32_t)));
    PLAN_CHECK(cudaMemsetAsync(p->adj_idx + p->nnz_node, 0, 8 * sizeof(int32_t), st));
    build_adjacency<1><<<grid_of(n_nodes, 128), 128, 0, st>>>(n_nodes, nn, p->conn, p->n2e_ptr, p->n2e_idx,
                                                                p->adj_ptr, p->adj_idx, d_flags);
    PLAN_CHECK(cudaGetLastError());

    // transposed scatter map (sources of every node block)
    {
        const int64_t nsrc = n_elem * nn * nn;
        int32_t* d_cur2 = nullptr;
        PLAN_CHECK(cudaMalloc(&p->src_ptr, (p->nnz_node + 1) * sizeof(int32_t)));
        PLAN_CHECK(cudaMalloc(&p->src, nsrc * sizeof(uint32_t)));
        PLAN_CHECK(cudaMalloc(&d_cur2, (p->nnz_node + 1) * sizeof(int32_t)));
        cudaError_t e1 = cudaMemsetAsync(p->src_ptr, 0, (p->nnz_node + 1) * sizeof(int32_t), st);
        cudaError_t e2 = cudaMemsetAsync(d_cur2, 0, (p->nnz_node + 1) * sizeof(int32_t), st);
        build_sources<0><<<grid_of(nsrc), 256, 0, st>>>(n_elem, nn, p->conn, p->adj_ptr, p->adj_idx, p->src_ptr, nullptr, nullptr);
        int rc2 = exclusive_scan(p->src_ptr, p->nnz_node + 1, st);
        build_sources<1><<<grid_of(nsrc), 256, 0, st>>>(n_elem, nn, p->conn, p->adj_ptr, p->adj_idx, p->src_ptr, d_cur2, p->src);
        sort_sources<<<grid_of(p->nnz_node), 256, 0, st>>>(p->nnz_node, p->src_ptr, p->src);
        // fixed-width (ELL) copy of the source lists when no node pair has more than 8 contributors: one
        // 16-byte load per 4 sources instead of a pointer pair + a dependent list walk
        p->ell_width = p->max_valence <= 4 ? 4 : (p->max_valence <= 8 ? 8 : 0);
        cudaError_t e4 = cudaSuccess;
        if (p->ell_width) {
            e4 = cudaMalloc(&p->src_ell, (size_t)p->nnz_node * p->ell_width * sizeof(uint32_t));
            if (e4 == cudaSuccess)
                pack_sources_ell<<<grid_of(p->nnz_node), 256, 0, st>>>(p->nnz_node, p->ell_width, p->src_ptr, p->src, p->src_ell);
        }
        cudaError_t e3 = cudaStreamSynchronize(st);
        cudaFree(d_cur2);
        PLAN_CHECK(e4);
        if (rc2 != FEMOL_OK) return fail(rc2);
        PLAN_CHECK(e1); PLAN_CHECK(e2); PLAN_CHECK(e3);
        PLAN_CHECK(cudaGetLastError());
    }
#undef PLAN_CHECK
    cudaFree(d_flags);
    cudaFree(d_cursor);
    *plan_out = p;
    return FEMOL_OK;
}

extern "C" int femol_plan_info(femol_plan_t p, int64_t* h) {
    FEMOL_REQUIRE(p && h, "femol_plan_info: null argument");
    h[0] = p->nnz_node; h[1] = p->nnz; h[2] = p->max_valence; h[3] = p->max_row_nodes;
    h[4] = p->n_elem; h[5] = p->n_nodes; h[6] = p->nn; h[7] = p->ndof;
    return FEMOL_OK;
}

extern "C" int femol_plan_export_csr(femol_plan_t p, int32_t* indptr, int32_t* indices, void* stream) {
    FEMOL_REQUIRE(p && indptr && indices, "femol_plan_export_csr: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    export_indptr<<<grid_of(p->n_nodes * p->ndof + 1), 256, 0, st>>>(p->n_nodes, p->ndof, p->adj_ptr, indptr);
    FEMOL_LAUNCH_CHECK();
    export_indices<<<grid_of(p->nnz_node), 256, 0, st>>>(p->n_nodes, p->ndof, p->adj_ptr, p->adj_idx, indices);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_plan_export_adjacency(femol_plan_t p, int32_t* adj_ptr, int32_t* adj_idx, void* stream) {
    FEMOL_REQUIRE(p && adj_ptr && adj_idx, "femol_plan_export_adjacency: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    FEMOL_CUDA_CHECK(cudaMemcpyAsync(adj_ptr, p->adj_ptr, (p->n_nodes + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    FEMOL_CUDA_CHECK(cudaMemcpyAsync(adj_idx, p->adj_idx, p->nnz_node * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return FEMOL_OK;
}

extern "C" int femol_plan_export_scatter_map(femol_plan_t p, int32_t* map, void* stream) {
    FEMOL_REQUIRE(p && map, "femol_plan_export_scatter_map: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    export_scatter<<<grid_of(p->n_elem * p->nn * p->nn), 256, 0, st>>>(p->n_elem, p->nn, p->conn, p->adj_ptr, p->adj_idx, map);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

namespace {
__global__ void check_grid_conn(int64_t nx, int64_t ny, const int32_t* __restrict__ conn, int* __restrict__ bad) {
    const int64_t ne = nx * ny;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / ny, j = e - i * ny;
        const int64_t a = i + (nx + 1) * j;
        if (conn[4 * e] != a || conn[4 * e + 1] != a + 1 || conn[4 * e + 2] != a + nx + 2 || conn[4 * e + 3] != a + nx + 1) *bad = 1;
    }
}
}  // namespace

// Declares the mesh to be an nx x ny rectangle_Q4 grid (nodes x-fastest, element e = i*ny + j with nodes
// [a, a+1, a+nx+2, a+nx+1], a = i + (nx+1) j: mesh.py:852-860 / meshzoo.rectangle_quad) after VERIFYING the
// connectivity on the device.  femol_assemble then derives the contributing elements of every node pair from grid
// coordinates instead of reading the source lists.  Synchronous (one flag read); nx = 0 clears.
extern "C" int femol_plan_set_grid(femol_plan_t p, int64_t nx, int64_t ny) {
    FEMOL_REQUIRE(p, "femol_plan_set_grid: null plan");
    p->grid_nx = p->grid_ny = 0;
    if (nx <= 0 || ny <= 0) return FEMOL_OK;
    FEMOL_REQUIRE(p->nn == 4 && p->n_elem == nx * ny && p->n_nodes == (nx + 1) * (ny + 1) && p->n_nodes < (int64_t)1 << 25 && nx >= 1,
                  "femol_plan_set_grid: plan is not an nx x ny 4-node grid (or has more than 2^25 nodes)");
    int* d_bad = nullptr;
    FEMOL_CUDA_CHECK(cudaMalloc(&d_bad, sizeof(int)));
    FEMOL_CUDA_CHECK(cudaMemset(d_bad, 0, sizeof(int)));
    check_grid_conn<<<femol_grid_for(p->n_elem, 256, p->sm_count), 256>>>(nx, ny, p->conn, d_bad);
    FEMOL_LAUNCH_CHECK();
    int bad = 0;
    const cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    FEMOL_CUDA_CHECK(e);
    if (bad) {
        femol_set_error("femol_plan_set_grid: connectivity is not in rectangle_Q4 order");
        return FEMOL_E_ARG;
    }
    p->grid_nx = nx; p->grid_ny = ny;
    return FEMOL_OK;
}

// A second handle on the same symbolic data with its own matrix-free descriptor (femol_plan_set_matrix_free): the
// sign-modulated operator S K S of the hourglass hierarchy lives on the same mesh as K.  The view must be destroyed
// before the plan it was made from.
extern "C" int femol_plan_create_view(femol_plan_t p, femol_plan_t* out) {
    FEMOL_REQUIRE(p && out, "femol_plan_create_view: null argument");
    femol_plan* v = new femol_plan(*p);
    v->mf = nullptr;
    v->is_view = true;
    *out = v;
    return FEMOL_OK;
}
