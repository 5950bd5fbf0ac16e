// Multi-GPU row-block partition: distributed block-Jacobi PCG.
//
// One process per GPU.  Rank r owns a contiguous range of nodes (structured meshes: strips of node rows) and
// holds a LOCAL mesh = its owned nodes first, then the ghost nodes of the one halo layer of elements.  The
// local plan / assembly are the single-GPU ones (halo elements are assembled redundantly, no communication);
// only the rows of owned nodes are used.  Per CG iteration:
//   halo exchange of p (pack kernel -> ncclSend/ncclRecv straight into the ghost segment of p)
//   q = K p on owned rows, local p.q          -> ncclAllReduce (1 double, in stream)
//   u += a p, r -= a q, z = M^-1 r, local r.z, r.r -> ncclAllReduce (2 doubles)
//   p = z + b p
// All scalars stay on the device; the host reads one flag per batch of iterations.  NCCL is loaded with
// dlopen at femol_dist_init so that libfemol_b200.so has no link-time NCCL dependency (it is the library
// torch already loaded in the process when there is one).
#include <dlfcn.h>
#include <math.h>
#include <string.h>
#include "common.cuh"
#include "spmv.cuh"
#include "dist.cuh"

extern "C" int femol_precond_setup(femol_plan_t p, const double* vals, int kind, double* dinv, void* stream);

namespace {

// minimal NCCL surface (nccl.h is not included: the .so must build where NCCL is absent)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclDouble_ = 8, ncclSum_ = 0 };  // ncclFloat64 = 8, ncclSum = 0 (stable since NCCL 2.0)

struct NcclApi {
    void* handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};
NcclApi g_nccl = {};

int load_nccl() {
    if (g_nccl.handle) return FEMOL_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { femol_set_error("femol_dist: cannot dlopen libnccl.so.2 (%s)", dlerror()); return FEMOL_E_ARG; }
#define NSYM(field, sym)                                                                  \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                                             \
    if (!g_nccl.field) { femol_set_error("femol_dist: symbol %s missing in libnccl", sym); return FEMOL_E_ARG; }
    NSYM(GetUniqueId, "ncclGetUniqueId") NSYM(CommInitRank, "ncclCommInitRank") NSYM(CommDestroy, "ncclCommDestroy")
    NSYM(AllReduce, "ncclAllReduce") NSYM(Send, "ncclSend") NSYM(Recv, "ncclRecv") NSYM(GroupStart, "ncclGroupStart")
    NSYM(GroupEnd, "ncclGroupEnd") NSYM(GetErrorString, "ncclGetErrorString")
#undef NSYM
    g_nccl.handle = h;
    return FEMOL_OK;
}

#define NCCL_CHECK(expr)                                                                              \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != 0) {                                                                                \
            femol_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return FEMOL_E_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

}  // namespace

// Peer-memory transport (femol_dist_p2p_open / _attach): every rank owns one cudaMalloc'ed WINDOW that all other ranks
// of the box map through CUDA IPC (NVLink / NVSwitch peer stores).  A window is an array of 16-byte elements, each one
// double with its own arrival flags (LLElem below):
//   [P2P_OFF_RED]     red[3 channels][2 parities][16 ranks][16]           partial sums of the scalar all-reduces (channel 0:
//                     the stand-alone all-reduce kernel; 1 and 2: the reductions fused into the CG kernels, below)
//   [P2P_OFF_HALO]    halo[2 (from lower, from upper)][2 parities][halo_doubles]
//   [off_gather]      gather[2 parities][gather_doubles]
// A transfer = the sender's kernel stores payload + epoch into the receiver's window (slot = epoch parity) with 16-byte
// stores; the receiver's kernel polls its own window until every element carries that epoch.  Epochs are DEVICE counters
// advanced by the kernels themselves, so the launch sequence is static (CUDA-graph capturable).  Two parities suffice: a
// rank can only push epoch k + 2 after it completed transfer k + 1, which needed the receiver's push k + 1, which the
// receiver's stream issues after its transfer-k kernel has finished reading slot k.
constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_RED_MAX = 16;
constexpr int P2P_RED_CHANNELS = 3;
constexpr size_t P2P_OFF_RED = 0, P2P_OFF_HALO = (size_t)P2P_RED_CHANNELS * 2 * P2P_MAX_RANKS * P2P_RED_MAX * 16;
enum { P2P_C_HALO_EP = 0, P2P_C_RED1_EP = 1, P2P_C_RED2_EP = 2, P2P_C_RED_EP = 3, P2P_C_GAT_EP = 4, P2P_C_GAT_DONE = 6, P2P_C_ERR = 7,
       P2P_C_N = 8 };

struct P2PPeers { char* w[P2P_MAX_RANKS]; };

struct femol_p2p {
    bool on;
    char* win;                       // my window
    size_t win_bytes, halo_doubles, gather_doubles, off_gather;
    P2PPeers peers;                  // every rank's window as mapped here (peers.w[rank] = win)
    unsigned long long* ctr;         // device counters (local): epochs, arrival tickets, error flag
    long long n_halo, n_red, n_gather;   // host-side statistics
};

struct femol_comm {
    ncclComm_t comm;
    int rank, nranks;
    femol_p2p p2p;
};

namespace {

enum { D_RZ0 = 0, D_RZ1 = 1, D_PQ = 2, D_RR = 3, D_FF = 4, D_DONE = 5, D_ITER = 6, D_TOL2 = 7, D_L0 = 8, D_L1 = 9, D_L2 = 10, D_NSLOT = 16 };
constexpr int DT = 256;

__global__ void pack_nodes(int64_t n_send, int nd, const int32_t* __restrict__ nodes, const double* __restrict__ x,
                           double* __restrict__ buf) {
    const int64_t total = n_send * nd;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / nd;
        buf[t] = x[(int64_t)nodes[k] * nd + (t - k * nd)];
    }
}

// z = M^-1 r for one node (block-Jacobi inverse or diagonal stored as full blocks)
template <int ND>
__device__ __forceinline__ void prec(const double* __restrict__ dinv, int64_t i, const double (&r)[ND], double (&z)[ND]) {
    const double* B = dinv + i * ND * ND;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < ND; ++b) s += __ldg(&B[a * ND + b]) * r[b];
        z[a] = s;
    }
}

// r = f - q ; z = M^-1 r ; p = z ; local sums (r.z, r.r, f.f) -> sc[D_L0..2]
template <int ND>
__global__ void __launch_bounds__(DT)
dpcg_init(int64_t n_owned, const double* __restrict__ f, const double* __restrict__ q, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ p, double* __restrict__ sc, double* __restrict__ red) {
    __shared__ double s_red[32 * 3];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_owned; i += (int64_t)gridDim.x * blockDim.x) {
        double rv[ND], zv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { const double fv = f[i * ND + d]; rv[d] = fv - q[i * ND + d]; acc[2] += fv * fv; }
        prec<ND>(dinv, i, rv, zv);
#pragma unroll
        for (int d = 0; d < ND; ++d) { r[i * ND + d] = rv[d]; p[i * ND + d] = zv[d]; acc[0] += rv[d] * zv[d]; acc[1] += rv[d] * rv[d]; }
    }
    grid_sum_finish<3>(acc, red, red_ticket(red), sc + D_L0, s_red);
}

__global__ void dpcg_init_finalize(double* sc, double rtol, int restart) {
    sc[D_RZ0] = sc[D_L0]; sc[D_RZ1] = sc[D_L0];
    sc[D_RR] = sc[D_L1]; sc[D_FF] = sc[D_L2];
    sc[D_TOL2] = rtol * rtol * sc[D_L2];
    if (!restart) sc[D_ITER] = 0.0;
    sc[D_DONE] = (sc[D_L1] <= rtol * rtol * sc[D_L2]) ? 1.0 : 0.0;
}

// u += a p ; r -= a q ; local (r.z, r.r) -> sc[D_L0..1]
template <int ND>
__global__ void __launch_bounds__(DT)
dpcg_update(int64_t n_owned, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv,
            double* __restrict__ u, double* __restrict__ r, double* __restrict__ sc, double* __restrict__ red) {
    __shared__ double s_red[32 * 2];
    const bool done = sc[D_DONE] != 0.0;
    const int cur = (((long long)sc[D_ITER]) & 1LL) ? D_RZ1 : D_RZ0;
    const double alpha = done ? 0.0 : sc[cur] / sc[D_PQ];
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_owned; i += (int64_t)gridDim.x * blockDim.x) {
        double rv[ND], zv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const int64_t o = i * ND + d;
            if (!done) u[o] += alpha * p[o];
            rv[d] = r[o] - alpha * q[o];
            if (!done) r[o] = rv[d];
        }
        prec<ND>(dinv, i, rv, zv);
#pragma unroll
        for (int d = 0; d < ND; ++d) { acc[0] += rv[d] * zv[d]; acc[1] += rv[d] * rv[d]; }
    }
    // every rank must take part in the collective that follows, so the sums are produced even when done
    grid_sum_finish<2>(acc, red, red_ticket(red), sc + D_L0, s_red);
}

__global__ void dpcg_update_finalize(double* sc) {
    if (sc[D_DONE] != 0.0) return;
    const double pq = sc[D_PQ];
    const int cur = (((long long)sc[D_ITER]) & 1LL) ? D_RZ1 : D_RZ0;
    const int oth = cur == D_RZ0 ? D_RZ1 : D_RZ0;
    sc[oth] = sc[D_L0];
    sc[D_RR] = sc[D_L1];
    sc[D_ITER] += 1.0;
    if (!(pq > 0.0) || !(sc[D_L1] == sc[D_L1])) sc[D_DONE] = 2.0;
    else if (sc[D_L1] <= sc[D_TOL2]) sc[D_DONE] = 1.0;
}

// p = z + beta p on owned nodes (beta = rz_new / rz_old; after the finalize the new value sits in the current slot)
template <int ND>
__global__ void __launch_bounds__(DT)
dpcg_direction(int64_t n_owned, const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p,
               const double* __restrict__ sc) {
    if (sc[D_DONE] != 0.0) return;
    const int cur = (((long long)sc[D_ITER]) & 1LL) ? D_RZ1 : D_RZ0;
    const int oth = cur == D_RZ0 ? D_RZ1 : D_RZ0;
    const double beta = sc[cur] / sc[oth];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_owned; i += (int64_t)gridDim.x * blockDim.x) {
        double rv[ND], zv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) rv[d] = r[i * ND + d];
        prec<ND>(dinv, i, rv, zv);
#pragma unroll
        for (int d = 0; d < ND; ++d) p[i * ND + d] = zv[d] + beta * p[i * ND + d];
    }
}

struct Halo {
    int n_nbr;
    const int* nbr_rank;
    const int64_t* send_ptr;      // [n_nbr+1] offsets into send_nodes
    const int32_t* send_nodes;    // device
    const int64_t* recv_start;    // [n_nbr] first local (ghost) node of each neighbour's segment
    const int64_t* recv_count;    // [n_nbr]
    double* sendbuf;              // device, send_ptr[n_nbr] * nd doubles
};

int halo_exchange(femol_comm* c, const Halo& h, int nd, double* x, cudaStream_t st) {
    const int64_t n_send = h.send_ptr[h.n_nbr];
    if (n_send > 0) {
        int grid = (int)((n_send * nd + 255) / 256);
        if (grid > 1184) grid = 1184;
        pack_nodes<<<grid, 256, 0, st>>>(n_send, nd, h.send_nodes, x, h.sendbuf);
        FEMOL_LAUNCH_CHECK();
    }
    NCCL_CHECK(g_nccl.GroupStart());
    for (int k = 0; k < h.n_nbr; ++k) {
        const int64_t cnt = h.send_ptr[k + 1] - h.send_ptr[k];
        if (cnt > 0) NCCL_CHECK(g_nccl.Send(h.sendbuf + h.send_ptr[k] * nd, (size_t)(cnt * nd), ncclDouble_, h.nbr_rank[k], c->comm, st));
        if (h.recv_count[k] > 0)
            NCCL_CHECK(g_nccl.Recv(x + h.recv_start[k] * nd, (size_t)(h.recv_count[k] * nd), ncclDouble_, h.nbr_rank[k], c->comm, st));
    }
    NCCL_CHECK(g_nccl.GroupEnd());
    return FEMOL_OK;
}

template <int ND>
int dpcg_run(femol_plan_t plan, femol_comm* c, int64_t n_owned, const Halo& h, const double* vals, const double* f, double* u,
             double rtol, int max_iter, int precond, double* ws, int* h_iters, double* h_relres, cudaStream_t st) {
    const int64_t n_local = plan->n_nodes * ND;
    double* r = ws;
    double* p = r + n_local;
    double* q = p + n_local;
    double* dinv = q + n_local;
    double* sc = dinv + plan->n_nodes * ND * ND;
    double* red = sc + D_NSLOT;
    const int vgrid = femol_grid_for(n_owned, DT, plan->sm_count, 8);
    FEMOL_CUDA_CHECK(cudaMemsetAsync(sc, 0, (D_NSLOT + FEMOL_RED_WS_DOUBLES) * sizeof(double), st));
    FEMOL_CUDA_CHECK(cudaMemsetAsync(p, 0, n_local * sizeof(double), st));
    int rc = femol_precond_setup(plan, vals, precond, dinv, st);
    if (rc != FEMOL_OK) return rc;
    // an SpMV restricted to the owned rows: a "view" plan with n_nodes = n_owned (owned nodes come first)
    femol_plan owned_view = *plan;
    owned_view.n_nodes = n_owned;

    double h_sc[D_NSLOT];
    int total_iters = 0, status = FEMOL_S_NOT_CONVERGED;
    double true_rel = 0.0, prev_true_rel = 1e300;
    const int batch = 25;
    for (int restart = 0; restart < 6; ++restart) {
        if ((rc = halo_exchange(c, h, ND, u, st)) != FEMOL_OK) return rc;
        if ((rc = femol_spmv_launch(&owned_view, vals, u, q, nullptr, nullptr, nullptr, st)) != FEMOL_OK) return rc;
        dpcg_init<ND><<<vgrid, DT, 0, st>>>(n_owned, f, q, dinv, r, p, sc, red);
        FEMOL_LAUNCH_CHECK();
        NCCL_CHECK(g_nccl.AllReduce(sc + D_L0, sc + D_L0, 3, ncclDouble_, ncclSum_, c->comm, st));
        dpcg_init_finalize<<<1, 1, 0, st>>>(sc, rtol, restart);
        FEMOL_LAUNCH_CHECK();
        FEMOL_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, sizeof(h_sc), cudaMemcpyDeviceToHost, st));
        FEMOL_CUDA_CHECK(cudaStreamSynchronize(st));
        const double ff = h_sc[D_FF];
        true_rel = ff > 0.0 ? sqrt(h_sc[D_RR] / ff) : 0.0;
        if (ff == 0.0) {  // zero right-hand side (global norm): the solution is zero whatever the warm start was
            FEMOL_CUDA_CHECK(cudaMemsetAsync(u, 0, (size_t)n_local * sizeof(double), st));
            status = FEMOL_OK; break;
        }
        if (h_sc[D_DONE] == 1.0) { status = FEMOL_OK; break; }
        if (restart > 0 && true_rel <= 10.0 * rtol) { status = FEMOL_OK; break; }
        if (restart > 1 && true_rel >= 0.5 * prev_true_rel) { status = FEMOL_S_STAGNATED; break; }
        prev_true_rel = true_rel;
        if (total_iters >= max_iter) break;
        bool finished = false;
        while (!finished && total_iters < max_iter) {
            const int nb = (max_iter - total_iters) < batch ? (max_iter - total_iters) : batch;
            for (int it = 0; it < nb; ++it) {  // every rank enqueues the same collectives, converged or not
                if ((rc = halo_exchange(c, h, ND, p, st)) != FEMOL_OK) return rc;
                if ((rc = femol_spmv_launch(&owned_view, vals, p, q, p, sc + D_L0, red, st)) != FEMOL_OK) return rc;
                NCCL_CHECK(g_nccl.AllReduce(sc + D_L0, sc + D_PQ, 1, ncclDouble_, ncclSum_, c->comm, st));
                dpcg_update<ND><<<vgrid, DT, 0, st>>>(n_owned, p, q, dinv, u, r, sc, red);
                FEMOL_LAUNCH_CHECK();
                NCCL_CHECK(g_nccl.AllReduce(sc + D_L0, sc + D_L0, 2, ncclDouble_, ncclSum_, c->comm, st));
                dpcg_update_finalize<<<1, 1, 0, st>>>(sc);
                dpcg_direction<ND><<<vgrid, DT, 0, st>>>(n_owned, r, dinv, p, sc);
                FEMOL_LAUNCH_CHECK();
            }
            FEMOL_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, sizeof(h_sc), cudaMemcpyDeviceToHost, st));
            FEMOL_CUDA_CHECK(cudaStreamSynchronize(st));
            total_iters = (int)h_sc[D_ITER];
            if (h_sc[D_DONE] == 2.0) { status = FEMOL_S_BREAKDOWN; finished = true; }
            else if (h_sc[D_DONE] == 1.0) finished = true;
        }
        if (status == FEMOL_S_BREAKDOWN || !finished) break;
    }
    if (h_iters) *h_iters = total_iters;
    if (h_relres) *h_relres = true_rel;
    return status;
}

}  // namespace

extern "C" int femol_dist_unique_id(void* out128) {
    FEMOL_REQUIRE(out128, "femol_dist_unique_id: null argument");
    int rc = load_nccl();
    if (rc != FEMOL_OK) return rc;
    ncclUniqueId id;
    NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return FEMOL_OK;
}

extern "C" int femol_dist_init(const void* unique_id128, int rank, int nranks, femol_comm** out) {
    FEMOL_REQUIRE(unique_id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, "femol_dist_init: bad argument");
    int rc = load_nccl();
    if (rc != FEMOL_OK) return rc;
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    femol_comm* c = new femol_comm();
    c->rank = rank; c->nranks = nranks; c->comm = nullptr;
    memset(&c->p2p, 0, sizeof(c->p2p));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) { femol_set_error("ncclCommInitRank -> %s", g_nccl.GetErrorString(r)); delete c; return FEMOL_E_CUDA; }
    *out = c;
    return FEMOL_OK;
}

extern "C" int femol_dist_destroy(femol_comm* c) {
    if (!c) return FEMOL_OK;
    if (c->p2p.win) {
        cudaDeviceSynchronize();
        for (int r = 0; r < c->nranks; ++r)
            if (r != c->rank && c->p2p.peers.w[r]) cudaIpcCloseMemHandle(c->p2p.peers.w[r]);
        cudaFree(c->p2p.win); cudaFree(c->p2p.ctr);
        cudaGetLastError();
    }
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return FEMOL_OK;
}

extern "C" int64_t femol_pcg_dist_workspace_doubles(femol_plan_t p, int64_t n_send_nodes) {
    if (!p) return 0;
    const int64_t n = p->n_nodes * p->ndof;
    return 3 * n + n * p->ndof + D_NSLOT + FEMOL_RED_WS_DOUBLES + n_send_nodes * p->ndof + 16;
}

extern "C" int femol_pcg_dist(femol_plan_t plan, femol_comm* comm, int64_t n_owned_nodes, int n_nbr, const int* h_nbr_rank,
                              const int64_t* h_send_ptr, const int32_t* d_send_nodes, const int64_t* h_recv_start,
                              const int64_t* h_recv_count, const double* vals, const double* f, double* u, double rtol,
                              int max_iter, int precond, double* ws, int* h_iters, double* h_relres, void* stream) {
    FEMOL_REQUIRE(plan && comm && vals && f && u && ws && n_nbr >= 0, "femol_pcg_dist: null argument");
    FEMOL_REQUIRE(n_owned_nodes >= 0 && n_owned_nodes <= plan->n_nodes, "femol_pcg_dist: bad owned node count");
    FEMOL_REQUIRE(n_nbr == 0 || (h_nbr_rank && h_send_ptr && h_recv_start && h_recv_count), "femol_pcg_dist: missing halo arrays");
    FEMOL_REQUIRE(precond >= 0 && precond <= 2 && rtol > 0.0, "femol_pcg_dist: bad preconditioner / tolerance");
    static const int64_t zero_ptr[1] = {0};
    Halo h;
    h.n_nbr = n_nbr; h.nbr_rank = h_nbr_rank; h.send_ptr = n_nbr ? h_send_ptr : zero_ptr; h.send_nodes = d_send_nodes;
    h.recv_start = h_recv_start; h.recv_count = h_recv_count;
    const int64_t n = plan->n_nodes * plan->ndof;
    h.sendbuf = ws + 3 * n + n * plan->ndof + D_NSLOT + FEMOL_RED_WS_DOUBLES;
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->ndof == 2) return dpcg_run<2>(plan, comm, n_owned_nodes, h, vals, f, u, rtol, max_iter, precond, ws, h_iters, h_relres, st);
    if (plan->ndof == 6) return dpcg_run<6>(plan, comm, n_owned_nodes, h, vals, f, u, rtol, max_iter, precond, ws, h_iters, h_relres, st);
    femol_set_error("femol_pcg_dist: unsupported ndof %d", plan->ndof);
    return FEMOL_E_ARG;
}

// =====================================================================================================================
// Strip-distributed multigrid PCG (gmg.cu: "Strip-distributed hierarchy").  Transport = NCCL: a halo exchange is one
// grouped ncclSend/ncclRecv pair per neighbour on contiguous row slices (local vectors are x-fastest lattice rows, so
// nothing is packed); reductions are ncclAllReduce on device scalars, in stream.
// =====================================================================================================================
int femol_comm_rank(const femol_comm* c) { return c->rank; }
int femol_comm_size(const femol_comm* c) { return c->nranks; }

namespace {

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One window element = one double with its own arrival flags: {low word, epoch, high word, epoch} written by ONE 16-byte
// store (each 8-byte half validates itself, so the element is safe even where only 8-byte stores are atomic).  The
// receiver polls the element until both flags equal the epoch it expects - no fence, no separate flag, no second
// NVLink traversal: a transfer costs one store latency.  (NCCL's LL protocol, for FP64 payloads.)
struct __align__(16) LLElem { unsigned lo, f0, hi, f1; };

__device__ __forceinline__ void ll_store(LLElem* p, double v, unsigned ep) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(ep), "r"(hi), "r"(ep) : "memory");
}
// spin until the element carries epoch `ep`; gives up after ~2 s (a peer died): raises the error counter, after which no
// wait blocks any more (results are garbage, the host reports the error)
__device__ __forceinline__ double ll_wait(const LLElem* p, unsigned ep, unsigned long long* ctr) {
    unsigned lo, f0, hi, f1;
    long long t0 = 0;
    unsigned spins = 0;
    while (true) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
        if (f0 == ep && f1 == ep) break;
        if ((++spins & 1023u) == 0u) {
            if (ld_volatile_u64(&ctr[P2P_C_ERR]) != 0ULL) break;
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000LL) { atomicExch(&ctr[P2P_C_ERR], 1ULL); break; }
        }
    }
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ unsigned epoch32(unsigned long long ep) { return (unsigned)(ep % 0xFFFFFFFFULL) + 1u; }   // never 0 (the memset value)

__device__ __forceinline__ LLElem* win_halo(char* w) { return reinterpret_cast<LLElem*>(w + P2P_OFF_HALO); }
__device__ __forceinline__ LLElem* win_red(char* w) { return reinterpret_cast<LLElem*>(w + P2P_OFF_RED); }

// ---- dot product + all-reduce in ONE kernel ------------------------------------------------------------------------
// The CG kernels that produce a partial sum store it straight into every rank's window from their last block (the one
// that completes the deterministic grid sum), and the kernel that consumes the scalar polls the W contributions and adds
// them in rank order - no all-reduce kernel in between.  One link = one reduction channel of the windows with its own
// device epoch counter: producers use epoch counter + 1, the consumer advances the counter once every one of its blocks
// has read it.  W = 0: not fused (single rank or NCCL transport; the caller runs an all-reduce).
struct RedLink {
    P2PPeers peers;
    unsigned long long* ctr;     // device counters of the communicator
    int W, me, channel, ctr_idx;
};
__device__ __forceinline__ LLElem* red_slot(char* w, int channel, size_t par, int rank, int k) {
    return win_red(w) + ((((size_t)channel * 2 + par) * P2P_MAX_RANKS + rank) * P2P_RED_MAX + k);
}
// one thread: my partial sum k of the link's next epoch -> every rank's window
__device__ __forceinline__ void red_push(const RedLink& L, int k, double v) {
    const unsigned long long ep64 = ld_volatile_u64(&L.ctr[L.ctr_idx]) + 1ULL;
    for (int r = 0; r < L.W; ++r) ll_store(red_slot(L.peers.w[(L.me + r) % L.W], L.channel, (size_t)(ep64 & 1ULL), L.me, k), v, epoch32(ep64));
}
// any thread: sum over the ranks (rank order) of partial sum k of the link's next epoch
__device__ __forceinline__ double red_poll(const RedLink& L, int k) {
    const unsigned long long ep64 = ld_volatile_u64(&L.ctr[L.ctr_idx]) + 1ULL;
    double sum = 0.0;
    for (int r = 0; r < L.W; ++r) sum += ll_wait(red_slot(L.peers.w[L.me], L.channel, (size_t)(ep64 & 1ULL), r, k), epoch32(ep64), L.ctr);
    return sum;
}
// one thread, after EVERY block of the consuming kernel has polled
__device__ __forceinline__ void red_advance(const RedLink& L) { L.ctr[L.ctr_idx] = ld_volatile_u64(&L.ctr[L.ctr_idx]) + 1ULL; }

// Halo exchange of one vector in ONE CTA (messages are a few lattice rows: 4 - 11 k doubles): store my boundary rows
// into the neighbours' mailboxes over NVLink, then poll my own mailbox element by element into the ghost rows.  Every
// microsecond of this kernel is on the critical path of the V-cycle (1 + 3 n_dist exchanges per CG iteration).
__global__ void __launch_bounds__(1024)
p2p_halo_kernel(double* x, int64_t rowd, int64_t own_lo, int64_t own_hi, int h_up, int h_down, int lower, int upper, int me,
                P2PPeers peers, size_t hmax, unsigned long long* ctr) {
    FEMOL_PDL_SYNC();
    const unsigned long long ep64 = ld_volatile_u64(&ctr[P2P_C_HALO_EP]) + 1ULL;
    const unsigned ep = epoch32(ep64);
    const size_t par = (size_t)(ep64 & 1ULL);
    const int n_dn = (int)(h_down * rowd), n_up = (int)(h_up * rowd);
    if (lower >= 0) {   // my bottom owned rows -> the lower neighbour's "from upper" slot
        LLElem* dst = win_halo(peers.w[lower]) + (2 + par) * hmax;
        const double* src = x + own_lo * rowd;
        for (int i = threadIdx.x; i < n_dn; i += blockDim.x) ll_store(dst + i, src[i], ep);
    }
    if (upper >= 0) {   // my top owned rows -> the upper neighbour's "from lower" slot
        LLElem* dst = win_halo(peers.w[upper]) + (0 + par) * hmax;
        const double* src = x + (own_hi - h_up) * rowd;
        for (int i = threadIdx.x; i < n_up; i += blockDim.x) ll_store(dst + i, src[i], ep);
    }
    const LLElem* box = win_halo(peers.w[me]);
    if (lower >= 0) {
        double* dst = x + (own_lo - h_up) * rowd;
        for (int i = threadIdx.x; i < n_up; i += blockDim.x) dst[i] = ll_wait(box + (0 + par) * hmax + i, ep, ctr);
    }
    if (upper >= 0) {
        double* dst = x + own_hi * rowd;
        for (int i = threadIdx.x; i < n_dn; i += blockDim.x) dst[i] = ll_wait(box + (2 + par) * hmax + i, ep, ctr);
    }
    __syncthreads();           // every thread has read the epoch counter
    if (threadIdx.x == 0) ctr[P2P_C_HALO_EP] = ep64;
}

// In-place sum of `count` (<= 16) doubles over all ranks: every rank stores its values into every window, then adds the
// W contributions in rank order (bit-identical result on all ranks).  One CTA.
__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(double* buf, int count, int W, int me, P2PPeers peers, unsigned long long* ctr) {
    FEMOL_PDL_SYNC();
    __shared__ double s_val[P2P_MAX_RANKS * P2P_RED_MAX];
    const unsigned long long ep64 = ld_volatile_u64(&ctr[P2P_C_RED_EP]) + 1ULL;
    const unsigned ep = epoch32(ep64);
    const size_t par = (size_t)(ep64 & 1ULL);
    const int t = threadIdx.x;
    if (t < W * count) {
        const int peer = t / count, k = t - peer * count;
        ll_store(red_slot(peers.w[peer], 0, par, me, k), buf[k], ep);
        s_val[t] = ll_wait(red_slot(peers.w[me], 0, par, peer, k), ep, ctr);   // rank `peer`'s k-th value
    }
    __syncthreads();
    if (t < count) {
        double sum = 0.0;
        for (int r = 0; r < W; ++r) sum += s_val[r * count + t];
        buf[t] = sum;
    }
    if (t == 0) ctr[P2P_C_RED_EP] = ep64;
}

// All-gather by pushes: my `count` doubles go to offset `goff` of the gather slot of EVERY window; the whole slot (total
// doubles) is polled into `out`.
__global__ void __launch_bounds__(256)
p2p_gather_kernel(const double* __restrict__ src, int64_t count, int64_t goff, int64_t total, double* __restrict__ out, int W, int me,
                  P2PPeers peers, size_t off_gather, size_t gmax, unsigned long long* ctr) {
    FEMOL_PDL_SYNC();
    const unsigned long long ep64 = ld_volatile_u64(&ctr[P2P_C_GAT_EP]) + 1ULL;
    const unsigned ep = epoch32(ep64);
    const size_t par = (size_t)(ep64 & 1ULL);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < W; ++r) {
        const int peer = (me + r) % W;      // spread the ranks' first targets over the switch
        LLElem* dst = reinterpret_cast<LLElem*>(peers.w[peer] + off_gather) + par * gmax + goff;
        for (int64_t i = tid; i < count; i += nth) ll_store(dst + i, src[i], ep);
    }
    const LLElem* box = reinterpret_cast<const LLElem*>(peers.w[me] + off_gather) + par * gmax;
    for (int64_t i = tid; i < total; i += nth) out[i] = ll_wait(box + i, ep, ctr);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(&ctr[P2P_C_GAT_DONE], 1ULL);
        if (t == (unsigned long long)gridDim.x - 1ULL) { ctr[P2P_C_GAT_DONE] = 0ULL; ctr[P2P_C_GAT_EP] = ep64; }
    }
}

}  // namespace

int femol_dist_exchange(femol_comm* c, double* x, int width, const femol_strip& s, cudaStream_t st) {
    if (c->nranks == 1) return FEMOL_OK;
    const int64_t rowd = s.row_nodes * width;
    const bool lower = c->rank > 0, upper = c->rank < c->nranks - 1;
    const int64_t hmaxrows = s.h_up > s.h_down ? s.h_up : s.h_down;
    if (c->p2p.on && (size_t)(hmaxrows * rowd) <= c->p2p.halo_doubles) {
        FEMOL_CUDA_CHECK(femol_launch_pdl(p2p_halo_kernel, dim3(1), dim3(1024), 0, st, x, rowd, s.own_lo, s.own_hi, s.h_up, s.h_down,
                                          lower ? c->rank - 1 : -1, upper ? c->rank + 1 : -1, c->rank, c->p2p.peers,
                                          c->p2p.halo_doubles, c->p2p.ctr));
        femol_count_launch();
        ++c->p2p.n_halo;
        return FEMOL_OK;
    }
    NCCL_CHECK(g_nccl.GroupStart());
    if (lower) {
        NCCL_CHECK(g_nccl.Send(x + s.own_lo * rowd, (size_t)(s.h_down * rowd), ncclDouble_, c->rank - 1, c->comm, st));
        NCCL_CHECK(g_nccl.Recv(x + (s.own_lo - s.h_up) * rowd, (size_t)(s.h_up * rowd), ncclDouble_, c->rank - 1, c->comm, st));
    }
    if (upper) {
        NCCL_CHECK(g_nccl.Send(x + (s.own_hi - s.h_up) * rowd, (size_t)(s.h_up * rowd), ncclDouble_, c->rank + 1, c->comm, st));
        NCCL_CHECK(g_nccl.Recv(x + s.own_hi * rowd, (size_t)(s.h_down * rowd), ncclDouble_, c->rank + 1, c->comm, st));
    }
    NCCL_CHECK(g_nccl.GroupEnd());
    return FEMOL_OK;
}

int femol_dist_allreduce(femol_comm* c, double* buf, int64_t count, cudaStream_t st) {
    if (c->nranks == 1 || count == 0) return FEMOL_OK;
    if (c->p2p.on && count <= P2P_RED_MAX && c->nranks * count <= 256) {
        FEMOL_CUDA_CHECK(femol_launch_pdl(p2p_allreduce_kernel, dim3(1), dim3(256), 0, st, buf, (int)count, c->nranks, c->rank,
                                          c->p2p.peers, c->p2p.ctr));
        femol_count_launch();
        ++c->p2p.n_red;
        return FEMOL_OK;
    }
    NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble_, ncclSum_, c->comm, st));
    return FEMOL_OK;
}

// out[0..total) = concatenation over ranks of their pieces (src: count doubles, stored at offset goff); the pieces
// tile [0, total).  P2P pushes when the transport is attached, else a zero-padded all-reduce.
int femol_dist_allgather(femol_comm* c, const double* src, int64_t count, int64_t goff, int64_t total, double* out, cudaStream_t st) {
    if (c->p2p.on && (size_t)total <= c->p2p.gather_doubles) {
        int64_t blocks = (total + 4 * 256 - 1) / (4 * 256);
        if (blocks > 32) blocks = 32;
        if (blocks < 1) blocks = 1;
        FEMOL_CUDA_CHECK(femol_launch_pdl(p2p_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, st, src, count, goff, total, out,
                                          c->nranks, c->rank, c->p2p.peers, c->p2p.off_gather, c->p2p.gather_doubles, c->p2p.ctr));
        femol_count_launch();
        ++c->p2p.n_gather;
        return FEMOL_OK;
    }
    FEMOL_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)total * sizeof(double), st));
    FEMOL_CUDA_CHECK(cudaMemcpyAsync(out + goff, src, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return femol_dist_allreduce(c, out, total, st);
}

// ---- window set-up -------------------------------------------------------------------------------------------------
// 1. every rank: femol_dist_p2p_open -> its 64-byte IPC handle;  2. the host all-gathers the handles (torch.distributed);
// 3. every rank: femol_dist_p2p_attach(all handles); 4. a host barrier before the first transfer.
extern "C" int femol_dist_p2p_open(femol_comm* c, int64_t halo_doubles, int64_t gather_doubles, void* h_handle64) {
    FEMOL_REQUIRE(c && h_handle64 && halo_doubles >= 0 && gather_doubles >= 0, "femol_dist_p2p_open: bad argument");
    FEMOL_REQUIRE(c->nranks <= P2P_MAX_RANKS, "femol_dist_p2p_open: too many ranks");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    femol_p2p& w = c->p2p;
    if (w.win) {   // a second hierarchy on the same communicator: the window is shared if it is large enough
        FEMOL_REQUIRE((size_t)halo_doubles <= w.halo_doubles && (size_t)gather_doubles <= w.gather_doubles,
                      "femol_dist_p2p_open: the communicator's window is open and too small for this hierarchy (use a new communicator)");
        cudaIpcMemHandle_t h0;
        FEMOL_CUDA_CHECK(cudaIpcGetMemHandle(&h0, w.win));
        memcpy(h_handle64, &h0, sizeof(h0));
        return FEMOL_OK;
    }
    w.halo_doubles = (size_t)((halo_doubles + 31) & ~31LL);
    w.gather_doubles = (size_t)((gather_doubles + 31) & ~31LL);
    w.off_gather = P2P_OFF_HALO + 4 * w.halo_doubles * 16;
    w.win_bytes = w.off_gather + 2 * w.gather_doubles * 16;
    cudaIpcMemHandle_t h;
    if (cudaMalloc(&w.win, w.win_bytes) != cudaSuccess || cudaMemset(w.win, 0, w.win_bytes) != cudaSuccess ||
        cudaMalloc(&w.ctr, P2P_C_N * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(w.ctr, 0, P2P_C_N * sizeof(unsigned long long)) != cudaSuccess || cudaIpcGetMemHandle(&h, w.win) != cudaSuccess) {
        femol_set_error("femol_dist_p2p_open: cannot allocate / export the window (%s)", cudaGetErrorString(cudaGetLastError()));
        cudaFree(w.win); cudaFree(w.ctr);
        w.win = nullptr; w.ctr = nullptr;
        cudaGetLastError();
        return FEMOL_E_CUDA;
    }
    memcpy(h_handle64, &h, sizeof(h));
    FEMOL_CUDA_CHECK(cudaDeviceSynchronize());
    return FEMOL_OK;
}

extern "C" int femol_dist_p2p_attach(femol_comm* c, const void* h_handles) {
    FEMOL_REQUIRE(c && h_handles && c->p2p.win, "femol_dist_p2p_attach: open the window first");
    femol_p2p& w = c->p2p;
    if (w.on) return FEMOL_OK;         // already attached (a shared window)
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) { w.peers.w[r] = w.win; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)h_handles + 64 * (size_t)r, sizeof(h));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            femol_set_error("femol_dist_p2p_attach: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            for (int q = 0; q < r; ++q)       // unmap what was mapped: the transport stays off (NCCL fallback)
                if (q != c->rank && w.peers.w[q]) { cudaIpcCloseMemHandle(w.peers.w[q]); w.peers.w[q] = nullptr; }
            return FEMOL_E_CUDA;
        }
        w.peers.w[r] = (char*)ptr;
    }
    w.on = true;
    return FEMOL_OK;
}

// 0 = healthy; 1 = a wait timed out (peer lost); h_stats (optional, 3 values): halo / reduce / gather kernels launched
extern "C" int femol_dist_p2p_status(femol_comm* c, int64_t* h_stats) {
    FEMOL_REQUIRE(c, "femol_dist_p2p_status: null communicator");
    if (h_stats) { h_stats[0] = c->p2p.n_halo; h_stats[1] = c->p2p.n_red; h_stats[2] = c->p2p.n_gather; }
    if (!c->p2p.on) return 0;
    unsigned long long err = 0;
    FEMOL_CUDA_CHECK(cudaMemcpy(&err, c->p2p.ctr + P2P_C_ERR, sizeof(err), cudaMemcpyDeviceToHost));
    return err ? 1 : 0;
}

extern "C" int femol_dist_allreduce_sum(femol_comm* c, double* buf, int64_t count, void* stream) {
    FEMOL_REQUIRE(c && buf && count >= 0, "femol_dist_allreduce_sum: bad argument");
    return femol_dist_allreduce(c, buf, count, (cudaStream_t)stream);
}

namespace {

// scalar slots of the distributed multigrid PCG (device)
enum { G_RZN = 0, G_RZO = 1, G_PQ = 2, G_RR = 3, G_FF = 4, G_DONE = 5, G_ITER = 6, G_TOL2 = 7, G_L0 = 8, G_L1 = 9, G_NSLOT = 16 };

// r = f - q on the owned range ; local (r.r, f.f) -> sc[G_L0..1]
__global__ void __launch_bounds__(DT)
dg_init(int64_t n, const double* __restrict__ f, const double* __restrict__ q, double* __restrict__ r, double* __restrict__ sc,
        double* __restrict__ red) {
    __shared__ double s_red[32 * 2];
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double fv = f[i], rv = fv - q[i];
        r[i] = rv;
        acc[0] += rv * rv; acc[1] += fv * fv;
    }
    grid_sum_finish<2>(acc, red, red_ticket(red), sc + G_L0, s_red);
}

__global__ void dg_init_finalize(double* sc, double rtol, int restart) {
    sc[G_RR] = sc[G_L0]; sc[G_FF] = sc[G_L1];
    sc[G_TOL2] = rtol * rtol * sc[G_L1];
    if (!restart) sc[G_ITER] = 0.0;
    sc[G_DONE] = (sc[G_L0] <= rtol * rtol * sc[G_L1]) ? 1.0 : 0.0;
}

// out[0] = a . b over the owned range (local part); fused link: the last block also stores it into every rank's window
// as partial sum k of the link's next epoch
__global__ void __launch_bounds__(DT)
dg_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, double* __restrict__ red,
       const RedLink link, int k) {
    FEMOL_PDL_SYNC();
    __shared__ double s_red[32];
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[0] += a[i] * b[i];
    if (grid_sum_last<1>(acc, red, red_ticket(red), s_red)) {
        out[0] = acc[0];
        if (link.W) red_push(link, k, acc[0]);
    }
}

// u += alpha p ; r -= alpha q ; local r.r -> sc[G_L1] ; z = omega Dinv r (the V-cycle's first sweep), owned nodes
template <int ND>
__global__ void __launch_bounds__(DT)
dg_update(int64_t n_nodes, const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ u, double* __restrict__ r,
          double* __restrict__ sc, double* __restrict__ red, const double* __restrict__ dinv, double omega, double* __restrict__ z,
          const RedLink pq_link, const RedLink rr_link) {
    FEMOL_PDL_SYNC();
    __shared__ double s_red[32];
    __shared__ double s_pq;
    if (pq_link.W) {          // fused all-reduce of p.Kp: every block adds the W contributions itself (rank order)
        if (threadIdx.x == 0) s_pq = red_poll(pq_link, 0);
        __syncthreads();
    }
    const double pq = pq_link.W ? s_pq : sc[G_PQ];
    const bool done = sc[G_DONE] != 0.0;
    const double alpha = done ? 0.0 : sc[G_RZN] / pq;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        double rv[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const int64_t o = i * ND + d;
            if (!done) u[o] += alpha * p[o];
            rv[d] = r[o] - alpha * q[o];
            if (!done) r[o] = rv[d];
            acc[0] += rv[d] * rv[d];
        }
        if (!done && dinv) {
            const double* B = dinv + i * ND * ND;
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < ND; ++b) s += __ldg(&B[a * ND + b]) * rv[b];
                z[i * ND + a] = omega * s;
            }
        }
    }
    // every rank takes part in the reduction that follows, converged or not
    if (grid_sum_last<1>(acc, red, red_ticket(red), s_red)) {      // the last block: every block has polled p.Kp by now
        sc[G_L1] = acc[0];
        if (pq_link.W) { sc[G_PQ] = pq; red_advance(pq_link); }
        if (rr_link.W) red_push(rr_link, 1, acc[0]);
    }
}

// after the all-reduce of (r.z, r.r): rotate the r.z slots, count the iteration, test convergence
__global__ void dg_iter_finalize(double* sc, int first, const RedLink link) {
    FEMOL_PDL_SYNC();
    if (link.W) {             // fused all-reduce of (r.z, r.r): partial sums 0 (dg_dot) and 1 (dg_update) of the link
        sc[G_L0] = red_poll(link, 0);
        if (!first) sc[G_L1] = red_poll(link, 1);
        red_advance(link);
    }
    if (sc[G_DONE] != 0.0) return;
    sc[G_RZO] = sc[G_RZN];
    sc[G_RZN] = sc[G_L0];
    if (first) return;
    sc[G_RR] = sc[G_L1];
    sc[G_ITER] += 1.0;
    const double pq = sc[G_PQ];
    if (!(pq > 0.0) || !(sc[G_L1] == sc[G_L1])) sc[G_DONE] = 2.0;
    else if (sc[G_L1] <= sc[G_TOL2]) sc[G_DONE] = 1.0;
}

// first: p = z ; else p = z + (rz_new / rz_old) p     (owned range)
__global__ void __launch_bounds__(DT)
dg_direction(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* __restrict__ sc, int first) {
    FEMOL_PDL_SYNC();
    if (sc[G_DONE] != 0.0) return;
    const double beta = first ? 0.0 : sc[G_RZN] / sc[G_RZO];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = first ? z[i] : z[i] + beta * p[i];
}

}  // namespace

bool femol_gmg_first_sweep(femol_gmg* g, const double** dinv, double* omega);
int femol_gmg_iteration_hint(femol_gmg* g, int set);
extern "C" int femol_dist_p2p_status(femol_comm* c, int64_t* h_stats);

extern "C" int64_t femol_pcg_gmg_dist_workspace_doubles(femol_plan_t p) {
    if (!p) return 0;
    return 4 * p->n_nodes * p->ndof + G_NSLOT + FEMOL_RED_WS_DOUBLES;
}

// K u = f on the strip partition with the distributed V-cycle as preconditioner.  plan / vals / f / u are the LOCAL
// arrays of the rank's strip (all local rows; u and the result are valid on owned rows, ghost rows of u are refreshed).
namespace {
int pcg_gmg_dist_run(femol_plan_t plan, femol_gmg* gmg, const double* vals, const double* f, double* u, double rtol,
                     int max_iter, double* ws, int* h_iters, double* h_relres, cudaStream_t st, cudaGraphExec_t& exec) {
    femol_comm* c = femol_gmg_dist_comm(gmg);
    const femol_strip s = *femol_gmg_dist_strip(gmg, 0);
    const int nd = plan->ndof;
    FEMOL_REQUIRE(s.row_nodes * s.n_rows == plan->n_nodes, "femol_pcg_gmg_dist: plan and strip disagree");
    const int64_t n = plan->n_nodes * nd;
    const int64_t off = s.own_lo * s.row_nodes * nd;                 // first owned entry
    const int64_t n_own_nodes = (s.own_hi - s.own_lo) * s.row_nodes;
    const int64_t n_own = n_own_nodes * nd;
    double* r = ws; double* p = r + n; double* q = p + n; double* z = q + n; double* sc = z + n; double* red = sc + G_NSLOT;
    const int vgrid = femol_grid_for(n_own, DT, plan->sm_count, 8);
    const int ngrid = femol_grid_for(n_own_nodes, DT, plan->sm_count, 8);
    const double* dinv0 = nullptr; double omega0 = 0.0;
    const bool first_sweep = femol_gmg_first_sweep(gmg, &dinv0, &omega0);
    FEMOL_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)(4 * n + G_NSLOT + FEMOL_RED_WS_DOUBLES) * sizeof(double), st));
    double h_sc[G_NSLOT];
    int total_iters = 0, status = FEMOL_S_NOT_CONVERGED, rc;
    double true_rel = 0.0, prev_true_rel = 1e300;
    const int batch = 4;
    auto true_residual = [&](int restart) -> int {
        if ((rc = femol_dist_exchange(c, u, nd, s, st)) != FEMOL_OK) return rc;
        if ((rc = femol_apply_launch(plan, vals, u, q, nullptr, nullptr, nullptr, st)) != FEMOL_OK) return rc;
        dg_init<<<vgrid, DT, 0, st>>>(n_own, f + off, q + off, r + off, sc, red);
        FEMOL_LAUNCH_CHECK();
        if ((rc = femol_dist_allreduce(c, sc + G_L0, 2, st)) != FEMOL_OK) return rc;
        dg_init_finalize<<<1, 1, 0, st>>>(sc, rtol, restart);
        FEMOL_LAUNCH_CHECK();
        FEMOL_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, sizeof(h_sc), cudaMemcpyDeviceToHost, st));
        FEMOL_CUDA_CHECK(cudaStreamSynchronize(st));
        true_rel = h_sc[G_FF] > 0.0 ? sqrt(h_sc[G_RR] / h_sc[G_FF]) : 0.0;
        return FEMOL_OK;
    };
    // reductions fused into the CG kernels when the peer-memory windows are attached (FEMOL_DIST_FUSED_REDUCE=0: separate
    // all-reduce kernels)
    RedLink pq_link, rz_link;
    memset(&pq_link, 0, sizeof(pq_link));
    memset(&rz_link, 0, sizeof(rz_link));
    {
        const char* ef = getenv("FEMOL_DIST_FUSED_REDUCE");
        if (c->p2p.on && c->nranks > 1 && !(ef && atoi(ef) == 0)) {
            pq_link.peers = c->p2p.peers; pq_link.ctr = c->p2p.ctr; pq_link.W = c->nranks; pq_link.me = c->rank;
            rz_link = pq_link;
            pq_link.channel = 1; pq_link.ctr_idx = P2P_C_RED1_EP;
            rz_link.channel = 2; rz_link.ctr_idx = P2P_C_RED2_EP;
        }
    }
    const bool fused_red = pq_link.W > 0;
    auto rz_and_direction = [&](int first) -> int {
        FEMOL_CUDA_CHECK(femol_launch_pdl(dg_dot, dim3(vgrid), dim3(DT), 0, st, n_own, (const double*)(r + off), (const double*)(z + off), sc + G_L0, red, rz_link, 0));
        femol_count_launch();
        if (fused_red) ++c->p2p.n_red;
        else if ((rc = femol_dist_allreduce(c, sc + G_L0, 2, st)) != FEMOL_OK) return rc;
        FEMOL_CUDA_CHECK(femol_launch_pdl(dg_iter_finalize, dim3(1), dim3(1), 0, st, sc, first, rz_link));
        femol_count_launch();
        FEMOL_CUDA_CHECK(femol_launch_pdl(dg_direction, dim3(vgrid), dim3(DT), 0, st, n_own, (const double*)(z + off), p + off, (const double*)sc, first));
        femol_count_launch();
        return FEMOL_OK;
    };
    auto iteration = [&]() -> int {
        if ((rc = femol_dist_exchange(c, p, nd, s, st)) != FEMOL_OK) return rc;
        if ((rc = femol_apply_launch(plan, vals, p, q, nullptr, nullptr, nullptr, st)) != FEMOL_OK) return rc;
        FEMOL_CUDA_CHECK(femol_launch_pdl(dg_dot, dim3(vgrid), dim3(DT), 0, st, n_own, (const double*)(p + off), (const double*)(q + off), sc + G_PQ, red, pq_link, 0));
        femol_count_launch();
        if (fused_red) ++c->p2p.n_red;
        else if ((rc = femol_dist_allreduce(c, sc + G_PQ, 1, st)) != FEMOL_OK) return rc;
        const double* dv = first_sweep ? dinv0 + s.own_lo * s.row_nodes * nd * nd : nullptr;
        if (nd == 2) FEMOL_CUDA_CHECK(femol_launch_pdl(dg_update<2>, dim3(ngrid), dim3(DT), 0, st, n_own_nodes, (const double*)(p + off), (const double*)(q + off), u + off, r + off, sc, red, dv, omega0, z + off, pq_link, rz_link));
        else FEMOL_CUDA_CHECK(femol_launch_pdl(dg_update<6>, dim3(ngrid), dim3(DT), 0, st, n_own_nodes, (const double*)(p + off), (const double*)(q + off), u + off, r + off, sc, red, dv, omega0, z + off, pq_link, rz_link));
        femol_count_launch();
        if ((rc = femol_gmg_dist_vcycle(gmg, r, z, st, first_sweep)) != FEMOL_OK) return rc;
        return rz_and_direction(0);
    };
    bool use_graph = getenv("FEMOL_NO_GRAPH") == nullptr && femol_comm_size(c) > 0;
    long long graph_kernels = 0, graph_p2p[3] = {0, 0, 0};
    for (int restart = 0; restart < 6; ++restart) {
        if ((rc = true_residual(restart)) != FEMOL_OK) return rc;
        if (h_sc[G_FF] == 0.0) {  // zero right-hand side (global norm): the solution is zero whatever the warm start was
            FEMOL_CUDA_CHECK(cudaMemsetAsync(u, 0, (size_t)n * sizeof(double), st));
            status = FEMOL_OK; break;
        }
        if (h_sc[G_DONE] == 1.0) { status = FEMOL_OK; break; }
        if (restart > 0 && true_rel <= 10.0 * rtol) { status = FEMOL_OK; break; }
        if (restart > 1 && true_rel >= 0.5 * prev_true_rel) { status = FEMOL_S_STAGNATED; break; }
        prev_true_rel = true_rel;
        if (total_iters >= max_iter) break;
        if ((rc = femol_gmg_dist_vcycle(gmg, r, z, st, false)) != FEMOL_OK) return rc;
        if ((rc = rz_and_direction(1)) != FEMOL_OK) return rc;
        if (use_graph && !exec) {   // one CG iteration (static launch sequence: device-side scalars and epochs) -> CUDA graph
            cudaGraph_t graph = nullptr;
            const long long before = femol_launch_count(0);
            const long long p2p_before[3] = {c->p2p.n_halo, c->p2p.n_red, c->p2p.n_gather};
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int crc = iteration();
                graph_kernels = femol_launch_count(0) - before;
                femol_count_launches(-graph_kernels);
                // transfers captured, not executed: they are counted per graph launch below
                graph_p2p[0] = c->p2p.n_halo - p2p_before[0]; graph_p2p[1] = c->p2p.n_red - p2p_before[1];
                graph_p2p[2] = c->p2p.n_gather - p2p_before[2];
                c->p2p.n_halo = p2p_before[0]; c->p2p.n_red = p2p_before[1]; c->p2p.n_gather = p2p_before[2];
                const cudaError_t e = cudaStreamEndCapture(st, &graph);
                if (crc != FEMOL_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
                if (graph) cudaGraphDestroy(graph);
            }
            cudaGetLastError();
            if (!exec) use_graph = false;
        }
        bool finished = false;
        int first_batch = restart == 0 ? femol_gmg_iteration_hint(gmg, -1) - 3 : batch;
        while (!finished && total_iters < max_iter) {
            int nb = first_batch > batch ? first_batch : batch;
            first_batch = 0;
            if (nb > max_iter - total_iters) nb = max_iter - total_iters;
            for (int it = 0; it < nb; ++it) {      // every rank enqueues the same sequence, converged or not
                if (exec) {
                    FEMOL_CUDA_CHECK(cudaGraphLaunch(exec, st));
                    femol_count_launches(graph_kernels);
                    c->p2p.n_halo += graph_p2p[0]; c->p2p.n_red += graph_p2p[1]; c->p2p.n_gather += graph_p2p[2];
                }
                else if ((rc = iteration()) != FEMOL_OK) return rc;
            }
            FEMOL_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, sizeof(h_sc), cudaMemcpyDeviceToHost, st));
            FEMOL_CUDA_CHECK(cudaStreamSynchronize(st));
            if (femol_dist_p2p_status(c, nullptr) != 0) {
                femol_set_error("femol_pcg_gmg_dist: a peer-memory wait timed out (a rank of the job is gone)");
                return FEMOL_E_CUDA;
            }
            total_iters = (int)h_sc[G_ITER];
            if (h_sc[G_DONE] == 2.0) { status = FEMOL_S_BREAKDOWN; finished = true; }
            else if (h_sc[G_DONE] == 1.0) finished = true;
        }
        if (status == FEMOL_S_BREAKDOWN || !finished) break;
    }
    femol_gmg_iteration_hint(gmg, total_iters);
    if (status == FEMOL_S_NOT_CONVERGED) {
        if ((rc = true_residual(1)) != FEMOL_OK) return rc;
        if (true_rel <= 10.0 * rtol) status = FEMOL_OK;
    }
    // the caller's sensitivities read u on the ghost rows of its elements
    if ((rc = femol_dist_exchange(c, u, nd, s, st)) != FEMOL_OK) return rc;
    if (h_iters) *h_iters = total_iters;
    if (h_relres) *h_relres = true_rel;
    return status;
}
}  // namespace

extern "C" int femol_pcg_gmg_dist(femol_plan_t plan, femol_gmg* gmg, const double* vals, const double* f, double* u, double rtol,
                                  int max_iter, double* ws, int* h_iters, double* h_relres, void* stream) {
    FEMOL_REQUIRE(plan && gmg && vals && f && u && ws, "femol_pcg_gmg_dist: null argument");
    FEMOL_REQUIRE(rtol > 0.0 && max_iter >= 0, "femol_pcg_gmg_dist: bad tolerance / iteration limit");
    FEMOL_REQUIRE(femol_gmg_is_dist(gmg), "femol_pcg_gmg_dist: the hierarchy is not distributed (femol_gmg_set_dist)");
    FEMOL_REQUIRE(plan->ndof == 2 || plan->ndof == 6, "femol_pcg_gmg_dist: ndof must be 2 or 6");
    // stream capture is not permitted on the legacy default stream: run on the library's side stream (as femol_pcg_gmg)
    cudaStream_t user = (cudaStream_t)stream, st = user;
    bool forked = false;
    if (getenv("FEMOL_NO_GRAPH") == nullptr && femol_stream_is_default(user)) {
        int frc = femol_stream_fork(user, &st);
        if (frc != FEMOL_OK) return frc;
        forked = true;
    }
    cudaGraphExec_t exec = nullptr;
    int rc = pcg_gmg_dist_run(plan, gmg, vals, f, u, rtol, max_iter, ws, h_iters, h_relres, st, exec);
    if (exec) cudaGraphExecDestroy(exec);
    if (forked) {
        const int jrc = femol_stream_join(st, user);
        if (rc >= 0 && jrc != FEMOL_OK) rc = jrc;
    }
    return rc;
}
