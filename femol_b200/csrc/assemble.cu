// K3: global K / M assembly into CSR values with fused SIMP density scaling and Dirichlet BCs.
//
// Replaces (reference olivecha/FEMOL): _K_structured_mesh_data_base/_coating FEM.py:550-606,
// _K_unstructured_* :608-695, _M_* :737-871 (the N_e*size^2 "data" vector), the Python BC scan
// _apply_boundary_conditions_to_matrix_data :902-930, and the value half of scipy's
// csr_matrix(...).sum_duplicates() :498-501, 532-534.
//
// Formulation: a sort-free, atomic-free GATHER.  Work item w = (node row i, dof da, neighbour k)
// owns the ndof contiguous CSR values vals[ndof*w .. ndof*w+ndof) — items are numbered so that the
// output is written as one linear, fully coalesced stream (16 B / 48 B per thread).  Each item walks
// the transposed scatter map (plan.src: the (element, a, b) triples that hit node block (i, k)),
// scales by x_e^p (precomputed once per element, never materialising the N_e*size^2 data vector)
// and applies the boundary conditions on the fly.  Summation order is fixed (ascending element id),
// so the result is bit-reproducible run to run.
//
// HBM traffic per assembly (algorithmic): 8*nnz written + 4*nnz_node (src_ptr) + 4*N_e*nn^2 (src)
// + 8*N_e (x^p) [+ 8*N_e*size^2 when element matrices differ per element].
#include "common.cuh"

namespace {

__global__ void density_power(int64_t n, const double* __restrict__ x, double expo, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = pow(x[i], expo);
}

constexpr int ASM_THREADS = 256;
constexpr int ASM_NODES = 64;  // nodes per CTA tile

template <int ND, int NN>
__global__ void __launch_bounds__(ASM_THREADS)
assemble_kernel(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                const int32_t* __restrict__ src_ptr, const uint32_t* __restrict__ src,
                const double* __restrict__ e_base, int base_pe, const double* __restrict__ e_coat, int coat_pe,
                const double* __restrict__ xpow, const uint8_t* __restrict__ fixed, double* __restrict__ vals) {
    constexpr int S = NN * ND;
    constexpr int S2 = S * S;
    __shared__ int32_t s_ptr[ASM_NODES + 1];
    extern __shared__ double s_mat[];  // [S2] base (if !base_pe) | [S2] coat (if e_coat && !coat_pe)
    double* s_base = s_mat;
    double* s_coat = s_mat + (base_pe ? 0 : S2);
    if (!base_pe)
        for (int i = threadIdx.x; i < S2; i += blockDim.x) s_base[i] = e_base[i];
    if (e_coat && !coat_pe)
        for (int i = threadIdx.x; i < S2; i += blockDim.x) s_coat[i] = e_coat[i];

    const int64_t n_tiles = (n_nodes + ASM_NODES - 1) / ASM_NODES;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * ASM_NODES;
        const int cnt = (int)min((int64_t)ASM_NODES, n_nodes - i0);
        __syncthreads();
        for (int i = threadIdx.x; i <= cnt; i += blockDim.x) s_ptr[i] = adj_ptr[i0 + i];
        __syncthreads();
        const int p0 = s_ptr[0];
        const int n_items = (s_ptr[cnt] - p0) * ND;
        for (int t = threadIdx.x; t < n_items; t += blockDim.x) {
            const int blk = p0 + t / ND;  // some node block of the owning node (used to locate the node)
            int lo = 0, hi = cnt;         // last node with s_ptr[node] <= blk
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_ptr[mid] <= blk) lo = mid; else hi = mid;
            }
            const int64_t i = i0 + lo;
            const int rp = s_ptr[lo];
            const int len = s_ptr[lo + 1] - rp;
            const int j = t - (rp - p0) * ND;  // item index within the node: (da, k)
            const int da = j / len, k = j - da * len;
            const int pos = rp + k;
            const int32_t jn = __ldg(&adj_idx[pos]);
            double acc[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) acc[d] = 0.0;
            const int sb = __ldg(&src_ptr[pos]), se = __ldg(&src_ptr[pos + 1]);
            for (int s = sb; s < se; ++s) {
                const uint32_t pk = __ldg(&src[s]);
                const int64_t e = pk >> 6;
                const int a = (pk >> 3) & 7, b = pk & 7;
                const double w = xpow ? __ldg(&xpow[e]) : 1.0;
                const int off = (a * ND + da) * S + b * ND;
                if (e_coat) {
                    const double* eb = base_pe ? e_base + e * S2 + off : s_base + off;
                    const double* ec = coat_pe ? e_coat + e * S2 + off : s_coat + off;
#pragma unroll
                    for (int d = 0; d < ND; ++d) acc[d] += eb[d] + w * ec[d];
                } else {
                    const double* eb = base_pe ? e_base + e * S2 + off : s_base + off;
#pragma unroll
                    for (int d = 0; d < ND; ++d) acc[d] += w * eb[d];
                }
            }
            if (fixed) {  // FEM.py:902-930: fixed rows and columns zeroed, unit diagonal
                const uint32_t fr = fixed[i], fc = fixed[jn];
                const bool row_fixed = (fr >> da) & 1u;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const bool col_fixed = (fc >> d) & 1u;
                    if (row_fixed || col_fixed) acc[d] = (row_fixed && jn == (int32_t)i && d == da) ? 1.0 : 0.0;
                }
            }
            double* o = vals + ((int64_t)p0 * ND + t) * ND;
            if (ND % 2 == 0) {
#pragma unroll
                for (int d = 0; d < ND; d += 2) *reinterpret_cast<double2*>(o + d) = make_double2(acc[d], acc[d + 1]);
            } else {
#pragma unroll
                for (int d = 0; d < ND; ++d) o[d] = acc[d];
            }
        }
    }
}

// Main kernel: thread (node i, slot k) owns the ND x ND block of node pair (i, adj[i][k]) — no search and
// no run-time division (SLOTS is a compile-time bound on the row length in nodes: 9 for structured Q4, so
// interior rows use every lane).  The first ncu capture showed the item kernel above to be issue-bound
// (451 M warp instructions, 16 % of DRAM peak; profiles/r01_*); this one also reads the contributing
// (element, a, b) triples from the fixed-width ELL copy with one 16-byte load (W = 4 or 8 sources; W = 0:
// pointer + list walk) and issues the x^p loads of all sources together.
template <int ND, int NN, int W>
struct BlockLoad {  // everything one node block needs from global memory, fetched before any use
    int pb, len;
    bool active;
    uint32_t pk[W ? W : 1];
    double w[W ? W : 1];
    int32_t jn;
    uint32_t fr, fc;
};

template <int ND, int NN, int W>
__device__ __forceinline__ void load_phase1(BlockLoad<ND, NN, W>& L, int64_t i, int k, bool valid,
                                            const int32_t* __restrict__ adj_ptr) {
    L.pb = valid ? __ldg(&adj_ptr[i]) : 0;
    L.len = valid ? __ldg(&adj_ptr[i + 1]) - L.pb : 0;
    L.active = valid && k < L.len;
}
template <int ND, int NN, int W>
__device__ __forceinline__ void load_phase2(BlockLoad<ND, NN, W>& L, int64_t i, int k, const int32_t* __restrict__ adj_idx,
                                            const uint32_t* __restrict__ src_ell, const uint8_t* __restrict__ fixed) {
    constexpr int WMAX = W ? W : 1;
#pragma unroll
    for (int q = 0; q < WMAX; ++q) L.pk[q] = 0xFFFFFFFFu;
    L.jn = 0; L.fr = 0; L.fc = 0;
    if (L.active) {
        const int64_t pos = (int64_t)L.pb + k;
        if (W) {
#pragma unroll
            for (int q = 0; q < WMAX / 4; ++q) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(src_ell + pos * WMAX) + q);
                L.pk[4 * q] = v.x; L.pk[4 * q + 1] = v.y; L.pk[4 * q + 2] = v.z; L.pk[4 * q + 3] = v.w;
            }
        }
        if (fixed) { L.jn = __ldg(&adj_idx[pos]); L.fr = fixed[i]; }
    }
}
template <int ND, int NN, int W>
__device__ __forceinline__ void load_phase3(BlockLoad<ND, NN, W>& L, const double* __restrict__ xpow,
                                            const uint8_t* __restrict__ fixed) {
    constexpr int WMAX = W ? W : 1;
#pragma unroll
    for (int q = 0; q < WMAX; ++q) L.w[q] = (L.pk[q] != 0xFFFFFFFFu && xpow) ? __ldg(&xpow[L.pk[q] >> 6]) : 1.0;
    if (L.active && fixed) L.fc = fixed[L.jn];
}

template <int ND, int NN, int W>
__device__ __forceinline__ void finish_block(const BlockLoad<ND, NN, W>& L, int64_t i, int k,
                                             const int32_t* __restrict__ src_ptr, const uint32_t* __restrict__ src,
                                             const double* __restrict__ e_base, int base_pe, const double* s_base,
                                             const double* __restrict__ e_coat, int coat_pe, const double* s_coat,
                                             const double* __restrict__ xpow, double* __restrict__ vals) {
    constexpr int S = NN * ND;
    constexpr int S2 = S * S;
    constexpr int WMAX = W ? W : 1;
    if (!L.active) return;
    double acc[ND][ND];
#pragma unroll
    for (int a = 0; a < ND; ++a)
#pragma unroll
        for (int b = 0; b < ND; ++b) acc[a][b] = 0.0;

    auto contribute = [&](uint32_t pk, double w) {
        const int64_t e = pk >> 6;
        const int off = (int)((pk >> 3) & 7) * ND * S + (int)(pk & 7) * ND;
        const double* eb = base_pe ? e_base + e * S2 + off : s_base + off;
        if (e_coat) {
            const double* ec = coat_pe ? e_coat + e * S2 + off : s_coat + off;
#pragma unroll
            for (int a = 0; a < ND; ++a)
#pragma unroll
                for (int b = 0; b < ND; ++b) acc[a][b] += eb[a * S + b] + w * ec[a * S + b];
        } else {
#pragma unroll
            for (int a = 0; a < ND; ++a)
#pragma unroll
                for (int b = 0; b < ND; ++b) acc[a][b] += w * eb[a * S + b];
        }
    };
    if (W) {
#pragma unroll
        for (int q = 0; q < WMAX; ++q)
            if (L.pk[q] != 0xFFFFFFFFu) contribute(L.pk[q], L.w[q]);
    } else {
        const int64_t pos = (int64_t)L.pb + k;
        const int sb = __ldg(&src_ptr[pos]), se = __ldg(&src_ptr[pos + 1]);
        for (int sidx = sb; sidx < se; ++sidx) {
            const uint32_t pk = __ldg(&src[sidx]);
            contribute(pk, xpow ? __ldg(&xpow[pk >> 6]) : 1.0);
        }
    }
    if (L.fr | L.fc) {  // FEM.py:902-930: fixed rows and columns zeroed, unit diagonal
        const bool diag = L.jn == (int32_t)i;
#pragma unroll
        for (int a = 0; a < ND; ++a)
#pragma unroll
            for (int b = 0; b < ND; ++b)
                if (((L.fr >> a) | (L.fc >> b)) & 1u) acc[a][b] = (diag && a == b) ? 1.0 : 0.0;
    }
    double* o = vals + (int64_t)ND * ND * L.pb + (int64_t)k * ND;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
        double2* o2 = reinterpret_cast<double2*>(o + (int64_t)a * L.len * ND);
#pragma unroll
        for (int b = 0; b < ND; b += 2) __stcs(&o2[b >> 1], make_double2(acc[a][b], acc[a][b + 1]));
    }
}

template <int ND, int NN, int SLOTS, int W>
__global__ void __launch_bounds__(256, ND == 2 ? 4 : 2)
assemble_slots_kernel(int64_t n_nodes, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                      const int32_t* __restrict__ src_ptr, const uint32_t* __restrict__ src,
                      const uint32_t* __restrict__ src_ell,
                      const double* __restrict__ e_base, int base_pe, const double* __restrict__ e_coat, int coat_pe,
                      const double* __restrict__ xpow, const uint8_t* __restrict__ fixed, double* __restrict__ vals) {
    constexpr int S2 = NN * ND * NN * ND;
    constexpr int NPB = 256 / SLOTS;  // nodes per block pass
    constexpr int ILP = ND == 2 ? 2 : 1;  // node blocks in flight per thread
    extern __shared__ double s_mat[];
    double* s_base = s_mat;
    double* s_coat = s_mat + (base_pe ? 0 : S2);
    if (!base_pe)
        for (int i = threadIdx.x; i < S2; i += blockDim.x) s_base[i] = e_base[i];
    if (e_coat && !coat_pe)
        for (int i = threadIdx.x; i < S2; i += blockDim.x) s_coat[i] = e_coat[i];
    __syncthreads();
    const int ln = threadIdx.x / SLOTS, k = threadIdx.x - ln * SLOTS;
    if (ln >= NPB) return;
    const int64_t step = (int64_t)gridDim.x * NPB;
    for (int64_t i0 = (int64_t)blockIdx.x * NPB + ln; i0 < n_nodes; i0 += ILP * step) {
        BlockLoad<ND, NN, W> L[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) load_phase1(L[u], i0 + u * step, k, i0 + u * step < n_nodes, adj_ptr);
#pragma unroll
        for (int u = 0; u < ILP; ++u) load_phase2(L[u], i0 + u * step, k, adj_idx, src_ell, fixed);
#pragma unroll
        for (int u = 0; u < ILP; ++u) load_phase3(L[u], xpow, fixed);
#pragma unroll
        for (int u = 0; u < ILP; ++u)
            finish_block<ND, NN, W>(L[u], i0 + u * step, k, src_ptr, src, e_base, base_pe, s_base, e_coat, coat_pe, s_coat,
                                    xpow, vals);
    }
}

template <int ND, int NN, int SLOTS, int W>
int launch_slots_w(femol_plan_t p, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                   const double* xpow, const uint8_t* fixed, double* vals, size_t smem, cudaStream_t st) {
    auto kern = assemble_slots_kernel<ND, NN, SLOTS, W>;
    if (smem > 48 * 1024) FEMOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int NPB = 256 / SLOTS;
    int64_t grid = (p->n_nodes + NPB - 1) / NPB;
    const int64_t cap = (int64_t)p->sm_count * 32;
    if (grid > cap) grid = cap;
    kern<<<(int)grid, 256, smem, st>>>(p->n_nodes, p->adj_ptr, p->adj_idx, p->src_ptr, p->src, p->src_ell, e_base, base_pe,
                                       e_coat, coat_pe, xpow, fixed, vals);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

template <int ND, int NN, int SLOTS>
int launch_slots(femol_plan_t p, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                 const double* xpow, const uint8_t* fixed, double* vals, size_t smem, cudaStream_t st) {
    if (p->ell_width == 4) return launch_slots_w<ND, NN, SLOTS, 4>(p, e_base, base_pe, e_coat, coat_pe, xpow, fixed, vals, smem, st);
    if (p->ell_width == 8) return launch_slots_w<ND, NN, SLOTS, 8>(p, e_base, base_pe, e_coat, coat_pe, xpow, fixed, vals, smem, st);
    return launch_slots_w<ND, NN, SLOTS, 0>(p, e_base, base_pe, e_coat, coat_pe, xpow, fixed, vals, smem, st);
}

template <int ND, int NN>
int launch_assemble(femol_plan_t p, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                    const double* xpow, const uint8_t* fixed, double* vals, cudaStream_t st) {
    constexpr int S2 = NN * ND * NN * ND;
    size_t smem = 0;
    if (!base_pe) smem += S2 * sizeof(double);
    if (e_coat && !coat_pe) smem += S2 * sizeof(double);
    const int mr = p->max_row_nodes;
#define SLOT_CASE(SL) \
    if (mr <= SL) return launch_slots<ND, NN, SL>(p, e_base, base_pe, e_coat, coat_pe, xpow, fixed, vals, smem, st);
    SLOT_CASE(7) SLOT_CASE(9) SLOT_CASE(13) SLOT_CASE(19) SLOT_CASE(21) SLOT_CASE(32)
#undef SLOT_CASE
    // very irregular meshes (a node with more than 32 neighbours): generic item kernel
    auto kern = assemble_kernel<ND, NN>;
    if (smem > 48 * 1024) FEMOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (p->n_nodes + ASM_NODES - 1) / ASM_NODES;
    int64_t grid = tiles;
    const int64_t cap = (int64_t)p->sm_count * 8;
    if (grid > cap) grid = cap;
    kern<<<(int)grid, ASM_THREADS, smem, st>>>(p->n_nodes, p->adj_ptr, p->adj_idx, p->src_ptr, p->src, e_base, base_pe,
                                               e_coat, coat_pe, xpow, fixed, vals);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}


// ---- structured rectangle_Q4 grids with a shared element matrix (femol_plan_set_grid) ---------------------------------
// One thread per node.  Which elements two nodes share, and with which local indices, follows from their grid
// coordinates, so nothing but the four element weights and the Dirichlet mask is read: the kernel is a pure 576 MB
// write stream on cfg 2 instead of three dependent gather levels (adjacency -> source list -> x^p).  Contributions
// are added in ascending element order, exactly like the list kernels: the values are bit-identical.
template <int ND, bool COAT>
__global__ void __launch_bounds__(128)
assemble_grid_kernel(int nx, int ny, const int32_t* __restrict__ adj_ptr, const double* __restrict__ e_base,
                     const double* __restrict__ e_coat, const double* __restrict__ xpow, const uint8_t* __restrict__ fixed,
                     double* __restrict__ vals) {
    constexpr int S = 4 * ND, S2 = S * S;
    constexpr int T = ND == 2 ? 1 : 4;          // threads per node (they split the nine neighbours)
    constexpr int NPP = 128 / T;                // nodes per CTA pass
    extern __shared__ __align__(16) double s_dyn[];
    double* s_mat = s_dyn;                      // [S2] base | [S2] coat
    double* s_out = s_dyn + (COAT ? 2 : 1) * S2;  // the pass's contiguous slice of vals: <= NPP * 9 * ND * ND doubles
    for (int t = threadIdx.x; t < S2; t += blockDim.x) { s_mat[t] = e_base[t]; if (COAT) s_mat[S2 + t] = e_coat[t]; }
    const int npx = nx + 1, n_nodes = npx * (ny + 1);
    const int ln = threadIdx.x / T, sub = threadIdx.x - ln * T;
    for (int i0 = blockIdx.x * NPP; i0 < n_nodes; i0 += gridDim.x * NPP) {
        __syncthreads();    // element matrices staged / previous pass copied out
        const int i1 = min(i0 + NPP, n_nodes);
        const int pb0 = __ldg(&adj_ptr[i0]), pb1 = __ldg(&adj_ptr[i1]);
        const int i = i0 + ln;
        if (i < i1) {
            const int yi = i / npx, xi = i - yi * npx;
            // weights of the four elements around the node, (ex, ey) -> element (xi - 1 + ex, yi - 1 + ey); absent: flagged
            double w[2][2]; bool in[2][2];
#pragma unroll
            for (int ex = 0; ex < 2; ++ex)
#pragma unroll
                for (int ey = 0; ey < 2; ++ey) {
                    const int gx = xi - 1 + ex, gy = yi - 1 + ey;
                    in[ex][ey] = gx >= 0 && gx < nx && gy >= 0 && gy < ny;
                    w[ex][ey] = (in[ex][ey] && xpow) ? __ldg(&xpow[gx * ny + gy]) : 1.0;
                }
            const uint32_t fr = fixed ? __ldg(&fixed[i]) : 0u;
            const int cx = 1 + (xi > 0) + (xi < nx), cy = 1 + (yi > 0) + (yi < ny);   // neighbours per direction
            const int len = cx * cy;
            double* out = s_out + (size_t)ND * ND * (__ldg(&adj_ptr[i]) - pb0);
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                if (yi + dj < 0 || yi + dj > ny) continue;
#pragma unroll
                for (int di = -1; di <= 1; ++di) {
                    if (xi + di < 0 || xi + di > nx) continue;
                    if (T > 1 && ((dj + 1) * 3 + di + 1) % T != sub) continue;
                    // position of this neighbour in the node's sorted adjacency
                    const int k = ((dj + 1) - (yi == 0)) * cx + (di + 1) - (xi == 0);
                    double acc[ND][ND];
#pragma unroll
                    for (int a = 0; a < ND; ++a)
#pragma unroll
                        for (int b = 0; b < ND; ++b) acc[a][b] = 0.0;
                    // shared elements: ex in {di, di + 1} & {0, 1}, ey likewise; ascending element id = ex-major
#pragma unroll
                    for (int ex = 0; ex < 2; ++ex) {
                        if (ex != di && ex != di + 1) continue;
#pragma unroll
                        for (int ey = 0; ey < 2; ++ey) {
                            if (ey != dj && ey != dj + 1) continue;
                            if (!in[ex][ey]) continue;
                            // local indices in that element (Q4 order (0,0) (1,0) (1,1) (0,1)): node i sits at (1-ex, 1-ey)
                            const int ax = 1 - ex, ay = 1 - ey, bx = ax + di, by = ay + dj;
                            const int la = ay == 0 ? ax : 3 - ax, lb = by == 0 ? bx : 3 - bx;
                            const double* eb = s_mat + la * ND * S + lb * ND;
                            const double ww = w[ex][ey];
#pragma unroll
                            for (int a = 0; a < ND; ++a)
#pragma unroll
                                for (int b = 0; b < ND; ++b)
                                    acc[a][b] += COAT ? eb[a * S + b] + ww * eb[S2 + a * S + b] : ww * eb[a * S + b];
                        }
                    }
                    const uint32_t fc = fixed ? __ldg(&fixed[i + dj * npx + di]) : 0u;
                    if (fr | fc) {  // FEM.py:902-930: fixed rows and columns zeroed, unit diagonal
                        const bool diag = di == 0 && dj == 0;
#pragma unroll
                        for (int a = 0; a < ND; ++a)
#pragma unroll
                            for (int b = 0; b < ND; ++b)
                                if (((fr >> a) | (fc >> b)) & 1u) acc[a][b] = (diag && a == b) ? 1.0 : 0.0;
                    }
#pragma unroll
                    for (int a = 0; a < ND; ++a) {
                        double2* o2 = reinterpret_cast<double2*>(out + (size_t)a * len * ND + (size_t)k * ND);
#pragma unroll
                        for (int b = 0; b < ND; b += 2) o2[b >> 1] = make_double2(acc[a][b], acc[a][b + 1]);
                    }
                }
            }
        }
        __syncthreads();
        // the slice is contiguous in vals: coalesced streaming stores
        const int n2 = (pb1 - pb0) * ND * ND / 2;
        double2* g2 = reinterpret_cast<double2*>(vals + (size_t)ND * ND * pb0);
        const double2* s2 = reinterpret_cast<const double2*>(s_out);
        for (int t = threadIdx.x; t < n2; t += blockDim.x) __stcs(&g2[t], s2[t]);
    }
}

template <int ND>
int launch_grid(femol_plan_t p, const double* e_base, const double* e_coat, const double* xpow, const uint8_t* fixed,
                double* vals, cudaStream_t st) {
    constexpr int NPP = 128 / (ND == 2 ? 1 : 4);
    const size_t smem = ((size_t)(e_coat ? 2 : 1) * 16 * ND * ND + (size_t)NPP * 9 * ND * ND) * sizeof(double);
    int64_t grid = (p->n_nodes + NPP - 1) / NPP;
    const int64_t cap = (int64_t)p->sm_count * 16;
    if (grid > cap) grid = cap;
    if (e_coat) {
        auto kern = assemble_grid_kernel<ND, true>;
        FEMOL_FUNC_ATTR_PER_DEVICE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
        kern<<<(int)grid, 128, smem, st>>>((int)p->grid_nx, (int)p->grid_ny, p->adj_ptr, e_base, e_coat, xpow, fixed, vals);
    } else {
        auto kern = assemble_grid_kernel<ND, false>;
        FEMOL_FUNC_ATTR_PER_DEVICE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
        kern<<<(int)grid, 128, smem, st>>>((int)p->grid_nx, (int)p->grid_ny, p->adj_ptr, e_base, nullptr, xpow, fixed, vals);
    }
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

}  // namespace

extern "C" int femol_assemble(femol_plan_t p, const double* e_base, int base_per_elem, const double* e_coat,
                              int coat_per_elem, const double* x, double expo, const uint8_t* fixed,
                              double* xpow_ws, double* vals, void* stream) {
    FEMOL_REQUIRE(p && e_base && vals, "femol_assemble: null argument");
    FEMOL_REQUIRE(!x || xpow_ws, "femol_assemble: x given without xpow workspace");
    cudaStream_t st = (cudaStream_t)stream;
    const double* xpow = nullptr;
    if (x) {
        density_power<<<femol_grid_for(p->n_elem, 256, p->sm_count), 256, 0, st>>>(p->n_elem, x, expo, xpow_ws);
        FEMOL_LAUNCH_CHECK();
        xpow = xpow_ws;
    }
    if (p->grid_nx > 0 && p->nn == 4 && !base_per_elem && !(e_coat && coat_per_elem)) {   // structured grid, shared Ke
        if (p->ndof == 2) return launch_grid<2>(p, e_base, e_coat, xpow, fixed, vals, st);
        if (p->ndof == 6) return launch_grid<6>(p, e_base, e_coat, xpow, fixed, vals, st);
    }
#define ASM_CASE(ND, NN) \
    if (p->ndof == ND && p->nn == NN) return launch_assemble<ND, NN>(p, e_base, base_per_elem, e_coat, coat_per_elem, xpow, fixed, vals, st);
    ASM_CASE(2, 3) ASM_CASE(2, 4) ASM_CASE(2, 6) ASM_CASE(2, 8)
    ASM_CASE(6, 3) ASM_CASE(6, 4) ASM_CASE(6, 6) ASM_CASE(6, 8)
#undef ASM_CASE
    femol_set_error("femol_assemble: unsupported (ndof=%d, nn=%d)", p->ndof, p->nn);
    return FEMOL_E_ARG;
}
