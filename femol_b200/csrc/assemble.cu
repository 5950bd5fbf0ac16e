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

template <int ND, int NN>
int launch_assemble(femol_plan_t p, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                    const double* xpow, const uint8_t* fixed, double* vals, cudaStream_t st) {
    constexpr int S2 = NN * ND * NN * ND;
    size_t smem = 0;
    if (!base_pe) smem += S2 * sizeof(double);
    if (e_coat && !coat_pe) smem += S2 * sizeof(double);
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
#define ASM_CASE(ND, NN) \
    if (p->ndof == ND && p->nn == NN) return launch_assemble<ND, NN>(p, e_base, base_per_elem, e_coat, coat_per_elem, xpow, fixed, vals, st);
    ASM_CASE(2, 3) ASM_CASE(2, 4) ASM_CASE(2, 6) ASM_CASE(2, 8)
    ASM_CASE(6, 3) ASM_CASE(6, 4) ASM_CASE(6, 6) ASM_CASE(6, 8)
#undef ASM_CASE
    femol_set_error("femol_assemble: unsupported (ndof=%d, nn=%d)", p->ndof, p->nn);
    return FEMOL_E_ARG;
}
