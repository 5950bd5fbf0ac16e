// Geometric multigrid preconditioner for structured Q4 meshes (plane stress and Reissner-Mindlin plate).
//
// Why: Jacobi / block-Jacobi PCG needs ~10^4 iterations on the 4 M-dof SIMP cantilever (profiles/r01): the
// iteration count, not the kernel bandwidth, decides SIMP iterations/s (SURVEY §7 "hard parts").  A V-cycle
// costs ~7 SpMV-equivalents and brings the count to O(10^1-10^2).
//
// Hierarchy: level l+1 has ceil(n/2) elements per direction (any n: a trailing coarse element may cover a
// single fine element); nodes are numbered x-fastest and elements e = i*ny + j on every level (the reference's
// meshzoo ordering, mesh.py:852-860), so every level reuses femol_plan_create / femol_assemble / the node-block
// SpMV unchanged.  Coarse operators are Galerkin, formed ELEMENT-WISE:
//      K_c = P^T K P = sum_E sum_{children c of E} P_c^T (x_c^p Ke) P_c ,
// i.e. one S x S matrix per coarse element (bilinear interpolation weights 0, 1/2, 1), then an ordinary
// per-element assembly.  Dirichlet dofs: fine fixed dofs neither interpolate nor restrict (their rows/columns
// are masked in the child matrices); a coarse node that coincides with a fixed fine node is fixed.
// Smoother: damped block-Jacobi (ndof x ndof node blocks), nu pre + nu post sweeps -> symmetric V-cycle, valid
// as a CG preconditioner.  Coarsest level: dense inverse computed on the device by one CTA.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <cooperative_groups.h>
#include "common.cuh"
#include "spmv.cuh"
#include "dist.cuh"

extern "C" int femol_precond_setup(femol_plan_t p, const double* vals, int kind, double* dinv, void* stream);

namespace cg = cooperative_groups;

namespace {

constexpr int GMG_MAX_LEVELS = 20;
constexpr int GMG_COARSE_MAX_DOF = 100;

struct Level {
    int64_t nx, ny, n_nodes, n_elem, n;
    femol_plan* plan;   // level 0: borrowed
    uint8_t* fixed;     // node masks (level 0: borrowed, may be NULL)
    double* ke;         // [n_elem][S][S] (levels >= 1)
    double* vals;       // CSR values (levels >= 1; level 0: set per setup)
    float* vals32;      // the same values in single precision (regular levels >= 1 of the mixed-precision V-cycle) or NULL
    double* dinv;       // [n_nodes][ND][ND]
    double *r, *e, *t;  // work vectors (level 0: r, e supplied by the caller through vcycle)
    double *r2, *e2;    // second coarse visit of the F / W cycle (levels >= 1)
    int32_t* lat_node;  // level 0 of a lattice hierarchy: lattice index (x-fastest) -> node id; NULL = identity
};

// Strip-distributed hierarchy (femol_gmg_set_dist): levels 0 .. n_dist-1 live on the rank's strip of lattice rows and
// exchange halo rows with the two neighbours; level n_dist is gathered and handled, together with everything below it,
// by `coarse`, a hierarchy every rank holds for the whole (small) level-n_dist grid.
struct GmgDist {
    femol_comm* comm;
    int n_dist;
    femol_strip strip[GMG_MAX_LEVELS];
    femol_gmg* coarse;            // borrowed
    femol_plan* coarse_plan;      // borrowed: plan of the global level-n_dist Q4 grid (level 0 of `coarse`)
    const uint8_t* coarse_fixed;  // borrowed: node mask of that grid
    int64_t g_nx, g_ny;           // elements of the global level-n_dist grid
    double *gke, *gvals;          // its element matrices (gathered) and assembled CSR values
    double *gr, *ge;              // gathered residual / correction
};

}  // namespace

struct femol_gmg {
    int nd, n_levels, nu;
    double omega;
    Level lv[GMG_MAX_LEVELS];
    double* coarse_inv;  // dense inverse of the coarsest operator [nc][nc]
    double* child_g;     // [2][4][S2] Galerkin images P_c^T Ke P_c of the shared element matrix (base | coat)
    int64_t nc;
    void* d_desc;        // device array of LevelDesc (all levels) for the fused coarse V-cycle
    int fused_from;      // first level handled by the single-CTA kernel (n_levels = none)
    int cluster_from;    // nu = 1: first level handled by the cluster kernel (n_levels = none), ...
    int solo_from;       // ... first level CTA 0 handles alone, ...
    int cluster_size;    // ... CTAs per cluster (0: unavailable)
    int hint_iters;      // CG iterations of the previous solve with this hierarchy (sizes the first batch)
    bool f32_levels;     // coarse-level operators stored in single precision for the smoother products (FEMOL_GMG_F32)
    int cycle;           // 0 V-cycle, 1 F-cycle (NOT symmetric: experiments only), 2 W-cycle (nu = 1 path; FEMOL_GMG_CYCLE = V | F | W)
    int cycle_depth;     // coarse levels 1 .. cycle_depth get the second visit, deeper ones a plain V-cycle
    femol_gmg* mod;      // hierarchy of the sign-modulated operator S K S (femol_gmg_attach_modulated) or NULL
    double *mod_r, *mod_z;  // S r and the correction of that hierarchy
    // lattice hierarchy (femol_gmg_create_lattice): level 0 is ANY mesh whose nodes form a complete lattice
    // (T3 / T6 / Q4 in any node numbering); level 1 is the Q4 grid over every second lattice line
    bool lattice;
    int32_t* node_lat;   // [n_nodes] lattice index of every level-0 node
    int32_t* child_ptr;  // [n_elem(level 1) + 1] level-0 elements inside each level-1 element ...
    int32_t* child_idx;  // ... ascending (deterministic summation order)
    bool capped;         // femol_gmg_truncate: the last level is a gather level without coarsest-level solve
    GmgDist* dist;       // strip-distributed hierarchy or NULL
    femol_gmg* parent;   // workspace view (femol_gmg_clone_workspace): operators borrowed from `parent`, own work vectors
};

namespace {

// ---- structured-grid helpers ------------------------------------------------------------------------------
__host__ __device__ inline int64_t coarse_of(int64_t n) { return (n + 1) / 2; }

// 1-D interpolation: fine node i (0..n) -> coarse nodes I0 (weight w0) and I1 (weight w1, may be 0)
__device__ __forceinline__ void cmap1d(int64_t i, int64_t n, int64_t& I0, double& w0, int64_t& I1, double& w1) {
    if ((i & 1) == 0) { I0 = i >> 1; w0 = 1.0; I1 = I0; w1 = 0.0; }
    else if (i == n) { I0 = (i + 1) >> 1; w0 = 1.0; I1 = I0; w1 = 0.0; }   // odd n: last fine node is a coarse node
    else { I0 = i >> 1; w0 = 0.5; I1 = I0 + 1; w1 = 0.5; }
}

__global__ void make_conn(int64_t nx, int64_t ny, int32_t* __restrict__ conn) {
    const int64_t ne = nx * ny;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / ny, j = e - i * ny;
        const int64_t a = i + (nx + 1) * j;
        conn[4 * e + 0] = (int32_t)a; conn[4 * e + 1] = (int32_t)(a + 1);
        conn[4 * e + 2] = (int32_t)(a + nx + 2); conn[4 * e + 3] = (int32_t)(a + nx + 1);
    }
}

__global__ void coarsen_mask(int64_t nxf, int64_t nyf, const uint8_t* __restrict__ ff, int64_t nxc, int64_t nyc,
                             uint8_t* __restrict__ fc, const int32_t* __restrict__ lat_node) {
    const int64_t nn = (nxc + 1) * (nyc + 1);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nn; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t J = t / (nxc + 1), I = t - J * (nxc + 1);
        const int64_t i = min(2 * I, nxf), j = min(2 * J, nyf);
        const int64_t lat = j * (nxf + 1) + i;
        fc[t] = ff ? ff[lat_node ? lat_node[lat] : lat] : 0;
    }
}

// ---- element-wise Galerkin coarsening ----------------------------------------------------------------------
// G[c] = P_c^T Ke P_c for the four children c = 2 ca + cb of a full parent element (no Dirichlet nodes), base and
// coating matrix: the level-0 fast path of coarsen_elements combines them with the element densities.
template <int ND>
__global__ void child_images(const double* __restrict__ e_base, const double* __restrict__ e_coat, double* __restrict__ G) {
    constexpr int S = 4 * ND, S2 = S * S;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 8 * S2; t += gridDim.x * blockDim.x) {
        const int mat = t / (4 * S2), c = (t / S2) & 3, idx = t % S2;
        const double* ke = mat ? e_coat : e_base;
        if (!ke) { G[t] = 0.0; continue; }
        const int ca = c >> 1, cb = c & 1;
        const double s0 = 0.5 * ca, s1 = 0.5 * (ca + 1), t0 = 0.5 * cb, t1 = 0.5 * (cb + 1);
        const double ps[4] = {s0, s1, s1, s0}, pt[4] = {t0, t0, t1, t1};
        const int r = idx / S, cc = idx - r * S;
        const int p = r / ND, da = r - p * ND, q = cc / ND, db = cc - q * ND;
        double v = 0.0;
        for (int a = 0; a < 4; ++a) {
            const double wa = p == 0 ? (1 - ps[a]) * (1 - pt[a]) : p == 1 ? ps[a] * (1 - pt[a]) : p == 2 ? ps[a] * pt[a] : (1 - ps[a]) * pt[a];
            for (int b = 0; b < 4; ++b) {
                const double wb = q == 0 ? (1 - ps[b]) * (1 - pt[b]) : q == 1 ? ps[b] * (1 - pt[b]) : q == 2 ? ps[b] * pt[b] : (1 - ps[b]) * pt[b];
                v += wa * wb * ke[(a * ND + da) * S + b * ND + db];
            }
        }
        G[t] = v;
    }
}

// One warp per coarse element.  LEVEL0: child matrices are x_c^p * Ke (+ coat) from the shared element matrix;
// otherwise they are the previous level's per-element matrices.
template <int ND, bool LEVEL0>
__global__ void __launch_bounds__(128)
coarsen_elements(int64_t nxf, int64_t nyf, int64_t nxc, int64_t nyc, const double* __restrict__ child_ke,
                 const double* __restrict__ e_base, const double* __restrict__ e_coat, const double* __restrict__ xpow,
                 const uint8_t* __restrict__ fixed_f, const double* __restrict__ child_g, double* __restrict__ out) {
    constexpr int S = 4 * ND, S2 = S * S;
    __shared__ double s_child[4][S2];  // one warp's child matrix at a time
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t nec = nxc * nyc;
    const int64_t warp0 = (int64_t)blockIdx.x * 4 + wid, nwarps = (int64_t)gridDim.x * 4;
    constexpr int NE = (S2 + 31) / 32;
    for (int64_t E = warp0; E < nec; E += nwarps) {
        const int64_t I = E / nyc, J = E - I * nyc;
        const bool fullx = 2 * I + 1 < nxf, fully = 2 * J + 1 < nyf;
        double acc[NE];
#pragma unroll
        for (int m = 0; m < NE; ++m) acc[m] = 0.0;
        if (LEVEL0 && fullx && fully) {
            // Fast path (all but the elements touching a Dirichlet node): the four children share one element
            // matrix, so the parent is  sum_c x_c^p G_c  with the four precomputed images G_c = P_c^T Ke P_c.
            uint32_t fm = 0u;
            if (fixed_f && lane < 9) fm = fixed_f[(2 * J + lane / 3) * (nxf + 1) + 2 * I + lane % 3];
            if (!__any_sync(0xffffffffu, fm != 0u)) {
                double w[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) w[c] = xpow ? __ldg(&xpow[(2 * I + (c >> 1)) * nyf + 2 * J + (c & 1)]) : 1.0;
                double* o = out + E * S2;
#pragma unroll
                for (int m = 0; m < NE; ++m) {
                    const int idx = lane + 32 * m;
                    if (idx < S2) {
                        double v = 0.0;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            v += e_coat ? __ldg(&child_g[c * S2 + idx]) + w[c] * __ldg(&child_g[(4 + c) * S2 + idx])
                                        : w[c] * __ldg(&child_g[c * S2 + idx]);
                        o[idx] = v;
                    }
                }
                continue;
            }
        }
        for (int ca = 0; ca <= (fullx ? 1 : 0); ++ca)
            for (int cb = 0; cb <= (fully ? 1 : 0); ++cb) {
                const int64_t ci = 2 * I + ca, cj = 2 * J + cb;
                const int64_t c = ci * nyf + cj;  // fine element id
                // fine node ids of the child (Q4 order) for the fixed mask
                const int64_t n00 = ci + (nxf + 1) * cj;
                const int64_t cn[4] = {n00, n00 + 1, n00 + nxf + 2, n00 + nxf + 1};
                uint32_t fm[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) fm[q] = fixed_f ? fixed_f[cn[q]] : 0u;
                __syncwarp();
                const double w = (LEVEL0 && xpow) ? xpow[c] : 1.0;
                for (int idx = lane; idx < S2; idx += 32) {
                    const int r = idx / S, cc = idx - r * S;
                    double v;
                    if (LEVEL0) v = e_coat ? e_base[idx] + w * e_coat[idx] : w * e_base[idx];
                    else v = child_ke[c * S2 + idx];
                    if (((fm[r / ND] >> (r % ND)) | (fm[cc / ND] >> (cc % ND))) & 1u) v = 0.0;
                    s_child[wid][idx] = v;
                }
                __syncwarp();
                // parametric position of the child's local nodes inside the parent: (s, t) in {0, 1/2, 1}
                const double s0 = fullx ? 0.5 * ca : 0.0, s1 = fullx ? 0.5 * (ca + 1) : 1.0;
                const double t0 = fully ? 0.5 * cb : 0.0, t1 = fully ? 0.5 * (cb + 1) : 1.0;
                const double ps[4] = {s0, s1, s1, s0}, pt[4] = {t0, t0, t1, t1};
                double W[4][4];  // W[child node][parent node]
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    W[q][0] = (1 - ps[q]) * (1 - pt[q]); W[q][1] = ps[q] * (1 - pt[q]);
                    W[q][2] = ps[q] * pt[q]; W[q][3] = (1 - ps[q]) * pt[q];
                }
#pragma unroll
                for (int m = 0; m < NE; ++m) {
                    const int idx = lane + 32 * m;
                    if (idx < S2) {
                        const int r = idx / S, cc = idx - r * S;
                        const int p = r / ND, da = r - p * ND, q = cc / ND, db = cc - q * ND;
                        double v = 0.0;
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const double wa = W[a][p];
                            if (wa != 0.0) {
#pragma unroll
                                for (int b = 0; b < 4; ++b) v += wa * W[b][q] * s_child[wid][(a * ND + da) * S + b * ND + db];
                            }
                        }
                        acc[m] += v;
                    }
                }
            }
        double* o = out + E * S2;
#pragma unroll
        for (int m = 0; m < NE; ++m) {
            const int idx = lane + 32 * m;
            if (idx < S2) o[idx] = acc[m];
        }
    }
}

// Galerkin coarsening of per-element matrices, register / shuffle version (levels >= 1 -> next).  The parent is
// sum_c P_c^T K_c P_c with P_c = W_c (x) I_ND (W_c: 4 x 4 bilinear weights of the child's nodes in the parent), so
// every dof pair (da, db) is an independent 4 x 4 triple product W_c^T K_c^{da,db} W_c.  A warp owns a coarse element;
// lane = (dd slot (4), r (4), q half (2)): stage 1 forms T[a = r][q] = sum_b K_c[(r,da),(b,db)] W_c[b][q] from four
// global loads, stage 2 forms out[p = r][q] = sum_a W_c[a][r] T[a][q] with four shuffles per q.  No shared memory,
// ~10 x fewer instructions than the generic kernel above (which stays for level 0, where child matrices are
// synthesised from the shared element matrix).
__device__ __forceinline__ double bilinear(int node, double s, double t) {
    return node == 0 ? (1 - s) * (1 - t) : node == 1 ? s * (1 - t) : node == 2 ? s * t : (1 - s) * t;
}

template <int ND>
__global__ void __launch_bounds__(128)
coarsen_elements_fast(int nxf, int nyf, int nxc, int nyc, const double* __restrict__ child_ke,
                      const uint8_t* __restrict__ fixed_f, double* __restrict__ out) {
    constexpr int S = 4 * ND, S2 = S * S, NDD = ND * ND;
    const int lane = threadIdx.x & 31;
    const int ddl = lane >> 3, r = (lane >> 1) & 3, qh = lane & 1;
    const int nec = nxc * nyc;
    const int warp0 = blockIdx.x * 4 + (threadIdx.x >> 5), nwarps = gridDim.x * 4;
    for (int E = warp0; E < nec; E += nwarps) {
        const int I = E / nyc, J = E - I * nyc;
        const bool fullx = 2 * I + 1 < nxf, fully = 2 * J + 1 < nyf;
        for (int dd0 = 0; dd0 < NDD; dd0 += 4) {
            const int dd = dd0 + ddl;            // NDD is a multiple of 4 (4 or 36): every lane is active
            const int da = dd / ND, db = dd - da * ND;
            double acc0 = 0.0, acc1 = 0.0;
            for (int ca = 0; ca <= (fullx ? 1 : 0); ++ca)
                for (int cb = 0; cb <= (fully ? 1 : 0); ++cb) {
                    const int ci = 2 * I + ca, cj = 2 * J + cb;
                    const size_t c = (size_t)ci * nyf + cj;                  // fine element id
                    const int n00 = ci + (nxf + 1) * cj;
                    const int cn[4] = {n00, n00 + 1, n00 + nxf + 2, n00 + nxf + 1};
                    // parametric position of the child's local nodes inside the parent: (s, t) in {0, 1/2, 1}
                    const double s0 = fullx ? 0.5 * ca : 0.0, s1 = fullx ? 0.5 * (ca + 1) : 1.0;
                    const double t0 = fully ? 0.5 * cb : 0.0, t1 = fully ? 0.5 * (cb + 1) : 1.0;
                    const double ps[4] = {s0, s1, s1, s0}, pt[4] = {t0, t0, t1, t1};
                    // stage 1: T[a = r][q], q = 2 qh, 2 qh + 1  (rows / columns of fixed fine dofs are zero)
                    const uint32_t fr = fixed_f ? fixed_f[cn[r]] : 0u;
                    double T0 = 0.0, T1 = 0.0;
                    const double* krow = child_ke + c * S2 + (size_t)(r * ND + da) * S + db;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const uint32_t fb = fixed_f ? fixed_f[cn[b]] : 0u;
                        double k = __ldg(krow + b * ND);
                        if (((fr >> da) | (fb >> db)) & 1u) k = 0.0;
                        T0 = fma(k, bilinear(2 * qh, ps[b], pt[b]), T0);
                        T1 = fma(k, bilinear(2 * qh + 1, ps[b], pt[b]), T1);
                    }
                    // stage 2: out[p = r][q] += sum_a W[a][r] T[a][q]   (T[a][q] lives in lane (ddl, a, qh))
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int src = (ddl << 3) | (a << 1) | qh;
                        const double ta0 = __shfl_sync(0xffffffffu, T0, src), ta1 = __shfl_sync(0xffffffffu, T1, src);
                        const double wa = bilinear(r, ps[a], pt[a]);
                        acc0 = fma(wa, ta0, acc0);
                        acc1 = fma(wa, ta1, acc1);
                    }
                }
            double* o = out + (size_t)E * S2 + (size_t)(r * ND + da) * S + db;
            o[(2 * qh) * ND] = acc0;
            o[(2 * qh + 1) * ND] = acc1;
        }
    }
}

// Lattice hierarchy, level 0 -> 1: the children of a coarse Q4 element are the level-0 elements (NN nodes each: T3,
// Q4, T6) that lie inside it; P_c holds the bilinear weights of the child's nodes, taken from their lattice
// coordinates.  One warp per coarse element; child matrix (density-scaled, Dirichlet rows / columns zeroed) and weights
// staged in shared memory; children are visited in ascending element order (deterministic).
template <int ND, int NN>
__global__ void __launch_bounds__(128)
coarsen_lattice(int nxf, int nyf, int nxc, int nyc, const int32_t* __restrict__ conn, const int32_t* __restrict__ node_lat,
                const int32_t* __restrict__ child_ptr, const int32_t* __restrict__ child_idx,
                const double* __restrict__ e_base, int base_pe, const double* __restrict__ e_coat, int coat_pe,
                const double* __restrict__ xpow, const uint8_t* __restrict__ fixed_f, double* __restrict__ out) {
    constexpr int S = 4 * ND, S2 = S * S, SC = NN * ND, SC2 = SC * SC;
    constexpr int NE = (S2 + 31) / 32;
    __shared__ double s_child[4][SC2];
    __shared__ double s_w[4][NN][4];
    __shared__ uint32_t s_fm[4][NN];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nec = nxc * nyc;
    for (int E = blockIdx.x * 4 + wid; E < nec; E += gridDim.x * 4) {
        const int I = E / nyc, J = E - I * nyc;
        const double sx = (2 * I + 1 < nxf) ? 0.5 : 1.0, sy = (2 * J + 1 < nyf) ? 0.5 : 1.0;   // trailing single-cell parent
        double acc[NE];
#pragma unroll
        for (int m = 0; m < NE; ++m) acc[m] = 0.0;
        for (int k = child_ptr[E]; k < child_ptr[E + 1]; ++k) {
            const int e = child_idx[k];
            __syncwarp();
            if (lane < NN) {
                const int node = conn[(size_t)e * NN + lane];
                const int lat = node_lat[node];
                const int j = lat / (nxf + 1), i = lat - j * (nxf + 1);
                const double ps = (i - 2 * I) * sx, pt = (j - 2 * J) * sy;
                s_w[wid][lane][0] = (1 - ps) * (1 - pt); s_w[wid][lane][1] = ps * (1 - pt);
                s_w[wid][lane][2] = ps * pt; s_w[wid][lane][3] = (1 - ps) * pt;
                s_fm[wid][lane] = fixed_f ? fixed_f[node] : 0u;
            }
            __syncwarp();
            const double w = xpow ? xpow[e] : 1.0;
            const double* kb = e_base + (base_pe ? (size_t)e * SC2 : 0);
            const double* kc = e_coat ? e_coat + (coat_pe ? (size_t)e * SC2 : 0) : nullptr;
            for (int idx = lane; idx < SC2; idx += 32) {
                const int r = idx / SC, c = idx - r * SC;
                double v = kc ? kb[idx] + w * kc[idx] : w * kb[idx];
                if (((s_fm[wid][r / ND] >> (r % ND)) | (s_fm[wid][c / ND] >> (c % ND))) & 1u) v = 0.0;
                s_child[wid][idx] = v;
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < NE; ++m) {
                const int idx = lane + 32 * m;
                if (idx < S2) {
                    const int r = idx / S, cc = idx - r * S;
                    const int p = r / ND, da = r - p * ND, q = cc / ND, db = cc - q * ND;
                    double v = 0.0;
                    for (int a = 0; a < NN; ++a) {
                        const double wa = s_w[wid][a][p];
                        if (wa != 0.0) {
                            double t = 0.0;
#pragma unroll
                            for (int b = 0; b < NN; ++b) t += s_w[wid][b][q] * s_child[wid][(a * ND + da) * SC + b * ND + db];
                            v += wa * t;
                        }
                    }
                    acc[m] += v;
                }
            }
        }
        double* o = out + (size_t)E * S2;
#pragma unroll
        for (int m = 0; m < NE; ++m) {
            const int idx = lane + 32 * m;
            if (idx < S2) o[idx] = acc[m];
        }
    }
}

// ---- transfer operators ---------------------------------------------------------------------------------------
// 1-D weight of fine node c + dc in the restriction onto the coarse node centred at fine index c (n fine cells):
// the transpose of cmap1d.  c is even except for the last node of an odd grid, which has no half-weight neighbours.
__device__ __forceinline__ double rweight1d(int c, int dc, int n) {
    if (dc == 0) return 1.0;
    const int i = c + dc;
    if (i < 0 || i > n || (c & 1)) return 0.0;
    if (dc == 1 && i == n) return 0.0;      // i is the (odd) last node: it is a coarse node of its own
    return 0.5;
}

// r_c = P^T r_f   (fine fixed dofs do not restrict; coarse fixed dofs receive 0); with dinv_c also the coarse
// level's first sweep e_c = omega Dinv_c r_c.  All nine fine loads of a thread are issued before they are used
// (the mask no longer gates the value load: the first version was a chain of dependent L2 round trips).
template <int ND>
__global__ void __launch_bounds__(256)
restrict_kernel(int64_t nxf64, int64_t nyf64, int64_t nxc64, int64_t nyc64, const uint8_t* __restrict__ ff,
                const uint8_t* __restrict__ fc, const double* __restrict__ rf, double* __restrict__ rc,
                const double* __restrict__ dinv_c, double omega, double* __restrict__ ec,
                const int32_t* __restrict__ lat_node) {
    FEMOL_PDL_SYNC();
    const int nxf = (int)nxf64, nyf = (int)nyf64, nxc = (int)nxc64, nyc = (int)nyc64;
    const int nn = (nxc + 1) * (nyc + 1);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += gridDim.x * blockDim.x) {
        const int J = t / (nxc + 1), I = t - J * (nxc + 1);
        const int ic = min(2 * I, nxf), jc = min(2 * J, nyf);
        double acc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = 0.0;
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            const double wy = rweight1d(jc, dj, nyf);
#pragma unroll
            for (int di = -1; di <= 1; ++di) {
                const double w = wy * rweight1d(ic, di, nxf);
                if (w == 0.0) continue;
                const int lat = (jc + dj) * (nxf + 1) + ic + di;
                const int nf = lat_node ? __ldg(&lat_node[lat]) : lat;    // lattice hierarchy: permuted fine level
                const uint32_t fm = ff ? __ldg(&ff[nf]) : 0u;
#pragma unroll
                for (int d = 0; d < ND; d += 2) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(rf + (size_t)nf * ND + d));
                    acc[d] += ((fm >> d) & 1u) ? 0.0 : w * v.x;
                    acc[d + 1] += ((fm >> (d + 1)) & 1u) ? 0.0 : w * v.y;
                }
            }
        }
        const uint32_t cm = fc ? fc[t] : 0u;
#pragma unroll
        for (int d = 0; d < ND; d += 2) {
            if ((cm >> d) & 1u) acc[d] = 0.0;
            if ((cm >> (d + 1)) & 1u) acc[d + 1] = 0.0;
            *reinterpret_cast<double2*>(rc + (size_t)t * ND + d) = make_double2(acc[d], acc[d + 1]);
        }
        if (dinv_c) {   // the coarse level's first smoothing sweep from a zero guess: e_c = omega Dinv r_c
            const double* B = dinv_c + (size_t)t * ND * ND;
#pragma unroll
            for (int a = 0; a < ND; a += 2) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int b = 0; b < ND; b += 2) {
                    const double2 b0 = __ldg(reinterpret_cast<const double2*>(B + a * ND + b));
                    const double2 b1 = __ldg(reinterpret_cast<const double2*>(B + (a + 1) * ND + b));
                    s0 += b0.x * acc[b] + b0.y * acc[b + 1];
                    s1 += b1.x * acc[b] + b1.y * acc[b + 1];
                }
                *reinterpret_cast<double2*>(ec + (size_t)t * ND + a) = make_double2(omega * s0, omega * s1);
            }
        }
    }
}

// out = e_f + P e_c   (fixed fine dofs copied; out may alias e_f)
template <int ND>
__global__ void __launch_bounds__(256)
prolong_add_kernel(int64_t nxf64, int64_t nyf64, int64_t nxc64, const uint8_t* __restrict__ ff,
                   const double* __restrict__ ec, const double* ef, double* out, const int32_t* __restrict__ lat_node) {
    FEMOL_PDL_SYNC();
    const int nxf = (int)nxf64, nyf = (int)nyf64, nxc = (int)nxc64;
    const int nn = (nxf + 1) * (nyf + 1);
    for (int lat = blockIdx.x * blockDim.x + threadIdx.x; lat < nn; lat += gridDim.x * blockDim.x) {
        const int j = lat / (nxf + 1), i = lat - j * (nxf + 1);
        const int t = lat_node ? __ldg(&lat_node[lat]) : lat;             // node id (lattice hierarchy: permuted)
        int64_t I0, I1, J0, J1; double u0, u1, v0, v1;
        cmap1d(i, nxf, I0, u0, I1, u1);      // I1 = I0 (weight 0) when the node coincides with a coarse node
        cmap1d(j, nyf, J0, v0, J1, v1);
        const uint32_t fm = ff ? __ldg(&ff[t]) : 0u;
        const int c00 = (int)J0 * (nxc + 1) + (int)I0, c10 = (int)J0 * (nxc + 1) + (int)I1;
        const int c01 = (int)J1 * (nxc + 1) + (int)I0, c11 = (int)J1 * (nxc + 1) + (int)I1;
        const double w00 = u0 * v0, w10 = u1 * v0, w01 = u0 * v1, w11 = u1 * v1;
#pragma unroll
        for (int d = 0; d < ND; d += 2) {
            const double2 a00 = __ldg(reinterpret_cast<const double2*>(ec + (size_t)c00 * ND + d));
            const double2 a10 = __ldg(reinterpret_cast<const double2*>(ec + (size_t)c10 * ND + d));
            const double2 a01 = __ldg(reinterpret_cast<const double2*>(ec + (size_t)c01 * ND + d));
            const double2 a11 = __ldg(reinterpret_cast<const double2*>(ec + (size_t)c11 * ND + d));
            double2 o = *reinterpret_cast<const double2*>(ef + (size_t)t * ND + d);
            const double vx = w00 * a00.x + w10 * a10.x + w01 * a01.x + w11 * a11.x;
            const double vy = w00 * a00.y + w10 * a10.y + w01 * a01.y + w11 * a11.y;
            if (!((fm >> d) & 1u)) o.x += vx;
            if (!((fm >> (d + 1)) & 1u)) o.y += vy;
            *reinterpret_cast<double2*>(out + (size_t)t * ND + d) = o;
        }
    }
}

// out (+)= S in, S = (-1)^(i + j) per grid node (nodes x-fastest, npx per row)
__global__ void __launch_bounds__(256)
sign_modulate(int64_t n_nodes, int npx, int nd, const double* __restrict__ in, double* __restrict__ out, int accumulate) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_nodes; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t / npx), i = (int)(t - (int64_t)j * npx);
        const double s = ((i + j) & 1) ? -1.0 : 1.0;
        for (int d = 0; d < nd; ++d) {
            const double v = s * in[t * nd + d];
            out[t * nd + d] = accumulate ? out[t * nd + d] + v : v;
        }
    }
}

__global__ void __launch_bounds__(256) axpy_kernel(int64_t n, const double* __restrict__ a, double* __restrict__ b) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] += a[i];
}

// ---- smoother ------------------------------------------------------------------------------------------------------
// FIRST: e = omega Dinv r ; else e += omega Dinv (r - Ke)
template <int ND, bool FIRST>
__global__ void __launch_bounds__(256)
jacobi_kernel(int64_t n_nodes, double omega, const double* __restrict__ dinv, const double* __restrict__ r,
              const double* __restrict__ Ke, double* __restrict__ e) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        double d[ND];
#pragma unroll
        for (int a = 0; a < ND; ++a) d[a] = FIRST ? r[i * ND + a] : r[i * ND + a] - Ke[i * ND + a];
        const double* B = dinv + i * ND * ND;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) s += __ldg(&B[a * ND + b]) * d[b];
            if (FIRST) e[i * ND + a] = omega * s; else e[i * ND + a] += omega * s;
        }
    }
}

__global__ void __launch_bounds__(256)
residual_kernel(int64_t n, const double* __restrict__ r, const double* __restrict__ Ke, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = r[i] - Ke[i];
}

// ---- coarsest level: dense inverse by Gauss-Jordan in one CTA ----------------------------------------------------
__global__ void csr_to_dense(int64_t n, const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx, int nd,
                             const double* __restrict__ vals, double* __restrict__ A) {
    const int64_t n_nodes = n / nd;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = r / nd;
        const int da = (int)(r - i * nd);
        const int pb = adj_ptr[i], len = adj_ptr[i + 1] - pb;
        const double* a = vals + (int64_t)nd * nd * pb + (int64_t)da * len * nd;
        for (int k = 0; k < len; ++k)
            for (int db = 0; db < nd; ++db) A[r * n + (int64_t)adj_idx[pb + k] * nd + db] = a[k * nd + db];
    }
    (void)n_nodes;
}

// In-place Gauss-Jordan inversion of the SPD matrix A (n x n) without pivoting; a non-positive pivot (a coarse
// dof that lost all its stiffness) is replaced by an identity row/column.
__global__ void __launch_bounds__(1024)
dense_inverse(int n, double* __restrict__ Ag) {
    extern __shared__ double A[];   // the whole matrix (n <= GMG_COARSE_MAX_DOF): ~4 n barriers, no global round trips
    __shared__ double s_scale;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) A[t] = Ag[t];
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(A[i * n + i]));
        s_scale = mx;
    }
    __syncthreads();
    const double tiny = 1e-13 * s_scale;
    for (int k = 0; k < n; ++k) {
        const double piv = A[k * n + k];
        __syncthreads();              // everybody has read the pivot before row k is rewritten
        const bool dead = !(piv > tiny);
        if (dead) {  // decouple dof k
            for (int j = threadIdx.x; j < n; j += blockDim.x) { A[k * n + j] = 0.0; A[j * n + k] = 0.0; }
            __syncthreads();
            continue;
        }
        const double ip = 1.0 / piv;
        // row k scaled (pivot element becomes 1/piv)
        for (int j = threadIdx.x; j < n; j += blockDim.x) A[k * n + j] = (j == k) ? ip : A[k * n + j] * ip;
        __syncthreads();
        // eliminate column k from every other row
        for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
            const int i = t / n, j = t - i * n;
            if (i == k || j == k) continue;
            A[t] -= A[i * n + k] * A[k * n + j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (i != k) A[i * n + k] = -A[i * n + k] * ip;
        __syncthreads();
    }
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) Ag[t] = A[t];
}

__global__ void __launch_bounds__(256)
dense_apply(int n, const double* __restrict__ Ainv, const double* __restrict__ r, double* __restrict__ e) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        double s = 0.0;
        for (int j = lane; j < n; j += 32) s += Ainv[(int64_t)i * n + j] * r[j];
        s = warp_sum(s);
        if (lane == 0) e[i] = s;
    }
}


// ---- fused coarse V-cycle: all levels with few nodes in ONE single-CTA kernel -----------------------------------
// Below a few thousand nodes every kernel of a level is launch-latency bound (~3.5 us each, 11 per level even inside a
// CUDA graph: 22 % of a CG iteration on cfg 2).  One CTA of 1024 threads walks those levels down and up with
// __syncthreads() between the phases instead.
struct LevelDesc {
    int64_t nx, ny, n_nodes;
    const int32_t *adj_ptr, *adj_idx;
    const double *vals, *dinv;
    const uint8_t* fixed;
    double *r, *e, *t;
};

template <int ND>
__device__ void cta_spmv(const LevelDesc& L, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t n_rows = L.n_nodes * ND;
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const int64_t i = r / ND;
        const int da = (int)(r - i * ND);
        const int pb = L.adj_ptr[i], len = L.adj_ptr[i + 1] - pb;
        const double* a = L.vals + (int64_t)ND * ND * pb + (int64_t)da * len * ND;
        double acc = 0.0;
        for (int k = 0; k < len; ++k) {
            const double* xv = x + (int64_t)L.adj_idx[pb + k] * ND;
#pragma unroll
            for (int d = 0; d < ND; ++d) acc += a[k * ND + d] * xv[d];
        }
        y[r] = acc;
    }
}

template <int ND, bool FIRST>
__device__ void cta_jacobi(const LevelDesc& L, double omega, const double* __restrict__ r, const double* __restrict__ Ke,
                           double* __restrict__ e) {
    for (int64_t i = threadIdx.x; i < L.n_nodes; i += blockDim.x) {
        double d[ND];
#pragma unroll
        for (int a = 0; a < ND; ++a) d[a] = FIRST ? r[i * ND + a] : r[i * ND + a] - Ke[i * ND + a];
        const double* B = L.dinv + i * ND * ND;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) s += B[a * ND + b] * d[b];
            if (FIRST) e[i * ND + a] = omega * s; else e[i * ND + a] += omega * s;
        }
    }
}

template <int ND>
__device__ void cta_restrict(const LevelDesc& F, const LevelDesc& C, const double* __restrict__ rf, double* __restrict__ rc) {
    for (int64_t t = threadIdx.x; t < C.n_nodes; t += blockDim.x) {
        const int64_t J = t / (C.nx + 1), I = t - J * (C.nx + 1);
        const int64_t ic = min(2 * I, F.nx), jc = min(2 * J, F.ny);
        double acc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = 0.0;
        for (int dj = -1; dj <= 1; ++dj) {
            const int64_t j = jc + dj;
            if (j < 0 || j > F.ny) continue;
            int64_t J0, J1; double v0, v1;
            cmap1d(j, F.ny, J0, v0, J1, v1);
            const double wy = (J0 == J ? v0 : 0.0) + (J1 == J && v1 != 0.0 ? v1 : 0.0);
            if (wy == 0.0) continue;
            for (int di = -1; di <= 1; ++di) {
                const int64_t i = ic + di;
                if (i < 0 || i > F.nx) continue;
                int64_t I0, I1; double u0, u1;
                cmap1d(i, F.nx, I0, u0, I1, u1);
                const double wx = (I0 == I ? u0 : 0.0) + (I1 == I && u1 != 0.0 ? u1 : 0.0);
                if (wx == 0.0) continue;
                const int64_t nf = j * (F.nx + 1) + i;
                const uint32_t fm = F.fixed ? F.fixed[nf] : 0u;
                const double w = wx * wy;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    if (!((fm >> d) & 1u)) acc[d] += w * rf[nf * ND + d];
            }
        }
        const uint32_t cm = C.fixed ? C.fixed[t] : 0u;
#pragma unroll
        for (int d = 0; d < ND; ++d) rc[t * ND + d] = ((cm >> d) & 1u) ? 0.0 : acc[d];
    }
}

template <int ND>
__device__ void cta_prolong_add(const LevelDesc& F, const LevelDesc& C, const double* __restrict__ ec, double* __restrict__ ef) {
    for (int64_t t = threadIdx.x; t < F.n_nodes; t += blockDim.x) {
        const int64_t j = t / (F.nx + 1), i = t - j * (F.nx + 1);
        int64_t I0, I1, J0, J1; double u0, u1, v0, v1;
        cmap1d(i, F.nx, I0, u0, I1, u1);
        cmap1d(j, F.ny, J0, v0, J1, v1);
        const uint32_t fm = F.fixed ? F.fixed[t] : 0u;
        const int64_t c00 = J0 * (C.nx + 1) + I0, c10 = J0 * (C.nx + 1) + I1, c01 = J1 * (C.nx + 1) + I0, c11 = J1 * (C.nx + 1) + I1;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double v = 0.0;
            if (!((fm >> d) & 1u)) {
                v = u0 * v0 * ec[c00 * ND + d];
                if (u1 != 0.0) v += u1 * v0 * ec[c10 * ND + d];
                if (v1 != 0.0) { v += u0 * v1 * ec[c01 * ND + d]; if (u1 != 0.0) v += u1 * v1 * ec[c11 * ND + d]; }
            }
            ef[t * ND + d] += v;
        }
    }
}

// levels l0 .. last; r/e of level l0 are passed (they belong to the caller), deeper levels use their own
template <int ND>
__global__ void __launch_bounds__(1024)
coarse_vcycle_kernel(const LevelDesc* __restrict__ lv, int l0, int last, int nu, double omega,
                     const double* __restrict__ coarse_inv, const double* r0, double* e0) {
    // down
    for (int l = l0; l < last; ++l) {
        const LevelDesc L = lv[l];
        const double* r = l == l0 ? r0 : L.r;
        double* e = l == l0 ? e0 : L.e;
        cta_jacobi<ND, true>(L, omega, r, nullptr, e);
        __syncthreads();
        for (int s = 1; s < nu; ++s) {
            cta_spmv<ND>(L, e, L.t);
            __syncthreads();
            cta_jacobi<ND, false>(L, omega, r, L.t, e);
            __syncthreads();
        }
        cta_spmv<ND>(L, e, L.t);
        __syncthreads();
        for (int64_t i = threadIdx.x; i < L.n_nodes * ND; i += blockDim.x) L.t[i] = r[i] - L.t[i];
        __syncthreads();
        cta_restrict<ND>(L, lv[l + 1], L.t, lv[l + 1].r);
        __syncthreads();
    }
    {   // coarsest: dense inverse
        const LevelDesc L = lv[last];
        const int n = (int)(L.n_nodes * ND);
        const double* r = last == l0 ? r0 : L.r;
        double* e = last == l0 ? e0 : L.e;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int i = warp; i < n; i += nw) {
            double s = 0.0;
            for (int j = lane; j < n; j += 32) s += coarse_inv[(int64_t)i * n + j] * r[j];
            s = warp_sum(s);
            if (lane == 0) e[i] = s;
        }
        __syncthreads();
    }
    // up
    for (int l = last - 1; l >= l0; --l) {
        const LevelDesc L = lv[l];
        const double* r = l == l0 ? r0 : L.r;
        double* e = l == l0 ? e0 : L.e;
        cta_prolong_add<ND>(L, lv[l + 1], lv[l + 1].e, e);
        __syncthreads();
        for (int s = 0; s < nu; ++s) {
            cta_spmv<ND>(L, e, L.t);
            __syncthreads();
            cta_jacobi<ND, false>(L, omega, r, L.t, e);
            __syncthreads();
        }
    }
}

// ---- cluster coarse V-cycle (nu = 1): the launch-bound levels in ONE kernel on a thread-block cluster -------------
// Levels whose operator is at most a few MB (cfg 2: 250 x 125 and below) cost ~6 us per kernel in launch latency and
// gaps, five kernels per level and iteration, plus 87 us for the single-CTA kernel above: a quarter of a CG
// iteration.  Here a cluster of up to 16 CTAs (16 SMs, hardware cluster barrier between the phases) walks levels
// [l0, solo) together; the smallest levels [solo, last], where even a cluster barrier costs more than the work, are
// done by CTA 0 alone with __syncthreads() while the others wait at one cluster barrier.  Per level: product with
// residual epilogue, restriction + first sweep of the next level, prolongation (out of place), product with Jacobi
// epilogue - four barriers.  A warp owns 32 / ND whole nodes, so the block-Jacobi epilogue is a few shuffles.
struct Team {
    int tid, stride;
    bool solo;
};

__device__ __forceinline__ void team_sync(const Team& tm) {
    if (tm.solo) __syncthreads();
    else cg::this_cluster().sync();
}

// MODE 1: y = rhs - K x ;  MODE 2: y = x + omega Dinv (rhs - K x)      (x, rhs, y: vectors written inside this kernel)
template <int ND, int MODE>
__device__ void team_product(const LevelDesc& L, const double* x, const double* rhs, double* y, double omega, const Team& tm) {
    constexpr int NPW = 32 / ND;
    const int lane = threadIdx.x & 31;
    const int nl = lane / ND, da = lane - nl * ND;
    const int64_t warp = tm.tid >> 5, n_warps = tm.stride >> 5;
    for (int64_t node0 = warp * NPW; node0 < L.n_nodes; node0 += n_warps * NPW) {
        const int64_t i = node0 + nl;
        const bool valid = nl < NPW && i < L.n_nodes;
        double res = 0.0;
        if (valid) {
            const int pb = __ldg(&L.adj_ptr[i]), len = __ldg(&L.adj_ptr[i + 1]) - pb;
            const double* a = L.vals + (int64_t)ND * ND * pb + (int64_t)da * len * ND;
            double acc = 0.0;
            for (int k = 0; k < len; ++k) {
                const double* xv = x + (int64_t)__ldg(&L.adj_idx[pb + k]) * ND;
#pragma unroll
                for (int d = 0; d < ND; d += 2) {
                    const double2 x2 = __ldcg(reinterpret_cast<const double2*>(xv + d));
                    acc = fma(__ldg(&a[k * ND + d]), x2.x, acc);
                    acc = fma(__ldg(&a[k * ND + d + 1]), x2.y, acc);
                }
            }
            res = __ldcg(&rhs[i * ND + da]) - acc;
        }
        if (MODE == 1) {
            if (valid) y[i * ND + da] = res;
        } else {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) {
                const double rb = __shfl_sync(0xffffffffu, res, nl * ND + b);
                if (valid) t = fma(__ldg(&L.dinv[(i * ND + da) * ND + b]), rb, t);
            }
            if (valid) y[i * ND + da] = __ldcg(&x[i * ND + da]) + omega * t;
        }
    }
}

// e = omega Dinv r
template <int ND>
__device__ void team_jacobi_first(const LevelDesc& L, double omega, const double* r, double* e, const Team& tm) {
    for (int64_t i = tm.tid; i < L.n_nodes; i += tm.stride) {
        double d[ND];
#pragma unroll
        for (int a = 0; a < ND; ++a) d[a] = __ldcg(&r[i * ND + a]);
        const double* B = L.dinv + i * ND * ND;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double sacc = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b) sacc += __ldg(&B[a * ND + b]) * d[b];
            e[i * ND + a] = omega * sacc;
        }
    }
}

// r_c = P^T r_f  and, if `smooth_c`, the coarse level's first sweep e_c = omega Dinv_c r_c
template <int ND>
__device__ void team_restrict(const LevelDesc& F, const LevelDesc& C, const double* rf, double* rc, bool smooth_c,
                              double omega, double* ec, const Team& tm) {
    for (int64_t t = tm.tid; t < C.n_nodes; t += tm.stride) {
        const int64_t J = t / (C.nx + 1), I = t - J * (C.nx + 1);
        const int64_t ic = min(2 * I, F.nx), jc = min(2 * J, F.ny);
        double acc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = 0.0;
        for (int dj = -1; dj <= 1; ++dj) {
            const int64_t j = jc + dj;
            if (j < 0 || j > F.ny) continue;
            int64_t J0, J1; double v0, v1;
            cmap1d(j, F.ny, J0, v0, J1, v1);
            const double wy = (J0 == J ? v0 : 0.0) + (J1 == J && v1 != 0.0 ? v1 : 0.0);
            if (wy == 0.0) continue;
            for (int di = -1; di <= 1; ++di) {
                const int64_t i = ic + di;
                if (i < 0 || i > F.nx) continue;
                int64_t I0, I1; double u0, u1;
                cmap1d(i, F.nx, I0, u0, I1, u1);
                const double wx = (I0 == I ? u0 : 0.0) + (I1 == I && u1 != 0.0 ? u1 : 0.0);
                if (wx == 0.0) continue;
                const int64_t nf = j * (F.nx + 1) + i;
                const uint32_t fm = F.fixed ? F.fixed[nf] : 0u;
                const double w = wx * wy;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    if (!((fm >> d) & 1u)) acc[d] += w * __ldcg(&rf[nf * ND + d]);
            }
        }
        const uint32_t cm = C.fixed ? C.fixed[t] : 0u;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            if ((cm >> d) & 1u) acc[d] = 0.0;
            rc[t * ND + d] = acc[d];
        }
        if (smooth_c) {
            const double* B = C.dinv + t * ND * ND;
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < ND; ++b) sacc += __ldg(&B[a * ND + b]) * acc[b];
                ec[t * ND + a] = omega * sacc;
            }
        }
    }
}

// out = e_f + P e_c
template <int ND>
__device__ void team_prolong(const LevelDesc& F, const LevelDesc& C, const double* ec, const double* ef, double* out, const Team& tm) {
    for (int64_t t = tm.tid; t < F.n_nodes; t += tm.stride) {
        const int64_t j = t / (F.nx + 1), i = t - j * (F.nx + 1);
        int64_t I0, I1, J0, J1; double u0, u1, v0, v1;
        cmap1d(i, F.nx, I0, u0, I1, u1);
        cmap1d(j, F.ny, J0, v0, J1, v1);
        const uint32_t fm = F.fixed ? F.fixed[t] : 0u;
        const int64_t c00 = J0 * (C.nx + 1) + I0, c10 = J0 * (C.nx + 1) + I1, c01 = J1 * (C.nx + 1) + I0, c11 = J1 * (C.nx + 1) + I1;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double v = 0.0;
            if (!((fm >> d) & 1u)) {
                v = u0 * v0 * __ldcg(&ec[c00 * ND + d]);
                if (u1 != 0.0) v += u1 * v0 * __ldcg(&ec[c10 * ND + d]);
                if (v1 != 0.0) { v += u0 * v1 * __ldcg(&ec[c01 * ND + d]); if (u1 != 0.0) v += u1 * v1 * __ldcg(&ec[c11 * ND + d]); }
            }
            out[t * ND + d] = __ldcg(&ef[t * ND + d]) + v;
        }
    }
}

template <int ND>
__device__ void team_down(const LevelDesc* lv, int l, int l0, int last, double omega, const double* r0, double* e0, const Team& tm) {
    const LevelDesc L = lv[l], C = lv[l + 1];
    const double* r = l == l0 ? r0 : L.r;
    double* e = l == l0 ? e0 : L.e;
    team_product<ND, 1>(L, e, r, L.t, 0.0, tm);                       // t = r - K e
    team_sync(tm);
    team_restrict<ND>(L, C, L.t, C.r, l + 1 != last, omega, C.e, tm);  // r_c (and e_c = omega Dinv r_c)
    team_sync(tm);
}

template <int ND>
__device__ void team_up(const LevelDesc* lv, int l, int l0, double omega, const double* r0, double* e0, const Team& tm) {
    const LevelDesc L = lv[l], C = lv[l + 1];
    const double* r = l == l0 ? r0 : L.r;
    double* e = l == l0 ? e0 : L.e;
    team_prolong<ND>(L, C, C.e, e, L.t, tm);                           // t = e + P e_c
    team_sync(tm);
    team_product<ND, 2>(L, L.t, r, e, omega, tm);                      // e = t + omega Dinv (r - K t)
    team_sync(tm);
}

// levels l0 .. last with l0 < last; e0 must already hold the first sweep omega Dinv r0 (the caller's restriction wrote
// it); [l0, solo) on the whole cluster, [solo, last] on CTA 0, l0 <= solo <= last
template <int ND>
__global__ void __launch_bounds__(1024)
cluster_vcycle_kernel(const LevelDesc* __restrict__ lv, int l0, int solo, int last, double omega,
                      const double* __restrict__ coarse_inv, const double* r0, double* e0) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), csz = (int)cluster.num_blocks();
    const Team all = {rank * (int)blockDim.x + (int)threadIdx.x, csz * (int)blockDim.x, false};
    const Team one = {(int)threadIdx.x, (int)blockDim.x, true};
    for (int l = l0; l < solo; ++l) team_down<ND>(lv, l, l0, last, omega, r0, e0, all);
    if (rank == 0) {
        for (int l = solo; l < last; ++l) team_down<ND>(lv, l, l0, last, omega, r0, e0, one);
        {   // coarsest: dense inverse
            const LevelDesc L = lv[last];
            const int n = (int)(L.n_nodes * ND);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int i = warp; i < n; i += nw) {
                double sacc = 0.0;
                for (int j = lane; j < n; j += 32) sacc += __ldg(&coarse_inv[(int64_t)i * n + j]) * __ldcg(&L.r[j]);
                sacc = warp_sum(sacc);
                if (lane == 0) L.e[i] = sacc;
            }
            __syncthreads();
        }
        for (int l = last - 1; l >= solo; --l) team_up<ND>(lv, l, l0, omega, r0, e0, one);
    }
    if (solo > l0) cluster.sync();      // the solo levels' correction is visible to the whole cluster
    for (int l = solo - 1; l >= l0; --l) team_up<ND>(lv, l, l0, omega, r0, e0, all);
}

int grid_for(int64_t n, int block = 256) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = (int64_t)femol_sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// largest cluster (16, 8, 4, 2, 1 CTAs of 1024 threads) the device can co-schedule for the cluster kernel
template <int ND>
int cluster_size_probe() {
    auto kern = cluster_vcycle_kernel<ND>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
    const char* env = getenv("FEMOL_GMG_CLUSTER_SIZE");
    int want = env ? atoi(env) : 16;
    for (int cs = want; cs >= 1; cs >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(1024);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1) return cs;
        cudaGetLastError();
    }
    return 0;
}
int cluster_size_for(int nd) {
    static int cached[2] = {-1, -1};
    int& c = cached[nd == 2 ? 0 : 1];
    if (c < 0) c = nd == 2 ? cluster_size_probe<2>() : cluster_size_probe<6>();
    return c;
}

template <int ND>
int launch_cluster_vcycle(femol_gmg* g, int l, const double* r, double* e, cudaStream_t st) {
    const int solo = g->solo_from < l ? l : g->solo_from;
    const int cs = solo > l ? g->cluster_size : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    FEMOL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, cluster_vcycle_kernel<ND>, (const LevelDesc*)g->d_desc, l, solo, g->n_levels - 1,
                                        g->omega, (const double*)g->coarse_inv, r, e));
    femol_count_launch();
    return FEMOL_OK;
}

template <int ND>
int smooth(femol_gmg* g, int l, const double* r, double* e, bool zero_start, cudaStream_t st) {
    Level& L = g->lv[l];
    const int grid = grid_for(L.n_nodes);
    for (int s = 0; s < g->nu; ++s) {
        if (s == 0 && zero_start) {
            jacobi_kernel<ND, true><<<grid, 256, 0, st>>>(L.n_nodes, g->omega, L.dinv, r, nullptr, e);
            FEMOL_LAUNCH_CHECK();
        } else {
            const femol_epilogue none = {FEMOL_EPI_NONE, nullptr, nullptr, 0.0, nullptr, nullptr};
            int rc = L.vals32 ? femol_spmv_f32_epilogue_launch(L.plan, L.vals32, e, L.t, none, st)
                              : femol_apply_launch(L.plan, L.vals, e, L.t, nullptr, nullptr, nullptr, st);
            if (rc != FEMOL_OK) return rc;
            jacobi_kernel<ND, false><<<grid, 256, 0, st>>>(L.n_nodes, g->omega, L.dinv, r, L.t, e);
            FEMOL_LAUNCH_CHECK();
        }
    }
    return FEMOL_OK;
}

// `presmoothed`: e already holds the first sweep omega Dinv r (written by the restriction that produced r)
template <int ND>
int vcycle(femol_gmg* g, int l, const double* r, double* e, cudaStream_t st, bool presmoothed = false,
           double* rz_out = nullptr, double* red_ws = nullptr, bool* rz_fused = nullptr, int cyc = -1) {
    if (cyc < 0) cyc = g->cycle;     // 0: V, 1: F (second coarse visit is a V-cycle), 2: W
    Level& L = g->lv[l];
    if (g->nu == 1 && l >= 1 && l >= g->cluster_from && presmoothed) return launch_cluster_vcycle<ND>(g, l, r, e, st);
    if (g->nu != 1 && l >= 1 && l >= g->fused_from && g->d_desc) {
        coarse_vcycle_kernel<ND><<<1, 1024, 0, st>>>((const LevelDesc*)g->d_desc, l, g->n_levels - 1, g->nu, g->omega, g->coarse_inv, r, e);
        FEMOL_LAUNCH_CHECK();
        return FEMOL_OK;
    }
    if (l == g->n_levels - 1) {
        dense_apply<<<grid_for((int64_t)L.n * 32), 256, 0, st>>>((int)L.n, g->coarse_inv, r, e);
        FEMOL_LAUNCH_CHECK();
        return FEMOL_OK;
    }
    Level& C = g->lv[l + 1];
    int rc;
    if (g->nu == 1) {
        // One pre- and one post-sweep with everything fused into the products (5 launches per level, was 7):
        //   e = omega Dinv r                     (level 0: jacobi kernel; below: epilogue of the restriction)
        //   t = r - K e                          (product, residual epilogue)
        //   r_c = P^T t, e_c = omega Dinv_c r_c  (restriction)
        //   ... coarse correction ...
        //   t = e + P e_c                        (prolongation, out of place)
        //   e = t + omega Dinv (r - K t)         (product, Jacobi epilogue)
        if (!presmoothed) {
            jacobi_kernel<ND, true><<<grid_for(L.n_nodes), 256, 0, st>>>(L.n_nodes, g->omega, L.dinv, r, nullptr, e);
            FEMOL_LAUNCH_CHECK();
        }
        const femol_epilogue res = {FEMOL_EPI_RESIDUAL, r, nullptr, 0.0, nullptr, nullptr};
        rc = L.vals32 ? femol_spmv_f32_epilogue_launch(L.plan, L.vals32, e, L.t, res, st)
                      : femol_apply_epilogue_launch(L.plan, L.vals, e, L.t, res, st);
        if (rc != FEMOL_OK) return rc;
        const bool coarse_regular = l + 1 != g->n_levels - 1;   // every smoothing level starts from omega Dinv r
        FEMOL_CUDA_CHECK(femol_launch_pdl(restrict_kernel<ND>, dim3(grid_for(C.n_nodes)), dim3(256), 0, st, L.nx, L.ny, C.nx, C.ny,
                                          (const uint8_t*)L.fixed, (const uint8_t*)C.fixed, (const double*)L.t, C.r,
                                          (const double*)(coarse_regular ? C.dinv : nullptr), g->omega, C.e,
                                          (const int32_t*)L.lat_node));
        femol_count_launch();
        rc = vcycle<ND>(g, l + 1, C.r, C.e, st, coarse_regular, nullptr, nullptr, nullptr, cyc);
        if (rc != FEMOL_OK) return rc;
        if (cyc > 0 && coarse_regular && l + 1 <= g->cycle_depth) {
            // second visit of the coarse level: e_c += Cycle(r_c - K_c e_c)
            const femol_epilogue res2 = {FEMOL_EPI_RESIDUAL, C.r, nullptr, 0.0, nullptr, nullptr};
            rc = C.vals32 ? femol_spmv_f32_epilogue_launch(C.plan, C.vals32, C.e, C.r2, res2, st)
                          : femol_apply_epilogue_launch(C.plan, C.vals, C.e, C.r2, res2, st);
            if (rc != FEMOL_OK) return rc;
            rc = vcycle<ND>(g, l + 1, C.r2, C.e2, st, false, nullptr, nullptr, nullptr, cyc == 1 ? 0 : cyc);
            if (rc != FEMOL_OK) return rc;
            axpy_kernel<<<grid_for(C.n), 256, 0, st>>>(C.n, C.e2, C.e);
            FEMOL_LAUNCH_CHECK();
        }
        if (femol_apply_has_jacobi(L.plan, L.vals)) {
            FEMOL_CUDA_CHECK(femol_launch_pdl(prolong_add_kernel<ND>, dim3(grid_for(L.n_nodes)), dim3(256), 0, st, L.nx, L.ny, C.nx,
                                              (const uint8_t*)L.fixed, (const double*)C.e, (const double*)e, L.t,
                                              (const int32_t*)L.lat_node));
            femol_count_launch();
            // level 0 inside CG: the last kernel of the V-cycle also delivers r . z
            // (the single-precision product has no fused dot: with a float copy of level 0 the caller computes r . z)
            const bool fuse_rz = rz_out && !L.vals32;
            const femol_epilogue jac = {FEMOL_EPI_JACOBI, r, L.dinv, g->omega, fuse_rz ? rz_out : nullptr, fuse_rz ? red_ws : nullptr};
            if (fuse_rz && rz_fused) *rz_fused = true;
            return L.vals32 ? femol_spmv_f32_epilogue_launch(L.plan, L.vals32, L.t, e, jac, st)
                            : femol_apply_epilogue_launch(L.plan, L.vals, L.t, e, jac, st);
        }
        prolong_add_kernel<ND><<<grid_for(L.n_nodes), 256, 0, st>>>(L.nx, L.ny, C.nx, L.fixed, C.e, e, e, L.lat_node);
        FEMOL_LAUNCH_CHECK();
        return smooth<ND>(g, l, r, e, false, st);
    }
    rc = smooth<ND>(g, l, r, e, true, st);
    if (rc != FEMOL_OK) return rc;
    rc = femol_apply_launch(L.plan, L.vals, e, L.t, nullptr, nullptr, nullptr, st);
    if (rc != FEMOL_OK) return rc;
    residual_kernel<<<grid_for(L.n), 256, 0, st>>>(L.n, r, L.t, L.t);
    FEMOL_LAUNCH_CHECK();
    restrict_kernel<ND><<<grid_for(C.n_nodes), 256, 0, st>>>(L.nx, L.ny, C.nx, C.ny, L.fixed, C.fixed, L.t, C.r, nullptr, 0.0, nullptr, L.lat_node);
    FEMOL_LAUNCH_CHECK();
    rc = vcycle<ND>(g, l + 1, C.r, C.e, st);
    if (rc != FEMOL_OK) return rc;
    prolong_add_kernel<ND><<<grid_for(L.n_nodes), 256, 0, st>>>(L.nx, L.ny, C.nx, L.fixed, C.e, e, e, L.lat_node);
    FEMOL_LAUNCH_CHECK();
    return smooth<ND>(g, l, r, e, false, st);
}

}  // namespace

// nu = 1: the caller may compute the V-cycle's first sweep z = omega Dinv r itself (level-0 block inverse) and
// pass presmoothed = true to femol_gmg_vcycle_internal
bool femol_gmg_first_sweep(femol_gmg* g, const double** dinv, double* omega) {
    if (g->nu != 1 || g->n_levels < 2) return false;
    *dinv = g->lv[0].dinv; *omega = g->omega;
    return true;
}

// Attach the hierarchy of S K S, S = diag((-1)^(i+j)) over the grid nodes (all dofs of a node share the sign): built
// by the caller with femol_gmg_create / femol_gmg_setup from the element matrix S_e Ke S_e on a plan VIEW whose
// matrix-free descriptor carries that element matrix.  NULL detaches.  `g` does not own `mod`.
extern "C" int femol_gmg_attach_modulated(femol_gmg* g, femol_gmg* mod) {
    FEMOL_REQUIRE(g, "femol_gmg_attach_modulated: null hierarchy");
    if (!mod) { g->mod = nullptr; return FEMOL_OK; }
    FEMOL_REQUIRE(mod->nd == g->nd && mod->lv[0].nx == g->lv[0].nx && mod->lv[0].ny == g->lv[0].ny && !mod->mod,
                  "femol_gmg_attach_modulated: hierarchies of different grids");
    if (!g->mod_r) {
        FEMOL_CUDA_CHECK(cudaMalloc(&g->mod_r, g->lv[0].n * sizeof(double)));
        FEMOL_CUDA_CHECK(cudaMalloc(&g->mod_z, g->lv[0].n * sizeof(double)));
    }
    g->mod = mod;
    return FEMOL_OK;
}

// remembers the CG iteration count of the last solve (set >= 0 stores, set < 0 only reads)
int femol_gmg_iteration_hint(femol_gmg* g, int set) {
    if (set >= 0) g->hint_iters = set;
    return g->hint_iters;
}

// z = V(r).  With rz_out (device scalar) the product r . z is requested as well; *rz_fused tells whether the V-cycle's
// last kernel delivered it (otherwise the caller computes it).
int femol_gmg_vcycle_internal(femol_gmg* g, const double* r, double* z, cudaStream_t st, double* rz_out, double* red_ws,
                              bool* rz_fused, bool presmoothed) {
    if (rz_fused) *rz_fused = false;
    if (g->parent) {          // the parent's fine-level operator pointer changes with every set-up
        g->lv[0].vals = g->parent->lv[0].vals;
        if (g->mod && g->parent->mod) g->mod->lv[0].vals = g->parent->mod->lv[0].vals;
    }
    if (g->dist) return femol_gmg_dist_vcycle(g, r, z, st, presmoothed);
    if (g->capped) { femol_set_error("femol_gmg: a truncated hierarchy needs femol_gmg_set_dist"); return FEMOL_E_ARG; }
    if (g->mod) {
        // z = V(r) + S V_h(S r): the second cycle works on the sign-modulated operator S K S, whose smooth functions
        // are the hourglass (checkerboard x smooth envelope) functions of K - the low-energy family of the
        // reference's selectively integrated Q4 plate that neither the smoother nor the coarse grids of K see
        int rc = g->nd == 2 ? vcycle<2>(g, 0, r, z, st, presmoothed) : vcycle<6>(g, 0, r, z, st, presmoothed);
        if (rc != FEMOL_OK) return rc;
        const Level& L0 = g->lv[0];
        const int grid = grid_for(L0.n_nodes);
        sign_modulate<<<grid, 256, 0, st>>>(L0.n_nodes, (int)L0.nx + 1, g->nd, r, g->mod_r, 0);
        FEMOL_LAUNCH_CHECK();
        rc = g->nd == 2 ? vcycle<2>(g->mod, 0, g->mod_r, g->mod_z, st, false) : vcycle<6>(g->mod, 0, g->mod_r, g->mod_z, st, false);
        if (rc != FEMOL_OK) return rc;
        sign_modulate<<<grid, 256, 0, st>>>(L0.n_nodes, (int)L0.nx + 1, g->nd, g->mod_z, z, 1);
        FEMOL_LAUNCH_CHECK();
        return FEMOL_OK;
    }
    return g->nd == 2 ? vcycle<2>(g, 0, r, z, st, presmoothed, rz_out, red_ws, rz_fused)
                      : vcycle<6>(g, 0, r, z, st, presmoothed, rz_out, red_ws, rz_fused);
}

extern "C" int femol_gmg_destroy(femol_gmg* g) {
    if (!g) return FEMOL_OK;
    if (g->parent) {          // workspace view: only the work vectors (and the view of the attached hierarchy) are its own
        for (int l = 0; l < g->n_levels; ++l) {
            Level& L = g->lv[l];
            if (l > 0) { cudaFree(L.r); cudaFree(L.e); cudaFree(L.r2); cudaFree(L.e2); }
            cudaFree(L.t);
        }
        cudaFree(g->mod_r); cudaFree(g->mod_z);
        if (g->mod) femol_gmg_destroy(g->mod);
        delete g;
        return FEMOL_OK;
    }
    for (int l = 0; l < g->n_levels; ++l) {
        Level& L = g->lv[l];
        if (l > 0) {
            femol_plan_destroy(L.plan);
            cudaFree(L.fixed); cudaFree(L.ke); cudaFree(L.vals); cudaFree(L.vals32); cudaFree(L.r); cudaFree(L.e); cudaFree(L.r2); cudaFree(L.e2);
        }
        else cudaFree(L.vals32);
        cudaFree(L.dinv); cudaFree(L.t);
    }
    cudaFree(g->coarse_inv);
    cudaFree(g->child_g);
    cudaFree(g->d_desc);
    cudaFree(g->mod_r); cudaFree(g->mod_z);
    cudaFree(g->lv[0].lat_node); cudaFree(g->node_lat); cudaFree(g->child_ptr); cudaFree(g->child_idx);
    if (g->dist) {
        cudaFree(g->dist->gke); cudaFree(g->dist->gvals); cudaFree(g->dist->gr); cudaFree(g->dist->ge);
        delete g->dist;
    }
    delete g;
    return FEMOL_OK;
}

namespace {
struct LatticeArgs { const int32_t *lat_node, *node_lat, *child_ptr, *child_idx; };
int gmg_create_impl(int64_t nx, int64_t ny, int ndof, femol_plan_t fine_plan, const uint8_t* fine_fixed, int nu, double omega,
                    const LatticeArgs* lat, cudaStream_t st, femol_gmg** out);
}  // namespace

extern "C" int femol_gmg_create(int64_t nx, int64_t ny, int ndof, femol_plan_t fine_plan, const uint8_t* fine_fixed,
                                int nu, double omega, void* stream, femol_gmg** out) {
    FEMOL_REQUIRE(out && fine_plan && nx >= 1 && ny >= 1, "femol_gmg_create: bad argument");
    FEMOL_REQUIRE(ndof == 2 || ndof == 6, "femol_gmg_create: ndof must be 2 or 6");
    FEMOL_REQUIRE(fine_plan->nn == 4 && fine_plan->ndof == ndof && fine_plan->n_elem == nx * ny &&
                  fine_plan->n_nodes == (nx + 1) * (ny + 1), "femol_gmg_create: plan is not an nx x ny structured Q4 mesh");
    return gmg_create_impl(nx, ny, ndof, fine_plan, fine_fixed, nu, omega, nullptr, (cudaStream_t)stream, out);
}

// Lattice hierarchy: the fine mesh is ANY single-cell-type mesh (T3, Q4, T6; any node / element numbering) whose nodes
// occupy every point of an (Nx + 1) x (Ny + 1) lattice exactly once and whose elements each lie inside one cell of the
// coarse grid over every second lattice line.  lat_node [(Nx+1)(Ny+1)]: lattice index (x-fastest) -> node; node_lat
// [n_nodes]: its inverse; child_ptr [NEc + 1] / child_idx [n_elem]: the fine elements inside each coarse element
// (NEc = ceil(Nx/2) ceil(Ny/2), coarse elements numbered I * ceil(Ny/2) + J), ascending.  Device arrays, copied.
extern "C" int femol_gmg_create_lattice(int64_t Nx, int64_t Ny, int ndof, femol_plan_t fine_plan, const uint8_t* fine_fixed,
                                        const int32_t* lat_node, const int32_t* node_lat, const int32_t* child_ptr,
                                        const int32_t* child_idx, int nu, double omega, void* stream, femol_gmg** out) {
    FEMOL_REQUIRE(out && fine_plan && Nx >= 1 && Ny >= 1 && lat_node && node_lat && child_ptr && child_idx,
                  "femol_gmg_create_lattice: bad argument");
    FEMOL_REQUIRE(ndof == 2 || ndof == 6, "femol_gmg_create_lattice: ndof must be 2 or 6");
    FEMOL_REQUIRE((fine_plan->nn == 3 || fine_plan->nn == 4 || fine_plan->nn == 6) && fine_plan->ndof == ndof &&
                  fine_plan->n_nodes == (Nx + 1) * (Ny + 1) && fine_plan->n_nodes < (1LL << 31),
                  "femol_gmg_create_lattice: the plan's nodes do not fill an (Nx+1) x (Ny+1) lattice");
    const LatticeArgs lat = {lat_node, node_lat, child_ptr, child_idx};
    return gmg_create_impl(Nx, Ny, ndof, fine_plan, fine_fixed, nu, omega, &lat, (cudaStream_t)stream, out);
}

namespace {
int gmg_create_impl(int64_t nx, int64_t ny, int ndof, femol_plan_t fine_plan, const uint8_t* fine_fixed, int nu, double omega,
                    const LatticeArgs* lat, cudaStream_t st, femol_gmg** out) {
    femol_gmg* g = new femol_gmg();
    memset(g, 0, sizeof(*g));
    g->nd = ndof; g->nu = nu > 0 ? nu : 2; g->omega = omega > 0 ? omega : 0.6;
    {   // mixed precision: only the stored coarse operators are rounded, every vector and sum stays in double
        const char* env = getenv("FEMOL_GMG_F32");
        g->f32_levels = g->nu == 1 && !(env && atoi(env) == 0);
    }
    auto fail = [&](int code) { femol_gmg_destroy(g); return code; };
#define GCHECK(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            femol_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return fail(FEMOL_E_CUDA);                                                            \
        }                                                                                         \
    } while (0)
    int l = 0;
    int64_t cx = nx, cy = ny;
    const int S2 = 16 * ndof * ndof;
    while (true) {
        Level& L = g->lv[l];
        L.nx = cx; L.ny = cy; L.n_nodes = (cx + 1) * (cy + 1); L.n_elem = cx * cy; L.n = L.n_nodes * ndof;
        g->n_levels = l + 1;
        GCHECK(cudaMalloc(&L.dinv, L.n_nodes * ndof * ndof * sizeof(double)));
        GCHECK(cudaMalloc(&L.t, L.n * sizeof(double)));
        if (l == 0) {
            L.plan = fine_plan;
            L.fixed = const_cast<uint8_t*>(fine_fixed);
            if (lat) {
                const int64_t nec = coarse_of(cx) * coarse_of(cy);
                g->lattice = true;
                GCHECK(cudaMalloc(&L.lat_node, L.n_nodes * sizeof(int32_t)));
                GCHECK(cudaMalloc(&g->node_lat, L.n_nodes * sizeof(int32_t)));
                GCHECK(cudaMalloc(&g->child_ptr, (nec + 1) * sizeof(int32_t)));
                GCHECK(cudaMalloc(&g->child_idx, fine_plan->n_elem * sizeof(int32_t)));
                GCHECK(cudaMemcpyAsync(L.lat_node, lat->lat_node, L.n_nodes * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
                GCHECK(cudaMemcpyAsync(g->node_lat, lat->node_lat, L.n_nodes * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
                GCHECK(cudaMemcpyAsync(g->child_ptr, lat->child_ptr, (nec + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
                GCHECK(cudaMemcpyAsync(g->child_idx, lat->child_idx, fine_plan->n_elem * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
                // Level 0 of a lattice hierarchy is an assembled CSR operator (no stencil): the two products of the V-cycle
                // (residual, post-smoothing) read a single-precision copy like the coarse levels do - 12 -> 8 bytes per
                // node-block entry pair, and on T6 meshes these two products are 2/3 of a CG iteration.  CG's own product
                // K p and the residual it converges on stay FP64.  FEMOL_GMG_F32_FINE=0 switches the copy off.
                const char* ef = getenv("FEMOL_GMG_F32_FINE");
                if (g->f32_levels && ndof == 2 && !(ef && atoi(ef) == 0)) GCHECK(cudaMalloc(&L.vals32, fine_plan->nnz * sizeof(float)));
            }
        } else {
            Level& F = g->lv[l - 1];
            int32_t* conn = nullptr;
            GCHECK(cudaMalloc(&conn, L.n_elem * 4 * sizeof(int32_t)));
            make_conn<<<grid_for(L.n_elem), 256, 0, st>>>(cx, cy, conn);
            int rc = femol_plan_create(L.n_elem, 4, ndof, L.n_nodes, conn, st, &L.plan);
            cudaFree(conn);
            if (rc != FEMOL_OK) return fail(rc);
            GCHECK(cudaMalloc(&L.fixed, L.n_nodes));
            coarsen_mask<<<grid_for(L.n_nodes), 256, 0, st>>>(F.nx, F.ny, F.fixed, cx, cy, L.fixed, F.lat_node);
            GCHECK(cudaGetLastError());
            GCHECK(cudaMalloc(&L.ke, L.n_elem * S2 * sizeof(double)));
            GCHECK(cudaMalloc(&L.vals, L.plan->nnz * sizeof(double)));
            L.vals32 = nullptr;
            if (g->f32_levels) GCHECK(cudaMalloc(&L.vals32, L.plan->nnz * sizeof(float)));
            GCHECK(cudaMalloc(&L.r, L.n * sizeof(double)));
            GCHECK(cudaMalloc(&L.e, L.n * sizeof(double)));
            GCHECK(cudaMalloc(&L.r2, L.n * sizeof(double)));
            GCHECK(cudaMalloc(&L.e2, L.n * sizeof(double)));
        }
        const bool small = L.n <= GMG_COARSE_MAX_DOF && l > 0;
        if (small || (cx == 1 && cy == 1) || l == GMG_MAX_LEVELS - 1) break;
        cx = coarse_of(cx); cy = coarse_of(cy);
        ++l;
    }
    Level& CL = g->lv[g->n_levels - 1];
    if (g->n_levels < 2 || CL.n > 4096) {
        femol_set_error("femol_gmg_create: mesh too small (%d levels) or coarsest level too large (%lld dofs)", g->n_levels, (long long)CL.n);
        return fail(FEMOL_E_ARG);
    }
    g->nc = CL.n;
    GCHECK(cudaMalloc(&g->coarse_inv, CL.n * CL.n * sizeof(double)));
    GCHECK(cudaMalloc(&g->child_g, (size_t)2 * 4 * 16 * g->nd * g->nd * sizeof(double)));
    {   // descriptors for the fused single-CTA coarse V-cycle (levels with at most a few thousand nodes)
        LevelDesc h[GMG_MAX_LEVELS];
        memset(h, 0, sizeof(h));
        g->fused_from = g->n_levels;
        const char* env = getenv("FEMOL_GMG_FUSE_NODES");
        const long long fuse_nodes = env ? atoll(env) : 3000;  // measured on cfg 2: 3000 -> -2 %, 10000 -> +13 %
        for (int k = 0; k < g->n_levels; ++k) {
            Level& L = g->lv[k];
            h[k].nx = L.nx; h[k].ny = L.ny; h[k].n_nodes = L.n_nodes;
            h[k].adj_ptr = L.plan->adj_ptr; h[k].adj_idx = L.plan->adj_idx;
            h[k].vals = L.vals; h[k].dinv = L.dinv; h[k].fixed = L.fixed;
            h[k].r = L.r; h[k].e = L.e; h[k].t = L.t;
            if (k >= 1 && L.n_nodes <= fuse_nodes && g->fused_from == g->n_levels) g->fused_from = k;
        }
        GCHECK(cudaMalloc(&g->d_desc, sizeof(h)));
        GCHECK(cudaMemcpyAsync(g->d_desc, h, sizeof(h), cudaMemcpyHostToDevice, st));
        // cluster kernel (nu = 1): levels whose operator is at most FEMOL_GMG_CLUSTER_MB, the smallest ones solo
        const int last = g->n_levels - 1;
        g->cluster_from = g->solo_from = g->n_levels;
        g->cluster_size = 0;
        g->hint_iters = 0;
        g->mod = nullptr; g->mod_r = nullptr; g->mod_z = nullptr;
        {
            const char* ec = getenv("FEMOL_GMG_CYCLE");
            g->cycle = (ec && (ec[0] == 'F' || ec[0] == 'f')) ? 1 : (ec && (ec[0] == 'W' || ec[0] == 'w')) ? 2 : 0;
            const char* ed = getenv("FEMOL_GMG_CYCLE_DEPTH");
            g->cycle_depth = ed ? atoi(ed) : 99;
        }
        const char* env_mb = getenv("FEMOL_GMG_CLUSTER_MB");
        const double cluster_mb = env_mb ? atof(env_mb) : 0.0;   // OFF by default: with PDL launches plain kernels win (cfg 2, it/s: off 62-67, 0.7 MB 60-63, 12 MB 47)
        const char* env_solo = getenv("FEMOL_GMG_SOLO_NODES");
        const long long solo_nodes = env_solo ? atoll(env_solo) : 700;
        if (g->nu == 1 && cluster_mb > 0) {
            for (int k = 1; k < last; ++k) {
                const double mb = (double)g->lv[k].plan->nnz_node * g->nd * g->nd * 8.0 / 1e6;
                if (mb <= cluster_mb) { g->cluster_from = k; break; }
            }
            if (g->cluster_from < last) {
                g->solo_from = last;
                for (int k = g->cluster_from; k < last; ++k)
                    if (g->lv[k].n_nodes <= solo_nodes) { g->solo_from = k; break; }
                g->cluster_size = cluster_size_for(g->nd);
                if (g->cluster_size == 0) g->cluster_from = g->solo_from = g->n_levels;
            } else {
                g->cluster_from = g->n_levels;
            }
        }
    }
    GCHECK(cudaStreamSynchronize(st));
#undef GCHECK
    *out = g;
    return FEMOL_OK;
}
}  // namespace

// (Re)build the hierarchy for the current densities: called once per assembled fine matrix.
namespace {
int gmg_setup_impl(femol_gmg* g, const double* e_base, int base_pe, const double* e_coat, int coat_pe, const double* xpow,
                   const double* fine_vals, cudaStream_t st);
}  // namespace

extern "C" int femol_gmg_setup(femol_gmg* g, const double* e_base, const double* e_coat, const double* xpow,
                               const double* fine_vals, void* stream) {
    FEMOL_REQUIRE(g && e_base && fine_vals, "femol_gmg_setup: null argument");
    FEMOL_REQUIRE(!g->lattice, "femol_gmg_setup: lattice hierarchies are refreshed with femol_gmg_setup_lattice");
    return gmg_setup_impl(g, e_base, 0, e_coat, 0, xpow, fine_vals, (cudaStream_t)stream);
}

// Lattice hierarchy: element matrices of the FINE mesh, shared (pe = 0: one S x S matrix) or one per element (pe = 1,
// the reference's unstructured path FEM.py:608-695), base [+ coating]; child = xpow_e * base, or base + xpow_e * coat.
extern "C" int femol_gmg_setup_lattice(femol_gmg* g, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                                       const double* xpow, const double* fine_vals, void* stream) {
    FEMOL_REQUIRE(g && e_base && fine_vals, "femol_gmg_setup_lattice: null argument");
    FEMOL_REQUIRE(g->lattice, "femol_gmg_setup_lattice: not a lattice hierarchy");
    return gmg_setup_impl(g, e_base, base_pe, e_coat, coat_pe, xpow, fine_vals, (cudaStream_t)stream);
}

namespace {
template <int ND, int NN>
void launch_coarsen_lattice(femol_gmg* g, int blocks, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                            const double* xpow, cudaStream_t st) {
    Level& F = g->lv[0];
    Level& C = g->lv[1];
    coarsen_lattice<ND, NN><<<blocks, 128, 0, st>>>((int)F.nx, (int)F.ny, (int)C.nx, (int)C.ny, F.plan->conn, g->node_lat,
                                                    g->child_ptr, g->child_idx, e_base, base_pe, e_coat, coat_pe, xpow,
                                                    F.fixed, C.ke);
}

int gmg_setup_impl(femol_gmg* g, const double* e_base, int base_pe, const double* e_coat, int coat_pe, const double* xpow,
                   const double* fine_vals, cudaStream_t st) {
    const int nd = g->nd;
    g->lv[0].vals = const_cast<double*>(fine_vals);
    if (g->lv[0].vals32) {
        int rc = femol_to_f32_launch(g->lv[0].plan, fine_vals, g->lv[0].vals32, st);
        if (rc != FEMOL_OK) return rc;
    }
    const char* env_old = getenv("FEMOL_GMG_OLD_COARSEN");
    const bool fast_coarsen = !(env_old && atoi(env_old) != 0);
    for (int l = 1; l < g->n_levels; ++l) {
        Level& F = g->lv[l - 1];
        Level& C = g->lv[l];
        int64_t blocks = (C.n_elem + 3) / 4;
        const int64_t cap = (int64_t)femol_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        if (l == 1 && g->lattice) {
            const int nn = F.plan->nn;
#define LAT_CASE(ND, NN) \
            if (nd == ND && nn == NN) launch_coarsen_lattice<ND, NN>(g, (int)blocks, e_base, base_pe, e_coat, coat_pe, xpow, st);
            LAT_CASE(2, 3) LAT_CASE(2, 4) LAT_CASE(2, 6) LAT_CASE(6, 3) LAT_CASE(6, 4) LAT_CASE(6, 6)
#undef LAT_CASE
        } else if (nd == 2) {
            if (l == 1) {
                child_images<2><<<2, 256, 0, st>>>(e_base, e_coat, g->child_g);
                coarsen_elements<2, true><<<(int)blocks, 128, 0, st>>>(F.nx, F.ny, C.nx, C.ny, nullptr, e_base, e_coat, xpow, F.fixed, g->child_g, C.ke);
            } else if (fast_coarsen) coarsen_elements_fast<2><<<(int)blocks, 128, 0, st>>>((int)F.nx, (int)F.ny, (int)C.nx, (int)C.ny, F.ke, F.fixed, C.ke);
            else coarsen_elements<2, false><<<(int)blocks, 128, 0, st>>>(F.nx, F.ny, C.nx, C.ny, F.ke, nullptr, nullptr, nullptr, F.fixed, nullptr, C.ke);
        } else {
            if (l == 1) {
                child_images<6><<<18, 256, 0, st>>>(e_base, e_coat, g->child_g);
                coarsen_elements<6, true><<<(int)blocks, 128, 0, st>>>(F.nx, F.ny, C.nx, C.ny, nullptr, e_base, e_coat, xpow, F.fixed, g->child_g, C.ke);
            } else if (fast_coarsen) coarsen_elements_fast<6><<<(int)blocks, 128, 0, st>>>((int)F.nx, (int)F.ny, (int)C.nx, (int)C.ny, F.ke, F.fixed, C.ke);
            else coarsen_elements<6, false><<<(int)blocks, 128, 0, st>>>(F.nx, F.ny, C.nx, C.ny, F.ke, nullptr, nullptr, nullptr, F.fixed, nullptr, C.ke);
        }
        FEMOL_LAUNCH_CHECK();
        if (g->capped && l == g->n_levels - 1) break;   // gather level: only its element matrices are needed
        int rc = femol_assemble(C.plan, C.ke, 1, nullptr, 0, nullptr, 1.0, C.fixed, nullptr, C.vals, st);
        if (rc != FEMOL_OK) return rc;
        if (C.vals32) {
            rc = femol_to_f32_launch(C.plan, C.vals, C.vals32, st);
            if (rc != FEMOL_OK) return rc;
        }
    }
    for (int l = 0; l < g->n_levels - 1; ++l) {
        int rc = femol_precond_setup(g->lv[l].plan, g->lv[l].vals, FEMOL_PRECOND_BLOCK_JACOBI, g->lv[l].dinv, st);
        if (rc != FEMOL_OK) return rc;
    }
    if (g->capped) return FEMOL_OK;
    Level& CL = g->lv[g->n_levels - 1];
    FEMOL_CUDA_CHECK(cudaMemsetAsync(g->coarse_inv, 0, CL.n * CL.n * sizeof(double), st));
    csr_to_dense<<<grid_for(CL.n), 256, 0, st>>>(CL.n, CL.plan->adj_ptr, CL.plan->adj_idx, nd, CL.vals, g->coarse_inv);
    FEMOL_LAUNCH_CHECK();
    FEMOL_FUNC_ATTR_PER_DEVICE(dense_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GMG_COARSE_MAX_DOF * GMG_COARSE_MAX_DOF * (int)sizeof(double));
    dense_inverse<<<1, 1024, CL.n * CL.n * sizeof(double), st>>>((int)CL.n, g->coarse_inv);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}
}  // namespace

// A second set of work vectors on the SAME hierarchy, so that several V-cycles (one per view, each on its own stream)
// can run concurrently - LOBPCG preconditions its block column by column, and a single multigrid-CG solve is a chain of
// small latency-bound kernels that leaves most of the GPU idle.  The view borrows every operator of `g` (plans, values,
// block inverses, masks, the coarsest inverse) and must be destroyed before it; it follows later femol_gmg_setup calls of
// `g` automatically.  Not for distributed hierarchies.
extern "C" int femol_gmg_clone_workspace(femol_gmg* g, femol_gmg** out) {
    FEMOL_REQUIRE(g && out, "femol_gmg_clone_workspace: null argument");
    FEMOL_REQUIRE(!g->dist && !g->capped && !g->parent, "femol_gmg_clone_workspace: plain hierarchies only");
    femol_gmg* v = new femol_gmg(*g);
    v->parent = g;
    v->mod = nullptr; v->mod_r = nullptr; v->mod_z = nullptr;
    v->fused_from = v->cluster_from = v->solo_from = v->n_levels;      // those kernels use the parent's descriptors
    v->hint_iters = g->hint_iters;
    for (int l = 0; l < v->n_levels; ++l) { Level& L = v->lv[l]; L.r = L.e = L.r2 = L.e2 = L.t = nullptr; }
    bool ok = true;
    for (int l = 0; l < v->n_levels && ok; ++l) {
        Level& L = v->lv[l];
        ok = cudaMalloc(&L.t, L.n * sizeof(double)) == cudaSuccess;
        if (l > 0) ok = ok && cudaMalloc(&L.r, L.n * sizeof(double)) == cudaSuccess && cudaMalloc(&L.e, L.n * sizeof(double)) == cudaSuccess &&
                         cudaMalloc(&L.r2, L.n * sizeof(double)) == cudaSuccess && cudaMalloc(&L.e2, L.n * sizeof(double)) == cudaSuccess;
    }
    if (ok && g->mod) {
        ok = femol_gmg_clone_workspace(g->mod, &v->mod) == FEMOL_OK &&
             cudaMalloc(&v->mod_r, g->lv[0].n * sizeof(double)) == cudaSuccess && cudaMalloc(&v->mod_z, g->lv[0].n * sizeof(double)) == cudaSuccess;
    }
    if (!ok) {
        femol_gmg_destroy(v);
        cudaGetLastError();
        femol_set_error("femol_gmg_clone_workspace: out of device memory");
        return FEMOL_E_CUDA;
    }
    *out = v;
    return FEMOL_OK;
}

extern "C" int femol_gmg_info(femol_gmg* g, int64_t* h_info /* [0]=levels, then (nx, ny) per level */) {
    FEMOL_REQUIRE(g && h_info, "femol_gmg_info: null argument");
    h_info[0] = g->n_levels;
    for (int l = 0; l < g->n_levels; ++l) { h_info[1 + 2 * l] = g->lv[l].nx; h_info[2 + 2 * l] = g->lv[l].ny; }
    return FEMOL_OK;
}

// z = V-cycle(r): one application of the preconditioner (exposed for tests and for LOBPCG)
extern "C" int femol_gmg_apply(femol_gmg* g, const double* r, double* z, void* stream) {
    FEMOL_REQUIRE(g && r && z, "femol_gmg_apply: null argument");
    return femol_gmg_vcycle_internal(g, r, z, (cudaStream_t)stream, nullptr, nullptr, nullptr, false);
}

// =====================================================================================================================
// Strip-distributed hierarchy (one large problem partitioned over the GPUs of a box).
//
// Every rank builds an ORDINARY hierarchy on its strip of lattice rows (femol_gmg_create / _create_lattice on the local
// mesh, truncated to n_dist + 1 levels): the partition cuts are multiples of 2^n_dist lattice rows and a strip starts
// 2^n_dist rows below its first owned row, so the local grids coarsen exactly like the global grid and every local
// element matrix of every level IS the global one (Galerkin sums of local children: no communication in the set-up of
// the distributed levels).  Assembled rows are complete for owned nodes only; a V-cycle therefore exchanges the halo
// rows of a vector before every operator that reads neighbours (4 per level: before the residual product, before the
// restriction, before the prolongation, before the post-smoothing product).  Level n_dist is small: its residual is
// gathered (owned rows, all-reduce of a zero-padded vector) and every rank runs the remaining levels redundantly on the
// whole grid with `coarse`, an ordinary single-GPU hierarchy whose level 0 carries one element matrix per element (the
// gathered Galerkin matrices of level n_dist) - no latency-bound halo traffic on the small levels.
// =====================================================================================================================
bool femol_gmg_is_dist(const femol_gmg* g) { return g && g->dist; }
const femol_strip* femol_gmg_dist_strip(const femol_gmg* g, int level) { return &g->dist->strip[level]; }
femol_comm* femol_gmg_dist_comm(const femol_gmg* g) { return g->dist->comm; }

// Keep the first n_levels levels; the last one becomes a gather level (element matrices only).
extern "C" int femol_gmg_truncate(femol_gmg* g, int n_levels) {
    FEMOL_REQUIRE(g && n_levels >= 2 && n_levels <= g->n_levels, "femol_gmg_truncate: bad level count");
    FEMOL_REQUIRE(g->nu == 1 && !g->mod, "femol_gmg_truncate: needs nu = 1 and no attached hierarchy");
    for (int l = n_levels; l < g->n_levels; ++l) {
        Level& L = g->lv[l];
        femol_plan_destroy(L.plan);
        cudaFree(L.fixed); cudaFree(L.ke); cudaFree(L.vals); cudaFree(L.vals32); cudaFree(L.r); cudaFree(L.e); cudaFree(L.r2); cudaFree(L.e2);
        cudaFree(L.dinv); cudaFree(L.t);
        memset(&L, 0, sizeof(L));
    }
    g->n_levels = n_levels;
    g->capped = true;
    g->fused_from = g->cluster_from = g->solo_from = n_levels;
    cudaFree(g->coarse_inv); g->coarse_inv = nullptr;
    return FEMOL_OK;
}

// h_row0 / h_own_begin / h_own_end: per level 0 .. n_dist, GLOBAL row index of the local grid's first row and of the
// owned row range [begin, end); h_up / h_down per level.  `coarse` = lattice hierarchy (Q4 cells, identity numbering)
// over the global level-n_dist grid with plan `coarse_plan` and node mask `coarse_fixed` (device); all borrowed.
extern "C" int femol_gmg_set_dist(femol_gmg* g, femol_comm* comm, int n_dist, const int64_t* h_row0, const int64_t* h_own_begin,
                                  const int64_t* h_own_end, const int* h_up, const int* h_down, femol_gmg* coarse,
                                  femol_plan_t coarse_plan, const uint8_t* coarse_fixed) {
    FEMOL_REQUIRE(g && comm && coarse && coarse_plan && h_row0 && h_own_begin && h_own_end && h_up && h_down,
                  "femol_gmg_set_dist: null argument");
    FEMOL_REQUIRE(g->capped && n_dist == g->n_levels - 1 && !g->dist, "femol_gmg_set_dist: truncate the hierarchy to n_dist + 1 levels first");
    FEMOL_REQUIRE(coarse->lattice && coarse->nd == g->nd && !coarse->capped && coarse_plan->nn == 4,
                  "femol_gmg_set_dist: `coarse` must be a complete Q4 lattice hierarchy of the same ndof");
    GmgDist* d = new GmgDist();
    memset(d, 0, sizeof(*d));
    d->comm = comm; d->n_dist = n_dist; d->coarse = coarse; d->coarse_plan = coarse_plan; d->coarse_fixed = coarse_fixed;
    for (int l = 0; l <= n_dist; ++l) {
        const Level& L = g->lv[l];
        femol_strip& s = d->strip[l];
        s.row_nodes = L.nx + 1; s.n_rows = L.ny + 1; s.row0 = h_row0[l];
        s.own_lo = h_own_begin[l] - h_row0[l]; s.own_hi = h_own_end[l] - h_row0[l];
        s.h_up = h_up[l]; s.h_down = h_down[l];
        const bool ok = s.own_lo >= 0 && s.own_hi > s.own_lo && s.own_hi <= s.n_rows &&
                        (femol_comm_rank(comm) == 0 || s.own_lo >= s.h_up) &&
                        (femol_comm_rank(comm) == femol_comm_size(comm) - 1 || s.own_hi + s.h_down <= s.n_rows) &&
                        s.own_hi - s.own_lo >= (s.h_up > s.h_down ? s.h_up : s.h_down);
        if (!ok) { delete d; femol_set_error("femol_gmg_set_dist: inconsistent strip at level %d", l); return FEMOL_E_ARG; }
    }
    const Level& G = g->lv[n_dist];
    d->g_nx = coarse->lv[0].nx; d->g_ny = coarse->lv[0].ny;
    if (d->g_nx != G.nx || d->strip[n_dist].row0 + G.ny > d->g_ny || coarse_plan->n_elem != d->g_nx * d->g_ny) {
        delete d; femol_set_error("femol_gmg_set_dist: the coarse grid does not match level %d of the strip", n_dist); return FEMOL_E_ARG;
    }
    const int S2 = 16 * g->nd * g->nd;
    const int64_t gn = coarse->lv[0].n;
    if (cudaMalloc(&d->gke, (size_t)d->g_nx * d->g_ny * S2 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d->gvals, (size_t)coarse_plan->nnz * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d->gr, (size_t)gn * sizeof(double)) != cudaSuccess || cudaMalloc(&d->ge, (size_t)gn * sizeof(double)) != cudaSuccess) {
        cudaFree(d->gke); cudaFree(d->gvals); cudaFree(d->gr); cudaFree(d->ge); delete d;
        femol_set_error("femol_gmg_set_dist: out of device memory"); cudaGetLastError(); return FEMOL_E_CUDA;
    }
    g->dist = d;
    return FEMOL_OK;
}

// After femol_gmg_setup[_lattice] of the strip hierarchy: gather the element matrices of level n_dist, assemble that
// level's operator and rebuild the replicated hierarchy below it.  (The block-Jacobi inverses need no exchange: they are
// only used on owned rows.)
extern "C" int femol_gmg_dist_refresh(femol_gmg* g, void* stream) {
    FEMOL_REQUIRE(g && g->dist, "femol_gmg_dist_refresh: not a distributed hierarchy");
    cudaStream_t st = (cudaStream_t)stream;
    GmgDist* d = g->dist;
    const int S2 = 16 * g->nd * g->nd;
    const Level& L = g->lv[d->n_dist];
    const femol_strip& s = d->strip[d->n_dist];
    // element row J (between node rows J and J + 1) belongs to the owner of node row J
    const int64_t J0 = s.row0 + s.own_lo;                                       // global
    int64_t J1 = s.row0 + s.own_hi; if (J1 > d->g_ny) J1 = d->g_ny;
    const size_t total = (size_t)d->g_nx * d->g_ny * S2;
    FEMOL_CUDA_CHECK(cudaMemsetAsync(d->gke, 0, total * sizeof(double), st));
    if (J1 > J0)
        FEMOL_CUDA_CHECK(cudaMemcpy2DAsync(d->gke + (size_t)J0 * S2, (size_t)d->g_ny * S2 * sizeof(double),
                                           L.ke + (size_t)(J0 - s.row0) * S2, (size_t)L.ny * S2 * sizeof(double),
                                           (size_t)(J1 - J0) * S2 * sizeof(double), (size_t)d->g_nx, cudaMemcpyDeviceToDevice, st));
    int rc = femol_dist_allreduce(d->comm, d->gke, (int64_t)total, st);
    if (rc != FEMOL_OK) return rc;
    rc = femol_assemble(d->coarse_plan, d->gke, 1, nullptr, 0, nullptr, 1.0, d->coarse_fixed, nullptr, d->gvals, st);
    if (rc != FEMOL_OK) return rc;
    return gmg_setup_impl(d->coarse, d->gke, 1, nullptr, 0, nullptr, d->gvals, st);
}

// h_out[0] = doubles of the largest halo message, h_out[1] = doubles of the gathered level (window sizes of the P2P transport)
extern "C" int femol_gmg_dist_p2p_sizes(femol_gmg* g, int64_t* h_out) {
    FEMOL_REQUIRE(g && g->dist && h_out, "femol_gmg_dist_p2p_sizes: not a distributed hierarchy");
    int64_t halo = 0;
    for (int l = 0; l < g->dist->n_dist; ++l) {
        const femol_strip& s = g->dist->strip[l];
        const int64_t rows = s.h_up > s.h_down ? s.h_up : s.h_down;
        if (rows * s.row_nodes * g->nd > halo) halo = rows * s.row_nodes * g->nd;
    }
    h_out[0] = halo; h_out[1] = g->dist->coarse->lv[0].n;
    return FEMOL_OK;
}

extern "C" int femol_gmg_dist_exchange(femol_gmg* g, int level, double* x, int width, void* stream) {
    FEMOL_REQUIRE(g && g->dist && x && level >= 0 && level <= g->dist->n_dist && width >= 1, "femol_gmg_dist_exchange: bad argument");
    return femol_dist_exchange(g->dist->comm, x, width, g->dist->strip[level], (cudaStream_t)stream);
}

namespace {

template <int ND>
int dist_gather_solve(femol_gmg* g, cudaStream_t st) {
    GmgDist* d = g->dist;
    Level& L = g->lv[d->n_dist];
    const femol_strip& s = d->strip[d->n_dist];
    const int64_t rowd = s.row_nodes * ND;
    const int64_t gn = d->coarse->lv[0].n;
    int rc = femol_dist_allgather(d->comm, L.r + s.own_lo * rowd, (s.own_hi - s.own_lo) * rowd, (s.row0 + s.own_lo) * rowd, gn, d->gr, st);
    if (rc != FEMOL_OK) return rc;
    rc = vcycle<ND>(d->coarse, 0, d->gr, d->ge, st, false);
    if (rc != FEMOL_OK) return rc;
    // every local row of the correction, ghost rows included (the level is replicated: no exchange)
    FEMOL_CUDA_CHECK(cudaMemcpyAsync(L.e, d->ge + s.row0 * rowd, (size_t)s.n_rows * rowd * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return FEMOL_OK;
}

template <int ND>
int dvcycle(femol_gmg* g, int l, const double* r, double* e, cudaStream_t st, bool presmoothed) {
    GmgDist* d = g->dist;
    Level& L = g->lv[l];
    Level& C = g->lv[l + 1];
    const femol_strip& s = d->strip[l];
    int rc;
    if (!presmoothed) {
        jacobi_kernel<ND, true><<<grid_for(L.n_nodes), 256, 0, st>>>(L.n_nodes, g->omega, L.dinv, r, nullptr, e);
        FEMOL_LAUNCH_CHECK();
    }
    if ((rc = femol_dist_exchange(d->comm, e, ND, s, st)) != FEMOL_OK) return rc;
    const femol_epilogue res = {FEMOL_EPI_RESIDUAL, r, nullptr, 0.0, nullptr, nullptr};
    rc = L.vals32 ? femol_spmv_f32_epilogue_launch(L.plan, L.vals32, e, L.t, res, st)
                  : femol_apply_epilogue_launch(L.plan, L.vals, e, L.t, res, st);
    if (rc != FEMOL_OK) return rc;
    if ((rc = femol_dist_exchange(d->comm, L.t, ND, s, st)) != FEMOL_OK) return rc;
    const bool coarse_dist = l + 1 < d->n_dist;
    FEMOL_CUDA_CHECK(femol_launch_pdl(restrict_kernel<ND>, dim3(grid_for(C.n_nodes)), dim3(256), 0, st, L.nx, L.ny, C.nx, C.ny,
                                      (const uint8_t*)L.fixed, (const uint8_t*)C.fixed, (const double*)L.t, C.r,
                                      (const double*)(coarse_dist ? C.dinv : nullptr), g->omega, C.e, (const int32_t*)L.lat_node));
    femol_count_launch();
    if (coarse_dist) {
        if ((rc = dvcycle<ND>(g, l + 1, C.r, C.e, st, true)) != FEMOL_OK) return rc;
        if ((rc = femol_dist_exchange(d->comm, C.e, ND, d->strip[l + 1], st)) != FEMOL_OK) return rc;
    } else if ((rc = dist_gather_solve<ND>(g, st)) != FEMOL_OK) return rc;
    if (femol_apply_has_jacobi(L.plan, L.vals)) {
        FEMOL_CUDA_CHECK(femol_launch_pdl(prolong_add_kernel<ND>, dim3(grid_for(L.n_nodes)), dim3(256), 0, st, L.nx, L.ny, C.nx,
                                          (const uint8_t*)L.fixed, (const double*)C.e, (const double*)e, L.t, (const int32_t*)L.lat_node));
        femol_count_launch();
        // no exchange here: the prolongation is computed on every local row, and on the halo rows both of its inputs are
        // valid (e since the exchange before the residual product, e_c since the exchange above / the replicated solve)
        const femol_epilogue jac = {FEMOL_EPI_JACOBI, r, L.dinv, g->omega, nullptr, nullptr};
        return L.vals32 ? femol_spmv_f32_epilogue_launch(L.plan, L.vals32, L.t, e, jac, st)
                        : femol_apply_epilogue_launch(L.plan, L.vals, L.t, e, jac, st);
    }
    prolong_add_kernel<ND><<<grid_for(L.n_nodes), 256, 0, st>>>(L.nx, L.ny, C.nx, L.fixed, C.e, e, e, L.lat_node);
    FEMOL_LAUNCH_CHECK();
    return smooth<ND>(g, l, r, e, false, st);      // (halo rows of e + P e_c are valid: see above)
}

}  // namespace

int femol_gmg_dist_vcycle(femol_gmg* g, const double* r, double* z, cudaStream_t st, bool presmoothed) {
    return g->nd == 2 ? dvcycle<2>(g, 0, r, z, st, presmoothed) : dvcycle<6>(g, 0, r, z, st, presmoothed);
}
