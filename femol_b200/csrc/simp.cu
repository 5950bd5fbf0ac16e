// K7-K11: SIMP optimiser kernels — element sensitivities, mesh-independency filter, optimality-criteria
// update with device-side bisection, modal assurance criterion.
//
// Replaces (reference olivecha/FEMOL) the per-element Python loops of
//   SIMP_COMP._solid/_coating_compliance_objective_function   TOM_SIMP.py:126-226
//   SIMP_VIBE._solid/_coating_max_eigs_objective_function     TOM.py:408-495
//   _filter_sensibility (O(N^2) all-pairs search)             TOM_SIMP.py:228-277, TOM.py:305-354
//   _get_new_x (30 numpy passes of bisection)                 TOM_SIMP.py:279-313, TOM.py:251-303
//   utils.MAC                                                 utils.py:436-447
#include <cooperative_groups.h>
#include <math.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SENS_WARPS = 8;

// ---- K7 / K8: one warp per element, lanes split the size^2 entries of the quadratic form ---------
// MODE 0: compliance (c and dc), MODE 1: eigenvalue sensitivity
template <int ND, int NN, int MODE>
__global__ void __launch_bounds__(SENS_WARPS * 32)
sens_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ e0, int pe0,
            const double* __restrict__ e1, int pe1, const double* __restrict__ u, const double* __restrict__ x,
            double c0, double c1, double ex0, double ex1, const double* __restrict__ lambda,
            double* __restrict__ out, double* __restrict__ c_out, double* __restrict__ red_ws) {
    constexpr int S = NN * ND, S2 = S * S;
    __shared__ double s_u[SENS_WARPS][S];
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * SENS_WARPS + wid, nwarps = (int64_t)gridDim.x * SENS_WARPS;
    const double lam = (MODE == 1) ? lambda[0] : 0.0;
    double csum[1] = {0.0};
    for (int64_t e = warp0; e < n_elem; e += nwarps) {
        __syncwarp();
        for (int t = lane; t < S; t += 32) {
            const int a = t / ND, d = t - a * ND;
            s_u[wid][t] = __ldg(&u[(int64_t)__ldg(&conn[e * NN + a]) * ND + d]);
        }
        __syncwarp();
        const double* m0 = pe0 ? e0 + e * S2 : e0;
        const double* m1 = e1 ? (pe1 ? e1 + e * S2 : e1) : nullptr;
        double q0 = 0.0, q1 = 0.0;
        for (int idx = lane; idx < S2; idx += 32) {
            const int r = idx / S, c = idx - r * S;
            const double uu = s_u[wid][r] * s_u[wid][c];
            q0 += __ldg(&m0[idx]) * uu;
            if (m1) q1 += __ldg(&m1[idx]) * uu;
        }
        q0 = warp_sum(q0);
        if (m1) q1 = warp_sum(q1);
        if (lane == 0) {
            const double xe = x[e];
            if (MODE == 0) {
                // solid: c += x^ex0 q0, dc = -c0 x^ex1 q0 ; coating: c += q0 + x^ex0 q1, dc = -c0 x^ex1 q1
                const double qs = m1 ? q1 : q0;
                csum[0] += (m1 ? q0 : 0.0) + pow(xe, ex0) * qs;
                out[e] = -c0 * pow(xe, ex1) * qs;
            } else {
                // dl = p x^ek (v'Kv) - lambda q x^em (v'Mv)   (TOM.py:430-431)
                out[e] = c0 * pow(xe, ex0) * q0 - lam * c1 * pow(xe, ex1) * q1;
            }
        }
    }
    if (MODE == 0 && c_out) grid_sum_finish<1>(csum, red_ws, red_ticket(red_ws), c_out, s_red);
}

// Structured meshes (one element matrix for all elements): one THREAD per element, the matrices broadcast
// from shared memory, u_e in registers.  (The warp-per-element kernel above spends ~300 warp instructions
// per element on the two pow() calls of its lane 0 — 1.44 ms at 2 M elements in the first ncu capture.)
template <int ND, int NN, int MODE>
__global__ void __launch_bounds__(128)
sens_shared_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ e0,
                   const double* __restrict__ e1, const double* __restrict__ u, const double* __restrict__ x,
                   double c0, double c1, double ex0, double ex1, const double* __restrict__ lambda,
                   double* __restrict__ out, double* __restrict__ c_out, double* __restrict__ red_ws) {
    constexpr int S = NN * ND, S2 = S * S;
    extern __shared__ double s_e[];  // [S2] e0 | [S2] e1
    __shared__ double s_red[32];
    for (int i = threadIdx.x; i < S2; i += blockDim.x) { s_e[i] = e0[i]; if (e1) s_e[S2 + i] = e1[i]; }
    __syncthreads();
    const double lam = (MODE == 1) ? lambda[0] : 0.0;
    double csum[1] = {0.0};
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_elem; e += (int64_t)gridDim.x * blockDim.x) {
        double ue[S];
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const int64_t node = __ldg(&conn[e * NN + a]);
#pragma unroll
            for (int d = 0; d < ND; d += 2) {
                const double2 t = __ldg(reinterpret_cast<const double2*>(u + node * ND + d));
                ue[a * ND + d] = t.x; ue[a * ND + d + 1] = t.y;
            }
        }
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int r = 0; r < S; ++r) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int c = 0; c < S; ++c) {
                t0 += s_e[r * S + c] * ue[c];
                if (e1) t1 += s_e[S2 + r * S + c] * ue[c];
            }
            q0 += ue[r] * t0;
            q1 += ue[r] * t1;
        }
        const double xe = x[e];
        if (MODE == 0) {
            const double qs = e1 ? q1 : q0;
            csum[0] += (e1 ? q0 : 0.0) + pow(xe, ex0) * qs;
            out[e] = -c0 * pow(xe, ex1) * qs;
        } else {
            out[e] = c0 * pow(xe, ex0) * q0 - lam * c1 * pow(xe, ex1) * q1;
        }
    }
    if (MODE == 0 && c_out) grid_sum_finish<1>(csum, red_ws, red_ticket(red_ws), c_out, s_red);
}

template <int MODE>
int launch_sens(femol_plan_t p, const double* e0, int pe0, const double* e1, int pe1, const double* u, const double* x,
                double c0, double c1, double ex0, double ex1, const double* lambda, double* out, double* c_out,
                double* ws, cudaStream_t st) {
    if (!pe0 && (!e1 || !pe1) && p->nn * p->ndof <= 24) {
        int64_t blocks = (p->n_elem + 127) / 128;
        int64_t cap = (int64_t)p->sm_count * 12;
        if (cap > FEMOL_RED_MAX_BLOCKS) cap = FEMOL_RED_MAX_BLOCKS;
        if (blocks > cap) blocks = cap;
#define SENS_SH(ND, NN)                                                                                              \
        if (p->ndof == ND && p->nn == NN) {                                                                          \
            const size_t smem = 2 * (size_t)(ND * NN) * (ND * NN) * sizeof(double);                                  \
            sens_shared_kernel<ND, NN, MODE><<<(int)blocks, 128, smem, st>>>(p->n_elem, p->conn, e0, e1, u, x, c0, c1, \
                                                                              ex0, ex1, lambda, out, c_out, ws);      \
            FEMOL_LAUNCH_CHECK();                                                                                    \
            return FEMOL_OK;                                                                                         \
        }
        SENS_SH(2, 3) SENS_SH(2, 4) SENS_SH(2, 6) SENS_SH(2, 8) SENS_SH(6, 4)
#undef SENS_SH
    }
    int64_t blocks = (p->n_elem + SENS_WARPS - 1) / SENS_WARPS;
    int64_t cap = (int64_t)p->sm_count * 8;
    if (cap > FEMOL_RED_MAX_BLOCKS) cap = FEMOL_RED_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
#define SENS_CASE(ND, NN)                                                                                          \
    if (p->ndof == ND && p->nn == NN) {                                                                            \
        sens_kernel<ND, NN, MODE><<<(int)blocks, SENS_WARPS * 32, 0, st>>>(p->n_elem, p->conn, e0, pe0, e1, pe1, u, x, \
                                                                            c0, c1, ex0, ex1, lambda, out, c_out, ws); \
        FEMOL_LAUNCH_CHECK();                                                                                      \
        return FEMOL_OK;                                                                                           \
    }
    SENS_CASE(2, 3) SENS_CASE(2, 4) SENS_CASE(2, 6) SENS_CASE(2, 8)
    SENS_CASE(6, 3) SENS_CASE(6, 4) SENS_CASE(6, 6) SENS_CASE(6, 8)
#undef SENS_CASE
    femol_set_error("sensitivity kernel: unsupported (ndof=%d, nn=%d)", p->ndof, p->nn);
    return FEMOL_E_ARG;
}

// ---- K9: filter ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
filter_grid_kernel(int64_t n_slow, int64_t n_fast, int rx, int ry, const double* __restrict__ w,
                   const double* __restrict__ x, const double* __restrict__ dc, double* __restrict__ out) {
    extern __shared__ double s_w[];
    const int wy = 2 * ry + 1, nw = (2 * rx + 1) * wy;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
    const int64_t n = n_slow * n_fast;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / n_fast, j = e - i * n_fast;
        double num = 0.0, den = 0.0;
        for (int di = -rx; di <= rx; ++di) {
            const int64_t ii = i + di;
            if (ii < 0 || ii >= n_slow) continue;
            for (int dj = -ry; dj <= ry; ++dj) {
                const int64_t jj = j + dj;
                if (jj < 0 || jj >= n_fast) continue;
                const double ww = s_w[(di + rx) * wy + (dj + ry)];
                if (ww > 0.0) {
                    const int64_t f = ii * n_fast + jj;
                    num += ww * __ldg(&x[f]) * __ldg(&dc[f]);
                    den += ww;
                }
            }
        }
        out[e] = num / (x[e] * den);
    }
}

__global__ void __launch_bounds__(256)
filter_list_kernel(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                   const double* __restrict__ w, const double* __restrict__ x, const double* __restrict__ dc,
                   double* __restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        double num = 0.0, den = 0.0;
        for (int s = ptr[e]; s < ptr[e + 1]; ++s) {
            const int32_t f = idx[s];
            const double ww = w[s];
            num += ww * __ldg(&x[f]) * __ldg(&dc[f]);
            den += ww;
        }
        out[e] = num / (x[e] * den);
    }
}

// ---- K10: optimality criteria, persistent cooperative kernel ---------------------------------------
// numpy's np.min / np.max propagate NaN (TOM_SIMP.py:293-297); fmin/fmax do not.
__device__ __forceinline__ double nmin(double a, double b) { return (a != a || b != b) ? nan("") : fmin(a, b); }
__device__ __forceinline__ double nmax(double a, double b) { return (a != a || b != b) ? nan("") : fmax(a, b); }

constexpr int OC_THREADS = 512;

// Every block sums the same list of partials in the same order -> identical decisions grid wide.
__device__ __forceinline__ double sum_partials(const double* part, int nblk, double* smem) {
    double v[1] = {0.0};
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) v[0] += __ldcg(&part[b]);
    block_sum<1>(v, smem);
    __shared__ double s_bcast;
    if (threadIdx.x == 0) s_bcast = v[0];
    __syncthreads();
    return s_bcast;
}
__device__ __forceinline__ double max_partials(const double* part, int nblk, double* smem) {
    double m = -INFINITY;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) m = fmax(m, __ldcg(&part[b]));
    m = warp_max(m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = m;
    __syncthreads();
    __shared__ double s_bmax;
    if (threadIdx.x == 0) {
        double t = -INFINITY;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, smem[w]);
        s_bmax = t;
    }
    __syncthreads();
    return s_bmax;
}
__device__ __forceinline__ double block_max(double m, double* smem) {
    m = warp_max(m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = m;
    __syncthreads();
    double t = -INFINITY;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, smem[w]);
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(OC_THREADS)
oc_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ dc, int mode, double volfrac,
          const double* __restrict__ areas, const uint8_t* __restrict__ mask, double* __restrict__ x_new,
          double* __restrict__ scal, double* __restrict__ ws) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double smem[64];
    double* partA = ws;                              // alternating partial buffers
    double* partB = ws + FEMOL_RED_MAX_BLOCKS;
    double* partC = ws + 2 * FEMOL_RED_MAX_BLOCKS;
    const int nblk = gridDim.x;
    const int64_t tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const bool vibe = (mode & 1) != 0, strict = (mode & 2) != 0;
    const double move = vibe ? 0.3 : 0.5;

    // ---- phase 0: sign convention (COMP) / clamp + rescale (VIBE) ------------------------------
    // TOM_SIMP.py:289-292 switches the sign convention of the WHOLE vector when any(dc > 0); the other
    // sign then becomes NaN through the fractional power.  strict: exactly that.  Default: the sign of
    // the entry of largest magnitude decides (identical whenever dc is one-signed, which is every case
    // the reference survives) and entries of the opposite sign — rounding noise of u'Ku on elements in
    // near rigid-body motion, ~1e-11 on a 4M-dof mesh — count as zero sensitivity instead of NaN.
    double sign = 1.0, scale = 1.0;
    bool clamp_opposite = false;
    if (!vibe) {
        double mxp = 0.0, mxn = 0.0;
        for (int64_t e = tid0; e < n; e += nth) { const double g = dc[e]; mxp = fmax(mxp, g); mxn = fmax(mxn, -g); }
        const double bp = block_max(mxp, smem);
        const double bn = block_max(mxn, smem);
        if (threadIdx.x == 0) { __stcg(&partA[blockIdx.x], bp); __stcg(&partB[blockIdx.x], bn); }
        grid.sync();
        const double gp = max_partials(partA, nblk, smem), gn = max_partials(partB, nblk, smem);
        if (strict) sign = gp > 0.0 ? 1.0 : -1.0;
        else { sign = gp > gn ? 1.0 : -1.0; clamp_opposite = true; }
    } else {
        double s[1] = {0.0}, mx = -INFINITY;
        for (int64_t e = tid0; e < n; e += nth) {
            double g = dc[e];
            if (g < 0.0) { g = 0.0; dc[e] = 0.0; }                   // TOM.py:258
            s[0] += g; mx = fmax(mx, g);
        }
        block_sum<1>(s, smem);
        const double bm = block_max(mx, smem);
        if (threadIdx.x == 0) { __stcg(&partA[blockIdx.x], s[0]); __stcg(&partB[blockIdx.x], bm); }
        grid.sync();
        const double tot = sum_partials(partA, nblk, smem);
        const double gmx = max_partials(partB, nblk, smem);
        if (tot < 1.0) {                                             // TOM.py:260-262
            scale = 1.0 / gmx;
            for (int64_t e = tid0; e < n; e += nth) dc[e] *= scale;
        }
    }
    grid.sync();  // partA/partB are reused below

    // ---- phase 1: B_e = x_e * g_e^0.3 (stored in x_new) -----------------------------------------
    for (int64_t e = tid0; e < n; e += nth) {
        double g = sign * dc[e];
        if (clamp_opposite && g < 0.0) g = 0.0;
        x_new[e] = x[e] * pow(g, 0.3);
    }

    // ---- phase 2: bisection on the Lagrange multiplier -----------------------------------------
    double l1 = 0.0, l2 = 100000.0, lmid = 0.0;
    double total_area = 0.0;
    int it = 0;
    if (areas) {
        double s[1] = {0.0};
        for (int64_t e = tid0; e < n; e += nth) s[0] += areas[e];
        block_sum<1>(s, smem);
        if (threadIdx.x == 0) __stcg(&partC[blockIdx.x], s[0]);
        grid.sync();
        total_area = sum_partials(partC, nblk, smem);
    }
    double vsum = 0.0;
    while ((l2 - l1) > 1e-4) {
        lmid = 0.5 * (l1 + l2);
        const double sfac = pow(lmid, -0.3);
        double s[1] = {0.0};
        for (int64_t e = tid0; e < n; e += nth) {
            const double xe = x[e];
            double xn = nmin(xe + move, x_new[e] * sfac);
            xn = nmin(1.0, xn);
            xn = nmax(xe - move, xn);
            xn = nmax(0.001, xn);
            if (mask) { const uint8_t m = mask[e]; if (m == 1) xn = 1.0; else if (m == 2) xn = 0.001; }
            s[0] += areas ? xn * areas[e] : xn;
        }
        block_sum<1>(s, smem);
        double* part = (it & 1) ? partB : partA;
        if (threadIdx.x == 0) __stcg(&part[blockIdx.x], s[0]);
        grid.sync();
        vsum = sum_partials(part, nblk, smem);
        const bool up = areas ? (vsum / total_area > volfrac) : ((vsum - volfrac * (double)n) > 0.0);
        if (up) l1 = lmid; else l2 = lmid;
        ++it;
    }
    grid.sync();

    // ---- phase 3: final densities with the last evaluated multiplier, change = max|dx| ----------
    const double sfac = pow(lmid, -0.3);
    double chg = 0.0;
    for (int64_t e = tid0; e < n; e += nth) {
        const double xe = x[e];
        double xn = nmin(xe + move, x_new[e] * sfac);
        xn = nmin(1.0, xn);
        xn = nmax(xe - move, xn);
        xn = nmax(0.001, xn);
        if (mask) { const uint8_t m = mask[e]; if (m == 1) xn = 1.0; else if (m == 2) xn = 0.001; }
        x_new[e] = xn;
        const double d = fabs(xn - xe);
        chg = (d != d) ? INFINITY : fmax(chg, d);
    }
    const double bm = block_max(chg, smem);
    if (threadIdx.x == 0) __stcg(&partC[blockIdx.x], bm);
    grid.sync();
    const double gchg = max_partials(partC, nblk, smem);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[0] = gchg; scal[1] = vsum; scal[2] = lmid; scal[3] = (double)it;
    }
}

// ---- K11: MAC ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mac_accumulate(int64_t n, const double* __restrict__ V, const double* __restrict__ vref, double* __restrict__ acc /*[k][3]*/) {
    const int m = blockIdx.y;
    const double* v = V + (int64_t)m * n;
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = v[i], b = vref[i];
        s[0] += a * b; s[1] += a * a; s[2] += b * b;
    }
    __shared__ double smem[32 * 3];
    block_sum<3>(s, smem);
    if (threadIdx.x == 0) { atomicAdd(&acc[m * 3 + 0], s[0]); atomicAdd(&acc[m * 3 + 1], s[1]); atomicAdd(&acc[m * 3 + 2], s[2]); }
}
__global__ void mac_finish(int k, const double* __restrict__ acc, double* __restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < k) out[m] = (acc[m * 3] * acc[m * 3]) / (acc[m * 3 + 1] * acc[m * 3 + 2]);
}

__global__ void __launch_bounds__(256)
dot_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
           double* __restrict__ red_ws) {
    __shared__ double smem[32];
    double s[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s[0] += a[i] * b[i];
    grid_sum_finish<1>(s, red_ws, red_ticket(red_ws), out, smem);
}


// ---- centroid stresses of Q4 plane-stress elements (Mesh.stress_from_displacement, mesh.py:387-425) ---------------
// One thread per element: gather the 4 nodes, B(0,0) from the Jacobian at the centroid, S = T B u with
// T = T_base, T_base x^3 or T_base + T_coat x^3 (exponent hard-coded in the reference), and the reference's
// "von Mises" combination.  Structure-of-arrays output [4][n_elem] = Sx | Sy | Sxy | Sv.
__global__ void __launch_bounds__(256)
stress_q4_kernel(int64_t n_elem, const int32_t* __restrict__ conn, const double* __restrict__ xy,
                 const double* __restrict__ u, int ndof, const double* __restrict__ t_base,
                 const double* __restrict__ t_coat, const double* __restrict__ x, double* __restrict__ out) {
    __shared__ double s_t[18];
    if (threadIdx.x < 9) { s_t[threadIdx.x] = t_base[threadIdx.x]; s_t[9 + threadIdx.x] = t_coat ? t_coat[threadIdx.x] : 0.0; }
    __syncthreads();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_elem; e += (int64_t)gridDim.x * blockDim.x) {
        const int4 c = __ldg(reinterpret_cast<const int4*>(conn) + e);
        const int nodes[4] = {c.x, c.y, c.z, c.w};
        double X[4], Y[4], ux[4], uy[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double2 p = __ldg(reinterpret_cast<const double2*>(xy) + nodes[a]);
            X[a] = p.x; Y[a] = p.y;
            const double2 d = __ldg(reinterpret_cast<const double2*>(u + (int64_t)nodes[a] * ndof));
            ux[a] = d.x; uy[a] = d.y;
        }
        // dN/dxi = (-1, 1, 1, -1)/4, dN/deta = (-1, -1, 1, 1)/4 at the centroid
        const double j00 = 0.25 * (-X[0] + X[1] + X[2] - X[3]), j01 = 0.25 * (-Y[0] + Y[1] + Y[2] - Y[3]);
        const double j10 = 0.25 * (-X[0] - X[1] + X[2] + X[3]), j11 = 0.25 * (-Y[0] - Y[1] + Y[2] + Y[3]);
        const double idet = 1.0 / (j00 * j11 - j01 * j10);
        const double dxi[4] = {-0.25, 0.25, 0.25, -0.25}, det[4] = {-0.25, -0.25, 0.25, 0.25};
        double exx = 0.0, eyy = 0.0, gxy = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double dx = idet * (j11 * dxi[a] - j01 * det[a]);
            const double dy = idet * (-j10 * dxi[a] + j00 * det[a]);
            exx += dx * ux[a]; eyy += dy * uy[a]; gxy += dy * ux[a] + dx * uy[a];
        }
        double w = 1.0;
        if (x) { const double xe = __ldg(&x[e]); w = xe * xe * xe; }
        double S[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double sb = s_t[3 * r] * exx + s_t[3 * r + 1] * eyy + s_t[3 * r + 2] * gxy;
            const double sc = s_t[9 + 3 * r] * exx + s_t[9 + 3 * r + 1] * eyy + s_t[9 + 3 * r + 2] * gxy;
            S[r] = t_coat ? sb + w * sc : w * sb;
        }
        const double a0 = S[0] - S[1], a1 = S[1] - S[2], a2 = S[2] - S[0];
        out[e] = S[0]; out[n_elem + e] = S[1]; out[2 * n_elem + e] = S[2];
        out[3 * n_elem + e] = sqrt(0.5 * (a0 * a0 + a1 * a1 + a2 * a2));
    }
}

}  // namespace

extern "C" int femol_sens_compliance(femol_plan_t p, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                                     const double* u, const double* x, double pexp, double pc, double pd, double* dc_out,
                                     double* c_out, double* ws, void* stream) {
    FEMOL_REQUIRE(p && e_base && u && x && dc_out && ws, "femol_sens_compliance: null argument");
    return launch_sens<0>(p, e_base, base_pe, e_coat, coat_pe, u, x, pexp, 0.0, pc, pd, nullptr, dc_out, c_out, ws,
                          (cudaStream_t)stream);
}

extern "C" int femol_sens_eig(femol_plan_t p, const double* ke, int ke_pe, const double* me, int me_pe, const double* v,
                              const double* x, double pexp, double qexp, double ek, double em, const double* lambda,
                              double* dl_out, void* stream) {
    FEMOL_REQUIRE(p && ke && me && v && x && lambda && dl_out, "femol_sens_eig: null argument");
    return launch_sens<1>(p, ke, ke_pe, me, me_pe, v, x, pexp, qexp, ek, em, lambda, dl_out, nullptr, nullptr,
                          (cudaStream_t)stream);
}

extern "C" int femol_filter_grid(int64_t n_slow, int64_t n_fast, int rx, int ry, const double* w, const double* x,
                                 const double* dc, double* out, void* stream) {
    FEMOL_REQUIRE(w && x && dc && out && n_slow > 0 && n_fast > 0 && rx >= 0 && ry >= 0, "femol_filter_grid: bad argument");
    FEMOL_REQUIRE(out != dc, "femol_filter_grid: out must not alias dc");
    const size_t smem = (size_t)(2 * rx + 1) * (2 * ry + 1) * sizeof(double);
    FEMOL_REQUIRE(smem <= 40 * 1024, "femol_filter_grid: filter radius too large");
    const int grid = femol_grid_for(n_slow * n_fast, 256, femol_sm_count(), 16);
    filter_grid_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(n_slow, n_fast, rx, ry, w, x, dc, out);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_filter_list(int64_t n_elem, const int32_t* nbr_ptr, const int32_t* nbr_idx, const double* nbr_w,
                                 const double* x, const double* dc, double* out, void* stream) {
    FEMOL_REQUIRE(nbr_ptr && nbr_idx && nbr_w && x && dc && out && n_elem > 0, "femol_filter_list: bad argument");
    FEMOL_REQUIRE(out != dc, "femol_filter_list: out must not alias dc");
    const int grid = femol_grid_for(n_elem, 256, femol_sm_count(), 16);
    filter_list_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_elem, nbr_ptr, nbr_idx, nbr_w, x, dc, out);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_oc_update(int64_t n_elem, const double* x, double* dc, int mode, double volfrac, const double* areas,
                               const uint8_t* mask, double* x_new, double* scalars_out, double* ws, void* stream) {
    FEMOL_REQUIRE(x && dc && x_new && scalars_out && ws && n_elem > 0, "femol_oc_update: bad argument");
    FEMOL_REQUIRE(mode >= 0 && mode <= 2, "femol_oc_update: mode must be 0 (COMP), 1 (VIBE) or 2 (COMP, strict sign rule)");
    int per_sm = 0;
    FEMOL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, oc_kernel, OC_THREADS, 0));
    FEMOL_REQUIRE(per_sm > 0, "femol_oc_update: kernel does not fit on an SM");
    int64_t grid = (int64_t)femol_sm_count() * (per_sm > 2 ? 2 : per_sm);
    const int64_t need = (n_elem + OC_THREADS - 1) / OC_THREADS;
    if (grid > need) grid = need;
    if (grid > FEMOL_RED_MAX_BLOCKS) grid = FEMOL_RED_MAX_BLOCKS;
    void* args[] = {(void*)&n_elem, (void*)&x, (void*)&dc, (void*)&mode, (void*)&volfrac, (void*)&areas,
                    (void*)&mask, (void*)&x_new, (void*)&scalars_out, (void*)&ws};
    FEMOL_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)oc_kernel, dim3((unsigned)grid), dim3(OC_THREADS), args, 0,
                                                 (cudaStream_t)stream));
    return FEMOL_OK;
}

extern "C" int femol_mac(int64_t n, int k, const double* V, const double* v_ref, double* mac_out, double* ws, void* stream) {
    FEMOL_REQUIRE(V && v_ref && mac_out && ws && n > 0 && k > 0, "femol_mac: bad argument");
    FEMOL_REQUIRE(3 * k <= FEMOL_RED_WS_DOUBLES, "femol_mac: too many modes");
    cudaStream_t st = (cudaStream_t)stream;
    FEMOL_CUDA_CHECK(cudaMemsetAsync(ws, 0, 3 * k * sizeof(double), st));
    int gx = femol_grid_for(n, 256, femol_sm_count(), 2);
    mac_accumulate<<<dim3(gx, k), 256, 0, st>>>(n, V, v_ref, ws);
    FEMOL_LAUNCH_CHECK();
    mac_finish<<<(k + 63) / 64, 64, 0, st>>>(k, ws, mac_out);
    FEMOL_LAUNCH_CHECK();
    // leave the reduction scratch clean for the ticket-based reducers
    FEMOL_CUDA_CHECK(cudaMemsetAsync(ws, 0, 3 * k * sizeof(double), st));
    return FEMOL_OK;
}

extern "C" int femol_dot(int64_t n, const double* a, const double* b, double* out, double* ws, void* stream) {
    FEMOL_REQUIRE(a && b && out && ws && n > 0, "femol_dot: bad argument");
    const int grid = femol_grid_for(n, 256, femol_sm_count(), 4);
    dot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, a, b, out, ws);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

extern "C" int femol_stress_q4(femol_plan_t p, const double* xy, const double* u, const double* t_base, const double* t_coat,
                               const double* x, double* out, void* stream) {
    FEMOL_REQUIRE(p && xy && u && t_base && out, "femol_stress_q4: null argument");
    FEMOL_REQUIRE(p->nn == 4 && p->ndof >= 2 && p->ndof % 2 == 0, "femol_stress_q4: 4-node elements with an even number of dofs per node only");
    if (p->n_elem == 0) return FEMOL_OK;
    const int grid = femol_grid_for(p->n_elem, 256, p->sm_count, 8);
    stress_q4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p->n_elem, p->conn, xy, u, p->ndof, t_base, t_coat, x, out);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}
