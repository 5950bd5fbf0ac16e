// K1 / K2: batched element stiffness and mass matrices, FP64 Gauss quadrature, one warp per element.
//
// Replaces (reference olivecha/FEMOL): T3.Ke elements.py:218-239, T6.Ke :560-621, Q4.Ke :987-1072,
// Q8 (stub :1351, standard serendipity element built from its shape tables :1183-1233 — parity
// unpinned), and the Me methods :241-274, :623-683, :1074-1142.
//
// No tensor cores: the contractions are (8 x S)^T (8 x 8) (8 x S) with S <= 48 per Gauss point, far
// below an MMA tile, and parity demands FP64.  The warp keeps the S x S accumulator distributed
// over its lanes' registers, stages the strain matrix B and C*B in shared memory, and writes the
// element matrix with fully coalesced 256-byte stores.
#include "common.cuh"

namespace {

template <int ET> struct ElemTraits;
template <> struct ElemTraits<FEMOL_T3> { static constexpr int NN = 3; };
template <> struct ElemTraits<FEMOL_Q4> { static constexpr int NN = 4; };
template <> struct ElemTraits<FEMOL_T6> { static constexpr int NN = 6; };
template <> struct ElemTraits<FEMOL_Q8> { static constexpr int NN = 8; };

struct QPoint { double xi, eta, w; int shear; };  // shear = 1: point of the transverse-shear rule

#define SQ3_3 0.57735026918962576451  // sqrt(3)/3
#define SQ_06 0.77459666924148337704  // sqrt(3/5)

// Quadrature tables (elements.py:50-59, 279-288, 694-712, 1164-1180)
template <int ET, int ND> struct Rule;
template <int ND> struct Rule<FEMOL_T3, ND> {
    static constexpr int N = 1;
    __device__ static QPoint get(int) { return {1.0 / 3, 1.0 / 3, 0.5, 0}; }  // area * B'CB
};
template <> struct Rule<FEMOL_Q4, 2> {
    static constexpr int N = 4;
    __device__ static QPoint get(int g) {
        const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
        return {sx[g] * SQ3_3, sy[g] * SQ3_3, 1.0, 0};
    }
};
template <> struct Rule<FEMOL_Q4, 6> {
    static constexpr int N = 5;  // 2x2 membrane+bending, 1-point shear with weight 4 (elements.py:1027-1033)
    __device__ static QPoint get(int g) {
        const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
        if (g < 4) return {sx[g] * SQ3_3, sy[g] * SQ3_3, 1.0, 0};
        return {0.0, 0.0, 4.0, 1};
    }
};
__device__ inline QPoint tri3_point(int g) {
    const double px[4] = {1.0 / 3, 1.0 / 5, 3.0 / 5, 1.0 / 5}, py[4] = {1.0 / 3, 1.0 / 5, 1.0 / 5, 3.0 / 5};
    const double w[4] = {-27.0 / 96, 25.0 / 96, 25.0 / 96, 25.0 / 96};
    return {px[g], py[g], w[g], 0};
}
__device__ inline QPoint tri2_point(int g, int shear) {
    const double px[3] = {2.0 / 3, 1.0 / 6, 1.0 / 6}, py[3] = {1.0 / 6, 1.0 / 6, 2.0 / 3};
    return {px[g], py[g], 1.0 / 3, shear};
}
template <> struct Rule<FEMOL_T6, 2> {
    static constexpr int N = 4;
    __device__ static QPoint get(int g) { return tri3_point(g); }
};
template <> struct Rule<FEMOL_T6, 6> {
    static constexpr int N = 7;  // 4-point membrane+bending, 3-point shear with sum-1 weights (elements.py:613-617)
    __device__ static QPoint get(int g) { return g < 4 ? tri3_point(g) : tri2_point(g - 4, 1); }
};
__device__ inline QPoint q8_point(int g) {
    const double c[3] = {-SQ_06, 0.0, SQ_06}, w[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};
    return {c[g % 3], c[g / 3], w[g % 3] * w[g / 3], 0};
}
template <> struct Rule<FEMOL_Q8, 2> {
    static constexpr int N = 9;
    __device__ static QPoint get(int g) { return q8_point(g); }
};
template <> struct Rule<FEMOL_Q8, 6> {
    static constexpr int N = 13;  // 3x3 membrane+bending, 2x2 reduced shear (unpinned)
    __device__ static QPoint get(int g) {
        if (g < 9) return q8_point(g);
        const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
        return {sx[g - 9] * SQ3_3, sy[g - 9] * SQ3_3, 1.0, 1};
    }
};

// Shape functions and parametric derivatives (elements.py:63-77, 291-330, 715-747, 1183-1233)
template <int ET>
__device__ __forceinline__ void shape(double xi, double eta, double* N, double* dxi, double* deta);

template <>
__device__ __forceinline__ void shape<FEMOL_T3>(double xi, double eta, double* N, double* dxi, double* deta) {
    N[0] = 1 - xi - eta; N[1] = xi; N[2] = eta;
    dxi[0] = -1; dxi[1] = 1; dxi[2] = 0;
    deta[0] = -1; deta[1] = 0; deta[2] = 1;
}
template <>
__device__ __forceinline__ void shape<FEMOL_Q4>(double xi, double eta, double* N, double* dxi, double* deta) {
    N[0] = 0.25 * (1 - xi) * (1 - eta); N[1] = 0.25 * (1 + xi) * (1 - eta);
    N[2] = 0.25 * (1 + xi) * (1 + eta); N[3] = 0.25 * (1 - xi) * (1 + eta);
    dxi[0] = 0.25 * (eta - 1); dxi[1] = 0.25 * (1 - eta); dxi[2] = 0.25 * (eta + 1); dxi[3] = 0.25 * (-eta - 1);
    deta[0] = 0.25 * (xi - 1); deta[1] = 0.25 * (-xi - 1); deta[2] = 0.25 * (xi + 1); deta[3] = 0.25 * (1 - xi);
}
template <>
__device__ __forceinline__ void shape<FEMOL_T6>(double xi, double eta, double* N, double* dxi, double* deta) {
    const double lm = 1 - xi - eta;
    N[0] = -lm * (1 - 2 * lm); N[1] = 4 * xi * lm; N[2] = -xi * (1 - 2 * xi);
    N[3] = 4 * xi * eta; N[4] = -eta * (1 - 2 * eta); N[5] = 4 * eta * lm;
    dxi[0] = 1 - 4 * lm; dxi[1] = 4 * (lm - xi); dxi[2] = -1 + 4 * xi; dxi[3] = 4 * eta; dxi[4] = 0; dxi[5] = -4 * eta;
    deta[0] = 1 - 4 * lm; deta[1] = -4 * xi; deta[2] = 0; deta[3] = 4 * xi; deta[4] = -1 + 4 * eta; deta[5] = 4 * (lm - eta);
}
template <>
__device__ __forceinline__ void shape<FEMOL_Q8>(double xi, double eta, double* N, double* dxi, double* deta) {
    N[0] = 0.25 * (1 - xi) * (1 - eta) * (-xi - eta - 1);
    N[1] = 0.25 * (1 + xi) * (1 - eta) * (xi - eta - 1);
    N[2] = 0.25 * (1 + xi) * (1 + eta) * (xi + eta - 1);
    N[3] = 0.25 * (1 - xi) * (1 + eta) * (-xi + eta - 1);
    N[4] = 0.50 * (1 - eta) * (1 + xi) * (1 - xi);
    N[5] = 0.50 * (1 + xi) * (1 + eta) * (1 - eta);
    N[6] = 0.50 * (1 + eta) * (1 + xi) * (1 - xi);
    N[7] = 0.50 * (1 - xi) * (1 + eta) * (1 - eta);
    dxi[0] = 0.25 * (1 - eta) * (2 * xi + eta);
    dxi[1] = 0.25 * (1 - eta) * (2 * xi - eta);
    dxi[2] = 0.25 * (1 + eta) * (2 * xi + eta);
    dxi[3] = 0.25 * (1 + eta) * (2 * xi - eta);
    dxi[4] = -xi * (1 - eta);
    dxi[5] = 0.5 * (1 - eta) * (1 + eta);
    dxi[6] = -xi * (1 + eta);
    dxi[7] = -0.5 * (1 - eta) * (1 + eta);
    deta[0] = 0.25 * (1 - xi) * (xi + 2 * eta);
    deta[1] = 0.25 * (1 + xi) * (2 * eta - xi);
    deta[2] = 0.25 * (1 + xi) * (xi + 2 * eta);
    deta[3] = 0.25 * (1 - xi) * (2 * eta - xi);
    deta[4] = -0.5 * (1 - xi) * (1 + xi);
    deta[5] = -eta * (1 + xi);
    deta[6] = 0.5 * (1 - xi) * (1 + xi);
    deta[7] = -eta * (1 - xi);
}

// Entry (row, node dof da) of the strain matrix B (elements.py:193-216 plane; 520-558, 947-977 plate).
// Rows 0-5: membrane + bending (Bp + Bb), rows 6-7: transverse shear (Bs).
template <int ND>
__device__ __forceinline__ double b_entry(int row, int da, double dNdx, double dNdy, double N, int shear) {
    if (ND == 2) {
        if (row == 0) return da == 0 ? dNdx : 0.0;
        if (row == 1) return da == 1 ? dNdy : 0.0;
        return da == 0 ? dNdy : dNdx;
    }
    if (!shear) {
        switch (row) {
            case 0: return da == 0 ? dNdx : 0.0;
            case 1: return da == 1 ? dNdy : 0.0;
            case 2: return da == 0 ? dNdy : (da == 1 ? dNdx : 0.0);
            case 3: return da == 3 ? -dNdx : 0.0;
            case 4: return da == 4 ? -dNdy : 0.0;
            case 5: return da == 4 ? -dNdx : (da == 3 ? -dNdy : 0.0);
            default: return 0.0;
        }
    }
    if (row == 6) return da == 2 ? dNdx : (da == 3 ? -N : 0.0);
    if (row == 7) return da == 2 ? dNdy : (da == 4 ? -N : 0.0);
    return 0.0;
}

constexpr int KE_WARPS = 4;

template <int ET, int ND>
__global__ void __launch_bounds__(KE_WARPS * 32)
ke_batch_kernel(int64_t n_elem, const double* __restrict__ xy, const int32_t* __restrict__ conn,
                const double* __restrict__ tens, double* __restrict__ out, int tens_stride, int accumulate) {
    constexpr int NN = ElemTraits<ET>::NN;
    constexpr int S = NN * ND;
    constexpr int NR = ND == 2 ? 3 : 8;
    constexpr int NE = (S * S + 31) / 32;
    // tens_stride = 0: one constitutive block for the whole batch (sCw[0]); 31: one per element, filled by its warp
    __shared__ double sCw[KE_WARPS][NR * NR];
    __shared__ double sB[KE_WARPS][NR * S];
    __shared__ double sCB[KE_WARPS][NR * S];
    __shared__ double sG[KE_WARPS][3][8];  // dN/dx, dN/dy, N at the current point
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    // constitutive block tensor: plane = A; plate = A + Bc + D (rows 0-5) and G (rows 6-7)
    auto block_entry = [](const double* __restrict__ t, int i) {
        const int r = i / NR, c = i % NR;
        double v = 0.0;
        if (ND == 2) {
            v = t[r * 3 + c];
        } else {
            if (r < 3 && c < 3) v = t[r * 3 + c];                          // A
            else if (r >= 3 && r < 6 && c >= 3 && c < 6) v = t[9 + (r - 3) * 3 + (c - 3)];  // D
            else if (r >= 6 && c >= 6) v = t[18 + (r - 6) * 2 + (c - 6)];  // G
            // membrane-bending coupling B (elements.py:1038-1039); the reference's T6 plate ignores it
            // (elements.py:587-621)
            else if (ET != FEMOL_T6 && r < 3 && c >= 3 && c < 6) v = t[22 + r * 3 + (c - 3)];
            else if (ET != FEMOL_T6 && r >= 3 && r < 6 && c < 3) v = t[22 + (r - 3) * 3 + c];
        }
        return v;
    };
    if (tens_stride == 0)
        for (int i = threadIdx.x; i < NR * NR; i += blockDim.x) sCw[0][i] = block_entry(tens, i);
    __syncthreads();

    const int64_t warp0 = (int64_t)blockIdx.x * KE_WARPS + wid;
    const int64_t nwarps = (int64_t)gridDim.x * KE_WARPS;
    for (int64_t e = warp0; e < n_elem; e += nwarps) {
        double X[NN], Y[NN];
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const int32_t node = __ldg(&conn[e * NN + a]);
            X[a] = __ldg(&xy[2 * (int64_t)node]);
            Y[a] = __ldg(&xy[2 * (int64_t)node + 1]);
        }
        double acc[NE];
#pragma unroll
        for (int m = 0; m < NE; ++m) acc[m] = 0.0;
        const double* sC = sCw[0];
        if (tens_stride) {   // this element's own tensors (variable-core laminates, FEM.py:697-731)
            __syncwarp();
            for (int i = lane; i < NR * NR; i += 32) sCw[wid][i] = block_entry(tens + e * tens_stride, i);
            __syncwarp();
            sC = sCw[wid];
        }

        for (int g = 0; g < Rule<ET, ND>::N; ++g) {
            const QPoint q = Rule<ET, ND>::get(g);
            double N[NN], dxi[NN], deta[NN];
            shape<ET>(q.xi, q.eta, N, dxi, deta);
            double J00 = 0, J01 = 0, J10 = 0, J11 = 0;  // [[dx/dxi, dy/dxi],[dx/deta, dy/deta]]
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                J00 += dxi[a] * X[a]; J01 += dxi[a] * Y[a];
                J10 += deta[a] * X[a]; J11 += deta[a] * Y[a];
            }
            const double det = J00 * J11 - J10 * J01;
            const double idet = 1.0 / det;
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    sG[wid][0][a] = idet * (J11 * dxi[a] - J01 * deta[a]);
                    sG[wid][1][a] = idet * (-J10 * dxi[a] + J00 * deta[a]);
                    sG[wid][2][a] = N[a];
                }
            }
            __syncwarp();
            for (int i = lane; i < NR * S; i += 32) {
                const int row = i / S, col = i % S, a = col / ND, da = col % ND;
                sB[wid][i] = b_entry<ND>(row, da, sG[wid][0][a], sG[wid][1][a], sG[wid][2][a], q.shear);
            }
            __syncwarp();
            for (int i = lane; i < NR * S; i += 32) {
                const int row = i / S, col = i % S;
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < NR; ++m) v += sC[row * NR + m] * sB[wid][m * S + col];
                sCB[wid][i] = v;
            }
            __syncwarp();
            const double wd = q.w * det;
#pragma unroll
            for (int m = 0; m < NE; ++m) {
                const int idx = lane + 32 * m;
                if (idx < S * S) {
                    const int r = idx / S, c = idx % S;
                    double v = 0.0;
#pragma unroll
                    for (int k = 0; k < NR; ++k) v += sB[wid][k * S + r] * sCB[wid][k * S + c];
                    acc[m] += wd * v;
                }
            }
        }
        double* o = out + e * (int64_t)(S * S);
#pragma unroll
        for (int m = 0; m < NE; ++m) {
            const int idx = lane + 32 * m;
            if (idx < S * S) {
                double v = acc[m];
                if (ND == 6) {  // drilling dof: diagonal *set* to 1 per element (elements.py:1071, 620)
                    const int r = idx / S, c = idx % S;
                    if (r == c && (r % 6) == 5) v = 1.0;
                }
                o[idx] = accumulate ? o[idx] + v : v;
            }
        }
    }
}

// 4x4 solve A^T n = b with partial pivoting (N = mono * inv(P): Q4.make_shape_xy, elements.py:792-825)
__device__ inline void solve4(double A[4][4], double b[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int p = k;
        double best = fabs(A[k][k]);
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            if (fabs(A[r][k]) > best) { best = fabs(A[r][k]); p = r; }
        }
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            if (p == r) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { double t = A[k][c]; A[k][c] = A[r][c]; A[r][c] = t; }
                double t = b[k]; b[k] = b[r]; b[r] = t;
            }
        }
        const double ip = 1.0 / A[k][k];
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            const double f = A[r][k] * ip;
#pragma unroll
            for (int c = k; c < 4; ++c) A[r][c] -= f * A[k][c];
            b[r] -= f * b[k];
        }
    }
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        double s = b[k];
#pragma unroll
        for (int c = k + 1; c < 4; ++c) s -= A[k][c] * b[c];
        b[k] = s / A[k][k];
    }
}

__device__ __forceinline__ double line2(double xa, double ya, double xb, double yb, double X, double Y) {
    return (xa - xb) * (Y - yb) - (ya - yb) * (X - xb);
}

// Physical-coordinate shape functions used by the reference mass matrices
template <int ET>
__device__ inline void shape_xy(const double* X, const double* Y, double px, double py, const double* Niso,
                                double* N) {
    constexpr int NN = ElemTraits<ET>::NN;
    if (ET == FEMOL_T3) {  // elements.py:170-191
        const double xi = X[0], xj = X[1], xk = X[2], yi = Y[0], yj = Y[1], yk = Y[2];
        const double A2 = xi * (yj - yk) + xj * (yk - yi) + xk * (yi - yj);
        N[0] = 1.0 / A2 * ((yk - yj) * (xj - px) - (xk - xj) * (yj - py));
        N[1] = 1.0 / A2 * ((yi - yk) * (xk - px) - (xi - xk) * (yk - py));
        N[2] = 1.0 / A2 * ((yj - yi) * (xi - px) - (xj - xi) * (yi - py));
    } else if (ET == FEMOL_Q4) {  // elements.py:792-825
        double A[4][4], b[4] = {1.0, px, py, px * py};
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // A = P^T, P rows [1, x, y, xy]
            A[0][k] = 1.0; A[1][k] = X[k]; A[2][k] = Y[k]; A[3][k] = X[k] * Y[k];
        }
        solve4(A, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) N[k] = b[k];
    } else if (ET == FEMOL_T6) {  // elements.py:342-402, node order [N1, N4, N2, N5, N3, N6]
        const double x1 = X[0], x2 = X[2], x3 = X[4], x4 = X[1], x5 = X[3], x6 = X[5];
        const double y1 = Y[0], y2 = Y[2], y3 = Y[4], y4 = Y[1], y5 = Y[3], y6 = Y[5];
#define PRODN(xa, ya, xb, yb, xc, yc, xd, yd, xn, yn)                                         \
    ((line2(xa, ya, xb, yb, px, py) * line2(xc, yc, xd, yd, px, py)) /                        \
     (line2(xa, ya, xb, yb, xn, yn) * line2(xc, yc, xd, yd, xn, yn)))
        N[0] = PRODN(x2, y2, x3, y3, x4, y4, x6, y6, x1, y1);
        N[2] = PRODN(x3, y3, x1, y1, x5, y5, x4, y4, x2, y2);
        N[4] = PRODN(x2, y2, x1, y1, x5, y5, x6, y6, x3, y3);
        N[1] = PRODN(x3, y3, x1, y1, x2, y2, x3, y3, x4, y4);
        N[3] = PRODN(x3, y3, x1, y1, x2, y2, x1, y1, x5, y5);
        N[5] = PRODN(x2, y2, x1, y1, x2, y2, x3, y3, x6, y6);
#undef PRODN
    } else {  // Q8: isoparametric (no reference implementation)
#pragma unroll
        for (int a = 0; a < NN; ++a) N[a] = Niso[a];
    }
}

template <int ET> struct MassRule;
template <> struct MassRule<FEMOL_T3> {
    static constexpr int N = 4;
    __device__ static QPoint get(int g) { return tri3_point(g); }
};
template <> struct MassRule<FEMOL_T6> {
    static constexpr int N = 3;
    __device__ static QPoint get(int g) { return tri2_point(g, 0); }
};
template <> struct MassRule<FEMOL_Q4> {
    static constexpr int N = 4;
    __device__ static QPoint get(int g) { return Rule<FEMOL_Q4, 2>::get(g); }
};
template <> struct MassRule<FEMOL_Q8> {
    static constexpr int N = 9;
    __device__ static QPoint get(int g) { return q8_point(g); }
};

template <int ET, int ND>
__global__ void __launch_bounds__(KE_WARPS * 32)
me_batch_kernel(int64_t n_elem, const double* __restrict__ xy, const int32_t* __restrict__ conn,
                const double* __restrict__ V, double* __restrict__ out) {
    constexpr int NN = ElemTraits<ET>::NN;
    constexpr int S = NN * ND;
    constexpr int NE = (S * S + 31) / 32;
    constexpr int NG = MassRule<ET>::N;
    __shared__ double sV[ND * ND];
    __shared__ double sN[KE_WARPS][NG][8];
    __shared__ double sW[KE_WARPS][NG];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ND * ND; i += blockDim.x) sV[i] = V[i];
    __syncthreads();

    const int64_t warp0 = (int64_t)blockIdx.x * KE_WARPS + wid;
    const int64_t nwarps = (int64_t)gridDim.x * KE_WARPS;
    for (int64_t e = warp0; e < n_elem; e += nwarps) {
        double X[NN], Y[NN];
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const int32_t node = __ldg(&conn[e * NN + a]);
            X[a] = __ldg(&xy[2 * (int64_t)node]);
            Y[a] = __ldg(&xy[2 * (int64_t)node + 1]);
        }
        __syncwarp();
        if (lane < NG) {  // one lane per Gauss point
            const QPoint q = MassRule<ET>::get(lane);
            double N[NN], dxi[NN], deta[NN], Np[NN];
            shape<ET>(q.xi, q.eta, N, dxi, deta);
            double J00 = 0, J01 = 0, J10 = 0, J11 = 0, px = 0, py = 0;
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                J00 += dxi[a] * X[a]; J01 += dxi[a] * Y[a];
                J10 += deta[a] * X[a]; J11 += deta[a] * Y[a];
                px += N[a] * X[a]; py += N[a] * Y[a];
            }
            shape_xy<ET>(X, Y, px, py, N, Np);
#pragma unroll
            for (int a = 0; a < NN; ++a) sN[wid][lane][a] = Np[a];
            sW[wid][lane] = q.w * (J00 * J11 - J10 * J01);
        }
        __syncwarp();
        double* o = out + e * (int64_t)(S * S);
#pragma unroll
        for (int m = 0; m < NE; ++m) {
            const int idx = lane + 32 * m;
            if (idx < S * S) {
                const int r = idx / S, c = idx % S;
                const int a = r / ND, da = r % ND, b = c / ND, db = c % ND;
                double v = 0.0;
#pragma unroll
                for (int g = 0; g < NG; ++g) v += sW[wid][g] * (sN[wid][g][a] * sV[da * ND + db] * sN[wid][g][b]);
                o[idx] = v;
            }
        }
    }
}

template <int ET, int ND>
int launch_ke(int64_t n_elem, const double* xy, const int32_t* conn, const double* tens, double* out,
              cudaStream_t st, int tens_stride = 0, int accumulate = 0) {
    int64_t blocks = (n_elem + KE_WARPS - 1) / KE_WARPS;
    const int64_t cap = (int64_t)femol_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    ke_batch_kernel<ET, ND><<<(int)blocks, KE_WARPS * 32, 0, st>>>(n_elem, xy, conn, tens, out, tens_stride, accumulate);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}
template <int ET, int ND>
int launch_me(int64_t n_elem, const double* xy, const int32_t* conn, const double* V, double* out,
              cudaStream_t st) {
    int64_t blocks = (n_elem + KE_WARPS - 1) / KE_WARPS;
    const int64_t cap = (int64_t)femol_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    me_batch_kernel<ET, ND><<<(int)blocks, KE_WARPS * 32, 0, st>>>(n_elem, xy, conn, V, out);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

}  // namespace

extern "C" int femol_ke_batch(int etype, int ndof, int64_t n_elem, const double* xy, const int32_t* conn,
                              const double* tensors, double* ke_out, void* stream) {
    FEMOL_REQUIRE(n_elem >= 0 && xy && conn && tensors && ke_out, "femol_ke_batch: null argument");
    FEMOL_REQUIRE(ndof == 2 || ndof == 6, "femol_ke_batch: ndof must be 2 (plane) or 6 (plate)");
    if (n_elem == 0) return FEMOL_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define KE_CASE(ET)                                                                   \
    case ET:                                                                          \
        return ndof == 2 ? launch_ke<ET, 2>(n_elem, xy, conn, tensors, ke_out, st)    \
                         : launch_ke<ET, 6>(n_elem, xy, conn, tensors, ke_out, st);
    switch (etype) {
        case FEMOL_T3:
            FEMOL_REQUIRE(ndof == 2, "femol_ke_batch: T3 supports plane stress only (reference elements.py:231)");
            return launch_ke<FEMOL_T3, 2>(n_elem, xy, conn, tensors, ke_out, st);
        KE_CASE(FEMOL_Q4)
        KE_CASE(FEMOL_T6)
        KE_CASE(FEMOL_Q8)
    }
#undef KE_CASE
    femol_set_error("femol_ke_batch: unknown element type %d", etype);
    return FEMOL_E_ARG;
}

// Ke with ONE TENSOR SET PER ELEMENT (tensors [n_elem][31]); accumulate != 0 adds to ke_out (the drilling diagonal, which
// is SET to 1 per call, adds up like in the reference's Ke_base + Ke_coat, FEM.py:728)
extern "C" int femol_ke_batch_var(int etype, int ndof, int64_t n_elem, const double* xy, const int32_t* conn,
                                  const double* tensors, int accumulate, double* ke_out, void* stream) {
    FEMOL_REQUIRE(n_elem >= 0 && xy && conn && tensors && ke_out, "femol_ke_batch_var: null argument");
    FEMOL_REQUIRE(ndof == 2 || ndof == 6, "femol_ke_batch_var: ndof must be 2 (plane) or 6 (plate)");
    if (n_elem == 0) return FEMOL_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define KEV_CASE(ET)                                                                                   \
    case ET:                                                                                           \
        return ndof == 2 ? launch_ke<ET, 2>(n_elem, xy, conn, tensors, ke_out, st, 31, accumulate)      \
                         : launch_ke<ET, 6>(n_elem, xy, conn, tensors, ke_out, st, 31, accumulate);
    switch (etype) {
        case FEMOL_T3:
            FEMOL_REQUIRE(ndof == 2, "femol_ke_batch_var: T3 supports plane stress only (reference elements.py:231)");
            return launch_ke<FEMOL_T3, 2>(n_elem, xy, conn, tensors, ke_out, st, 31, accumulate);
        KEV_CASE(FEMOL_Q4)
        KEV_CASE(FEMOL_T6)
        KEV_CASE(FEMOL_Q8)
    }
#undef KEV_CASE
    femol_set_error("femol_ke_batch_var: unknown element type %d", etype);
    return FEMOL_E_ARG;
}

extern "C" int femol_me_batch(int etype, int ndof, int64_t n_elem, const double* xy, const int32_t* conn,
                              const double* V, double* me_out, void* stream) {
    FEMOL_REQUIRE(n_elem >= 0 && xy && conn && V && me_out, "femol_me_batch: null argument");
    FEMOL_REQUIRE(ndof == 2 || ndof == 6, "femol_me_batch: ndof must be 2 (plane) or 6 (plate)");
    if (n_elem == 0) return FEMOL_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define ME_CASE(ET)                                                              \
    case ET:                                                                     \
        return ndof == 2 ? launch_me<ET, 2>(n_elem, xy, conn, V, me_out, st)     \
                         : launch_me<ET, 6>(n_elem, xy, conn, V, me_out, st);
    switch (etype) {
        case FEMOL_T3:
            FEMOL_REQUIRE(ndof == 2, "femol_me_batch: T3 supports plane stress only (reference elements.py:253)");
            return launch_me<FEMOL_T3, 2>(n_elem, xy, conn, V, me_out, st);
        ME_CASE(FEMOL_Q4)
        ME_CASE(FEMOL_T6)
        ME_CASE(FEMOL_Q8)
    }
#undef ME_CASE
    femol_set_error("femol_me_batch: unknown element type %d", etype);
    return FEMOL_E_ARG;
}
