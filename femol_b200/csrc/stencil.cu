// Matrix-free application of K on structured rectangle_Q4 meshes:  y = sum_e (x_e^p Ke [+ coat]) u_e  with the
// Dirichlet rows / columns treated exactly like the assembled matrix (FEM.py:902-930).
//
// The assembled CSR K stays the product of femol_assemble (it is what `FEM_Problem.K` returns and what the
// parity tests compare), but INSIDE the solvers every fine-level product K p costs 8*nnz + ... = 720 MB of HBM
// traffic per launch on cfg 2, while the information content is one element matrix (512 B), one density per
// element and the vector: 8*N_e + 16*n = 80 MB.  This kernel moves only that: one thread per node, the 3 x 3
// node patch of u in registers, the element matrices broadcast from shared memory.  It is selected by the
// solvers (femol_apply_launch) when the plan carries a matrix-free descriptor for the value array in use.
#include "common.cuh"
#include "spmv.cuh"

namespace {

template <int ND, bool COAT>
__global__ void __launch_bounds__(128)
stencil_apply_kernel(int64_t nx, int64_t ny, const double* __restrict__ e_base, const double* __restrict__ e_coat,
                     const double* __restrict__ xpow, const uint8_t* __restrict__ fixed, const double* __restrict__ u,
                     double* __restrict__ y, const double* __restrict__ dot_with, double* __restrict__ dot_out,
                     double* __restrict__ red_ws) {
    constexpr int S = 4 * ND, S2 = S * S;
    extern __shared__ double s_k[];  // [S2] base | [S2] coat
    __shared__ double s_red[32];
    for (int i = threadIdx.x; i < S2; i += blockDim.x) { s_k[i] = e_base[i]; if (COAT) s_k[S2 + i] = e_coat[i]; }
    __syncthreads();
    const int64_t npx = nx + 1, n_nodes = npx * (ny + 1);
    double dot[1] = {0.0};
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = n / npx, i = n - j * npx;
        // 3 x 3 patch of u with fixed dofs zeroed (columns of fixed dofs are zero in K)
        double up[3][3][ND];
#pragma unroll
        for (int dj = 0; dj < 3; ++dj)
#pragma unroll
            for (int di = 0; di < 3; ++di) {
                const int64_t ii = i + di - 1, jj = j + dj - 1;
                const bool ok = ii >= 0 && ii <= nx && jj >= 0 && jj <= ny;
                const int64_t m = jj * npx + ii;
                const uint32_t fm = (ok && fixed) ? fixed[m] : 0u;
#pragma unroll
                for (int d = 0; d < ND; d += 2) {
                    double2 v = ok ? __ldg(reinterpret_cast<const double2*>(u + m * ND + d)) : make_double2(0.0, 0.0);
                    if ((fm >> d) & 1u) v.x = 0.0;
                    if ((fm >> (d + 1)) & 1u) v.y = 0.0;
                    up[dj][di][d] = v.x; up[dj][di][d + 1] = v.y;
                }
            }
        double acc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = 0.0;
        // the four elements around node (i, j): element (ei, ej) = (i - 1 + ex, j - 1 + ey)
#pragma unroll
        for (int ey = 0; ey < 2; ++ey)
#pragma unroll
            for (int ex = 0; ex < 2; ++ex) {
                const int64_t ei = i - 1 + ex, ej = j - 1 + ey;
                if (ei < 0 || ei >= nx || ej < 0 || ej >= ny) continue;
                const double w = xpow ? __ldg(&xpow[ei * ny + ej]) : 1.0;
                // local index of node (i, j) in that element: Q4 order (0,0) (1,0) (1,1) (0,1)
                const int lx = 1 - ex, ly = 1 - ey;
                const int a = ly == 0 ? lx : 3 - lx;
                double part[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) part[d] = 0.0;
                double partc[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) partc[d] = 0.0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int bx = (b == 1 || b == 2) ? 1 : 0, by = (b >= 2) ? 1 : 0;
                    const double* ub = up[ey + by][ex + bx];   // node (ei + bx, ej + by) in patch coordinates
#pragma unroll
                    for (int da = 0; da < ND; ++da) {
                        const int row = (a * ND + da) * S + b * ND;
#pragma unroll
                        for (int db = 0; db < ND; ++db) {
                            part[da] += s_k[row + db] * ub[db];
                            if (COAT) partc[da] += s_k[S2 + row + db] * ub[db];
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) acc[d] += COAT ? part[d] + w * partc[d] : w * part[d];
            }
        // rows of fixed dofs are the identity
        const uint32_t fr = fixed ? fixed[n] : 0u;
#pragma unroll
        for (int d = 0; d < ND; d += 2) {
            double2 o = make_double2(acc[d], acc[d + 1]);
            if (fr) {
                const double2 orig = __ldg(reinterpret_cast<const double2*>(u + n * ND + d));
                if ((fr >> d) & 1u) o.x = orig.x;
                if ((fr >> (d + 1)) & 1u) o.y = orig.y;
            }
            *reinterpret_cast<double2*>(y + n * ND + d) = o;
            if (dot_with) dot[0] += o.x * dot_with[n * ND + d] + o.y * dot_with[n * ND + d + 1];
        }
    }
    if (dot_out) grid_sum_finish<1>(dot, red_ws, red_ticket(red_ws), dot_out, s_red);
}

template <int ND>
int launch_stencil(const femol_matrix_free* mf, int sm_count, const double* x, double* y, const double* dot_with,
                   double* dot_out, double* red_ws, cudaStream_t st) {
    const int64_t n_nodes = (mf->nx + 1) * (mf->ny + 1);
    int64_t blocks = (n_nodes + 127) / 128;
    int64_t cap = (int64_t)sm_count * 12;
    if (cap > FEMOL_RED_MAX_BLOCKS) cap = FEMOL_RED_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)(mf->e_coat ? 2 : 1) * 16 * ND * ND * sizeof(double);
    if (mf->e_coat)
        stencil_apply_kernel<ND, true><<<(int)blocks, 128, smem, st>>>(mf->nx, mf->ny, mf->e_base, mf->e_coat, mf->xpow,
                                                                       mf->fixed, x, y, dot_with, dot_out, red_ws);
    else
        stencil_apply_kernel<ND, false><<<(int)blocks, 128, smem, st>>>(mf->nx, mf->ny, mf->e_base, nullptr, mf->xpow,
                                                                        mf->fixed, x, y, dot_with, dot_out, red_ws);
    FEMOL_LAUNCH_CHECK();
    return FEMOL_OK;
}

}  // namespace

// K x for the solvers: the matrix-free stencil when the plan carries a descriptor for exactly this value array,
// the CSR node-block kernel otherwise.
int femol_apply_launch(femol_plan_t p, const double* vals, const double* x, double* y, const double* dot_with,
                       double* dot_out, double* red_ws, cudaStream_t st) {
    const femol_matrix_free* mf = p->mf;
    if (mf && mf->vals == vals) {
        if (p->ndof == 2) return launch_stencil<2>(mf, p->sm_count, x, y, dot_with, dot_out, red_ws, st);
        if (p->ndof == 6) return launch_stencil<6>(mf, p->sm_count, x, y, dot_with, dot_out, red_ws, st);
    }
    return femol_spmv_launch(p, vals, x, y, dot_with, dot_out, red_ws, st);
}

extern "C" int femol_plan_set_matrix_free(femol_plan_t p, int64_t nx, int64_t ny, const double* e_base,
                                          const double* e_coat, const double* xpow, const uint8_t* fixed,
                                          const double* vals) {
    FEMOL_REQUIRE(p, "femol_plan_set_matrix_free: null plan");
    if (!e_base || !vals) {  // clear
        delete p->mf;
        p->mf = nullptr;
        return FEMOL_OK;
    }
    FEMOL_REQUIRE(p->nn == 4 && (p->ndof == 2 || p->ndof == 6) && p->n_elem == nx * ny && p->n_nodes == (nx + 1) * (ny + 1),
                  "femol_plan_set_matrix_free: plan is not an nx x ny structured Q4 mesh");
    if (!p->mf) p->mf = new femol_matrix_free();
    p->mf->nx = nx; p->mf->ny = ny; p->mf->e_base = e_base; p->mf->e_coat = e_coat; p->mf->xpow = xpow;
    p->mf->fixed = fixed; p->mf->vals = vals;
    return FEMOL_OK;
}

// y = K x through whichever operator the solvers would use (tests compare it with the CSR product)
extern "C" int femol_apply(femol_plan_t p, const double* vals, const double* x, double* y, void* stream) {
    FEMOL_REQUIRE(p && vals && x && y, "femol_apply: null argument");
    return femol_apply_launch(p, vals, x, y, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
