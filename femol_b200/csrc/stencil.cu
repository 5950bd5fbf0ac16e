// Matrix-free application of K on structured rectangle_Q4 meshes:  y = sum_e (x_e^p Ke [+ coat]) u_e  with the
// Dirichlet rows / columns treated exactly like the assembled matrix (FEM.py:902-930).
//
// The assembled CSR K stays the product of femol_assemble (it is what `FEM_Problem.K` returns and what the
// parity tests compare), but INSIDE the solvers every fine-level product K p costs 8*nnz + ... = 720 MB of HBM
// traffic per launch on cfg 2, while the information content is one element matrix (512 B), one density per
// element and the vector: 8*N_e + 16*n = 80 MB.  This kernel moves only that: a CTA per 32 x 16 node tile staged
// in shared memory, one thread per node, the element matrix as constant-bank operands.  It is selected by the
// solvers (femol_apply_launch) when the plan carries a matrix-free descriptor for the value array in use.
#include "common.cuh"
#include "spmv.cuh"
#include <cstring>

extern "C" int femol_dot(int64_t n, const double* a, const double* b, double* out, double* ws, void* stream);

namespace {

// Nodes are numbered x-fastest (n = j (nx+1) + i) but elements y-fastest (e = i ny + j, mesh.py:852-860), so a
// thread-per-node sweep reads the densities with stride ny.  Each CTA therefore owns a TI x TJ tile of nodes and
// stages what the tile needs through shared memory with both arrays read along their own fast axis:
//   s_u  (TJ+2) x (TI+2) nodes of u, rows along i, fixed dofs already zeroed (columns of fixed dofs are 0 in K)
//   s_w  (TI+1) x (TJ+1) element weights x^p, rows along j, 0 (coating: -1) outside the mesh
//
// The element matrix travels as a kernel parameter: every thread multiplies by the same entries, so they are
// constant-bank operands of the DFMAs (measured on B200, tools/ubench/fp64_rate.cu: one broadcast LDS.128 per DFMA
// caps the FP64 pipe at 9 of its 33.7 TFLOP/s; with constant operands the loop is the 64 (ND = 2) / 576 (ND = 6)
// DFMAs per node and nothing else).
template <int ND, bool COAT>
struct StencilMat {
    double base[16 * ND * ND];
    double coat[COAT ? 16 * ND * ND : 1];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// A CTA owns NT consecutive tiles.  It first issues the asynchronous copies (cp.async, zero-filled outside the
// mesh) of all NT patches - NT x 15-19 KB in flight per CTA, which is what hides the DRAM latency: a version
// with one tile per CTA and ordinary loads ran at 1.4 TB/s - then consumes them in order.  Everything is
// straight-line (no loop around the multiply), so the element-matrix operands are read from the constant bank
// where they are used instead of being hoisted into, and spilled from, registers.
template <int ND, bool COAT, int TI, int TJ, int NT, int NPT, int MODE>
__global__ void __launch_bounds__(TI * TJ / NPT, ND == 2 ? (COAT ? 4 : 5) : (COAT ? 1 : 2))
stencil_tile_kernel(const __grid_constant__ StencilMat<ND, COAT> km, int nx, int ny, int tiles_i, int n_tiles, int cta0,
                    const double* __restrict__ xpow, const uint8_t* __restrict__ fixed,
                    const double* __restrict__ u, double* __restrict__ y, const double* __restrict__ dot_with,
                    double* __restrict__ dot_out, double* __restrict__ red_ws, const double* __restrict__ rhs,
                    const double* __restrict__ dinv, double omega) {
    // 32-bit index arithmetic throughout (femol_plan_set_matrix_free checks n_dof < 2^31): the first version of
    // this kernel spent more than half of its issue slots on 64-bit address computation
    constexpr int S = 4 * ND, PI = TI + 2, PJ = TJ + 2, WI = TI + 1, WJ = TJ + 1, H = ND / 2, NTH = TI * TJ / NPT;
    static_assert(TJ % NPT == 0, "a thread owns NPT vertically adjacent nodes");
    constexpr int STAGE = PJ * PI * ND + WI * WJ + ((WI * WJ) & 1);   // doubles per tile buffer, 16-byte multiple
    constexpr int NPN = (PJ * PI + NTH - 1) / NTH;                    // patch nodes per thread (mask handling)
    extern __shared__ __align__(16) double s_mem[];
    __shared__ double s_red[32];
    FEMOL_PDL_SYNC();
    const int npx = nx + 1;
    const int first = (cta0 + (int)blockIdx.x) * NT;
    const int tj0 = first / tiles_i, ti0 = first - tj0 * tiles_i;
    uint32_t fm[NT][NPN];
    {
        int ti = ti0, tj = tj0;
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            if (first + k < n_tiles) {
                const int i0 = ti * TI, j0 = tj * TJ;
                double* s_u = s_mem + k * STAGE;          // [PJ][PI][ND]
                double* s_w = s_u + PJ * PI * ND;         // [WI][WJ]
#pragma unroll
                for (int it = threadIdx.x; it < PJ * PI * H; it += NTH) {
                    const int r = it / (PI * H), c = it - r * (PI * H);
                    const int ci = c / H, d = (c - ci * H) * 2;
                    const int ii = i0 - 1 + ci, jj = j0 - 1 + r;
                    const bool in = (unsigned)ii <= (unsigned)nx && (unsigned)jj <= (unsigned)ny;
                    cp_async16(s_u + it * 2, u + (in ? (jj * npx + ii) * ND + d : 0), in ? 16 : 0);
                }
#pragma unroll
                for (int it = threadIdx.x; it < WI * WJ; it += NTH) {
                    const int ri = it / WJ, rj = it - ri * WJ;
                    const int ei = i0 - 1 + ri, ej = j0 - 1 + rj;
                    const bool in = (unsigned)ei < (unsigned)nx && (unsigned)ej < (unsigned)ny;
                    if (xpow) cp_async8(s_w + it, xpow + (in ? ei * ny + ej : 0), in ? 8 : 0);
                    else s_w[it] = in ? 1.0 : 0.0;
                }
#pragma unroll
                for (int q = 0; q < NPN; ++q) {
                    const int pn = threadIdx.x + q * NTH;
                    const int r = pn / PI, ci = pn - r * PI;
                    const int ii = i0 - 1 + ci, jj = j0 - 1 + r;
                    fm[k][q] = (fixed && pn < PJ * PI && (unsigned)ii <= (unsigned)nx && (unsigned)jj <= (unsigned)ny)
                                   ? __ldg(&fixed[jj * npx + ii]) : 0u;
                }
            }
            cp_async_commit();
            if (++ti == tiles_i) { ti = 0; ++tj; }
        }
    }
    double dot[1] = {0.0};
    // thread -> NPT vertically adjacent nodes (li, lj0 .. lj0 + NPT - 1): the element-matrix operand of every DFMA
    // (one uniform constant-bank load) is then shared by NPT nodes, and neighbouring patches overlap in registers
    const int lp = threadIdx.x / TI, li = threadIdx.x - lp * TI, lj0 = lp * NPT;
    int ti = ti0, tj = tj0;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        if (first + k >= n_tiles) break;     // uniform across the CTA
        switch (NT - 1 - k) {       // groups that may still be in flight while tile k is consumed (k is a constant here)
            case 7: cp_async_wait<7>(); break;
            case 6: cp_async_wait<6>(); break;
            case 5: cp_async_wait<5>(); break;
            case 4: cp_async_wait<4>(); break;
            case 3: cp_async_wait<3>(); break;
            case 2: cp_async_wait<2>(); break;
            case 1: cp_async_wait<1>(); break;
            default: cp_async_wait<0>(); break;
        }
        double* s_u = s_mem + k * STAGE;
        const double* s_w = s_u + PJ * PI * ND;
        uint32_t mine = 0u;
#pragma unroll
        for (int q = 0; q < NPN; ++q) mine |= fm[k][q];
        if (__syncthreads_or(mine != 0u)) {
            // columns of fixed dofs are zero in K: clear them in the staged patch (rare: Dirichlet boundary tiles)
#pragma unroll
            for (int q = 0; q < NPN; ++q)
                if (fm[k][q]) {
                    const int pn = threadIdx.x + q * NTH;
#pragma unroll
                    for (int d = 0; d < ND; ++d)
                        if ((fm[k][q] >> d) & 1u) s_u[pn * ND + d] = 0.0;
                }
            __syncthreads();
        }
        const int i = ti * TI + li, jb = tj * TJ + lj0;
        if (++ti == tiles_i) { ti = 0; ++tj; }
        if (i > nx || jb > ny) continue;
        double acc[NPT][ND];
#pragma unroll
        for (int q = 0; q < NPT; ++q)
#pragma unroll
            for (int d = 0; d < ND; ++d) acc[q][d] = 0.0;
        // the four elements around a node, by relative position (ex, ey): element (i - 1 + ex, j - 1 + ey)
#pragma unroll
        for (int ey = 0; ey < 2; ++ey)
#pragma unroll
            for (int ex = 0; ex < 2; ++ex) {
                // local index of the node in that element: Q4 order (0,0) (1,0) (1,1) (0,1)
                const int lx = 1 - ex, ly = 1 - ey;
                const int a = ly == 0 ? lx : 3 - lx;
                double part[NPT][ND], partc[NPT][ND];
#pragma unroll
                for (int q = 0; q < NPT; ++q)
#pragma unroll
                    for (int d = 0; d < ND; ++d) { part[q][d] = 0.0; partc[q][d] = 0.0; }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int bx = (b == 1 || b == 2) ? 1 : 0, by = (b >= 2) ? 1 : 0;
                    double ub[NPT][ND];                         // node (ei + bx, ej + by) of the staged patch, per owned node
#pragma unroll
                    for (int q = 0; q < NPT; ++q)
#pragma unroll
                        for (int d = 0; d < ND; d += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(s_u + ((lj0 + q + ey + by) * PI + li + ex + bx) * ND + d);
                            ub[q][d] = v.x; ub[q][d + 1] = v.y;
                        }
#pragma unroll
                    for (int da = 0; da < ND; ++da) {
                        const int row = (a * ND + da) * S + b * ND;
#pragma unroll
                        for (int db = 0; db < ND; ++db) {
#pragma unroll
                            for (int q = 0; q < NPT; ++q) {
                                part[q][da] = fma(km.base[row + db], ub[q][db], part[q][da]);
                                if (COAT) partc[q][da] = fma(km.coat[row + db], ub[q][db], partc[q][da]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < NPT; ++q) {
                    const double w = s_w[(li + ex) * WJ + lj0 + q + ey];
                    // outside the mesh w is 0 (zero fill); the coating's base part is unweighted, so test explicitly
                    const bool inside = (unsigned)(i - 1 + ex) < (unsigned)nx && (unsigned)(jb + q - 1 + ey) < (unsigned)ny;
#pragma unroll
                    for (int d = 0; d < ND; ++d) acc[q][d] += COAT ? (inside ? fma(w, partc[q][d], part[q][d]) : 0.0) : w * part[q][d];
                }
            }
#pragma unroll
        for (int q = 0; q < NPT; ++q) {
            const int j = jb + q;
            if (j > ny) break;
            const int n = j * npx + i;
            // rows of fixed dofs are the identity
            const uint32_t frow = fixed ? __ldg(&fixed[n]) : 0u;
            const double* un = u + (size_t)n * ND;
            double* yn = y + (size_t)n * ND;
            if (frow) {
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    if ((frow >> d) & 1u) acc[q][d] = __ldg(un + d);
            }
            if (MODE == 1) {            // y = rhs - K u
#pragma unroll
                for (int d = 0; d < ND; d += 2) {
                    const double2 r2 = __ldg(reinterpret_cast<const double2*>(rhs + (size_t)n * ND + d));
                    acc[q][d] = r2.x - acc[q][d]; acc[q][d + 1] = r2.y - acc[q][d + 1];
                }
            } else if (MODE == 2) {     // y = u + omega Dinv (rhs - K u)
                double res[ND];
#pragma unroll
                for (int d = 0; d < ND; d += 2) {
                    const double2 r2 = __ldg(reinterpret_cast<const double2*>(rhs + (size_t)n * ND + d));
                    res[d] = r2.x - acc[q][d]; res[d + 1] = r2.y - acc[q][d + 1];
                }
                const double* B = dinv + (size_t)n * ND * ND;
#pragma unroll
                for (int da = 0; da < ND; ++da) {
                    double t = 0.0;
#pragma unroll
                    for (int db = 0; db < ND; db += 2) {
                        const double2 b2 = __ldg(reinterpret_cast<const double2*>(B + da * ND + db));
                        t = fma(b2.x, res[db], t); t = fma(b2.y, res[db + 1], t);
                    }
                    acc[q][da] = __ldg(un + da) + omega * t;
                }
                if (dot_out) {      // fused y . rhs (the r.z of preconditioned CG)
#pragma unroll
                    for (int d = 0; d < ND; d += 2) {
                        const double2 r2 = __ldg(reinterpret_cast<const double2*>(rhs + (size_t)n * ND + d));
                        dot[0] = fma(acc[q][d], r2.x, dot[0]); dot[0] = fma(acc[q][d + 1], r2.y, dot[0]);
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < ND; d += 2) {
                const double2 o = make_double2(acc[q][d], acc[q][d + 1]);
                *reinterpret_cast<double2*>(yn + d) = o;
                if (MODE == 0 && dot_with) {
                    const double2 w2 = __ldg(reinterpret_cast<const double2*>(dot_with + (size_t)n * ND + d));
                    dot[0] += o.x * w2.x + o.y * w2.y;
                }
            }
        }
    }
    cp_async_wait<0>();
    if (dot_out) grid_sum_finish<1>(dot, red_ws, red_ticket(red_ws), dot_out, s_red);
}

template <int ND, bool COAT, int MODE>
int launch_stencil_c(femol_plan_t p, const femol_matrix_free* mf, const double* x, double* y, const double* dot_with,
                     double* dot_out, double* red_ws, const femol_epilogue& ep, cudaStream_t st) {
    constexpr int TI = 32, TJ = 8, NT = 2, NPT = ND == 2 ? 2 : 1;
    constexpr int WW = (TI + 1) * (TJ + 1);
    constexpr size_t smem = (size_t)NT * ((TJ + 2) * (TI + 2) * ND + WW + (WW & 1)) * sizeof(double);
    auto kern = stencil_tile_kernel<ND, COAT, TI, TJ, NT, NPT, MODE>;
    FEMOL_FUNC_ATTR_PER_DEVICE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles_i = (int)((mf->nx + 1 + TI - 1) / TI);
    const int64_t n_tiles = (int64_t)tiles_i * ((mf->ny + 1 + TJ - 1) / TJ);
    const int64_t n_cta = (n_tiles + NT - 1) / NT;
    StencilMat<ND, COAT> km;
    memcpy(km.base, mf->e_base, sizeof(km.base));
    if (COAT) memcpy(km.coat, mf->e_coat, sizeof(km.coat));
    // the one-channel grid sum has FEMOL_RED_CHANNELS * FEMOL_RED_MAX_BLOCKS slots; larger grids go in chunks and
    // the dot product becomes its own pass
    const int64_t max_grid = (int64_t)FEMOL_RED_CHANNELS * FEMOL_RED_MAX_BLOCKS;
    const bool fused = n_cta <= max_grid;
    for (int64_t c0 = 0; c0 < n_cta; c0 += max_grid) {
        const int64_t nb = n_cta - c0 < max_grid ? n_cta - c0 : max_grid;
        FEMOL_CUDA_CHECK(femol_launch_pdl(kern, dim3((unsigned)nb), dim3(TI * TJ / NPT), smem, st, km, (int)mf->nx, (int)mf->ny, tiles_i,
                                          (int)n_tiles, (int)c0, mf->xpow, mf->fixed, x, y, fused ? dot_with : nullptr,
                                          fused ? dot_out : nullptr, red_ws, ep.rhs, ep.dinv, ep.omega));
        femol_count_launch();
    }
    // (mode 2 fuses y . rhs, the r.z of preconditioned CG; mode 0 fuses y . dot_with)
    if (!fused && dot_out) return femol_dot(p->n_nodes * p->ndof, y, MODE == 2 ? ep.rhs : dot_with, dot_out, red_ws, st);
    return FEMOL_OK;
}

template <int ND, int MODE>
int launch_stencil(femol_plan_t p, const femol_matrix_free* mf, const double* x, double* y, const double* dot_with,
                   double* dot_out, double* red_ws, const femol_epilogue& ep, cudaStream_t st) {
    return mf->has_coat ? launch_stencil_c<ND, true, MODE>(p, mf, x, y, dot_with, dot_out, red_ws, ep, st)
                        : launch_stencil_c<ND, false, MODE>(p, mf, x, y, dot_with, dot_out, red_ws, ep, st);
}

template <int ND>
int launch_stencil_mode(femol_plan_t p, const femol_matrix_free* mf, const double* x, double* y, const double* dot_with,
                        double* dot_out, double* red_ws, const femol_epilogue& ep, cudaStream_t st) {
    if (ep.mode == FEMOL_EPI_RESIDUAL) return launch_stencil<ND, 1>(p, mf, x, y, nullptr, nullptr, red_ws, ep, st);
    if (ep.mode == FEMOL_EPI_JACOBI) return launch_stencil<ND, 2>(p, mf, x, y, nullptr, ep.dot_out, ep.red_ws, ep, st);
    return launch_stencil<ND, 0>(p, mf, x, y, dot_with, dot_out, red_ws, ep, st);
}

}  // namespace

// K x for the solvers: the matrix-free stencil when the plan carries a descriptor for exactly this value array,
// the CSR node-block kernel otherwise.
static inline const femol_matrix_free* stencil_for(femol_plan_t p, const double* vals) {
    const femol_matrix_free* mf = p->mf;
    return (mf && mf->vals == vals && (p->ndof == 2 || p->ndof == 6)) ? mf : nullptr;
}

int femol_apply_launch(femol_plan_t p, const double* vals, const double* x, double* y, const double* dot_with,
                       double* dot_out, double* red_ws, cudaStream_t st) {
    const femol_epilogue none = {FEMOL_EPI_NONE, nullptr, nullptr, 0.0, nullptr, nullptr};
    if (const femol_matrix_free* mf = stencil_for(p, vals)) {
        if (p->ndof == 2) return launch_stencil_mode<2>(p, mf, x, y, dot_with, dot_out, red_ws, none, st);
        return launch_stencil_mode<6>(p, mf, x, y, dot_with, dot_out, red_ws, none, st);
    }
    return femol_spmv_launch(p, vals, x, y, dot_with, dot_out, red_ws, st);
}

bool femol_apply_has_jacobi(femol_plan_t p, const double* vals) { return stencil_for(p, vals) != nullptr || p->ndof == 2; }

int femol_apply_epilogue_launch(femol_plan_t p, const double* vals, const double* x, double* y, const femol_epilogue& ep,
                                cudaStream_t st) {
    if (ep.mode == FEMOL_EPI_NONE) return femol_apply_launch(p, vals, x, y, nullptr, nullptr, nullptr, st);
    if (const femol_matrix_free* mf = stencil_for(p, vals)) {
        if (p->ndof == 2) return launch_stencil_mode<2>(p, mf, x, y, nullptr, nullptr, nullptr, ep, st);
        return launch_stencil_mode<6>(p, mf, x, y, nullptr, nullptr, nullptr, ep, st);
    }
    if (ep.mode == FEMOL_EPI_RESIDUAL) {
        if (p->ndof == 2) return spmv_node_launch<2, 1>(p, vals, x, y, nullptr, nullptr, nullptr, st, ep.rhs);
        if (p->ndof == 6) return spmv_node_launch<6, 1>(p, vals, x, y, nullptr, nullptr, nullptr, st, ep.rhs);
    } else if (ep.mode == FEMOL_EPI_JACOBI && p->ndof == 2) {
        return spmv_node_launch<2, 2>(p, vals, x, y, nullptr, ep.dot_out, ep.red_ws, st, ep.rhs, ep.dinv, ep.omega);
    }
    femol_set_error("femol_apply_epilogue_launch: mode %d not available for ndof %d", ep.mode, p->ndof);
    return FEMOL_E_ARG;
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
    FEMOL_REQUIRE(p->n_nodes * p->ndof < (int64_t)2147483647, "femol_plan_set_matrix_free: more than 2^31 dofs");
    if (!p->mf) p->mf = new femol_matrix_free();
    const size_t s2 = (size_t)16 * p->ndof * p->ndof;
    p->mf->nx = nx; p->mf->ny = ny; p->mf->xpow = xpow; p->mf->fixed = fixed; p->mf->vals = vals;
    memcpy(p->mf->e_base, e_base, s2 * sizeof(double));
    p->mf->has_coat = e_coat != nullptr;
    if (e_coat) memcpy(p->mf->e_coat, e_coat, s2 * sizeof(double));
    return FEMOL_OK;
}

// y = op(K x) with the multigrid epilogues (mode 0: K x; 1: rhs - K x; 2: x + omega Dinv (rhs - K x)) through the
// operator the solvers use; exposed for the parity tests and the per-kernel timings of bench.py
extern "C" int femol_apply_epilogue(femol_plan_t p, const double* vals, const double* x, double* y, int mode,
                                    const double* rhs, const double* dinv, double omega, void* stream) {
    FEMOL_REQUIRE(p && vals && x && y && x != y, "femol_apply_epilogue: null or aliased argument");
    FEMOL_REQUIRE(mode >= 0 && mode <= 2 && (mode == 0 || rhs) && (mode != 2 || dinv), "femol_apply_epilogue: bad mode / missing operand");
    FEMOL_REQUIRE(mode != 2 || femol_apply_has_jacobi(p, vals), "femol_apply_epilogue: mode 2 needs ndof = 2 or the matrix-free operator");
    const femol_epilogue ep = {mode, rhs, dinv, omega, nullptr, nullptr};
    return femol_apply_epilogue_launch(p, vals, x, y, ep, (cudaStream_t)stream);
}

// y = K x through whichever operator the solvers would use (tests compare it with the CSR product)
extern "C" int femol_apply(femol_plan_t p, const double* vals, const double* x, double* y, void* stream) {
    FEMOL_REQUIRE(p && vals && x && y, "femol_apply: null argument");
    return femol_apply_launch(p, vals, x, y, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
