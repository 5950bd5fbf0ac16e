/* femol_b200 — C-ABI of the B200-native FEMOL hot path (sm_100a).
 *
 * Drop-in boundary.  The reference (olivecha/FEMOL) is pure Python and has no
 * FFI; the seams this library is bound behind are the Python methods cited on each entry point.
 * The host-side mirror of those methods lives in `femol_b200/` (Python, like the reference) and
 * calls these functions through ctypes; INTEGRATION.md shows the stub a FEMOL maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (torch tensors in our host code)
 *     unless the parameter name starts with `h_` (host pointer, written synchronously);
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous w.r.t. the host unless
 *     the function returns host scalars (`h_*`), in which case it synchronises that stream;
 *   - return value: 0 = ok; <0 = error (FEMOL_E_*; FEMOL_E_CUDA: see femol_last_error());
 *     >0 = solver status (FEMOL_S_*);
 *   - no exceptions cross the boundary; no global mutable state besides the last-error string;
 *   - a plan may be used from one stream at a time;
 *   - doubles are IEEE FP64, indices are int32 (as scipy CSR), element/node counts are int64;
 *   - element matrices are row-major, node-major / dof-minor, dofs (Ux,Uy,Uz,Tx,Ty,Tz)
 *     (reference mesh.py:17, 297-301).
 */
#ifndef FEMOL_B200_H
#define FEMOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEMOL_OK 0
#define FEMOL_E_ARG (-1)      /* bad argument */
#define FEMOL_E_CUDA (-2)     /* CUDA runtime / launch failure */
#define FEMOL_E_LIMIT (-3)    /* mesh exceeds a compiled limit (node valence, row length) */
#define FEMOL_E_NOGPU (-4)    /* no CUDA device */
#define FEMOL_S_NOT_CONVERGED 1
#define FEMOL_S_BREAKDOWN 2   /* PCG/LOBPCG breakdown (p'Ap <= 0 or NaN) */
#define FEMOL_S_STAGNATED 3   /* recurrence converged, true residual at its attainable floor (~eps*cond) */

/* element types: reference FEMOL/elements.py classes T3 :43, T6 :277, Q4 :686, Q8 :1156 */
#define FEMOL_T3 0
#define FEMOL_Q4 1
#define FEMOL_T6 2
#define FEMOL_Q8 3

/* preconditioners for femol_pcg / femol_lobpcg */
#define FEMOL_PRECOND_NONE 0
#define FEMOL_PRECOND_JACOBI 1
#define FEMOL_PRECOND_BLOCK_JACOBI 2 /* ndof x ndof node blocks */

typedef struct femol_plan* femol_plan_t;

const char* femol_last_error(void);
const char* femol_version(void);
int femol_device_count(void);
/* number of femol kernels launched by this process so far (optionally reset); bench.py reports it */
int64_t femol_launch_count(int reset);

/* ---- K1/K2: batched element matrices --------------------------------------------------------
 * Replaces Element.Ke(*tensors) (elements.py:218-239 T3, 560-621 T6, 987-1072 Q4, Q8 stub 1351) and
 * Element.Me(material, thickness[, V]) (elements.py:241-274, 623-683, 1074-1142), evaluated for a
 * batch of elements.  xy: [n_nodes][2]; conn: [n_elem][nn].
 * tensors: 31 doubles = A(3x3) | D(3x3) | G(2x2) | B(3x3) row-major (plane stress uses A only).
 * V: ndof x ndof mass tensor (the T6 0.45/0.5 factors, elements.py:649,673, are folded in by the host).
 * out: [n_elem][size][size], size = nn*ndof. */
int femol_ke_batch(int etype, int ndof, int64_t n_elem, const double* xy, const int32_t* conn,
                   const double* tensors, double* ke_out, void* stream);

/* Element stiffness matrices with one tensor set PER ELEMENT: the variable-core laminate re-assembly of
 * FEM_Problem._K_variable_core_laminate_data (FEM.py:697-731, `solve(solve_from_zc=True)`).  tensors [n_elem][31] in
 * the layout of femol_ke_batch.  accumulate != 0: ke_out += Ke (base laminate first, coating laminate second; the
 * drilling diagonal is SET to 1 per call, so it adds up to 2 exactly as in the reference's Ke_base + Ke_coat). */
int femol_ke_batch_var(int etype, int ndof, int64_t n_elem, const double* xy, const int32_t* conn,
                       const double* tensors, int accumulate, double* ke_out, void* stream);
int femol_me_batch(int etype, int ndof, int64_t n_elem, const double* xy, const int32_t* conn,
                   const double* V, double* me_out, void* stream);

/* ---- symbolic plan: sparsity pattern + element<->CSR maps -----------------------------------
 * Replaces Mesh.global_matrix_indexes (mesh.py:284-308) + the pattern part of
 * csr_matrix(...).sum_duplicates() (FEM.py:498-501): sorted int32 indices, explicit zeros kept,
 * pattern = mesh connectivity x ndof^2.  The plan copies `conn` and owns the node adjacency and the
 * node->element incidence; CSR arrays are exported into caller-owned buffers. */
int femol_plan_create(int64_t n_elem, int nn, int ndof, int64_t n_nodes, const int32_t* conn,
                      void* stream, femol_plan_t* plan_out);
int femol_plan_destroy(femol_plan_t plan);
/* h_info[0]=nnz_node, [1]=nnz (= ndof^2 * nnz_node), [2]=max node valence, [3]=max nodes per row,
 * [4]=n_elem, [5]=n_nodes, [6]=nn, [7]=ndof */
int femol_plan_info(femol_plan_t plan, int64_t* h_info);
int femol_plan_export_csr(femol_plan_t plan, int32_t* indptr, int32_t* indices, void* stream);
/* node adjacency (block-CSR view of the same pattern): adj_ptr [n_nodes+1], adj_idx [nnz_node] */
int femol_plan_export_adjacency(femol_plan_t plan, int32_t* adj_ptr, int32_t* adj_idx, void* stream);
/* element -> CSR scatter map: map[e][a][b] = node-block position of (conn[e][a], conn[e][b]) in
 * adj_idx; the CSR position of dof pair (da,db) is indptr[node_a*ndof+da] + rank*ndof + db. */
int femol_plan_export_scatter_map(femol_plan_t plan, int32_t* map, void* stream);

/* ---- K3: numeric assembly with fused SIMP scaling and boundary conditions --------------------
 * Replaces _K/_M_*_mesh_data_* (FEM.py:550-695, 737-871), _apply_boundary_conditions_to_matrix_data
 * (FEM.py:902-930) and the value part of csr_matrix + sum_duplicates (FEM.py:498-501, 532-534).
 *   solid   (e_coat == NULL):  vals = sum_e  x_e^expo * E_e
 *   coating (e_coat != NULL):  vals = sum_e (E_e + x_e^expo * Ecoat_e)
 * x == NULL means no scaling.  e_base/e_coat are one size x size matrix (per_elem = 0: structured
 * mesh, every element equal) or [n_elem][size][size] (per_elem = 1).  fixed: [n_nodes] bytes, bit d of
 * fixed[node] set = dof d of that node is a Dirichlet dof (row and column zeroed, diagonal set to 1);
 * may be NULL.
 * xpow_ws: n_elem doubles of workspace (receives x^expo).  vals: [nnz], fully overwritten. */
int femol_assemble(femol_plan_t plan, const double* e_base, int base_per_elem, const double* e_coat,
                   int coat_per_elem, const double* x, double expo, const uint8_t* fixed,
                   double* xpow_ws, double* vals, void* stream);

/* ---- K4: SpMV ------------------------------------------------------------------------------- */
/* y = A x with A's values in plan CSR order, using the node-block adjacency for column indices */
int femol_spmv(femol_plan_t plan, const double* vals, const double* x, double* y, void* stream);
/* Single-precision copy of an assembled operator (vals32: nnz floats; ndof = 2: one float4 per node pair in adjacency
 * order) and the product from it, as the multigrid levels use it (vectors and sums stay FP64): mode 0 y = K x,
 * 1 y = rhs - K x, 2 y = x + omega Dinv (rhs - K x) (ndof = 2).  Exposed for the parity tests and the bench. */
int femol_to_f32(femol_plan_t plan, const double* vals, float* vals32, void* stream);
int femol_apply_f32(femol_plan_t plan, const float* vals32, const double* x, double* y, int mode, const double* rhs,
                    const double* dinv, double omega, void* stream);
/* generic CSR (any matrix, e.g. a scipy one) */
int femol_spmv_csr(int64_t n_rows, const int32_t* indptr, const int32_t* indices, const double* vals,
                   const double* x, double* y, void* stream);

/* Declares (and verifies on the device) that the plan's mesh is an nx x ny rectangle_Q4 grid: nodes x-fastest, element
 * e = i*ny + j with nodes [a, a+1, a+nx+2, a+nx+1], a = i + (nx+1) j (mesh.py:852-860).  femol_assemble then derives the
 * elements shared by a node pair from grid coordinates instead of reading the scatter lists (bit-identical values).
 * Synchronous; nx = 0 clears; FEMOL_E_ARG if the connectivity is in another order. */
int femol_plan_set_grid(femol_plan_t plan, int64_t nx, int64_t ny);

/* A second handle on the same symbolic plan with its own matrix-free descriptor (destroy it before the plan). */
int femol_plan_create_view(femol_plan_t plan, femol_plan_t* view_out);

/* Matrix-free K for the solvers on structured rectangle_Q4 meshes (nodes x-fastest, element e = i*ny + j):
 * declares that the value array `vals` equals sum_e (xpow_e * e_base [or e_base + xpow_e * e_coat]) with the
 * Dirichlet mask `fixed` applied — i.e. the arguments of the femol_assemble call that produced it.  femol_pcg /
 * femol_pcg_gmg then apply K from these 80 MB instead of streaming the 720 MB of CSR values (cfg 2).
 * e_base_host / e_coat_host: the shared (4 ndof)^2 element matrices in HOST memory (copied; they become kernel
 * parameters).  xpow, fixed, vals: device pointers, borrowed, must outlive the solves.  e_base_host = NULL clears.
 * femol_apply = K x through the operator the solvers use (femol_spmv always uses the CSR values). */
int femol_plan_set_matrix_free(femol_plan_t plan, int64_t nx, int64_t ny, const double* e_base_host,
                               const double* e_coat_host, const double* xpow, const uint8_t* fixed, const double* vals);
int femol_apply(femol_plan_t plan, const double* vals, const double* x, double* y, void* stream);
/* The same product with the multigrid epilogues fused in (y must not alias x):
 *   mode 0  y = K x          mode 1  y = rhs - K x          mode 2  y = x + omega * Dinv (rhs - K x)
 * Dinv = femol_precond_setup(..., FEMOL_PRECOND_BLOCK_JACOBI) ([n_nodes][ndof][ndof]).  Mode 2 is available for
 * ndof = 2 and, with the matrix-free operator, for ndof = 6. */
int femol_apply_epilogue(femol_plan_t plan, const double* vals, const double* x, double* y, int mode,
                         const double* rhs, const double* dinv, double omega, void* stream);

/* ---- K5: preconditioned conjugate gradients (replaces spsolve, FEM.py:258-265) --------------
 * Solves A u = f from the initial guess in u.  ws: femol_pcg_workspace_doubles(plan) doubles.
 * Stops when ||r|| <= rtol*||f|| (recurrence residual; the true residual is re-checked and the
 * recurrence restarted if it drifted).  h_iters / h_relres receive the iteration count and the final
 * TRUE relative residual ||f - A u|| / ||f||. */
int64_t femol_pcg_workspace_doubles(femol_plan_t plan);
int femol_pcg(femol_plan_t plan, const double* vals, const double* f, double* u, double rtol,
              int max_iter, int precond, double* ws, int* h_iters, double* h_relres, void* stream);

/* ---- K5b: geometric-multigrid preconditioned CG for structured Q4 meshes -----------------------------------
 * (SURVEY §7 "hard parts": the PCG iteration count, not bandwidth, decides SIMP iterations/s.)
 * The mesh must be the reference's rectangle_Q4 ordering (mesh.py:852-860): nodes x-fastest, element e = i*ny+j.
 * femol_gmg_create builds the level hierarchy (ceil(n/2) coarsening, plans, Dirichlet masks) once;
 * femol_gmg_setup rebuilds the Galerkin coarse operators element-wise for the current densities:
 *   solid: child matrices xpow_e * e_base ; coating: e_base + xpow_e * e_coat (xpow = x^p as left in the
 *   femol_assemble workspace; NULL = no scaling); fine_vals = the assembled fine CSR values.
 * femol_pcg_gmg = femol_pcg with one V-cycle (nu pre/post damped block-Jacobi sweeps) as preconditioner. */
typedef struct femol_gmg* femol_gmg_t;
int femol_gmg_create(int64_t nx, int64_t ny, int ndof, femol_plan_t fine_plan, const uint8_t* fine_fixed,
                     int nu, double omega, void* stream, femol_gmg_t* out);
int femol_gmg_destroy(femol_gmg_t gmg);
int femol_gmg_setup(femol_gmg_t gmg, const double* e_base, const double* e_coat, const double* xpow,
                    const double* fine_vals, void* stream);
int femol_gmg_info(femol_gmg_t gmg, int64_t* h_info);

/* Lattice hierarchy: the same multigrid for ANY single-cell-type fine mesh (T3, T6, Q4 in any node / element
 * numbering; the reference's unstructured assembly path FEM.py:608-695) whose nodes occupy every point of an
 * (Nx+1) x (Ny+1) lattice exactly once - rectangle_T3 (lattice = corner nodes) and rectangle_T6 (corner + mid-side
 * nodes: Nx = 2 nelx) - and whose elements each lie inside one cell of the coarse grid over every second lattice line.
 * Level 0 is the assembled operator of that mesh (block-Jacobi smoothing on the CSR values); level 1 is the bilinear Q4
 * grid ceil(Nx/2) x ceil(Ny/2), its element matrices the Galerkin images of the fine elements inside each coarse
 * element; the levels below are those of femol_gmg_create.  Device int32 arrays (copied): lat_node [(Nx+1)(Ny+1)]
 * lattice index (x-fastest) -> node, node_lat [n_nodes] its inverse, child_ptr [NEc+1] / child_idx [n_elem] the fine
 * elements inside coarse element I * ceil(Ny/2) + J, ascending.  femol_gmg_setup_lattice takes the fine element
 * matrices, shared (pe = 0) or one per element (pe = 1); femol_pcg_gmg / femol_gmg_apply work unchanged. */
int femol_gmg_create_lattice(int64_t Nx, int64_t Ny, int ndof, femol_plan_t fine_plan, const uint8_t* fine_fixed,
                             const int32_t* lat_node, const int32_t* node_lat, const int32_t* child_ptr,
                             const int32_t* child_idx, int nu, double omega, void* stream, femol_gmg_t* out);
int femol_gmg_setup_lattice(femol_gmg_t gmg, const double* e_base, int base_pe, const double* e_coat, int coat_pe,
                            const double* xpow, const double* fine_vals, void* stream);

/* Hourglass correction for the reference's Q4 Mindlin plate (shear integrated with one point, elements.py:1029-1054):
 * its assembled operator has a family of low-energy modes w = (-1)^(i+j) x smooth envelope that are only restrained
 * through the boundary conditions (plates with free edges), invisible to the smoother and to bilinear coarse grids;
 * multigrid-CG then needs O(n_x) iterations.  `mod` is a second hierarchy built with femol_gmg_create / _setup for
 * the sign-modulated operator S K S (element matrix S_e Ke S_e, S_e = diag(+,-,+,-) over the element's nodes) on a
 * plan view; the preconditioner becomes z = V(r) + S V_mod(S r) and the iteration count is mesh-independent again.
 * NULL detaches; the caller owns `mod`. */
int femol_gmg_attach_modulated(femol_gmg_t gmg, femol_gmg_t mod);

/* A second set of work vectors on the same hierarchy: V-cycles / femol_pcg_gmg solves on different views (each on its own
 * stream, from its own host thread) run concurrently.  A view borrows every operator of `gmg`, follows its later set-ups
 * and must be destroyed (femol_gmg_destroy) before it.  Used by LOBPCG, which preconditions its block column by column. */
int femol_gmg_clone_workspace(femol_gmg_t gmg, femol_gmg_t* out);

/* [0] = levels, then (nx, ny) per level */
int femol_gmg_apply(femol_gmg_t gmg, const double* r, double* z, void* stream); /* z = V-cycle(r) */
int64_t femol_pcg_gmg_workspace_doubles(femol_plan_t plan);
int femol_pcg_gmg(femol_plan_t plan, femol_gmg_t gmg, const double* vals, const double* f, double* u, double rtol,
                  int max_iter, double* ws, int* h_iters, double* h_relres, void* stream);
/* One CG iteration of femol_pcg_gmg is captured into a CUDA graph per solve.  Capture is not permitted on the legacy
 * default stream, so a solve called with stream 0 runs on a library-owned non-blocking stream that is ordered after
 * `stream` on entry and joined back on exit.  h_out[0] captures attempted, [1] graphs instantiated, [2] graph
 * launches, [3] solves moved to the side stream (process totals; reset != 0 clears them). */
int femol_graph_stats(int64_t* h_out, int reset);

/* ---- multi-GPU row-block partition: distributed block-Jacobi PCG --------------------------------
 * One process per GPU.  `plan` is the plan of the rank's LOCAL mesh: owned nodes first (a contiguous range of the
 * global numbering), then the ghost nodes of its one halo layer of elements, grouped by owning rank.  Only the
 * rows of owned nodes are used.  Halo: for neighbour k, the owned local nodes send_nodes[send_ptr[k]..send_ptr[k+1])
 * are packed and sent, and recv_count[k] nodes are received straight into the ghost segment starting at local
 * node recv_start[k].  Dot products are ncclAllReduce'd in stream; scalars never leave the device inside a batch.
 * NCCL is dlopen'ed (libnccl.so.2): rank 0 calls femol_dist_unique_id, the 128 bytes are broadcast by the host
 * (torch.distributed), every rank calls femol_dist_init.  vals/f/u are local arrays (u: owned + ghost entries). */
typedef struct femol_comm* femol_comm_t;
int femol_dist_unique_id(void* h_out128);
int femol_dist_init(const void* h_unique_id128, int rank, int nranks, femol_comm_t* out);
int femol_dist_destroy(femol_comm_t comm);
int64_t femol_pcg_dist_workspace_doubles(femol_plan_t plan, int64_t n_send_nodes);
int femol_pcg_dist(femol_plan_t plan, femol_comm_t comm, int64_t n_owned_nodes, int n_nbr, const int* h_nbr_rank,
                   const int64_t* h_send_ptr, const int32_t* d_send_nodes, const int64_t* h_recv_start,
                   const int64_t* h_recv_count, const double* vals, const double* f, double* u, double rtol,
                   int max_iter, int precond, double* ws, int* h_iters, double* h_relres, void* stream);

/* ---- multi-GPU strip partition with multigrid (BASELINE config 5) ------------------------------------
 * One LARGE problem on the GPUs of a box, for meshes whose nodes form a lattice (rectangle_Q4 / T3 / T6): rank r owns
 * the lattice rows [cut_r, cut_r+1), cuts being multiples of 2^n_dist rows, and holds the rows [cut_r - 2^n_dist,
 * cut_r+1] as an ordinary LOCAL mesh (femol_b200/dist.py StripPartition).  Its hierarchy is the single-GPU one
 * (femol_gmg_create / _create_lattice + femol_gmg_setup / _setup_lattice on the local mesh), cut after n_dist + 1
 * levels by femol_gmg_truncate: the alignment of the cuts makes every local element matrix of every level the global
 * one, so the set-up of the distributed levels needs no communication.  femol_gmg_set_dist attaches the strip
 * description (per level 0..n_dist: global row of the local grid's first row, owned global rows [begin, end), rows moved
 * up / down by a halo exchange) and `coarse`, a complete lattice hierarchy (Q4, identity numbering, plan `coarse_plan`,
 * node mask `coarse_fixed`) that EVERY rank holds for the whole level-n_dist grid; all borrowed.
 * femol_gmg_dist_refresh (after each femol_gmg_setup*) all-reduces the element matrices of level n_dist, assembles
 * that level and rebuilds `coarse`.  A V-cycle exchanges halo rows before every operator that reads neighbours
 * (4 per distributed level) and gathers the level-n_dist residual for the replicated coarse solve.
 * femol_pcg_gmg_dist: multigrid PCG on the strip (local plan / vals / f / u over ALL local rows; results valid on the
 * owned rows, the ghost rows of u are refreshed at the end); every dot product is an all-reduce of device scalars.
 * femol_gmg_dist_exchange: fill the ghost rows of a level-`level` vector (width doubles per node). */
int femol_gmg_truncate(femol_gmg_t gmg, int n_levels);
int femol_gmg_set_dist(femol_gmg_t gmg, femol_comm_t comm, int n_dist, const int64_t* h_row0, const int64_t* h_own_begin,
                       const int64_t* h_own_end, const int* h_up, const int* h_down, femol_gmg_t coarse,
                       femol_plan_t coarse_plan, const uint8_t* coarse_fixed);
int femol_gmg_dist_refresh(femol_gmg_t gmg, void* stream);
int femol_gmg_dist_exchange(femol_gmg_t gmg, int level, double* x, int width, void* stream);
int femol_dist_allreduce_sum(femol_comm_t comm, double* buf, int64_t count, void* stream);
/* Peer-memory transport (default; NCCL send/recv + all-reduce is the fallback and the baseline it is measured against):
 * every rank opens a WINDOW (cudaMalloc + CUDA IPC handle, 64 bytes) sized by femol_gmg_dist_p2p_sizes (h_out[0] doubles
 * of the largest halo message, h_out[1] doubles of the gathered level), the host all-gathers the handles
 * (torch.distributed) and every rank attaches them.  From then on a halo exchange is ONE single-CTA kernel per rank: it
 * stores its boundary rows straight into the neighbours' windows over NVLink as 16-byte elements {low word, epoch, high
 * word, epoch} and polls its own window until every element carries the epoch it expects - no fence, no separate flag,
 * one store latency per transfer (NCCL's LL protocol for FP64 payloads).  The two scalar all-reduces of a CG iteration
 * are fused into the CG kernels (the dot-product kernel's last block pushes the local sum to every window, the consuming
 * kernel polls and adds in rank order: bit-identical on all ranks; FEMOL_DIST_FUSED_REDUCE=0: a separate all-reduce
 * kernel), and the gather of the coarse residual is an all-to-all push.  Epochs are device counters, so the launch sequence is static (CUDA-graph capturable).
 * femol_dist_p2p_status: 0 healthy, 1 a wait timed out (~2 s: peer lost); h_stats[3] (optional) = halo / all-reduce /
 * gather kernels launched. */
int femol_gmg_dist_p2p_sizes(femol_gmg_t gmg, int64_t* h_out);
int femol_dist_p2p_open(femol_comm_t comm, int64_t halo_doubles, int64_t gather_doubles, void* h_handle64);
int femol_dist_p2p_attach(femol_comm_t comm, const void* h_handles);
int femol_dist_p2p_status(femol_comm_t comm, int64_t* h_stats);
int64_t femol_pcg_gmg_dist_workspace_doubles(femol_plan_t plan);
int femol_pcg_gmg_dist(femol_plan_t plan, femol_gmg_t gmg, const double* vals, const double* f, double* u, double rtol,
                       int max_iter, double* ws, int* h_iters, double* h_relres, void* stream);

/* ---- K6: block kernels for LOBPCG (replaces the dense eigh(K, M) of FEM.py:277 / eigsh of FEM.py:234) ----
 * The Rayleigh-Ritz outer iteration (3m x 3m pencil, m <= 16) is host logic (femol_b200/lobpcg.py); all
 * O(n) work is here.  Block vectors are row-major [n][ld] (the m values of one dof contiguous).
 *   femol_spmm            Y[:, :m] = A X[:, :m]   (matrix values read once for the m vectors; m in 1,2,4,8,12,16)
 *   femol_gram            out[qa][qb] = A^T B over n rows (deterministic); ws: femol_gram_workspace_doubles(qa,qb)
 *   femol_block_gemm      Y[n][mc] = alpha S[n][q] C[q][mc] + beta Y   (C: device, row-major; Y may alias
 *                         columns of S over the same rows)
 *   femol_block_residual  R = KX - MX diag(lambda)             (lambda: device [m])
 *   femol_precond_setup   dinv[n_nodes][ndof][ndof] = inverse of diag / diagonal blocks / identity
 *   femol_precond_apply   W = dinv R with fixed dofs (node bit mask, as femol_assemble) zeroed */
int femol_spmm(femol_plan_t plan, const double* vals, const double* X, int ldx, double* Y, int ldy, int m,
               void* stream);
int64_t femol_gram_workspace_doubles(int qa, int qb);
int femol_gram(int64_t n, int qa, int qb, const double* A, int lda, const double* B, int ldb, double* out,
               double* ws, void* stream);
int femol_block_gemm(int64_t n, int q, int mc, const double* S, int lds, const double* C, double* Y, int ldy,
                     double alpha, double beta, void* stream);
int femol_block_residual(int64_t n, int m, const double* KX, int ldk, const double* MX, int ldm,
                         const double* lambda, double* R, int ldr, void* stream);
int femol_precond_setup(femol_plan_t plan, const double* vals, int kind, double* dinv, void* stream);
int femol_precond_apply(femol_plan_t plan, const double* dinv, const uint8_t* fixed, const double* R, int ldr,
                        double* W, int ldw, int m, void* stream);
/* W[:, :m] ~= K^-1 R[:, :m]: `iters` block-CG iterations (one recurrence per column on a shared SpMM,
 * block-Jacobi preconditioned through dinv, fixed dofs kept at zero).  No host synchronisation: scalars stay
 * on the device and a column freezes once its preconditioned residual dropped by rel_tol.  This is the
 * "T ~ K^-1" preconditioner of LOBPCG.  W may alias R.  ws: femol_block_pcg_workspace_doubles(plan, m). */
int64_t femol_block_pcg_workspace_doubles(femol_plan_t plan, int m);
int femol_block_pcg(femol_plan_t plan, const double* vals, const double* dinv, const uint8_t* fixed,
                    const double* R, int ldr, double* W, int ldw, int m, int iters, double rel_tol,
                    double* ws, void* stream);

/* ---- K7/K8: element sensitivities -----------------------------------------------------------
 * compliance (TOM_SIMP.py:126-226):  q_e = u_e' E_e u_e (and qc_e with Ecoat)
 *   solid:   c = sum x^pc q ;           dc_e = -p x^pd q
 *   coating: c = sum (q + x^pc qc) ;    dc_e = -p x^pd qc
 * pc/pd are passed explicitly: structured pc = p, pd = p-1; unstructured pc = 2p, pd = 2p-1 (the
 * reference re-uses the already penalised element matrices, FEM.py:630-631, 673-675).
 * c_out: device scalar.  ws: femol_reduce_workspace_doubles() doubles. */
int64_t femol_reduce_workspace_doubles(void);
int femol_sens_compliance(femol_plan_t plan, const double* e_base, int base_per_elem, const double* e_coat,
                          int coat_per_elem, const double* u, const double* x, double p, double pc,
                          double pd, double* dc_out, double* c_out, double* ws, void* stream);
/* eigenvalue (TOM.py:408-495): dl_e = v_e'(p x^ek K_e - lambda q x^em M_e) v_e ; structured ek = p-1,
 * em = q-1 ; unstructured ek = 2p-1, em = 2q-1.  lambda: device scalar pointer (Hz, FEM.py:256). */
int femol_sens_eig(femol_plan_t plan, const double* ke, int ke_per_elem, const double* me, int me_per_elem,
                   const double* v, const double* x, double p, double q, double ek, double em,
                   const double* lambda, double* dl_out, void* stream);

/* ---- K9: mesh-independency filter (TOM_SIMP.py:228-277 == TOM.py:305-354) --------------------
 * out_i = sum_{j != i} H_ij x_j dc_j / (x_i sum_{j != i} H_ij),  H_ij = max(0, rmin - d_ij / h).
 * grid: structured mesh, element e = i*n_fast + j; w: [(2rx+1)][(2ry+1)] weight table (centre = 0).
 * list: general neighbour list (CSR) with weights. */
int femol_filter_grid(int64_t n_slow, int64_t n_fast, int rx, int ry, const double* w, const double* x,
                      const double* dc, double* out, void* stream);
int femol_filter_list(int64_t n_elem, const int32_t* nbr_ptr, const int32_t* nbr_idx, const double* nbr_w,
                      const double* x, const double* dc, double* out, void* stream);

/* ---- K10: optimality-criteria update with device-side bisection ------------------------------
 * mode 0 = SIMP_COMP._get_new_x (TOM_SIMP.py:279-313): move 0.5; the sign convention follows the entry of
 *          largest magnitude and entries of the opposite sign count as zero (identical to the reference for
 *          one-signed dc; robust to rounding-level sign noise)
 * mode 2 = as mode 0 with the reference's literal rule: any(dc>0) switches the sign of the whole vector and
 *          the other sign becomes NaN (TOM_SIMP.py:289-297)
 * mode 1 = SIMP_VIBE._get_new_x (TOM.py:251-303): move 0.3, dl<0 -> 0, rescale by 1/max if sum(dl) < 1
 *          (dc is modified in place like the reference's self.dlmbd)
 * areas: NULL -> volume test by count (sum x - f N > 0), else area weighted (TOM.py:294).
 * mask: per element 0 = free, 1 = solid (x=1), 2 = void (x=0.001); may be NULL.
 * scalars_out (device, 4 doubles): [0] change = max|x_new - x|, [1] sum(x_new), [2] last lmid, [3] #bisections.
 * ws: femol_reduce_workspace_doubles() doubles. */
int femol_oc_update(int64_t n_elem, const double* x, double* dc, int mode, double volfrac,
                    const double* areas, const uint8_t* mask, double* x_new, double* scalars_out,
                    double* ws, void* stream);

/* Centroid stresses of Q4 plane-stress elements.  Replaces Mesh.stress_from_displacement + Q4.stress
 * (mesh.py:387-425, elements.py:1144-1152): S = T B(0,0) u_e with T = t_base (x NULL), t_base * x^3 (t_coat NULL) or
 * t_base + t_coat * x^3 (the exponent 3 is the reference's), and Sv = sqrt(((Sx-Sy)^2 + (Sy-Sxy)^2 + (Sxy-Sx)^2) / 2)
 * exactly as the reference combines them.  xy [n_nodes][2], u [n_nodes][ndof] (Ux, Uy first), t_* [3][3] device,
 * out [4][n_elem] = Sx | Sy | Sxy | Sv. */
int femol_stress_q4(femol_plan_t plan, const double* xy, const double* u, const double* t_base, const double* t_coat,
                    const double* x, double* out, void* stream);

/* ---- K11: modal assurance criterion (utils.py:436-447; TOM.py:383-386) ----------------------
 * mac_out[i] = (V_i . v_ref)^2 / ((V_i . V_i)(v_ref . v_ref)), V: [k][n] row-major; device output [k]. */
int femol_mac(int64_t n, int k, const double* V, const double* v_ref, double* mac_out, double* ws,
              void* stream);

/* ---- small vector helpers used by the host mirror (all device pointers) ---------------------- */
int femol_dot(int64_t n, const double* a, const double* b, double* out, double* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEMOL_B200_H */
