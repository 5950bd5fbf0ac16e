// Strip partition shared by the distributed multigrid (gmg.cu) and its transport (dist.cu).
//
// One process per GPU; rank r holds the lattice rows [row0, row0 + n_rows) of every distributed level as an ordinary
// LOCAL rectangular grid (nodes x-fastest), of which it OWNS the rows [own_lo, own_hi) (local numbering).  The rows
// below own_lo / above own_hi - 1 are ghost rows: an exchange fills the `h_up` rows just below own_lo from the lower
// neighbour's top owned rows and the `h_down` rows starting at own_hi from the upper neighbour's bottom owned rows
// (the partition cuts are the same on all ranks, so h_up / h_down are properties of the level, not of the rank).
// Ghost rows beyond those exist only so that the local grids coarsen exactly like the global one (cuts are multiples of
// 2^levels lattice rows); they hold bounded garbage that never reaches an owned row.
#pragma once
#include "common.cuh"

struct femol_comm;

struct femol_strip {
    int64_t row_nodes;        // nodes per lattice row
    int64_t n_rows;           // local rows
    int64_t own_lo, own_hi;   // owned rows [own_lo, own_hi), local numbering
    int64_t row0;             // global index of local row 0
    int h_up, h_down;         // rows an exchange moves to the upper / lower neighbour
};

int femol_comm_rank(const femol_comm* c);
int femol_comm_size(const femol_comm* c);

// Fill the ghost rows of x (width doubles per node) from the neighbours' owned rows; in stream order.
int femol_dist_exchange(femol_comm* c, double* x, int width, const femol_strip& s, cudaStream_t st);
// buf[0..count) = sum over ranks, in place, in stream order (deterministic: same result on every rank)
int femol_dist_allreduce(femol_comm* c, double* buf, int64_t count, cudaStream_t st);
// out[0..total) = the ranks' pieces side by side (mine: src[0..count) at offset goff); the pieces tile [0, total)
int femol_dist_allgather(femol_comm* c, const double* src, int64_t count, int64_t goff, int64_t total, double* out, cudaStream_t st);

// ---- distributed multigrid entry points of gmg.cu used by the distributed PCG (dist.cu) ---------------------------
struct femol_gmg;
bool femol_gmg_is_dist(const femol_gmg* g);
const femol_strip* femol_gmg_dist_strip(const femol_gmg* g, int level);
femol_comm* femol_gmg_dist_comm(const femol_gmg* g);
// z = V(r) on the strip hierarchy (r, z: level-0 local vectors; valid on owned rows)
int femol_gmg_dist_vcycle(femol_gmg* g, const double* r, double* z, cudaStream_t st, bool presmoothed);
