// vdn_comm.h -- inter-rank helpers used by the multigrid and the field ghost fills (implemented in vdn_comm.cu)
#pragma once
#include "vdn_ctx.h"

// Fill ng ghost layers of an array along the split directions in dmask: faces, edges and corners of the neighbour ranks in one phase.
// nodal: face-centred direction of the array (-1: cell-centred); carry_n: multigrid level arrays -- transverse ranges of directions that
// are not split also carry index n (the level layout keeps the high boundary / periodic-seam face coefficient there).
void comm_halo(vdn_ctx *c, View v, const int *n, int dim, int ng, int nc, int nodal, int dmask, bool carry_n);
void comm_exchange_field(vdn_ctx *c, int field);          // multifab_fill_boundary between ranks, every split direction at once
double *comm_sym_alloc(vdn_ctx *c, size_t bytes, bool *owned);   // symmetric-heap allocation (peer-memory transport) or cudaMalloc
int comm_rank(const vdn_ctx *c);
int comm_nranks(const vdn_ctx *c);
const int *comm_pgrid(const vdn_ctx *c);
const int *comm_pcoord(const vdn_ctx *c);
bool comm_has_neighbor(const vdn_ctx *c, int d, int s);
void comm_allgather(vdn_ctx *c, const double *send, double *recv, size_t count);
void comm_coord_of(const vdn_ctx *c, int r, int *pc);
// peer-memory tables for kernels that write into neighbour ranks' arrays themselves (see vdn_comm.cu); false: not available (NCCL transport)
bool comm_peer_tables(vdn_ctx *c, const double *const *arrs, int narr, int dmask, long *delta27, unsigned *mask27,
                      unsigned long long **pub27, const unsigned long long **wait27, unsigned long long **mine, unsigned long long *epoch);
int comm_mg_xchg(const vdn_ctx *c);                         // exchange style of the fused multigrid levels (vdn_ctx.h: comm_mode)
// push form of comm_halo for up to 3 level arrays of one level (peer-memory transport): boundary layers stored into the neighbours' ghost layers
void comm_push(vdn_ctx *c, double *const *arrs, int narr, long off, int sy, int sz, const int *n, int dim, int ng, int dmask);
bool comm_peer_mode(const vdn_ctx *c);                     // the peer-memory transport is up (symmetric heap mapped by every rank)
long comm_halo_volume(vdn_ctx *c, const int *n, int dim, int ng, int dmask);      // cells an exchange of depth ng would move (accounting)
void comm_allreduce_max_dev(vdn_ctx *c, double *d_v);       // ncclAllReduce(MAX) of one device double, in place, asynchronous
