// vdn_comm.h -- inter-rank helpers used by the multigrid (implemented in vdn_comm.cu)
#pragma once
#include "vdn_ctx.h"

// incl_n: transverse ranges of directions that are not split also carry index n (multigrid level arrays keep the high boundary /
// periodic-seam face coefficient there); dmask_all: every split direction of the array (defaults to dmask)
void comm_halo(vdn_ctx *c, View v, const int *n, int dim, int ng, int nc, int fdir, int dmask, bool grow_prev, bool incl_n = false, int dmask_all = -1);
void comm_halo_deep(vdn_ctx *c, View v, const int *n, int dim, int ng, int dmask);
int comm_rank(const vdn_ctx *c);
int comm_nranks(const vdn_ctx *c);
const int *comm_pgrid(const vdn_ctx *c);
const int *comm_pcoord(const vdn_ctx *c);
bool comm_has_neighbor(const vdn_ctx *c, int d, int s);
void comm_allgather(vdn_ctx *c, const double *send, double *recv, size_t count);
void comm_coord_of(const vdn_ctx *c, int r, int *pc);
