// vdn_comm.cu -- inter-rank plumbing (halo exchange, scalar all-reduces).  Single-rank contexts never
// reach the exchange; the reductions are the identity.  The multi-GPU implementation (NCCL send/recv over
// NVLink, one rank per GPU) replaces FBoxLib's MPI-based multifab_fill_boundary / parallel_reduce.
#include "vdn_ctx.h"

struct Comm { int rank = 0, nranks = 1; };

void comm_destroy(Comm *cm) { delete cm; }
void comm_exchange(vdn_ctx *c, int field, int d) { (void)c; (void)field; (void)d; }
double comm_allreduce_max(vdn_ctx *c, double v) { (void)c; return v; }
double comm_allreduce_sum(vdn_ctx *c, double v) { (void)c; return v; }

extern "C" int vdn_ctx_set_comm(vdn_ctx *ctx, int rank, int nranks, const int *region_lo, const int *region_hi, const void *nccl_unique_id)
{
    (void)region_lo; (void)region_hi; (void)nccl_unique_id; (void)rank;
    if (!ctx) return 1;
    if (nranks == 1) return 0;
    ctx->err = "multi-rank contexts are not built in this library version";
    return 1;
}
