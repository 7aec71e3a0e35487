// vdn_comm.cu -- inter-rank plumbing of the hot path: one rank per GPU, NCCL over NVLink/NVSwitch.
//
// Replaces FBoxLib's MPI layer for this path:
//   multifab_fill_boundary between boxes of different ranks  -> pack kernel + grouped ncclSend/ncclRecv + unpack kernel,
//       direction by direction (x, then y over the x-ghosted range, then z) so edge/corner ghosts need no diagonal messages
//   norm_inf / parallel_reduce (macproject.f90:65,203; F_MG residual norms) -> ncclAllReduce(double, MAX/SUM)
//   F_MG coarse levels -> agglomeration: ncclAllGather of the coarse right-hand side / coefficients, every GPU then
//       finishes the V-cycle locally on the whole coarse domain (no scatter step needed)
// The decomposition must be a tensor-product grid of rectangular regions (what boxarray_maxsize + a block map gives).
#include "vdn_ctx.h"
#include "vdn_comm.h"
#include <nccl.h>
#include <algorithm>

#define VDN_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    char b_[512]; snprintf(b_, sizeof b_, "%s:%d NCCL error %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
    throw VdnError(b_); } } while (0)

struct Comm {
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
    int nbr[3][2];                 // neighbour rank per (dir, side); -1 = none (physical boundary); == rank = periodic self
    int pgrid[3] = {1, 1, 1}, pcoord[3] = {0, 0, 0};
    std::vector<int> rlo, rhi;     // all regions [nranks][3]
    std::vector<int> coord2rank;   // process-grid coordinates (x fastest) -> rank
    double *sbuf = nullptr, *rbuf = nullptr; size_t buf_doubles = 0;
    double *d_scal = nullptr;
};

// ---- host-only planning (testable without a GPU) ----
extern "C" int vdn_comm_plan(int dim, int rank, int nranks, const int *region_lo, const int *region_hi,
                             const int *dom_lo, const int *dom_hi, const int *phys_bc, int *nbr, int *pgrid, int *pcoord)
{
    if (rank < 0 || rank >= nranks) return 1;
    const int *mlo = &region_lo[rank * 3], *mhi = &region_hi[rank * 3];
    for (int d = 0; d < 3; ++d) { pgrid[d] = 1; pcoord[d] = 0; nbr[d * 2] = nbr[d * 2 + 1] = -1; }
    for (int d = 0; d < dim; ++d) {
        // process-grid extent and my coordinate along d: distinct lower bounds among regions sharing my transverse extents
        std::vector<int> los;
        for (int r = 0; r < nranks; ++r) {
            bool same = true;
            for (int t = 0; t < dim; ++t) if (t != d && (region_lo[r * 3 + t] != mlo[t] || region_hi[r * 3 + t] != mhi[t])) same = false;
            if (same) los.push_back(region_lo[r * 3 + d]);
        }
        std::sort(los.begin(), los.end());
        pgrid[d] = (int)los.size();
        for (int q = 0; q < pgrid[d]; ++q) if (los[q] == mlo[d]) pcoord[d] = q;
        const bool per = phys_bc[d * 2] == BC_PERIODIC;
        for (int s = 0; s < 2; ++s) {
            int want;                               // the lower (s=1) / upper (s=0) bound the neighbour must have along d
            if (s == 1) { want = mhi[d] + 1; if (mhi[d] == dom_hi[d]) { if (!per) continue; want = dom_lo[d]; } }
            else        { want = mlo[d] - 1; if (mlo[d] == dom_lo[d]) { if (!per) continue; want = dom_hi[d]; } }
            int found = -1;
            for (int r = 0; r < nranks; ++r) {
                bool same = true;
                for (int t = 0; t < dim; ++t) if (t != d && (region_lo[r * 3 + t] != mlo[t] || region_hi[r * 3 + t] != mhi[t])) same = false;
                if (!same) continue;
                if (s == 1 ? region_lo[r * 3 + d] == want : region_hi[r * 3 + d] == want) found = r;
            }
            if (found < 0) return 2;               // not a tensor-product decomposition
            nbr[d * 2 + s] = found;
        }
    }
    long cells = 0, dom = 1;
    for (int r = 0; r < nranks; ++r) { long c = 1; for (int d = 0; d < dim; ++d) c *= region_hi[r * 3 + d] - region_lo[r * 3 + d] + 1; cells += c; }
    for (int d = 0; d < dim; ++d) dom *= dom_hi[d] - dom_lo[d] + 1;
    if (cells != dom || pgrid[0] * pgrid[1] * pgrid[2] != nranks) return 2;
    return 0;
}

extern "C" int vdn_nccl_unique_id(void *out128)
{
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return 0;
}

// ---- pack / unpack of up to 6 rectangular segments in one launch ----
struct Seg { int lo[3], n[3]; long off; };
struct PackArgs { View v; int nc; int nseg; Seg seg[26]; double *buf; int unpack; };
__global__ void k_pack(PackArgs a)
{
    const Seg &s = a.seg[blockIdx.y];
    const long per = (long)s.n[0] * s.n[1] * s.n[2];
    const long tot = per * a.nc;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t / per); const long q = t - (long)c * per;
        const int i = s.lo[0] + (int)(q % s.n[0]), j = s.lo[1] + (int)((q / s.n[0]) % s.n[1]), k = s.lo[2] + (int)(q / ((long)s.n[0] * s.n[1]));
        if (a.unpack) a.v(i, j, k, c) = a.buf[s.off + t]; else a.buf[s.off + t] = a.v(i, j, k, c);
    }
}

static void launch_pack(vdn_ctx *c, PackArgs &a)
{
    if (a.nseg == 0) return;
    long mx = 0;
    for (int q = 0; q < a.nseg; ++q) mx = std::max(mx, (long)a.seg[q].n[0] * a.seg[q].n[1] * a.seg[q].n[2] * a.nc);
    dim3 gr((unsigned)std::min<long>(1184, (mx + 255) / 256), a.nseg);
    k_pack<<<gr, 256, 0, c->stream>>>(a);
    VDN_CUDA(cudaGetLastError());
}

static void ensure_buf(vdn_ctx *c, size_t doubles)
{
    Comm *cm = c->comm;
    if (doubles <= cm->buf_doubles) return;
    VDN_CUDA(cudaStreamSynchronize(c->stream));
    if (cm->sbuf) cudaFree(cm->sbuf);
    if (cm->rbuf) cudaFree(cm->rbuf);
    cm->buf_doubles = doubles + doubles / 4;
    VDN_CUDA(cudaMalloc(&cm->sbuf, sizeof(double) * cm->buf_doubles));
    VDN_CUDA(cudaMalloc(&cm->rbuf, sizeof(double) * cm->buf_doubles));
}

// Exchange along the directions in `dirs` (bit mask) of an array described by (v, n, ng, nc, fdir).
// grow_prev: transverse range includes ghosts in directions < d (the x->y->z cascade that fills corners).
void comm_halo(vdn_ctx *c, View v, const int *n, int dim, int ng, int nc, int fdir, int dmask, bool grow_prev, bool incl_n, int dmask_all)
{
    if (dmask_all < 0) dmask_all = dmask;
    Comm *cm = c->comm;
    if (!cm || ng == 0) return;
    // the x -> y -> z cascade: the slab sent along d carries the ghost cells of the directions before it, so those directions must be
    // complete (received AND unpacked) before the slab is packed -- one phase per direction, not one pack for all of them
    if (grow_prev && (dmask & (dmask - 1)) != 0) {
        for (int d = 0; d < dim; ++d) if ((dmask >> d) & 1) comm_halo(c, v, n, dim, ng, nc, fdir, 1 << d, true, incl_n, dmask_all);
        return;
    }
    PackArgs ps; ps.v = v; ps.nc = nc; ps.nseg = 0; ps.unpack = 0;
    PackArgs pu = ps; pu.unpack = 1;
    struct Msg { int peer; long off, cnt; };
    std::vector<Msg> sends, recvs;
    long soff = 0, roff = 0;
    for (int d = 0; d < dim; ++d) {
        if (!((dmask >> d) & 1)) continue;
        const int nod = (fdir == d) ? 1 : 0;
        int tlo[3], tn[3];
        for (int t = 0; t < 3; ++t) {
            if (t >= dim) { tlo[t] = 0; tn[t] = 1; continue; }
            const int ext = n[t] + (fdir == t ? 1 : 0);
            if (grow_prev && t < d) { tlo[t] = -ng; tn[t] = ext + 2 * ng; }
            else { tlo[t] = 0; tn[t] = ext + ((incl_n && t != d && !(((dmask_all >> t) & 1) && cm->pgrid[t] > 1)) ? 1 : 0); }
        }
        Msg rcv_of_side[2]; bool has[2] = { false, false };
        for (int s = 0; s < 2; ++s) {
            const int peer = cm->nbr[d][s];
            if (peer < 0 || peer == cm->rank) continue;
            Seg snd, rcv;
            for (int t = 0; t < 3; ++t) { snd.lo[t] = rcv.lo[t] = tlo[t]; snd.n[t] = rcv.n[t] = tn[t]; }
            snd.n[d] = rcv.n[d] = ng;
            if (s == 0) { snd.lo[d] = nod ? 1 : 0;  rcv.lo[d] = -ng; }                       // to/from the lo neighbour
            else        { snd.lo[d] = n[d] - ng;    rcv.lo[d] = n[d] + nod; }                // to/from the hi neighbour
            const long cnt = (long)snd.n[0] * snd.n[1] * snd.n[2] * nc;
            snd.off = soff; rcv.off = roff;
            ps.seg[ps.nseg++] = snd; pu.seg[pu.nseg++] = rcv;
            sends.push_back({ peer, soff, cnt });
            rcv_of_side[s] = { peer, roff, cnt }; has[s] = true;
            soff += cnt; roff += cnt;
        }
        // messages to one peer are matched in issue order: sends go (lo, hi), receives (hi, lo), so that with a single
        // peer on both sides (2 ranks along a periodic direction) its lo slab lands in my hi ghost and vice versa
        if (has[1]) recvs.push_back(rcv_of_side[1]);
        if (has[0]) recvs.push_back(rcv_of_side[0]);
    }
    if (ps.nseg == 0) return;
    ensure_buf(c, (size_t)std::max(soff, roff));
    c->comm_bytes += 8LL * soff;
    ps.buf = cm->sbuf; pu.buf = cm->rbuf;
    launch_pack(c, ps);
    VDN_NCCL(ncclGroupStart());
    for (const Msg &m : sends) VDN_NCCL(ncclSend(cm->sbuf + m.off, (size_t)m.cnt, ncclDouble, m.peer, cm->nccl, c->stream));
    for (const Msg &m : recvs) VDN_NCCL(ncclRecv(cm->rbuf + m.off, (size_t)m.cnt, ncclDouble, m.peer, cm->nccl, c->stream));
    VDN_NCCL(ncclGroupEnd());
    launch_pack(c, pu);
}

// Ghost layers of depth ng of a cell-centred array in ONE phase: faces, edges and corners travel as separate messages to the
// (up to 26) neighbour ranks of the process grid inside one NCCL group, bracketed by one pack and one unpack launch.  The
// direction-by-direction cascade above needs one pack / group / unpack per split direction; the fused multigrid smoother
// calls this once per launch, so the latency of a phase is what limits multi-GPU scaling.
// Message matching: NCCL pairs the sends and receives of two ranks in issue order.  Every rank issues its sends in
// lexicographic order of the offset vector o (neighbour = my coordinates + o) and its receives in the reverse order -- the
// message a peer sent for its offset o' is the one I receive for my offset -o', and negation reverses the order -- so the
// pairing also holds when one peer is my neighbour for several offsets (two ranks along a periodic direction).
// host-only message plan of the single-phase exchange (testable without a GPU: tests/test_halo_plan.py).
//   pgrid / pcoord: process grid and this rank's coordinates; periodic[d]: the domain is periodic along d;
//   coord2rank[x + pgrid[0]*(y + pgrid[1]*z)]: rank at those coordinates; n: local cells; dmask: split directions of the array.
// Outputs (up to 26 entries each): peer rank and the inclusive-lo / extent boxes (local indices) of what is sent and of the ghost
// region that is received, in ISSUE order (sends lexicographic in the neighbour offset, receives in the reverse order).
extern "C" int vdn_halo_plan(int dim, const int *pgrid, const int *pcoord, const int *periodic, const int *coord2rank, const int *n, int ng, int dmask,
                             int *nsend, int *send_peer, int *send_lo, int *send_n, int *nrecv, int *recv_peer, int *recv_lo, int *recv_n)
{
    *nsend = 0; *nrecv = 0;
    auto peer_of = [&](const int *o) -> int {          // rank at my process-grid coordinates + o, or -1
        int pc[3];
        for (int d = 0; d < 3; ++d) {
            pc[d] = pcoord[d] + o[d];
            if (o[d] == 0) continue;
            if (pc[d] < 0 || pc[d] >= pgrid[d]) {
                if (!periodic[d]) return -1;
                pc[d] = (pc[d] + pgrid[d]) % pgrid[d];
            }
        }
        return coord2rank[pc[0] + pgrid[0] * (pc[1] + pgrid[1] * pc[2])];
    };
    int o[3];
    for (int pass = 0; pass < 2; ++pass)                // pass 0: sends (lexicographic), pass 1: receives (reverse)
        for (int q = 0; q < 27; ++q) {
            const int qq = pass == 0 ? q : 26 - q;
            o[0] = qq % 3 - 1; o[1] = (qq / 3) % 3 - 1; o[2] = qq / 9 - 1;       // z slowest: lexicographic in (z, y, x)
            if (o[0] == 0 && o[1] == 0 && o[2] == 0) continue;
            bool ok = true;
            for (int d = 0; d < 3; ++d) if (o[d] != 0 && (d >= dim || !((dmask >> d) & 1) || pgrid[d] == 1)) ok = false;
            if (!ok) continue;
            const int peer = peer_of(o);
            if (peer < 0) continue;
            int lo[3], ext[3];
            for (int d = 0; d < 3; ++d) {
                if (d >= dim) { lo[d] = 0; ext[d] = 1; continue; }
                // along a direction that is not split the slab also carries index n: the level arrays keep the coefficient of the high
                // boundary / periodic-seam face there, and the ghost planes are relaxed with it (split directions: index n is the first
                // ghost cell and comes with the edge message)
                if (o[d] == 0) { lo[d] = 0; ext[d] = n[d] + ((((dmask >> d) & 1) && pgrid[d] > 1) ? 0 : 1); }
                else if (pass == 0) { lo[d] = o[d] < 0 ? 0 : n[d] - ng; ext[d] = ng; }       // my cells next to that neighbour
                else                { lo[d] = o[d] < 0 ? -ng : n[d];    ext[d] = ng; }       // my ghost cells on that side
            }
            int &cnt = pass == 0 ? *nsend : *nrecv;
            if (cnt >= 26) return 1;
            int *pp = pass == 0 ? send_peer : recv_peer, *pl = pass == 0 ? send_lo : recv_lo, *pn = pass == 0 ? send_n : recv_n;
            pp[cnt] = peer;
            for (int d = 0; d < 3; ++d) { pl[3 * cnt + d] = lo[d]; pn[3 * cnt + d] = ext[d]; }
            ++cnt;
        }
    return 0;
}

void comm_halo_deep(vdn_ctx *c, View v, const int *n, int dim, int ng, int dmask)
{
    Comm *cm = c->comm;
    if (!cm || ng == 0) return;
    int ns = 0, nr = 0, speer[26], rpeer[26], slo[78], sn[78], rlo[78], rn[78], per[3];
    for (int d = 0; d < 3; ++d) per[d] = c->dom_bc[d][0] == BC_PERIODIC ? 1 : 0;
    VDN_REQUIRE(vdn_halo_plan(dim, cm->pgrid, cm->pcoord, per, cm->coord2rank.data(), n, ng, dmask, &ns, speer, slo, sn, &nr, rpeer, rlo, rn) == 0,
                "too many halo segments");
    if (ns == 0 && nr == 0) return;
    PackArgs ps; ps.v = v; ps.nc = 1; ps.nseg = 0; ps.unpack = 0;
    PackArgs pu = ps; pu.unpack = 1;
    struct Msg { int peer; long off, cnt; };
    std::vector<Msg> sends, recvs;
    long soff = 0, roff = 0;
    for (int q = 0; q < ns; ++q) {
        Seg sg; long cnt = 1;
        for (int d = 0; d < 3; ++d) { sg.lo[d] = slo[3 * q + d]; sg.n[d] = sn[3 * q + d]; cnt *= sg.n[d]; }
        sg.off = soff; ps.seg[ps.nseg++] = sg; sends.push_back({ speer[q], soff, cnt }); soff += cnt;
    }
    for (int q = 0; q < nr; ++q) {
        Seg sg; long cnt = 1;
        for (int d = 0; d < 3; ++d) { sg.lo[d] = rlo[3 * q + d]; sg.n[d] = rn[3 * q + d]; cnt *= sg.n[d]; }
        sg.off = roff; pu.seg[pu.nseg++] = sg; recvs.push_back({ rpeer[q], roff, cnt }); roff += cnt;
    }
    ensure_buf(c, (size_t)std::max(soff, roff));
    c->comm_bytes += 8LL * soff;
    ps.buf = cm->sbuf; pu.buf = cm->rbuf;
    launch_pack(c, ps);
    VDN_NCCL(ncclGroupStart());
    for (const Msg &m : sends) VDN_NCCL(ncclSend(cm->sbuf + m.off, (size_t)m.cnt, ncclDouble, m.peer, cm->nccl, c->stream));
    for (const Msg &m : recvs) VDN_NCCL(ncclRecv(cm->rbuf + m.off, (size_t)m.cnt, ncclDouble, m.peer, cm->nccl, c->stream));
    VDN_NCCL(ncclGroupEnd());
    launch_pack(c, pu);
}

void comm_exchange(vdn_ctx *c, int field, int d)
{
    DField &f = c->f[field];
    LaunchScope ls(c, "halo_exchange", 0.0, 2);
    comm_halo(c, f.view(), c->geo.n, c->dim, f.ng, f.nc, f.fdir, 1 << d, true);
}

static double allreduce(vdn_ctx *c, double v, ncclRedOp_t op)
{
    Comm *cm = c->comm;
    if (!cm || cm->nranks == 1) return v;
    c->h_pin[8] = v;
    VDN_CUDA(cudaMemcpyAsync(cm->d_scal, &c->h_pin[8], 8, cudaMemcpyHostToDevice, c->stream));
    VDN_NCCL(ncclAllReduce(cm->d_scal, cm->d_scal, 1, ncclDouble, op, cm->nccl, c->stream));
    VDN_CUDA(cudaMemcpyAsync(&c->h_pin[9], cm->d_scal, 8, cudaMemcpyDeviceToHost, c->stream));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
    return c->h_pin[9];
}
double comm_allreduce_max(vdn_ctx *c, double v) { return allreduce(c, v, ncclMax); }
double comm_allreduce_sum(vdn_ctx *c, double v) { return allreduce(c, v, ncclSum); }

int comm_rank(const vdn_ctx *c) { return c->comm ? c->comm->rank : 0; }
int comm_nranks(const vdn_ctx *c) { return c->comm ? c->comm->nranks : 1; }
const int *comm_pgrid(const vdn_ctx *c) { return c->comm->pgrid; }
const int *comm_pcoord(const vdn_ctx *c) { return c->comm->pcoord; }
bool comm_has_neighbor(const vdn_ctx *c, int d, int s) { return c->comm && c->comm->nbr[d][s] >= 0 && c->comm->nbr[d][s] != c->comm->rank; }

// all-gather equal-sized blocks (count doubles per rank): send -> recv[nranks*count]
void comm_allgather(vdn_ctx *c, const double *send, double *recv, size_t count)
{
    c->comm_bytes += 8LL * (long long)count * (c->comm->nranks - 1);
    VDN_NCCL(ncclAllGather(send, recv, count, ncclDouble, c->comm->nccl, c->stream));
}
// coordinates of rank r in the process grid
void comm_coord_of(const vdn_ctx *c, int r, int *pc)
{
    const Comm *cm = c->comm;
    for (int d = 0; d < 3; ++d) {
        if (d >= c->dim) { pc[d] = 0; continue; }
        const int nloc = c->geo.n[d];
        pc[d] = (cm->rlo[r * 3 + d] - c->dom_lo[d]) / nloc;
    }
}

void comm_destroy(Comm *cm)
{
    if (!cm) return;
    if (cm->nccl) ncclCommDestroy(cm->nccl);
    if (cm->sbuf) cudaFree(cm->sbuf);
    if (cm->rbuf) cudaFree(cm->rbuf);
    if (cm->d_scal) cudaFree(cm->d_scal);
    delete cm;
}

void ctx_rebuild_bc(vdn_ctx *c);      // vdn_ctx.cu

extern "C" int vdn_ctx_set_comm(vdn_ctx *ctx, int rank, int nranks, const int *region_lo, const int *region_hi, const void *nccl_unique_id)
{
    if (!ctx) return 1;
    try {
        VDN_CUDA(cudaSetDevice(ctx->device));
        if (nranks == 1) return 0;
        VDN_REQUIRE(!ctx->comm, "communicator already set");
        VDN_REQUIRE(!ctx->mg, "vdn_ctx_set_comm must be called before the first solve");
        for (int d = 0; d < ctx->dim; ++d)
            VDN_REQUIRE(region_lo[rank * 3 + d] == ctx->rlo[d] && region_hi[rank * 3 + d] == ctx->rhi[d], "region of this rank does not match its boxes");
        Comm *cm = new Comm();
        cm->rank = rank; cm->nranks = nranks;
        cm->rlo.assign(region_lo, region_lo + 3 * nranks); cm->rhi.assign(region_hi, region_hi + 3 * nranks);
        int pbc[6]; for (int d = 0; d < 3; ++d) { pbc[d * 2] = ctx->dom_bc[d][0]; pbc[d * 2 + 1] = ctx->dom_bc[d][1]; }
        int nb[6];
        int rc = vdn_comm_plan(ctx->dim, rank, nranks, region_lo, region_hi, ctx->dom_lo, ctx->dom_hi, pbc, nb, cm->pgrid, cm->pcoord);
        if (rc) { delete cm; throw VdnError("regions are not a tensor-product decomposition of the domain"); }
        for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) cm->nbr[d][s] = nb[d * 2 + s];
        for (int r = 0; r < nranks; ++r) for (int d = 0; d < ctx->dim; ++d)
            if (region_hi[r * 3 + d] - region_lo[r * 3 + d] + 1 != ctx->geo.n[d]) { delete cm; throw VdnError("all ranks must own regions of equal size"); }
        ncclUniqueId id; memcpy(&id, nccl_unique_id, 128);
        ctx->comm = cm;
        VDN_NCCL(ncclCommInitRank(&cm->nccl, nranks, id, rank));
        VDN_CUDA(cudaMalloc(&cm->d_scal, 64));
        // region faces with a rank neighbour are INTERIOR; periodic directions spanned by this rank alone keep wrapping
        for (int d = 0; d < ctx->dim; ++d) {
            for (int s = 0; s < 2; ++s) {
                const bool at_dom = s == 0 ? (ctx->rlo[d] == ctx->dom_lo[d]) : (ctx->rhi[d] == ctx->dom_hi[d]);
                ctx->geo.pbc[d][s] = at_dom ? ctx->dom_bc[d][s] : BC_INTERIOR;
            }
            ctx->wrap[d] = ctx->dom_bc[d][0] == BC_PERIODIC && cm->nbr[d][0] == rank;
        }
        cm->coord2rank.assign(nranks, -1);
        for (int r = 0; r < nranks; ++r) { int pc[3]; comm_coord_of(ctx, r, pc); cm->coord2rank[pc[0] + cm->pgrid[0] * (pc[1] + cm->pgrid[1] * pc[2])] = r; }
        for (int r = 0; r < nranks; ++r) VDN_REQUIRE(cm->coord2rank[r] >= 0, "process grid has holes");
        ctx_rebuild_bc(ctx);
        return 0;
    } catch (const std::exception &e) { ctx->err = e.what(); return 1; }
}
