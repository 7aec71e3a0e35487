// vdn_comm.cu -- inter-rank plumbing of the hot path: one rank per GPU over NVLink / NVSwitch.
//
// Replaces FBoxLib's MPI layer for this path:
//   multifab_fill_boundary between boxes of different ranks and the ghost-layer fills inside the multigrid -> ONE exchange plan
//       (faces, edges and corners of the up-to-26 neighbour ranks in a single phase) executed by one of two transports:
//         * peer memory (default inside a node): every rank allocates its exchanged arrays from one "symmetric heap" whose CUDA-IPC handle
//           all ranks map; an exchange is ONE kernel that announces "my stream has reached exchange e" in a flag of its own heap, spins on
//           the flags of the ranks it reads from and then copies their boundary cells straight out of their memory into its own ghost cells
//           -- no pack buffers, no NCCL launch, ~1 launch instead of 3 + a group of up to 34 messages;
//         * NCCL (fallback when IPC mapping is not possible): pack kernel + grouped ncclSend/ncclRecv + unpack kernel of the same plan.
//   norm_inf / parallel_reduce (macproject.f90:65,203; F_MG residual norms) -> ncclAllReduce(double, MAX/SUM)
//   F_MG coarse levels -> agglomeration: ncclAllGather of the coarse right-hand side / coefficients, every GPU then
//       finishes the V-cycle locally on the whole coarse domain (no scatter step needed)
// The decomposition must be a tensor-product grid of equal rectangular regions (what boxarray_maxsize + a block map gives).
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
    // peer-memory transport
    bool p2p = false;
    char *heap = nullptr; size_t heap_bytes = 0, heap_used = 0;     // symmetric heap: same layout on every rank
    std::vector<char *> peer_base;                                  // base of rank r's heap in MY address space (own rank: heap)
    char **d_peer_base = nullptr;
    unsigned long long epoch = 0;                                   // exchanges issued so far (the same sequence on every rank)
    unsigned int *d_push_ctr = nullptr;                             // block counter of k_halo_push
    int mg_xchg = 0;                                                // how the fused multigrid levels exchange (vdn_comm_tune): 0 push inside the sweep, 2 pull kernels, 3 push kernels
};
constexpr size_t HEAP_RESERVED = 1024;                              // header of every heap: word 0 = "my stream has reached exchange e", words 1 + r = "the data rank r pushed for exchange e has arrived"

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

// ------------------------------------------------------------------------------------------
// The exchange plan (host only; tests/test_halo_plan.py plays every rank of a process grid through it on the CPU).
//   pgrid / pcoord: process grid and this rank's coordinates; periodic[d]: the domain is periodic along d;
//   coord2rank[x + pgrid[0]*(y + pgrid[1]*z)]: rank at those coordinates; n: local cells; ng: ghost layers to fill;
//   dmask: split directions to exchange; nodal: direction in which the array is face-centred (-1: cell-centred);
//   carry_n: multigrid level arrays -- along a direction that is NOT split, or that is split with this rank at the physical high end of the
//            domain, a slab also carries index n (they keep the coefficient of the high boundary / periodic-seam face there, and the ghost
//            planes are relaxed with it: a Dirichlet face uses its coefficient).
// Outputs (up to 26 entries each, in ISSUE order): peer rank, inclusive-lo / extent boxes (local indices) of what is sent and of the ghost
// region that is received, and for a received box the index shift into the PEER's local numbering (peer index = my index + shift).
// NCCL matches the messages of a pair of ranks first-in first-out: sends are issued in lexicographic order of the neighbour offset o and
// receives in the reverse order -- the message a peer sent for its offset o' is the one I receive for my offset -o', and negation reverses
// the order -- so the pairing also holds when one peer is my neighbour for several offsets (two ranks along a periodic direction).
// ------------------------------------------------------------------------------------------
static int halo_plan_core(int dim, const int *pgrid, const int *pcoord, const int *periodic, const int *coord2rank, const int *n, int ng,
                          int dmask, int nodal, int carry_n,
                          int *nsend, int *send_peer, int *send_lo, int *send_n, int *nrecv, int *recv_peer, int *recv_lo, int *recv_n, int *recv_shift,
                          int *send_shift)
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
            int lo[3], ext[3], sh[3];
            for (int d = 0; d < 3; ++d) {
                sh[d] = 0;
                if (d >= dim) { lo[d] = 0; ext[d] = 1; continue; }
                const int nod = (d == nodal) ? 1 : 0;
                const bool split = ((dmask >> d) & 1) && pgrid[d] > 1;
                // index n travels along a direction where no neighbour rank supplies it: not split, or split with this rank at the physical
                // high end (the peer of a transverse message sits at the same coordinate, so both sides agree on the extent)
                const bool hi_end = split && pcoord[d] == pgrid[d] - 1 && !periodic[d];
                if (o[d] == 0) { lo[d] = 0; ext[d] = n[d] + nod + ((carry_n && (!split || hi_end)) ? 1 : 0); }
                else if (pass == 0) { lo[d] = o[d] < 0 ? nod : n[d] - ng; ext[d] = ng; sh[d] = -o[d] * n[d]; }   // my cells / faces next to that neighbour (its ghosts)
                else                { lo[d] = o[d] < 0 ? -ng : n[d] + nod; ext[d] = ng; sh[d] = -o[d] * n[d]; }   // my ghosts on that side
            }
            int &cnt = pass == 0 ? *nsend : *nrecv;
            if (cnt >= 26) return 1;
            int *pp = pass == 0 ? send_peer : recv_peer, *pl = pass == 0 ? send_lo : recv_lo, *pn = pass == 0 ? send_n : recv_n;
            pp[cnt] = peer;
            for (int d = 0; d < 3; ++d) { pl[3 * cnt + d] = lo[d]; pn[3 * cnt + d] = ext[d]; if (pass == 1 && recv_shift) recv_shift[3 * cnt + d] = sh[d];
                                          if (pass == 0 && send_shift) send_shift[3 * cnt + d] = sh[d]; }
            ++cnt;
        }
    return 0;
}
extern "C" int vdn_halo_plan_ex(int dim, const int *pgrid, const int *pcoord, const int *periodic, const int *coord2rank, const int *n, int ng,
                                int dmask, int nodal, int carry_n,
                                int *nsend, int *send_peer, int *send_lo, int *send_n, int *nrecv, int *recv_peer, int *recv_lo, int *recv_n, int *recv_shift)
{
    return halo_plan_core(dim, pgrid, pcoord, periodic, coord2rank, n, ng, dmask, nodal, carry_n, nsend, send_peer, send_lo, send_n,
                          nrecv, recv_peer, recv_lo, recv_n, recv_shift, nullptr);
}
// the multigrid form (cell-centred level arrays): kept as the entry point the round-1 tests bind
extern "C" int vdn_halo_plan(int dim, const int *pgrid, const int *pcoord, const int *periodic, const int *coord2rank, const int *n, int ng, int dmask,
                             int *nsend, int *send_peer, int *send_lo, int *send_n, int *nrecv, int *recv_peer, int *recv_lo, int *recv_n)
{
    return vdn_halo_plan_ex(dim, pgrid, pcoord, periodic, coord2rank, n, ng, dmask, -1, 1, nsend, send_peer, send_lo, send_n, nrecv, recv_peer, recv_lo, recv_n, nullptr);
}

// ---- peer-memory transport: one kernel per exchange ----
struct PullSeg { int peer; int lo[3], n[3], shift[3]; int blk0; };      // blk0: first block of the flattened grid that works on this segment
struct PullArgs {
    long arr_off;                   // byte offset, inside the symmetric heap, of the array's local element (0,0,0), comp 0
    int sy, sz, cs, nc;
    int nseg; PullSeg seg[26];
    char *const *peer_base; int me;
    unsigned long long epoch;
};
constexpr int PULL_NT = 256, PULL_UNR = 4;      // a block moves PULL_NT * PULL_UNR elements: all loads in flight before the first store
// Reads over NVLink are latency-bound (~2 us round trip): the grid covers every element at once (one short block per 1024 elements) so
// that the whole exchange costs about two round trips -- the flag, then the data -- instead of one per loop iteration.
__global__ void __launch_bounds__(PULL_NT) k_halo_pull(PullArgs a)
{
    int q = 0;
#pragma unroll 1
    for (int t = 1; t < a.nseg; ++t) if ((int)blockIdx.x >= a.seg[t].blk0) q = t;
    const PullSeg &s = a.seg[q];
    if (threadIdx.x == 0) {
        // everything my stream produced before this kernel is complete: publish it, then wait until the rank this block reads from has
        // published the same exchange (its producing kernels are complete too)
        volatile unsigned long long *mine = (volatile unsigned long long *)a.peer_base[a.me];
        __threadfence_system();
        *mine = a.epoch;
        volatile unsigned long long *theirs = (volatile unsigned long long *)a.peer_base[s.peer];
        while (*theirs < a.epoch) { }
        __threadfence_system();
    }
    __syncthreads();
    const double *src = (const double *)(a.peer_base[s.peer] + a.arr_off);
    double *dst = (double *)(a.peer_base[a.me] + a.arr_off);
    const long per = (long)s.n[0] * s.n[1] * s.n[2];
    const long tot = per * a.nc;
    const long t0 = (long)((int)blockIdx.x - s.blk0) * (PULL_NT * PULL_UNR) + threadIdx.x;
    double v[PULL_UNR]; long dix[PULL_UNR]; bool ok[PULL_UNR];        // (ghost cells have negative offsets: validity is a flag of its own)
#pragma unroll
    for (int u = 0; u < PULL_UNR; ++u) {
        const long t = t0 + (long)u * PULL_NT;
        dix[u] = 0; v[u] = 0.0;
        ok[u] = t < tot;
        if (ok[u]) {
            const int c = (int)(t / per); const long r = t - (long)c * per;
            const int i = s.lo[0] + (int)(r % s.n[0]), j = s.lo[1] + (int)((r / s.n[0]) % s.n[1]), k = s.lo[2] + (int)(r / ((long)s.n[0] * s.n[1]));
            dix[u] = i + (long)a.sy * j + (long)a.sz * k + (long)a.cs * c;
            const long s_ix = (i + s.shift[0]) + (long)a.sy * (j + s.shift[1]) + (long)a.sz * (k + s.shift[2]) + (long)a.cs * c;
            v[u] = __ldcv(src + s_ix);              // peer memory: never through a stale L1 line
        }
    }
#pragma unroll
    for (int u = 0; u < PULL_UNR; ++u) if (ok[u]) dst[dix[u]] = v[u];
}

static bool in_heap(const Comm *cm, const void *p)
{
    return cm->p2p && (const char *)p >= cm->heap && (const char *)p < cm->heap + cm->heap_bytes;
}

// ---- peer-memory transport, push form: one kernel stores this rank's boundary layers of up to 3 arrays of one level straight into the ghost
// layers of the neighbours' arrays (local reads, posted NVLink writes), then tells every neighbour "my data of exchange e has arrived" in the
// neighbour's own heap header (word 1 + my rank) and waits until the neighbours have told it the same.  Used for the ping-pong arrays of the
// fused multigrid levels, where the receiver is never still reading the ghost layers that are being overwritten (it read the OTHER buffer in the
// sweep it may still be running), so no "ready to receive" handshake is needed and the exchange costs one flag latency. ----
struct PushSeg { int peer; int lo[3], n[3], shift[3]; int blk0; };
struct PushArgs {
    long arr_off[3]; int narr;      // byte offsets, inside the symmetric heap, of each array's local element (0,0,0)
    int sy, sz;
    int nseg; PushSeg seg[26];
    int npeer; int peers[26];       // distinct neighbour ranks
    char *const *peer_base; int me;
    unsigned long long epoch;
    unsigned int *ctr; int nblk;    // blocks that have stored (and fenced) their part
};
constexpr int PUSH_NT = 256, PUSH_UNR = 4;
__global__ void __launch_bounds__(PUSH_NT) k_halo_push(PushArgs a)
{
    int q = 0;
#pragma unroll 1
    for (int t = 1; t < a.nseg; ++t) if ((int)blockIdx.x >= a.seg[t].blk0) q = t;
    const PushSeg &s = a.seg[q];
    const long per = (long)s.n[0] * s.n[1] * s.n[2];
    const long t0 = (long)((int)blockIdx.x - s.blk0) * (PUSH_NT * PUSH_UNR) + threadIdx.x;
    for (int ar = 0; ar < a.narr; ++ar) {
        const double *src = (const double *)(a.peer_base[a.me] + a.arr_off[ar]);
        double *dst = (double *)(a.peer_base[s.peer] + a.arr_off[ar]);
        double v[PUSH_UNR]; long dix[PUSH_UNR]; bool ok[PUSH_UNR];
#pragma unroll
        for (int u = 0; u < PUSH_UNR; ++u) {
            const long t = t0 + (long)u * PUSH_NT;
            dix[u] = 0; v[u] = 0.0;
            ok[u] = t < per;
            if (ok[u]) {
                const int i = s.lo[0] + (int)(t % s.n[0]), j = s.lo[1] + (int)((t / s.n[0]) % s.n[1]), k = s.lo[2] + (int)(t / ((long)s.n[0] * s.n[1]));
                v[u] = src[i + (long)a.sy * j + (long)a.sz * k];
                dix[u] = (i + s.shift[0]) + (long)a.sy * (j + s.shift[1]) + (long)a.sz * (k + s.shift[2]);
            }
        }
#pragma unroll
        for (int u = 0; u < PUSH_UNR; ++u) if (ok[u]) dst[dix[u]] = v[u];
    }
    // the last block to finish signals the neighbours and waits for their signals
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(a.ctr, 1u) == (unsigned)(a.nblk - 1));
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *a.ctr = 0;                       // for the next exchange (stream-ordered after this kernel)
    if ((int)threadIdx.x < a.npeer) {
        const int p = a.peers[threadIdx.x];
        __threadfence_system();
        *(volatile unsigned long long *)(a.peer_base[p] + 8 * (1 + a.me)) = a.epoch;
        const volatile unsigned long long *f = (const volatile unsigned long long *)(a.peer_base[a.me] + 8 * (1 + p));
        while (*f < a.epoch) { }
        __threadfence_system();
    }
}

// Fill ng ghost layers of an array along the split directions in dmask (faces, edges and corners in one phase).
// nodal: face-centred direction of the array or -1; carry_n: multigrid level arrays (see vdn_halo_plan_ex).
void comm_halo(vdn_ctx *c, View v, const int *n, int dim, int ng, int nc, int nodal, int dmask, bool carry_n)
{
    Comm *cm = c->comm;
    if (!cm || ng == 0) return;
    int ns = 0, nr = 0, speer[26], rpeer[26], slo[78], sn[78], rlo[78], rn[78], rsh[78], per[3];
    for (int d = 0; d < 3; ++d) per[d] = c->dom_bc[d][0] == BC_PERIODIC ? 1 : 0;
    VDN_REQUIRE(vdn_halo_plan_ex(dim, cm->pgrid, cm->pcoord, per, cm->coord2rank.data(), n, ng, dmask, nodal, carry_n ? 1 : 0,
                                 &ns, speer, slo, sn, &nr, rpeer, rlo, rn, rsh) == 0, "too many halo segments");
    if (ns == 0 && nr == 0) return;
    if (in_heap(cm, v.p)) {
        PullArgs a;
        a.arr_off = (long)((const char *)v.p - cm->heap); a.sy = v.sy; a.sz = v.sz; a.cs = v.cs; a.nc = nc;
        a.nseg = nr; a.peer_base = cm->d_peer_base; a.me = cm->rank; a.epoch = ++cm->epoch;
        long bytes = 0; int nblk = 0;
        for (int q = 0; q < nr; ++q) {
            a.seg[q].peer = rpeer[q];
            long cnt = nc;
            for (int d = 0; d < 3; ++d) { a.seg[q].lo[d] = rlo[3 * q + d]; a.seg[q].n[d] = rn[3 * q + d]; a.seg[q].shift[d] = rsh[3 * q + d]; cnt *= rn[3 * q + d]; }
            a.seg[q].blk0 = nblk;
            nblk += (int)((cnt + PULL_NT * PULL_UNR - 1) / (PULL_NT * PULL_UNR));
            bytes += 8 * cnt;
        }
        c->comm_bytes += bytes;                      // pulled = what the peers would have sent (equal regions: symmetric)
        k_halo_pull<<<nblk, PULL_NT, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
        return;
    }
    PackArgs ps; ps.v = v; ps.nc = nc; ps.nseg = 0; ps.unpack = 0;
    PackArgs pu = ps; pu.unpack = 1;
    struct Msg { int peer; long off, cnt; };
    std::vector<Msg> sends, recvs;
    long soff = 0, roff = 0;
    for (int q = 0; q < ns; ++q) {
        Seg sg; long cnt = nc;
        for (int d = 0; d < 3; ++d) { sg.lo[d] = slo[3 * q + d]; sg.n[d] = sn[3 * q + d]; cnt *= sg.n[d]; }
        sg.off = soff; ps.seg[ps.nseg++] = sg; sends.push_back({ speer[q], soff, cnt }); soff += cnt;
    }
    for (int q = 0; q < nr; ++q) {
        Seg sg; long cnt = nc;
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

// push form of comm_halo for up to 3 multigrid level arrays of one level (all in the symmetric heap): see k_halo_push
void comm_push(vdn_ctx *c, double *const *arrs, int narr, long off, int sy, int sz, const int *n, int dim, int ng, int dmask)
{
    Comm *cm = c->comm;
    if (!cm || ng == 0 || narr == 0) return;
    VDN_REQUIRE(cm->p2p && narr <= 3, "comm_push needs the peer-memory transport");
    int ns = 0, nr = 0, speer[26], rpeer[26], slo[78], sn[78], rlo[78], rn[78], rsh[78], ssh[78], per[3];
    for (int d = 0; d < 3; ++d) per[d] = c->dom_bc[d][0] == BC_PERIODIC ? 1 : 0;
    VDN_REQUIRE(halo_plan_core(dim, cm->pgrid, cm->pcoord, per, cm->coord2rank.data(), n, ng, dmask, -1, 1,
                               &ns, speer, slo, sn, &nr, rpeer, rlo, rn, rsh, ssh) == 0, "too many halo segments");
    if (ns == 0) return;
    PushArgs a;
    a.narr = narr;
    for (int q = 0; q < narr; ++q) { VDN_REQUIRE(in_heap(cm, arrs[q]), "comm_push: array outside the symmetric heap"); a.arr_off[q] = (long)((const char *)(arrs[q] + off) - cm->heap); }
    a.sy = sy; a.sz = sz; a.nseg = ns; a.peer_base = cm->d_peer_base; a.me = cm->rank; a.epoch = ++cm->epoch;
    a.npeer = 0;
    long bytes = 0; int nblk = 0;
    for (int q = 0; q < ns; ++q) {
        a.seg[q].peer = speer[q];
        long cnt = 1;
        for (int d = 0; d < 3; ++d) { a.seg[q].lo[d] = slo[3 * q + d]; a.seg[q].n[d] = sn[3 * q + d]; a.seg[q].shift[d] = ssh[3 * q + d]; cnt *= sn[3 * q + d]; }
        a.seg[q].blk0 = nblk;
        nblk += (int)((cnt + PUSH_NT * PUSH_UNR - 1) / (PUSH_NT * PUSH_UNR));
        bytes += 8 * cnt * narr;
        bool seen = false;
        for (int t = 0; t < a.npeer; ++t) if (a.peers[t] == speer[q]) seen = true;
        if (!seen) a.peers[a.npeer++] = speer[q];
    }
    a.ctr = cm->d_push_ctr; a.nblk = nblk;
    c->comm_bytes += bytes;
    k_halo_push<<<nblk, PUSH_NT, 0, c->stream>>>(a);
    VDN_CUDA(cudaGetLastError());
}

long comm_halo_volume(vdn_ctx *c, const int *n, int dim, int ng, int dmask)
{
    Comm *cm = c->comm;
    if (!cm) return 0;
    int ns = 0, nr = 0, speer[26], rpeer[26], slo[78], sn[78], rlo[78], rn[78], rsh[78], per[3];
    for (int d = 0; d < 3; ++d) per[d] = c->dom_bc[d][0] == BC_PERIODIC ? 1 : 0;
    if (vdn_halo_plan_ex(dim, cm->pgrid, cm->pcoord, per, cm->coord2rank.data(), n, ng, dmask, -1, 1, &ns, speer, slo, sn, &nr, rpeer, rlo, rn, rsh)) return 0;
    long v = 0;
    for (int q = 0; q < nr; ++q) v += (long)rn[3 * q] * rn[3 * q + 1] * rn[3 * q + 2];
    return v;
}

// Tables for a kernel that stores into its neighbours' ghost layers itself (the fused smoother in peer-memory mode): for the rank at every
// process-grid offset (ox, oy, oz) -> index (ox+1) + 3 (oy+1) + 9 (oz+1) -- the byte distance from this rank's symmetric heap to that rank's
// (one layout on every rank: local pointer + distance = the same array there) and its flag word; this rank's flag word and the epoch of this
// launch (one epoch per exchange-like event, the same sequence on every rank).  arrs: the arrays the kernel will push (all must live in the heap).
// dmask: split directions of the level.  Returns false when the peer-memory transport is not available for these arrays.
constexpr int HDR_REACHED = 64;         // heap header words HDR_REACHED + r: "rank r's stream has reached epoch e", stored there BY rank r
bool comm_peer_tables(vdn_ctx *c, const double *const *arrs, int narr, int dmask, long *delta27, unsigned *mask27,
                      unsigned long long **pub27, const unsigned long long **wait27, unsigned long long **mine, unsigned long long *epoch)
{
    Comm *cm = c->comm;
    if (!cm) return false;
    for (int q = 0; q < narr; ++q) if (arrs[q] && !in_heap(cm, arrs[q])) return false;
    *mask27 = 0;
    for (int q = 0; q < 27; ++q) {
        delta27[q] = 0; pub27[q] = nullptr; wait27[q] = nullptr;
        const int o[3] = { q % 3 - 1, (q / 3) % 3 - 1, q / 9 - 1 };
        int pc[3]; bool ok = true;
        for (int d = 0; d < 3; ++d) {
            pc[d] = cm->pcoord[d] + o[d];
            if (o[d] == 0) continue;
            if (d >= c->dim || !((dmask >> d) & 1) || cm->pgrid[d] == 1) { ok = false; break; }
            if (pc[d] < 0 || pc[d] >= cm->pgrid[d]) {
                if (c->dom_bc[d][0] != BC_PERIODIC) { ok = false; break; }
                pc[d] = (pc[d] + cm->pgrid[d]) % cm->pgrid[d];
            }
        }
        if (!ok) continue;
        const int r = cm->coord2rank[pc[0] + cm->pgrid[0] * (pc[1] + cm->pgrid[1] * pc[2])];
        delta27[q] = (long)(cm->peer_base[r] - cm->heap);
        if (q != 13) *mask27 |= 1u << q;
        if (r != cm->rank) {
            pub27[q] = (unsigned long long *)cm->peer_base[r] + HDR_REACHED + cm->rank;
            wait27[q] = (const unsigned long long *)cm->heap + HDR_REACHED + r;
        }
    }
    *mine = (unsigned long long *)cm->heap;
    *epoch = ++cm->epoch;
    return true;
}

// multifab_fill_boundary between ranks for a field: every split direction at once (vdn_stream.cu then wraps the periodic directions this
// rank owns alone over the ghosted range, which completes the edge and corner ghosts)
void comm_exchange_field(vdn_ctx *c, int field)
{
    DField &f = c->f[field];
    Comm *cm = c->comm;
    if (!cm) return;
    int dmask = 0;
    for (int d = 0; d < c->dim; ++d) if (cm->pgrid[d] > 1) dmask |= 1 << d;
    if (!dmask) return;
    LaunchScope ls(c, "halo_exchange", 0.0, in_heap(cm, f.base) ? 1 : 2);
    comm_halo(c, f.view(), c->geo.n, c->dim, f.ng, f.nc, f.fdir, dmask, false);
}

// symmetric allocation: from the heap when the peer-memory transport is up (every rank must allocate the same sequence of sizes), else
// a plain cudaMalloc.  *owned tells the caller whether it has to cudaFree the block.
double *comm_sym_alloc(vdn_ctx *c, size_t bytes, bool *owned)
{
    Comm *cm = c->comm;
    if (cm && cm->p2p) {
        const size_t al = (bytes + 255) & ~(size_t)255;
        if (cm->heap_used + al <= cm->heap_bytes) {
            double *p = (double *)(cm->heap + cm->heap_used);
            cm->heap_used += al;
            *owned = false;
            return p;
        }
        throw VdnError("symmetric heap exhausted");
    }
    double *p; VDN_CUDA(cudaMalloc(&p, bytes));
    *owned = true;
    return p;
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
// the same on a DEVICE scalar, in place and asynchronous on the context's stream (no host round trip)
void comm_allreduce_max_dev(vdn_ctx *c, double *d_v)
{
    Comm *cm = c->comm;
    if (!cm || cm->nranks == 1) return;
    VDN_NCCL(ncclAllReduce(d_v, d_v, 1, ncclDouble, ncclMax, cm->nccl, c->stream));
}
double comm_allreduce_sum(vdn_ctx *c, double v) { return allreduce(c, v, ncclSum); }

bool comm_peer_mode(const vdn_ctx *c) { return c->comm && c->comm->p2p; }
int comm_mg_xchg(const vdn_ctx *c) { return c->comm ? c->comm->mg_xchg : 0; }
int comm_rank(const vdn_ctx *c) { return c->comm ? c->comm->rank : 0; }
int comm_nranks(const vdn_ctx *c) { return c->comm ? c->comm->nranks : 1; }
const int *comm_pgrid(const vdn_ctx *c) { return c->comm->pgrid; }
const int *comm_pgrid_or_null(const vdn_ctx *c) { return c->comm ? c->comm->pgrid : nullptr; }
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
    for (int r = 0; r < (int)cm->peer_base.size(); ++r) if (r != cm->rank && cm->peer_base[r]) cudaIpcCloseMemHandle(cm->peer_base[r]);
    if (cm->d_peer_base) cudaFree(cm->d_peer_base);
    if (cm->d_push_ctr) cudaFree(cm->d_push_ctr);
    if (cm->heap) cudaFree(cm->heap);
    delete cm;
}

void ctx_rebuild_bc(vdn_ctx *c);      // vdn_ctx.cu

// Peer-memory transport: one symmetric heap per rank holding every array that is exchanged (all fields + the distributed multigrid levels),
// its CUDA-IPC handle gathered over NCCL and mapped by every rank.  Any failure leaves the NCCL transport in place (all ranks agree).
static void p2p_setup(vdn_ctx *c)
{
    Comm *cm = c->comm;
    if (c->comm_mode == 1) return;
    cm->mg_xchg = c->comm_mode;
    // size: the fields + the multigrid arrays that are not aliases of fields (level 0: 2 arrays; coarser distributed levels: 7 each)
    size_t need = HEAP_RESERVED;
    for (int i = 0; i < VDN_NFIELDS; ++i) {
        if (i >= VDN_SEDGE_X && i <= VDN_SEDGE_Z) continue;
        if (c->f[i].base) need += (c->f[i].bytes + 255) & ~(size_t)255;
    }
    {
        int nn[3] = { c->geo.n[0], c->geo.n[1], c->geo.n[2] };
        for (int l = 0; l < 32; ++l) {
            size_t tot = 1;
            for (int d = 0; d < c->dim; ++d) tot *= (size_t)(nn[d] + 2 * MG_PAD);
            need += (l == 0 ? 2 : 7) * ((tot * 8 + 255) & ~(size_t)255);
            bool ok = true;
            for (int d = 0; d < c->dim; ++d) if (nn[d] % 2 != 0 || nn[d] / 2 < 2) ok = false;
            if (!ok) break;
            for (int d = 0; d < c->dim; ++d) nn[d] /= 2;
        }
    }
    need += 1 << 20;
    int ok = 1;
    char *heap = nullptr;
    if (cudaMalloc(&heap, need) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine, heap) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    // gather (ok flag, handle) of every rank
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const int rec = 64 + 8;
    std::vector<char> all((size_t)rec * cm->nranks, 0);
    char *d_all = nullptr;
    VDN_CUDA(cudaMalloc(&d_all, all.size()));
    { char mrec[72]; memset(mrec, 0, sizeof mrec); memcpy(mrec, &mine, 64); mrec[64] = (char)ok;
      VDN_CUDA(cudaMemcpyAsync(d_all + (size_t)rec * cm->rank, mrec, rec, cudaMemcpyHostToDevice, c->stream)); }
    VDN_NCCL(ncclAllGather(d_all + (size_t)rec * cm->rank, d_all, rec, ncclChar, cm->nccl, c->stream));
    VDN_CUDA(cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, c->stream));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < cm->nranks; ++r) if (!all[(size_t)rec * r + 64]) ok = 0;
    std::vector<char *> peer(cm->nranks, nullptr);
    if (ok) {
        for (int r = 0; r < cm->nranks && ok; ++r) {
            if (r == cm->rank) { peer[r] = heap; continue; }
            cudaIpcMemHandle_t h; memcpy(&h, &all[(size_t)rec * r], 64);
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            peer[r] = (char *)p;
        }
    }
    // every rank must have mapped every heap
    { double v = ok ? 1.0 : 0.0;
      VDN_CUDA(cudaMemcpyAsync(cm->d_scal, &v, 8, cudaMemcpyHostToDevice, c->stream));
      VDN_NCCL(ncclAllReduce(cm->d_scal, cm->d_scal, 1, ncclDouble, ncclMin, cm->nccl, c->stream));
      VDN_CUDA(cudaMemcpyAsync(&v, cm->d_scal, 8, cudaMemcpyDeviceToHost, c->stream));
      VDN_CUDA(cudaStreamSynchronize(c->stream));
      ok = v > 0.5; }
    cudaFree(d_all);
    if (!ok) {
        for (int r = 0; r < cm->nranks; ++r) if (r != cm->rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
        if (heap) cudaFree(heap);
        return;
    }
    cm->heap = heap; cm->heap_bytes = need; cm->heap_used = HEAP_RESERVED; cm->peer_base = peer; cm->p2p = true;
    VDN_CUDA(cudaMemsetAsync(heap, 0, HEAP_RESERVED, c->stream));
    VDN_REQUIRE(cm->nranks <= 32 && 8 * (size_t)(HDR_REACHED + cm->nranks) <= HEAP_RESERVED, "too many ranks for the heap header");
    VDN_CUDA(cudaMalloc(&cm->d_push_ctr, 64)); VDN_CUDA(cudaMemsetAsync(cm->d_push_ctr, 0, 64, c->stream));
    VDN_CUDA(cudaMalloc(&cm->d_peer_base, sizeof(char *) * cm->nranks));
    VDN_CUDA(cudaMemcpyAsync(cm->d_peer_base, peer.data(), sizeof(char *) * cm->nranks, cudaMemcpyHostToDevice, c->stream));
    // move the fields into the heap (nothing has been uploaded yet: contents = the initial values of vdn_ctx_create)
    for (int i = 0; i < VDN_NFIELDS; ++i) {
        if (i >= VDN_SEDGE_X && i <= VDN_SEDGE_Z) continue;
        DField &f = c->f[i];
        if (!f.base) continue;
        bool owned;
        double *p = comm_sym_alloc(c, f.bytes, &owned);
        VDN_CUDA(cudaMemcpyAsync(p, f.base, f.bytes, cudaMemcpyDeviceToDevice, c->stream));
        VDN_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(f.base);
        f.base = p; f.in_heap = true;
    }
    for (int d = 0; d < c->dim; ++d) {          // SEDGE_* alias UEDGE_*
        const int nc = c->f[VDN_SEDGE_X + d].nc;
        c->f[VDN_SEDGE_X + d] = c->f[VDN_UEDGE_X + d]; c->f[VDN_SEDGE_X + d].nc = nc;
    }
    // all ranks finish mapping before anyone starts exchanging
    VDN_NCCL(ncclAllReduce(cm->d_scal, cm->d_scal, 1, ncclDouble, ncclMin, cm->nccl, c->stream));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
}

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
        p2p_setup(ctx);
        return 0;
    } catch (const std::exception &e) { ctx->err = e.what(); return 1; }
}
