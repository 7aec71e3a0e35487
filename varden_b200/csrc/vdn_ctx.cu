// vdn_ctx.cu -- context, device field registry, path-boundary copies, profiler and the extern "C" ABI (include/vdn.h).
#include "vdn_ctx.h"
#include <algorithm>

static std::string g_create_err;

// ------------------------------------------------------------------------------------------
// profiler: CUDA events on the launching stream, one (start, stop) pair per bracketed launch group
// ------------------------------------------------------------------------------------------
LaunchScope::LaunchScope(vdn_ctx *ctx, const char *name, double alg_bytes, int nlaunch) : c(ctx)
{
    c->launches += nlaunch;
    if (!c->prof_on) return;
    auto it = c->prof_idx.find(name);
    if (it == c->prof_idx.end()) {
        idx = (int)c->prof.size();
        c->prof_idx[name] = idx;
        ProfEntry e; e.name = name; c->prof.push_back(e);
    } else idx = it->second;
    c->prof[idx].launches += nlaunch;
    c->prof[idx].bytes += alg_bytes;
    auto get = [&]() { cudaEvent_t e; if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); } else VDN_CUDA(cudaEventCreate(&e)); return e; };
    e0 = get(); e1 = get();
    cudaEventRecord(e0, c->stream);
}
LaunchScope::~LaunchScope()
{
    if (idx < 0) return;
    cudaEventRecord(e1, c->stream);
    c->prof[idx].pending.emplace_back(e0, e1);
    if (c->prof[idx].pending.size() > 4096) prof_collect(c);
}
void prof_collect(vdn_ctx *c)
{
    cudaStreamSynchronize(c->stream);
    for (auto &p : c->prof) {
        for (auto &pr : p.pending) {
            float ms = 0.f; cudaEventElapsedTime(&ms, pr.first, pr.second); p.ms += ms;
            c->ev_pool.push_back(pr.first); c->ev_pool.push_back(pr.second);
        }
        p.pending.clear();
    }
}

// ------------------------------------------------------------------------------------------
// BC tables (define_bc_tower.f90:158-340), evaluated on the REGION faces
// ------------------------------------------------------------------------------------------
static void build_bc_tables(vdn_ctx *c)
{
    const int dm = c->dim, nscal = c->prm.nscal;
    const int press = dm + nscal, extrap = press + 1;
    for (int q = 0; q < 16; ++q) for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) c->adv_bc[q][d][s] = BC_INTERIOR;
    for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) c->ell_bc[d][s] = ELL_INT;
    for (int d = 0; d < dm; ++d) for (int s = 0; s < 2; ++s) {
        const int p = c->geo.pbc[d][s];
        auto set_vel = [&](int v) { for (int q = 0; q < dm; ++q) c->adv_bc[q][d][s] = v; };
        auto set_scal = [&](int v) { for (int q = 0; q < nscal; ++q) c->adv_bc[dm + q][d][s] = v; };
        if (p == BC_SLIP_WALL) {
            set_vel(BC_HOEXTRAP); c->adv_bc[d][d][s] = BC_EXT_DIR; set_scal(BC_HOEXTRAP);
            c->adv_bc[press][d][s] = BC_FOEXTRAP; c->adv_bc[extrap][d][s] = BC_FOEXTRAP; c->ell_bc[d][s] = ELL_NEU;
        } else if (p == BC_NO_SLIP_WALL) {
            set_vel(BC_EXT_DIR); set_scal(BC_HOEXTRAP);
            c->adv_bc[press][d][s] = BC_FOEXTRAP; c->adv_bc[extrap][d][s] = BC_FOEXTRAP; c->ell_bc[d][s] = ELL_NEU;
        } else if (p == BC_INLET) {
            set_vel(BC_EXT_DIR); set_scal(BC_EXT_DIR);
            c->adv_bc[press][d][s] = BC_FOEXTRAP; c->adv_bc[extrap][d][s] = BC_FOEXTRAP; c->ell_bc[d][s] = ELL_NEU;
        } else if (p == BC_OUTLET) {
            set_vel(BC_FOEXTRAP); set_scal(BC_FOEXTRAP);
            c->adv_bc[press][d][s] = BC_EXT_DIR; c->adv_bc[extrap][d][s] = BC_FOEXTRAP; c->ell_bc[d][s] = ELL_DIR;
        } else if (p == BC_SYMMETRY) {
            set_vel(BC_REFLECT_EVEN); c->adv_bc[d][d][s] = BC_REFLECT_ODD; set_scal(BC_REFLECT_EVEN);
            c->adv_bc[press][d][s] = BC_EXT_DIR; c->adv_bc[extrap][d][s] = BC_REFLECT_EVEN; c->ell_bc[d][s] = ELL_NEU;
        } else if (p == BC_PERIODIC) {
            c->ell_bc[d][s] = ELL_PER;
        }
    }
}

void ctx_rebuild_bc(vdn_ctx *c) { build_bc_tables(c); }

// sng = storage ghost width (>= ng).  PHI, RH and BETA_* are stored in the multigrid's padded layout (sng = MG_PAD, extent n+2*MG_PAD
// in every direction, which also holds the n+1 faces) so that MG level 0 can alias them without copies.
static void alloc_field(vdn_ctx *c, int id, int ng, int nc, int fdir, int sng = -1)
{
    DField &f = c->f[id];
    f.ng = ng; f.nc = nc; f.fdir = fdir;
    const bool padded = sng >= 0;
    if (sng < 0) sng = ng;
    for (int d = 0; d < 3; ++d) {
        f.ngd[d] = d < c->dim ? sng : 0;
        f.ext[d] = d < c->dim ? c->geo.n[d] + 2 * sng + ((d == fdir && !padded) ? 1 : 0) : 1;
    }
    f.sy = f.ext[0]; f.sz = (long)f.ext[0] * f.ext[1]; f.cs = f.sz * f.ext[2];
    f.bytes = (size_t)f.cs * nc * sizeof(double);
    VDN_REQUIRE(f.cs * (long)std::max(nc, 3) < (1L << 31), "field too large for 32-bit element offsets (2^31 elements)");
    VDN_CUDA(cudaMalloc(&f.base, f.bytes));
    VDN_CUDA(cudaMemsetAsync(f.base, 0, f.bytes, c->stream));
}

static void ctx_build(vdn_ctx *c, const vdn_params *prm, int dim, int nboxes, const int *box_lo, const int *box_hi,
                      const int *dom_lo, const int *dom_hi, const int *phys_bc, const double *dx, int device)
{
    VDN_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    VDN_REQUIRE(nboxes >= 1, "need at least one box");
    VDN_REQUIRE(prm->nscal >= 1 && prm->nscal <= 8, "nscal out of range");
    VDN_REQUIRE(prm->stencil_order == 2, "only stencil_order = 2 is implemented");
    VDN_REQUIRE(prm->slope_order == 0 || prm->slope_order == 2 || prm->slope_order == 4, "slope_order must be 0, 2 or 4");
    c->prm = *prm; c->dim = dim; c->device = device; c->nboxes = nboxes;
    int ndev = 0;
    VDN_CUDA(cudaGetDeviceCount(&ndev));
    VDN_REQUIRE(ndev > 0, "no CUDA device: the hot path has no CPU fallback");
    VDN_REQUIRE(device >= 0 && device < ndev, "bad device ordinal");
    VDN_CUDA(cudaSetDevice(device));
    VDN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int d = 0; d < 3; ++d) {
        c->dom_lo[d] = d < dim ? dom_lo[d] : 0; c->dom_hi[d] = d < dim ? dom_hi[d] : 0;
        c->dom_bc[d][0] = d < dim ? phys_bc[d * 2] : BC_INTERIOR; c->dom_bc[d][1] = d < dim ? phys_bc[d * 2 + 1] : BC_INTERIOR;
        c->rlo[d] = 1 << 30; c->rhi[d] = -(1 << 30);
    }
    c->box_lo.resize(nboxes); c->box_hi.resize(nboxes);
    long cells = 0;
    for (int b = 0; b < nboxes; ++b) {
        long bc = 1;
        for (int d = 0; d < 3; ++d) {
            c->box_lo[b][d] = d < dim ? box_lo[b * 3 + d] : 0; c->box_hi[b][d] = d < dim ? box_hi[b * 3 + d] : 0;
            VDN_REQUIRE(c->box_hi[b][d] >= c->box_lo[b][d], "empty box");
            c->rlo[d] = std::min(c->rlo[d], c->box_lo[b][d]); c->rhi[d] = std::max(c->rhi[d], c->box_hi[b][d]);
            bc *= c->box_hi[b][d] - c->box_lo[b][d] + 1;
        }
        cells += bc;
    }
    Geo &g = c->geo;
    memset(&g, 0, sizeof g);
    g.dim = dim;
    long rc = 1;
    for (int d = 0; d < 3; ++d) { g.n[d] = c->rhi[d] - c->rlo[d] + 1; rc *= g.n[d]; g.h[d] = d < dim ? dx[d] : 1.0; }
    VDN_REQUIRE(rc == cells, "the local boxes must tile their bounding box (tensor-product layout)");
    // tensor-product cuts
    for (int d = 0; d < 3; ++d) {
        std::vector<int> cuts;
        for (int b = 0; b < nboxes; ++b) cuts.push_back(c->box_lo[b][d] - c->rlo[d]);
        std::sort(cuts.begin(), cuts.end()); cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        VDN_REQUIRE((int)cuts.size() <= VDN_MAXCUT, "too many boxes per direction on one rank");
        g.nb[d] = (int)cuts.size();
        for (int q = 0; q < g.nb[d]; ++q) g.cut[d][q] = cuts[q];
        g.cut[d][g.nb[d]] = g.n[d];
    }
    VDN_REQUIRE(g.nb[0] * g.nb[1] * g.nb[2] == nboxes, "boxes are not a tensor-product chop of the region");
    for (int b = 0; b < nboxes; ++b) for (int d = 0; d < 3; ++d) {
        int l = c->box_lo[b][d] - c->rlo[d], h = c->box_hi[b][d] - c->rlo[d];
        int q = g.box1(d, l);
        VDN_REQUIRE(g.cut[d][q] == l && g.cut[d][q + 1] == h + 1, "boxes are not a tensor-product chop of the region");
    }
    // region-face BCs (single rank: the region is the domain; vdn_ctx_set_comm refines this)
    for (int d = 0; d < 3; ++d) {
        for (int s = 0; s < 2; ++s) {
            bool at_dom = s == 0 ? (c->rlo[d] == c->dom_lo[d]) : (c->rhi[d] == c->dom_hi[d]);
            g.pbc[d][s] = (d < dim && at_dom) ? c->dom_bc[d][s] : BC_INTERIOR;
        }
        bool per = d < dim && c->dom_bc[d][0] == BC_PERIODIC;
        if (per) VDN_REQUIRE(c->dom_bc[d][1] == BC_PERIODIC, "periodic BC must be set on both sides");
        c->wrap[d] = per && c->rlo[d] == c->dom_lo[d] && c->rhi[d] == c->dom_hi[d];
        if (d < dim) VDN_REQUIRE(g.n[d] >= 4, "region must be at least 4 cells wide in every direction");
    }
    build_bc_tables(c);

    const int dm = dim, ns = prm->nscal;
    alloc_field(c, VDN_UOLD, 3, dm, -1); alloc_field(c, VDN_SOLD, 3, ns, -1);
    alloc_field(c, VDN_UNEW, 3, dm, -1); alloc_field(c, VDN_SNEW, 3, ns, -1);
    alloc_field(c, VDN_GP, 1, dm, -1);
    alloc_field(c, VDN_EXT_VEL_FORCE, 1, dm, -1); alloc_field(c, VDN_EXT_SCAL_FORCE, 1, ns, -1);
    alloc_field(c, VDN_LAPU, 0, dm, -1);
    for (int d = 0; d < dm; ++d) alloc_field(c, VDN_UMAC_X + d, 1, 1, d);
    alloc_field(c, VDN_MAC_RHS, 1, 1, -1); alloc_field(c, VDN_RHOHALF, 1, 1, -1);
    alloc_field(c, VDN_VEL_FORCE, 1, dm, -1); alloc_field(c, VDN_SCAL_FORCE, 1, ns, -1);
    alloc_field(c, VDN_RH, 0, 1, -1, MG_PAD); alloc_field(c, VDN_PHI, 1, 1, -1, MG_PAD);
    for (int d = 0; d < dm; ++d) alloc_field(c, VDN_BETA_X + d, 0, 1, d, MG_PAD);
    // edge states: the scalar and velocity phases never overlap, so SEDGE aliases the first nscal comps of UEDGE
    const int ne = std::max(dm, ns);
    for (int d = 0; d < dm; ++d) {
        alloc_field(c, VDN_UEDGE_X + d, 0, ne, d);
        c->f[VDN_SEDGE_X + d] = c->f[VDN_UEDGE_X + d]; c->f[VDN_SEDGE_X + d].nc = ns;
        c->f[VDN_UEDGE_X + d].nc = dm;
        alloc_field(c, VDN_SFLUX_X + d, 0, 1, d);        // only the conservative comp (density) carries a flux
    }
    // Godunov scratch arena of the staged 2-D kernels (S-layout: cells -1..n, faces 0..n); the 3-D plane-marching kernels keep
    // every intermediate on chip and need none
    if (dim == 2) {
        c->nscr = 36;
        c->s_sy = g.n[0] + 2; c->s_sz = (long)c->s_sy * (g.n[1] + 2);
        c->s_n = c->s_sz;
        c->s_off = 1 + c->s_sy;
        VDN_CUDA(cudaMalloc(&c->scratch, sizeof(double) * c->s_n * c->nscr));
        VDN_CUDA(cudaMemsetAsync(c->scratch, 0, sizeof(double) * c->s_n * c->nscr, c->stream));
    }
    VDN_CUDA(cudaMalloc(&c->d_eps, sizeof(double) * std::max(nboxes, 1)));
    VDN_CUDA(cudaMalloc(&c->d_red, sizeof(double) * 64));
    VDN_CUDA(cudaMallocHost(&c->h_pin, sizeof(double) * 64));
    // umac = 1.d20 (advance_timestep.f90:76-77)
    for (int d = 0; d < dm; ++d) st_setval(c, VDN_UMAC_X + d, 1.0e20);
    VDN_CUDA(cudaStreamSynchronize(c->stream));
}

static void ctx_free(vdn_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->mg) mg_destroy(c->mg);
    if (c->mgh) mg_destroy(c->mgh);
    if (c->comm) comm_destroy(c->comm);
    for (int i = 0; i < VDN_NFIELDS; ++i) {
        if (i >= VDN_SEDGE_X && i <= VDN_SEDGE_Z) continue;      // aliases UEDGE
        if (c->f[i].base && !c->f[i].in_heap) cudaFree(c->f[i].base);
    }
    if (c->scratch) cudaFree(c->scratch);
    if (c->d_eps) cudaFree(c->d_eps);
    if (c->d_red) cudaFree(c->d_red);
    if (c->d_dbg) cudaFree(c->d_dbg);
    for (int q = 0; q < 4; ++q) if (c->xstage[q]) cudaFree(c->xstage[q]);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->stage) cudaFreeHost(c->stage);
    for (auto &p : c->prof) for (auto &pr : p.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h); for (int i = 0; i < VDN_NFIELDS; ++i) { cudaEventDestroy(c->ev_up[i]); cudaEventDestroy(c->ev_fin[i]); } }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// copy between a host box array and the region array
static void box_copy(vdn_ctx *c, int field, int ibox, double *host, int ng, int ncomp, bool upload, cudaStream_t stream = nullptr, bool wait = true)
{
    if (!stream) stream = c->stream;
    VDN_REQUIRE(field >= 0 && field < VDN_NFIELDS, "bad field id");
    VDN_REQUIRE(ibox >= 0 && ibox < c->nboxes, "bad box index");
    DField &f = c->f[field];
    VDN_REQUIRE(f.base != nullptr, "field not allocated for this dimension");
    VDN_REQUIRE(ng == f.ng && ncomp == f.nc, "host (ng, ncomp) does not match the field's fixed layout");
    if (upload && field >= VDN_UMAC_X && field <= VDN_UMAC_Z) ++c->umac_epoch;
    if (upload && field == VDN_LAPU) c->lapu_set = true;
    VDN_CUDA(cudaSetDevice(c->device));
    int hext[3], clo[3], chi[3], hlo[3];
    for (int d = 0; d < 3; ++d) {
        const int lo = c->box_lo[ibox][d], hi = c->box_hi[ibox][d];
        if (d >= c->dim) { hext[d] = 1; clo[d] = 0; chi[d] = 0; hlo[d] = 0; continue; }
        const int nod = (d == f.fdir) ? 1 : 0;
        hext[d] = hi - lo + 1 + 2 * ng + nod; hlo[d] = lo - ng;
        // upload: the box's valid cells plus the ghost cells that lie outside the region's valid area (inner ghosts are other boxes'
        // valid cells).  download: the whole ghosted extent -- the region array holds the neighbour boxes' valid cells and the filled
        // outer ghosts there, which is what ml_restrict_and_fill leaves in the reference's per-box ghosts (update.f90:103-107).
        clo[d] = lo - ((!upload || lo == c->rlo[d]) ? ng : 0);
        chi[d] = hi + nod + ((!upload || hi == c->rhi[d]) ? ng : 0);
    }
    // one box covering the region, stored with the host's own ghost width: host and device layouts coincide -> one flat copy
    bool flat = c->nboxes == 1;
    for (int d = 0; d < c->dim; ++d) if (f.ngd[d] != ng || f.ext[d] != hext[d]) flat = false;
    if (flat) {
        const size_t nb = sizeof(double) * (size_t)f.cs * ncomp;      // not f.bytes: SEDGE_* alias the (larger) UEDGE_* storage
        if (upload) VDN_CUDA(cudaMemcpyAsync(f.base, host, nb, cudaMemcpyHostToDevice, stream));
        else        VDN_CUDA(cudaMemcpyAsync(host, f.base, nb, cudaMemcpyDeviceToHost, stream));
        if (!upload && wait) VDN_CUDA(cudaStreamSynchronize(stream));
        return;
    }
    // several boxes per region (or a different ghost width): the host box travels as ONE flat copy through a device staging buffer and a
    // kernel scatters / gathers its rows into / out of the region array.  (cudaMemcpy3DAsync moves 2 KB rows at ~21 GB/s; a flat copy of
    // pinned memory runs at the PCIe rate, ~55 GB/s, and the device-side pass costs ~2 % of it.)
    const size_t hcount = (size_t)hext[0] * hext[1] * hext[2] * ncomp;
    const int slot = (upload ? 0 : 2) + (stream == c->stream ? 0 : 1);       // one staging buffer per (direction, stream): stream order protects it
    if (c->xstage_bytes[slot] < hcount * sizeof(double)) {
        VDN_CUDA(cudaStreamSynchronize(stream));
        if (c->xstage[slot]) VDN_CUDA(cudaFree(c->xstage[slot]));
        // sized at once for the largest box of any field (widest ghost zone, most components, face-centred): no regrowth from field to field
        int ngmax = 0, ncmax = 1;
        for (int i = 0; i < VDN_NFIELDS; ++i) if (c->f[i].base) { ngmax = std::max(ngmax, c->f[i].ng); ncmax = std::max(ncmax, c->f[i].nc); }
        size_t big = 0;
        for (int b = 0; b < c->nboxes; ++b) {
            size_t v = sizeof(double) * ncmax;
            for (int d = 0; d < c->dim; ++d) v *= (size_t)(c->box_hi[b][d] - c->box_lo[b][d] + 2 + 2 * ngmax);
            big = std::max(big, v);
        }
        const size_t want = std::max(hcount * sizeof(double), big);
        VDN_CUDA(cudaMalloc(&c->xstage[slot], want));
        c->xstage_bytes[slot] = want;
    }
    double *stg = c->xstage[slot];
    BoxCopyArgs a;
    a.stage = stg; a.base = f.base; a.cs = f.cs; a.ncomp = ncomp; a.upload = upload ? 1 : 0;
    for (int d = 0; d < 3; ++d) {
        a.hext[d] = hext[d]; a.hofs[d] = clo[d] - hlo[d]; a.n[d] = chi[d] - clo[d] + 1;
        a.dofs[d] = d < c->dim ? clo[d] - c->rlo[d] + f.ngd[d] : 0;
    }
    a.dext0 = f.ext[0]; a.dext1 = f.ext[1];
    if (upload) {
        VDN_CUDA(cudaMemcpyAsync(stg, host, hcount * sizeof(double), cudaMemcpyHostToDevice, stream));
        st_box_copy(a, stream);
    } else {
        st_box_copy(a, stream);
        VDN_CUDA(cudaMemcpyAsync(host, stg, hcount * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (wait) VDN_CUDA(cudaStreamSynchronize(stream));
    }
}

// one pass of the hot path, advance_timestep.f90:95-124
// vdn_advance_host hooks: a stage waits for the upload of its own inputs; an output starts its way back as soon as it is final
static void io_need(vdn_ctx *c, int field) { if (c->hio) VDN_CUDA(cudaStreamWaitEvent(c->stream, c->ev_up[field], 0)); }
static void io_final(vdn_ctx *c, int field, double *const *host)
{
    if (!c->hio) return;
    VDN_CUDA(cudaEventRecord(c->ev_fin[field], c->stream));
    VDN_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_fin[field], 0));
    for (int b = 0; b < c->nboxes; ++b) box_copy(c, field, b, host[b], c->f[field].ng, c->f[field].nc, false, c->s_d2h, false);
}

// A region that does not own a whole direction of the domain needs its neighbours: without a communicator the rank-boundary ghost
// cells would silently stay stale (st_fill_boundary / the multigrid halo do nothing).
void ctx_require_comm(const vdn_ctx *c)
{
    if (c->comm) return;
    for (int d = 0; d < c->dim; ++d) {
        const bool whole = c->rlo[d] == c->dom_lo[d] && c->rhi[d] == c->dom_hi[d];
        VDN_REQUIRE(whole, "the context's region is a strict part of the domain: call vdn_ctx_set_comm (one context per rank) before any stage");
    }
}

static void advance_impl(vdn_ctx *c, double dt, double mac_rel_eps, int *cycles, double *resnorm)
{
    ctx_require_comm(c);
    // the explicit viscous term Lu is computed by the reference Fortran (get_explicit_diffusive_term, advance_timestep.f90:84-88)
    // and must have been handed over; it is never computed here
    VDN_REQUIRE(c->prm.visc_coef == 0.0 || c->lapu_set, "visc_coef > 0 needs the LAPU field (vdn_field_upload / vdn_host_state.lapu)");
    // advance_timestep.f90:76-77 builds umac = 1.d20 every step; here the 1.d20 poison is set once at context creation:
    // the only faces that keep it (ghost faces outside non-periodic boundaries) are never written afterwards.
    // advance_premac (advance_premac.f90:44-51)
    // phase brackets = the reference's own timers (advance_timestep.f90:97-131, printed at :160-164)
    int rc;
    {
        LaunchScope ph(c, "phase:advance_premac", 0.0, 0);
        io_need(c, VDN_EXT_VEL_FORCE); io_need(c, VDN_GP); io_need(c, VDN_SOLD);
        if (c->hio && c->hio->lapu) io_need(c, VDN_LAPU);
        st_mkvelforce(c, VDN_SOLD, 1.0);
        io_need(c, VDN_UOLD);
        st_velpred(c, dt);
    }
    {
        // macproject (macproject.f90:20-133)
        LaunchScope ph(c, "phase:MAC_Project", 0.0, 0);
        if (c->hio && c->hio->mac_rhs) io_need(c, VDN_MAC_RHS);
        const double bn = st_divumac(c, true);          // rh and |rh|_inf in one pass
        st_mk_mac_coeffs(c);
        st_setval(c, VDN_PHI, 0.0);
        rc = st_mac_solve(c, mac_rel_eps > 0 ? mac_rel_eps : 1.0e-10, -1.0, cycles, resnorm, true, bn);
        st_mkumac(c);
    }
    {
        // scalar_advance (scalar_advance.f90:96-119)
        LaunchScope ph(c, "phase:Scalar_update", 0.0, 0);
        io_need(c, VDN_EXT_SCAL_FORCE);
        st_mkscalforce(c, 1.0);
        st_mkflux(c, 0, dt);
        st_mkscalforce(c, 0.0);
        st_update(c, 0, dt);
    }
    if (c->hio) io_final(c, VDN_SNEW, c->hio->snew);
    {
        // make_at_halftime (advance_timestep.f90:114)
        LaunchScope ph(c, "phase:make_at_halftime", 0.0, 0);
        st_make_at_halftime(c);
    }
    if (c->hio) io_final(c, VDN_RHOHALF, c->hio->rhohalf);
    {
        // velocity_advance (velocity_advance.f90:70-93)
        LaunchScope ph(c, "phase:Velocity_update", 0.0, 0);
        // velocity_advance.f90:70 evaluates mkvelforce(rho^n, visc_fac = 1) again; VEL_FORCE still holds exactly that from advance_premac
        // (same ext_vel_force, gp, sold, lapu; nothing in between writes it), so the fused path does not recompute it
        st_mkflux(c, 1, dt);
        st_mkvelforce(c, VDN_RHOHALF, 0.0);
        st_update(c, 1, dt);
    }
    if (c->hio) io_final(c, VDN_UNEW, c->hio->unew);
    if (rc != 0) throw VdnError("MAC multigrid did not converge within mg_max_cycles");
}

// ------------------------------------------------------------------------------------------
// extern "C" ABI
// ------------------------------------------------------------------------------------------
#define VDN_TRY(ctx, body) \
    if (!(ctx)) return 1; \
    try { VDN_CUDA(cudaSetDevice((ctx)->device)); body; return 0; } \
    catch (const std::exception &e) { (ctx)->err = e.what(); return 1; } \
    catch (...) { (ctx)->err = "unknown error"; return 1; }

extern "C" {

void vdn_params_default(vdn_params *p)
{
    memset(p, 0, sizeof *p);
    p->nscal = 2; p->slope_order = 4; p->use_minion = 0; p->boussinesq = 0; p->stencil_order = 2;
    p->mg_verbose = 0; p->mg_nu1 = 2; p->mg_nu2 = 2; p->mg_max_cycles = 100; p->mg_max_bottom_iter = 100;
    p->mg_bottom_eps = 1.0e-3; p->visc_coef = 0.0; p->diff_coef = 0.0;
}

int vdn_ctx_create(const vdn_params *prm, int dim, int nboxes, const int *box_lo, const int *box_hi,
                   const int *dom_lo, const int *dom_hi, const int *phys_bc, const double *dx, int device, vdn_ctx **out)
{
    if (!out) return 1;
    *out = nullptr;
    vdn_ctx *c = new vdn_ctx();
    try { ctx_build(c, prm, dim, nboxes, box_lo, box_hi, dom_lo, dom_hi, phys_bc, dx, device); *out = c; return 0; }
    catch (const std::exception &e) { g_create_err = e.what(); ctx_free(c); return 1; }
}
void vdn_ctx_destroy(vdn_ctx *ctx) { ctx_free(ctx); }
int vdn_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
const char *vdn_last_error(const vdn_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int vdn_field_upload(vdn_ctx *ctx, int field, int ibox, const double *host, int ng, int ncomp)
{ VDN_TRY(ctx, box_copy(ctx, field, ibox, const_cast<double *>(host), ng, ncomp, true)) }
int vdn_field_download(vdn_ctx *ctx, int field, int ibox, double *host, int ng, int ncomp)
{ VDN_TRY(ctx, box_copy(ctx, field, ibox, host, ng, ncomp, false)) }
int vdn_field_setval(vdn_ctx *ctx, int field, double val)
{ VDN_TRY(ctx, { VDN_REQUIRE(field >= 0 && field < VDN_NFIELDS && ctx->f[field].base, "bad field id"); st_setval(ctx, field, val); }) }
int vdn_sync(vdn_ctx *ctx) { VDN_TRY(ctx, VDN_CUDA(cudaStreamSynchronize(ctx->stream))) }
void *vdn_get_stream(vdn_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int vdn_fill_boundary(vdn_ctx *ctx, int field)
{ VDN_TRY(ctx, { VDN_REQUIRE(field >= 0 && field < VDN_NFIELDS && ctx->f[field].base, "bad field id"); st_fill_boundary(ctx, field); }) }
int vdn_fill_and_physbc(vdn_ctx *ctx, int field, int bccomp, int same_boundary)
{ VDN_TRY(ctx, { VDN_REQUIRE(field >= 0 && field < VDN_NFIELDS && ctx->f[field].base, "bad field id");
                 VDN_REQUIRE(bccomp >= 0 && bccomp + (same_boundary ? 1 : ctx->f[field].nc) <= ctx->dim + ctx->prm.nscal + 2, "bccomp out of range");
                 st_fill_boundary(ctx, field); st_physbc(ctx, field, bccomp, same_boundary != 0); }) }

int vdn_mkvelforce(vdn_ctx *ctx, int rho_field, double visc_fac) { VDN_TRY(ctx, st_mkvelforce(ctx, rho_field, visc_fac)) }
int vdn_mkscalforce(vdn_ctx *ctx, double diff_fac) { VDN_TRY(ctx, st_mkscalforce(ctx, diff_fac)) }
int vdn_velpred(vdn_ctx *ctx, double dt) { VDN_TRY(ctx, st_velpred(ctx, dt)) }
int vdn_mkflux(vdn_ctx *ctx, int is_vel, double dt) { VDN_TRY(ctx, st_mkflux(ctx, is_vel, dt)) }
int vdn_update(vdn_ctx *ctx, int is_vel, double dt) { VDN_TRY(ctx, st_update(ctx, is_vel, dt)) }
int vdn_make_at_halftime(vdn_ctx *ctx) { VDN_TRY(ctx, st_make_at_halftime(ctx)) }

int vdn_divumac(vdn_ctx *ctx, double *rhmax)
{ VDN_TRY(ctx, { double v = st_divumac(ctx, rhmax != nullptr); if (rhmax) *rhmax = v; }) }
int vdn_mk_mac_coeffs(vdn_ctx *ctx) { VDN_TRY(ctx, st_mk_mac_coeffs(ctx)) }
int vdn_mkumac(vdn_ctx *ctx) { VDN_TRY(ctx, st_mkumac(ctx)) }

int vdn_mac_solve(vdn_ctx *ctx, double rel_eps, double abs_eps, int *ncycles, double *resnorm)
{
    if (!ctx) return 1;
    try {
        VDN_CUDA(cudaSetDevice(ctx->device));
        int rc = st_mac_solve(ctx, rel_eps > 0 ? rel_eps : 1.0e-10, abs_eps, ncycles, resnorm);
        if (rc) { ctx->err = "MAC multigrid did not converge within mg_max_cycles"; return 2; }
        return 0;
    } catch (const std::exception &e) { ctx->err = e.what(); return 1; }
}

int vdn_macproject(vdn_ctx *ctx, double rel_eps, double abs_eps, int *ncycles, double *resnorm)
{
    if (!ctx) return 1;
    try {
        VDN_CUDA(cudaSetDevice(ctx->device));
        // macproject.f90:60-67 computes umac_norm only for the (overridden) abs tolerance; skipped like the reference's "HACK"
        const double bn = st_divumac(ctx, true);
        st_mk_mac_coeffs(ctx);
        st_setval(ctx, VDN_PHI, 0.0);
        int rc = st_mac_solve(ctx, rel_eps > 0 ? rel_eps : 1.0e-10, abs_eps, ncycles, resnorm, true, bn);
        st_mkumac(ctx);
        if (rc) { ctx->err = "MAC multigrid did not converge within mg_max_cycles"; return 2; }
        return 0;
    } catch (const std::exception &e) { ctx->err = e.what(); return 1; }
}

int vdn_advance(vdn_ctx *ctx, double dt, double mac_rel_eps, int *mac_cycles, double *mac_resnorm)
{ VDN_TRY(ctx, advance_impl(ctx, dt, mac_rel_eps, mac_cycles, mac_resnorm)) }

static void advance_host_impl(vdn_ctx *c, double dt, double mac_rel_eps, const vdn_host_state *hs, int *cycles, double *resnorm)
{
    VDN_REQUIRE(hs && hs->uold && hs->sold && hs->gp && hs->ext_vel_force && hs->ext_scal_force && hs->unew && hs->snew && hs->rhohalf,
                "vdn_advance_host: every host pointer array must be given");
    if (!c->s_h2d) {
        VDN_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        VDN_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < VDN_NFIELDS; ++i) {
            VDN_CUDA(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
            VDN_CUDA(cudaEventCreateWithFlags(&c->ev_fin[i], cudaEventDisableTiming));
        }
    }
    // the previous pass (and whatever the caller enqueued on the context's stream) must be done with the input fields
    VDN_CUDA(cudaEventRecord(c->ev_fin[VDN_UOLD], c->stream));
    VDN_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_fin[VDN_UOLD], 0));
    VDN_REQUIRE(c->prm.visc_coef == 0.0 || hs->lapu, "vdn_advance_host: visc_coef > 0 needs vdn_host_state.lapu");
    const struct { int field; const double *const *host; } up[7] = {
        { VDN_EXT_VEL_FORCE, hs->ext_vel_force }, { VDN_GP, hs->gp }, { VDN_SOLD, hs->sold }, { VDN_LAPU, hs->lapu }, { VDN_UOLD, hs->uold },
        { VDN_MAC_RHS, hs->mac_rhs }, { VDN_EXT_SCAL_FORCE, hs->ext_scal_force } };
    for (const auto &u : up) {
        if (!u.host) continue;                       // optional inputs (lapu, mac_rhs)
        for (int b = 0; b < c->nboxes; ++b)
            box_copy(c, u.field, b, const_cast<double *>(u.host[b]), c->f[u.field].ng, c->f[u.field].nc, true, c->s_h2d);
        VDN_CUDA(cudaEventRecord(c->ev_up[u.field], c->s_h2d));
    }
    c->hio = hs;
    try { advance_impl(c, dt, mac_rel_eps, cycles, resnorm); }
    catch (...) { c->hio = nullptr; cudaStreamSynchronize(c->s_d2h); throw; }
    c->hio = nullptr;
    VDN_CUDA(cudaStreamSynchronize(c->s_d2h));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
}

int vdn_advance_host(vdn_ctx *ctx, double dt, double mac_rel_eps, const vdn_host_state *hs, int *mac_cycles, double *mac_resnorm)
{ VDN_TRY(ctx, advance_host_impl(ctx, dt, mac_rel_eps, hs, mac_cycles, mac_resnorm)) }

int vdn_estdt(vdn_ctx *ctx, double dtold, double cflfac, double max_dt_growth, double *dt)
{ VDN_TRY(ctx, { ctx_require_comm(ctx); VDN_REQUIRE(dt != nullptr, "dt is null"); *dt = st_estdt(ctx, dtold, cflfac, max_dt_growth); }) }

int vdn_field_copy(vdn_ctx *ctx, int dst_field, int src_field)
{ VDN_TRY(ctx, { VDN_REQUIRE(dst_field >= 0 && dst_field < VDN_NFIELDS && src_field >= 0 && src_field < VDN_NFIELDS, "bad field id");
                 st_field_copy(ctx, dst_field, src_field); }) }

int vdn_visc_solve(vdn_ctx *ctx, double mu, int diffusion_type, int *ncycles, double *resnorm)
{ VDN_TRY(ctx, { int rc = st_visc_solve(ctx, mu, diffusion_type, ncycles, resnorm);
                 if (rc) { ctx->err = "Helmholtz multigrid (visc_solve) did not converge within mg_max_cycles"; return 2; } }) }

int vdn_diff_scalar_solve(vdn_ctx *ctx, double mu, int icomp, int diffusion_type, int *ncycles, double *resnorm)
{ VDN_TRY(ctx, { int rc = st_diff_scalar_solve(ctx, mu, icomp, diffusion_type, ncycles, resnorm);
                 if (rc) { ctx->err = "Helmholtz multigrid (diff_scalar_solve) did not converge within mg_max_cycles"; return 2; } }) }

int vdn_mg_tune(vdn_ctx *ctx, int fuse_min, int tile)
{ VDN_TRY(ctx, { VDN_REQUIRE(tile >= -1 && tile < 5, "tile shape out of range");
                 VDN_CUDA(cudaStreamSynchronize(ctx->stream));
                 if (ctx->mg) { mg_destroy(ctx->mg); ctx->mg = nullptr; }
                 ctx->mg_fuse_min = fuse_min; ctx->mg_tile_force = tile; }) }

int vdn_comm_tune(vdn_ctx *ctx, int force_nccl)
{ VDN_TRY(ctx, { VDN_REQUIRE(!ctx->comm, "vdn_comm_tune must be called before vdn_ctx_set_comm"); VDN_REQUIRE(force_nccl >= 0 && force_nccl <= 3, "transport mode out of range"); ctx->comm_mode = force_nccl; }) }

// measurement hook: 32 device counters (the fused smoother's flag-wait accounting: 4 per kernel family -- ns waited, max ns, waits, unused);
// the first call allocates them (so the kernels start counting), every call returns and resets them
int vdn_debug_counters(vdn_ctx *ctx, unsigned long long *out32)
{ VDN_TRY(ctx, {
      VDN_CUDA(cudaStreamSynchronize(ctx->stream));
      if (!ctx->d_dbg) { VDN_CUDA(cudaMalloc(&ctx->d_dbg, 32 * 8)); VDN_CUDA(cudaMemset(ctx->d_dbg, 0, 32 * 8)); }
      if (out32) VDN_CUDA(cudaMemcpy(out32, ctx->d_dbg, 32 * 8, cudaMemcpyDeviceToHost));
      VDN_CUDA(cudaMemset(ctx->d_dbg, 0, 32 * 8)); }) }

int vdn_prof_enable(vdn_ctx *ctx, int on)
{ VDN_TRY(ctx, { prof_collect(ctx); ctx->prof.clear(); ctx->prof_idx.clear(); ctx->prof_on = on != 0; }) }
int vdn_prof_count(vdn_ctx *ctx) { if (!ctx) return 0; prof_collect(ctx); return (int)ctx->prof.size(); }
int vdn_prof_get(vdn_ctx *ctx, int idx, char *name, long long *launches, double *ms, double *alg_bytes)
{
    if (!ctx || idx < 0 || idx >= (int)ctx->prof.size()) return 1;
    prof_collect(ctx);
    const ProfEntry &p = ctx->prof[idx];
    if (name) { strncpy(name, p.name.c_str(), 63); name[63] = 0; }
    if (launches) *launches = p.launches;
    if (ms) *ms = p.ms;
    if (alg_bytes) *alg_bytes = p.bytes;
    return 0;
}
long long vdn_launch_count(vdn_ctx *ctx) { return ctx ? ctx->launches : 0; }
long long vdn_comm_bytes(vdn_ctx *ctx) { return ctx ? ctx->comm_bytes : 0; }

} // extern "C"
