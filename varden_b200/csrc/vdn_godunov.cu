// vdn_godunov.cu -- Godunov edge-state predictor on the device (compiled with -fmad=false).
//
// Replaces, for the rank's merged region:
//   slope.f90     slopex_2d :148 / slopey_2d :291 / slopez_3d :437
//   velpred.f90   velpred_3d :1776-2765 (production form incl. its hi-x OUTLET quirk :2075), velpred_2d :125-524
//   mkflux.f90    mkflux_3d :1186-2567, mkflux_2d :152-691
// 3-D: ONE plane-marching kernel per routine (vdn_godunov_march.cuh): a CTA marches a tile of columns along z with every
// intermediate in registers / shared memory, so a field is read once and the edge states are written once -- no scratch arena.
// 2-D (config 1, tiny): direction-generic staged kernels (vdn_godunov_kernels.cuh), one stage per launch, intermediates in a small
// "S-layout" arena (cells -1..n, faces 0..n).  Operation order follows the reference line by line, so both are bit-comparable
// with the CPU oracle.  eps is per reference box.
#include "vdn_ctx.h"

#include "vdn_godunov_kernels.cuh"
#include "vdn_godunov_march.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// per-box eps: 1e-8 * max|.| over the box's valid cells (velpred) / valid faces (mkflux)
// ------------------------------------------------------------------------------------------
struct MaxArgs { View v[3]; int nv; int ncomp; Range r[3]; double *out; };
__global__ void k_absmax_box(MaxArgs a)      // one block-group per box: grid.x blocks, atomicMax on the bit pattern
{
    double m = 0.0;
    for (int q = 0; q < a.nv; ++q) {
        const Range &r = a.r[q];
        const long nx = r.hi[0] - r.lo[0] + 1, ny = r.hi[1] - r.lo[1] + 1, nz = r.hi[2] - r.lo[2] + 1;
        const long tot = nx * ny * nz;
        for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
            const int i = r.lo[0] + (int)(t % nx), j = r.lo[1] + (int)((t / nx) % ny), k = r.lo[2] + (int)(t / (nx * ny));
            for (int c = 0; c < a.ncomp; ++c) m = fmax(m, fabs(a.v[q](i, j, k, c)));
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double sm[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sm[w] = m;
    __syncthreads();
    if (w == 0) {
        m = (l < (blockDim.x >> 5)) ? sm[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        // non-negative doubles order like their bit patterns
        if (l == 0) atomicMax((unsigned long long *)a.out, (unsigned long long)__double_as_longlong(m));
    }
}
__global__ void k_eps_finish(double *e, int n)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) { double um = e[t]; e[t] = (um == 0.0) ? 1.0e-8 : 1.0e-8 * um; }
}

const dim3 BLK(64, 4, 1);

void box_local(const vdn_ctx *c, int ib, int lo[3], int hi[3])
{
    for (int d = 0; d < 3; ++d) { lo[d] = c->box_lo[ib][d] - c->rlo[d]; hi[d] = c->box_hi[ib][d] - c->rlo[d]; }
}
// index of box ib in the eps table = geo.box(lo)
int box_slot(const vdn_ctx *c, int ib)
{
    int lo[3], hi[3]; box_local(c, ib, lo, hi);
    return c->geo.box(lo[0], lo[1], lo[2]);
}

void compute_eps(vdn_ctx *c, bool from_umac)
{
    // the scalar and the velocity mkflux of one step see the same projected umac: reduce it once
    if (from_umac && c->eps_epoch == c->umac_epoch) return;
    c->eps_epoch = from_umac ? c->umac_epoch : -1;
    LaunchScope ls(c, from_umac ? "eps_umac" : "eps_u", (double)c->ncells() * 8.0 * c->dim, c->nboxes + 2);
    VDN_CUDA(cudaMemsetAsync(c->d_eps, 0, sizeof(double) * c->nboxes, c->stream));
    for (int ib = 0; ib < c->nboxes; ++ib) {
        int lo[3], hi[3]; box_local(c, ib, lo, hi);
        MaxArgs a; a.out = c->d_eps + box_slot(c, ib);
        if (!from_umac) {
            a.nv = 1; a.ncomp = c->dim; a.v[0] = c->f[VDN_UOLD].view();
            a.r[0] = mk_range(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
        } else {
            a.nv = c->dim; a.ncomp = 1;
            for (int d = 0; d < c->dim; ++d) {
                a.v[d] = c->f[VDN_UMAC_X + d].view();
                a.r[d] = mk_range(lo[0], hi[0] + (d == 0), lo[1], hi[1] + (d == 1), lo[2], hi[2] + (d == 2));
            }
        }
        long tot = (long)(hi[0] - lo[0] + 2) * (hi[1] - lo[1] + 2) * (hi[2] - lo[2] + 2);
        int nblk = (int)std::min<long>(1184, (tot + 255) / 256);
        k_absmax_box<<<nblk, 256, 0, c->stream>>>(a);
    }
    k_eps_finish<<<cdiv(c->nboxes, 64), 64, 0, c->stream>>>(c->d_eps, c->nboxes);
    VDN_CUDA(cudaGetLastError());
}


// S-layout scratch slots
enum { SL0 = 0 /* slopes: 9 */, UL0 = 9, UR0 = 18, UI0 = 27, NSCR_NEEDED = 36 };
// mkflux scratch slots: per component
enum { MSL0 = 0 /* 3 slopes */, ML0 = 3, MR0 = 6, MI0 = 9, MX0 = 12 /* 9 slots, X[d][t] at d*3+t */, MF_NSCR = 21 };

// launcher handed to the stage orchestration in vdn_godunov_kernels.cuh: CUDA launch on the context's stream,
// profiling brackets from the context's CUDA-event profiler
struct CudaLauncher {
    vdn_ctx *c;
    LaunchScope scope(const char *name, double alg_bytes, int nlaunch) { return LaunchScope(c, name, alg_bytes, nlaunch); }
    template <class A> void operator()(void (*k)(A), const Range &r, const A &a) { k<<<grid3(r, BLK), BLK, 0, c->stream>>>(a); }
    // plane-marching kernels: explicit grid / block / dynamic shared memory
    template <class A> void run(void (*k)(A), dim3 grid, dim3 block, size_t smem, const A &a)
    {
        if (smem > 48 * 1024) VDN_CUDA(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, block, smem, c->stream>>>(a);
    }
    // resident CTAs of kernel k on this device (the z-chunking fills whole waves of them)
    template <class A> int slots(void (*k)(A), int nthreads, size_t smem)
    {
        static std::map<const void *, int> cache;
        auto it = cache.find((const void *)k);
        if (it != cache.end()) return it->second;
        if (smem > 48 * 1024) VDN_CUDA(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        VDN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k, nthreads, smem));
        VDN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        const int v = std::max(1, per_sm) * std::max(1, sms);
        cache[(const void *)k] = v;
        return v;
    }
};

void velpred_march_impl(vdn_ctx *c, double dt)
{
    compute_eps(c, false);
    View out[3];
    for (int d = 0; d < 3; ++d) out[d] = c->f[VDN_UMAC_X + d].view();
    CudaLauncher L{c};
    march::velpred_march(L, c->geo, c->f[VDN_UOLD].view(), c->f[VDN_VEL_FORCE].view(), out, c->d_eps, dt,
                         c->prm.slope_order, c->prm.use_minion, c->adv_bc, 0);
    VDN_CUDA(cudaGetLastError());
}

void mkflux_march_impl(vdn_ctx *c, int is_vel, double dt)
{
    // scalar_advance passes divu == 0 as mac_rhs (scalar_advance.f90:102) and no velocity component is conservative
    // (velocity_advance.f90:51), so the s*mac_rhs term of mkflux.f90:2340-2345 never contributes: x - dt2*s*0 == x exactly.
    compute_eps(c, true);
    const int ncomp = is_vel ? 3 : c->prm.nscal;
    View mac[3], sedge[3], flux[3];
    for (int d = 0; d < 3; ++d) {
        mac[d] = c->f[VDN_UMAC_X + d].view();
        sedge[d] = c->f[(is_vel ? VDN_UEDGE_X : VDN_SEDGE_X) + d].view();
        flux[d] = c->f[VDN_SFLUX_X + d].view();
    }
    CudaLauncher L{c};
    march::mkflux_march(L, c->geo, c->f[is_vel ? VDN_UOLD : VDN_SOLD].view(), c->f[is_vel ? VDN_VEL_FORCE : VDN_SCAL_FORCE].view(),
                        mac, sedge, flux, c->d_eps, dt, is_vel, ncomp, c->prm.slope_order, c->prm.use_minion,
                        c->adv_bc + (is_vel ? 0 : 3), 0);
    VDN_CUDA(cudaGetLastError());
}

template <int DIM>
void velpred_impl(vdn_ctx *c, double dt)
{
    compute_eps(c, false);
    VpArgs a; a.g = c->geo; a.u = c->f[VDN_UOLD].view(); a.force = c->f[VDN_VEL_FORCE].view();
    a.eps = c->d_eps; a.dt = dt; a.use_minion = c->prm.use_minion; a.order = c->prm.slope_order;
    for (int cc = 0; cc < 3; ++cc) for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) a.sbc[cc][d][s] = c->adv_bc[cc][d][s];
    for (int d = 0; d < 3; ++d) {
        const int q = d < DIM ? d : 0;
        a.sl[d] = c->S(SL0 + 3 * q); a.ul[d] = c->S(UL0 + 3 * q); a.ur[d] = c->S(UR0 + 3 * q); a.uimh[d] = c->S(UI0 + 3 * q);
        a.out[d] = c->f[VDN_UMAC_X + q].view();
        for (int t = 0; t < 3; ++t) a.X[d][t] = c->S(SL0 + d * 3 + t);     // transverse states reuse the slope slots
    }
    CudaLauncher L{c};
    velpred_stages<DIM>(L, a);
    VDN_CUDA(cudaGetLastError());
}

template <int DIM>
void mkflux_impl(vdn_ctx *c, int is_vel, double dt)
{
    const int ncomp = is_vel ? DIM : c->prm.nscal;
    const int bccomp = is_vel ? 0 : DIM;
    const DField &sfld = c->f[is_vel ? VDN_UOLD : VDN_SOLD];
    const DField &ffld = c->f[is_vel ? VDN_VEL_FORCE : VDN_SCAL_FORCE];
    compute_eps(c, true);
    // scalar_advance passes divu == 0 as mac_rhs (scalar_advance.f90:102) => the s*div(u) term is skipped (use_rhs = 0,
    // x - dt2*s*0 == x exactly); velocity_advance passes mac_rhs (:76) but no velocity comp is conservative.
    MfArgs a; a.g = c->geo; a.mac_rhs = c->f[VDN_MAC_RHS].view(); a.eps = c->d_eps; a.dt = dt;
    a.use_minion = c->prm.use_minion; a.is_vel = is_vel; a.use_rhs = is_vel; a.order = c->prm.slope_order;
    for (int d = 0; d < 3; ++d) {
        const int q = d < DIM ? d : 0;
        a.mac[d] = c->f[VDN_UMAC_X + q].view();
        a.sl[d] = c->S(MSL0 + q); a.l[d] = c->S(ML0 + q); a.rr[d] = c->S(MR0 + q); a.simh[d] = c->S(MI0 + q);
        for (int t = 0; t < 3; ++t) a.X[d][t] = c->S(MX0 + d * 3 + t);
    }
    CudaLauncher L{c};
    for (int comp = 0; comp < ncomp; ++comp) {
        a.cons = (!is_vel && comp == 0) ? 1 : 0;     // scalar_advance.f90:54-57, velocity_advance.f90:51
        a.s = sfld.view().comp(comp); a.force = ffld.view().comp(comp); a.comp = comp;
        for (int d = 0; d < 3; ++d) for (int sd = 0; sd < 2; ++sd) a.sbc[d][sd] = c->adv_bc[bccomp + comp][d][sd];
        for (int d = 0; d < 3; ++d) {
            const int q = d < DIM ? d : 0;
            a.sedge[d] = c->f[(is_vel ? VDN_UEDGE_X : VDN_SEDGE_X) + q].view().comp(comp);
            a.flux[d] = is_vel ? a.sedge[d] : c->f[VDN_SFLUX_X + q].view().comp(comp);
        }
        mkflux_stages<DIM>(L, a);
    }
    VDN_CUDA(cudaGetLastError());
}

} // namespace

void st_velpred(vdn_ctx *c, double dt)
{
    if (c->dim == 3) velpred_march_impl(c, dt);
    else { VDN_REQUIRE(c->nscr >= NSCR_NEEDED, "scratch arena too small for velpred"); velpred_impl<2>(c, dt); }
    ++c->umac_epoch;
    for (int d = 0; d < c->dim; ++d) st_fill_boundary(c, VDN_UMAC_X + d);      // velpred.f90:107-112
}

void st_mkflux(vdn_ctx *c, int is_vel, double dt)
{
    if (c->dim == 3) mkflux_march_impl(c, is_vel, dt);
    else { VDN_REQUIRE(c->nscr >= MF_NSCR, "scratch arena too small for mkflux"); mkflux_impl<2>(c, is_vel, dt); }
}
