// vdn_stream.cu -- streaming / glue kernels of the hot path (compiled with -fmad=false):
//   update_3d/_2d          update.f90:186-278, :113-184
//   mkvelforce/mkscalforce mkforce.f90:144-236, :333-402 (valid cells; ghosts by fill + FOEXTRAP, :75-76)
//   make_at_halftime       make_at_halftime.f90:95-115
//   divumac                macproject.f90:137-225, kernel :250-278
//   mk_mac_coeffs          macproject.f90:280-336, kernel :361-401
//   mkumac                 macproject.f90:403-505, kernel :578-645
//   multifab_fill_boundary (FBoxLib; periodic wrap inside the rank's region, NCCL halo between ranks)
//   multifab_physbc        multifab_physbc.f90:238-561
// All of them are pure HBM streaming: one thread per cell, i fastest (coalesced 256 B per warp row).
#include "vdn_ctx.h"
#include <type_traits>

namespace {

constexpr double HALF = 0.5, ZERO = 0.0, TWO = 2.0;
const dim3 BLK(64, 4, 1);

#define THREAD_IJK(r)                                                            \
    const int i = (r).lo[0] + blockIdx.x * blockDim.x + threadIdx.x;             \
    const int j = (r).lo[1] + blockIdx.y * blockDim.y + threadIdx.y;             \
    const int k = (r).lo[2] + blockIdx.z;                                        \
    if (i > (r).hi[0] || j > (r).hi[1]) return;

// ---------------- update ----------------
struct UpdArgs { Range r; int dim, ncomp, is_vel, cons_mask; View sold, snew, force, mac[3], sedge[3], flux[3]; double dt, dx[3]; };
// NC components per thread, every load issued (read-only path) before the first store so the memory system sees them all at once
template <int DIM, int NC, bool VEL>
__global__ void __launch_bounds__(256) k_update(UpdArgs a)
{
    THREAD_IJK(a.r)
    const long sx = 1, sy = a.mac[1].sy, sz = a.mac[2].sz;
    const double *m0 = &a.mac[0](i, j, k), *m1 = &a.mac[1](i, j, k), *m2 = &a.mac[DIM == 3 ? 2 : 0](i, j, k);
    const double u0 = __ldg(m0), u1 = __ldg(m0 + sx), v0 = __ldg(m1), v1 = __ldg(m1 + sy);
    const double w0 = DIM == 3 ? __ldg(m2) : ZERO, w1 = DIM == 3 ? __ldg(m2 + sz) : ZERO;
    double xl[NC], xh[NC], yl[NC], yh[NC], zl[NC], zh[NC], so[NC], fo[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const bool cons = !VEL && ((a.cons_mask >> c) & 1);
        const View &fx = cons ? a.flux[0] : a.sedge[0], &fy = cons ? a.flux[1] : a.sedge[1], &fz = cons ? a.flux[2] : a.sedge[2];
        const double *px = &fx(i, j, k, c), *py = &fy(i, j, k, c);
        xl[c] = __ldg(px); xh[c] = __ldg(px + 1); yl[c] = __ldg(py); yh[c] = __ldg(py + fy.sy);
        if (DIM == 3) { const double *pz = &fz(i, j, k, c); zl[c] = __ldg(pz); zh[c] = __ldg(pz + fz.sz); } else { zl[c] = ZERO; zh[c] = ZERO; }
        so[c] = __ldg(&a.sold(i, j, k, c)); fo[c] = __ldg(&a.force(i, j, k, c));
    }
    const double ubar = HALF * (u0 + u1);
    const double vbar = HALF * (v0 + v1);
    const double wbar = DIM == 3 ? HALF * (w0 + w1) : ZERO;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double adv;
        if (!VEL && ((a.cons_mask >> c) & 1)) {
            adv = (xh[c] - xl[c]) / a.dx[0] + (yh[c] - yl[c]) / a.dx[1];
            if (DIM == 3) adv = adv + (zh[c] - zl[c]) / a.dx[2];
        } else {
            adv = ubar * (xh[c] - xl[c]) / a.dx[0] + vbar * (yh[c] - yl[c]) / a.dx[1];
            if (DIM == 3) adv = adv + wbar * (zh[c] - zl[c]) / a.dx[2];
        }
        a.snew(i, j, k, c) = so[c] - a.dt * adv + a.dt * fo[c];
    }
}

// ---------------- forces ----------------
struct VfArgs { Range r; int dim, nscal_s, boussinesq; View vf, ext, gp, s, lapu; double visc; };
__global__ void k_mkvelforce(VfArgs a)
{
    THREAD_IJK(a.r)
    const double rho = a.s(i, j, k, 0);
    const double tr = a.nscal_s > 1 ? a.s(i, j, k, 1) : ZERO;
    for (int c = 0; c < a.dim; ++c) {
        double f = a.boussinesq == 1 ? tr * a.ext(i, j, k, c) : a.ext(i, j, k, c);
        double ll = a.visc * a.lapu(i, j, k, c);      // visc = visc_coef*visc_fac
        a.vf(i, j, k, c) = f + (ll - a.gp(i, j, k, c)) / rho;
    }
}
struct SfArgs { Range r; int nscal; View sf, ext; };
__global__ void k_mkscalforce(SfArgs a)        // laps == 0 on this path (diff_coef == 0): ext + 0
{
    THREAD_IJK(a.r)
    a.sf(i, j, k, 0) = ZERO;
    for (int c = 1; c < a.nscal; ++c) a.sf(i, j, k, c) = a.ext(i, j, k, c) + ZERO;
}
struct HtArgs { Range r; View rh, ro, rn; };
__global__ void k_halftime(HtArgs a)
{
    THREAD_IJK(a.r)
    a.rh(i, j, k) = HALF * (a.ro(i, j, k) + a.rn(i, j, k));
}

// ---------------- macproject glue ----------------
struct DivArgs { Range r; int dim; View mac[3], mrhs, rh; double dxinv[3]; double *nrm; };
template <int DIM>
__global__ void k_divumac(DivArgs a)
{
    const int i = a.r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.z;
    double v = ZERO;
    if (i <= a.r.hi[0] && j <= a.r.hi[1]) {
        double d = (a.mac[0](i + 1, j, k) - a.mac[0](i, j, k)) * a.dxinv[0]
                 + (a.mac[1](i, j + 1, k) - a.mac[1](i, j, k)) * a.dxinv[1];
        if (DIM == 3) d = d + (a.mac[2](i, j, k + 1) - a.mac[2](i, j, k)) * a.dxinv[2];
        v = d * (-1.0) + a.mrhs(i, j, k);
        a.rh(i, j, k) = v;
        v = fabs(v);
    }
    if (a.nrm) block_atomic_max(v, a.nrm);
}
struct CoefArgs { Range r; int d; View rho, beta; };
__global__ void k_mk_mac_coeffs(CoefArgs a)
{
    THREAD_IJK(a.r)
    const double *p = &a.rho(i, j, k);
    a.beta(i, j, k) = TWO / (p[0] + p[-a.rho.st(a.d)]);
}
struct UmacArgs { Range r; int d, n, bclo, bchi; View mac, phi, beta; double dx; };
__global__ void k_mkumac(UmacArgs a)
{
    THREAD_IJK(a.r)
    const int ix[3] = { i, j, k };
    const long st = a.phi.st(a.d);
    const double *p = &a.phi(i, j, k);
    double g;
    if (ix[a.d] == 0 && a.bclo == ELL_NEU) return;
    if (ix[a.d] == a.n && a.bchi == ELL_NEU) return;
    if (ix[a.d] == 0 && a.bclo == ELL_DIR)      g = (3.0 * p[0] - p[st] / 3.0) / a.dx;
    else if (ix[a.d] == a.n && a.bchi == ELL_DIR) g = -(3.0 * p[-st] - p[-2 * st] / 3.0) / a.dx;
    else                                         g = (p[0] - p[-st]) / a.dx;
    a.mac(i, j, k) = a.mac(i, j, k) - a.beta(i, j, k) * g;
}

// ---------------- ghost fills ----------------
// periodic wrap along d inside the region (the rank owns the whole periodic extent in d).
// Cell data: ghost -g <- n-g, n-1+g <- g-1.  Face data normal to d (nodal in d): faces 0..n valid,
// ghost face -g <- face n-g, ghost face n+g <- face g.
struct WrapArgs { Range r; int d, n, ng, nodal, ncomp; View v; };
__global__ void k_wrap(WrapArgs a)     // range covers the transverse extents and g = 1..ng along d (index lo[d]..hi[d] = 1..ng)
{
    THREAD_IJK(a.r)
    int ix[3] = { i, j, k };
    const int g = ix[a.d];
    const long st = a.v.st(a.d);
    ix[a.d] = 0;
    double *p0 = &a.v(ix[0], ix[1], ix[2]);
    for (int c = 0; c < a.ncomp; ++c) {
        double *p = p0 + a.v.cs * c;
        if (!a.nodal) { p[-(long)g * st] = p[(long)(a.n - g) * st]; p[(long)(a.n - 1 + g) * st] = p[(long)(g - 1) * st]; }
        else          { p[-(long)g * st] = p[(long)(a.n - g) * st]; p[(long)(a.n + g) * st] = p[(long)g * st]; }
    }
}
// the same wrap along x: ghost columns are strided in memory, so threads run along j (one row each) and ghost layer g
__global__ void k_wrap_x(WrapArgs a)    // blockDim (64, 4): x -> j, y -> g-1 (ng <= 4); grid (rows/64, planes)
{
    const int j = a.r.lo[1] + blockIdx.x * blockDim.x + threadIdx.x;
    const int g = 1 + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.y;
    if (j > a.r.hi[1] || g > a.ng) return;
    double *p0 = &a.v(0, j, k);
    for (int c = 0; c < a.ncomp; ++c) {
        double *p = p0 + a.v.cs * c;
        if (!a.nodal) { p[-g] = p[a.n - g]; p[a.n - 1 + g] = p[g - 1]; }
        else          { p[-g] = p[a.n - g]; p[a.n + g] = p[g]; }
    }
}
// physical BC ghost fill along d (both sides) for one comp; ranges as multifab_physbc.f90
struct PbcArgs { Range r; int d, n, ng; int bc[2]; int full_lo[3], full_hi[3]; double val[2]; View v; };
__global__ void k_physbc(PbcArgs a)    // range: transverse extents (full ghosted); index along d unused (lo=hi=0)
{
    THREAD_IJK(a.r)
    int ix[3] = { i, j, k };
    ix[a.d] = 0;
    const long st = a.v.st(a.d);
    double *p0 = &a.v(ix[0], ix[1], ix[2]);
    // restricted transverse range for the extrapolating / reflecting types: skip ghost rows in directions > d on
    // non-interior sides (encoded by the caller in full_lo/full_hi = the restricted range)
    bool in_restricted = true;
    for (int t = 0; t < 3; ++t) if (t != a.d && (ix[t] < a.full_lo[t] || ix[t] > a.full_hi[t])) in_restricted = false;
    for (int side = 0; side < 2; ++side) {
        const int b = a.bc[side];
        if (b == BC_INTERIOR || b == BC_PERIODIC) continue;
        double *e = side == 0 ? p0 : p0 + (long)(a.n - 1) * st;     // first interior cell
        const long o = side == 0 ? -st : st;                         // outward
        if (b == BC_EXT_DIR) {
            for (int g = 1; g <= a.ng; ++g) e[o * g] = a.val[side];
        } else if (!in_restricted) {
            continue;
        } else if (b == BC_FOEXTRAP) {
            const double v = e[0];
            for (int g = 1; g <= a.ng; ++g) e[o * g] = v;
        } else if (b == BC_HOEXTRAP) {
            const double v = (15.0 * e[0] - 10.0 * e[-o] + 3.0 * e[-2 * o]) * 0.125;
            for (int g = 1; g <= a.ng; ++g) e[o * g] = v;
        } else if (b == BC_REFLECT_EVEN) {
            for (int g = 1; g <= a.ng; ++g) e[o * g] = e[-o * (g - 1)];
        } else if (b == BC_REFLECT_ODD) {
            for (int g = 1; g <= a.ng; ++g) e[o * g] = -e[-o * (g - 1)];
        }
    }
}

struct SetArgs { double *p; long n; double v; };
__global__ void k_setval(SetArgs a)
{
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < a.n; t += (long)gridDim.x * blockDim.x) a.p[t] = a.v;
}
struct AmaxArgs { Range r; int ncomp; View v; double *out; };
__global__ void k_absmax(AmaxArgs a)
{
    const int i = a.r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.z;
    double m = ZERO;
    if (i <= a.r.hi[0] && j <= a.r.hi[1])
        for (int c = 0; c < a.ncomp; ++c) m = fmax(m, fabs(a.v(i, j, k, c)));
    block_atomic_max(m, a.out);
}

// estdt_2d / estdt_3d (estdt.f90:89-181): max |u_d| and max |gp_d / rho - f_d| over the valid cells, six block reductions + atomics
struct EstArgs { Range r; int dim; View u, s, gp, f; double *out; };
__global__ void k_estdt(EstArgs a)
{
    const int i = a.r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.z;
    double um[3] = { ZERO, ZERO, ZERO }, fm[3] = { ZERO, ZERO, ZERO };
    if (i <= a.r.hi[0] && j <= a.r.hi[1]) {
        const double rho = a.s(i, j, k, 0);
        for (int d = 0; d < a.dim; ++d) {
            um[d] = fabs(a.u(i, j, k, d));
            fm[d] = fabs(a.gp(i, j, k, d) / rho - a.f(i, j, k, d));
        }
    }
    for (int d = 0; d < a.dim; ++d) { block_atomic_max(um[d], a.out + d); __syncthreads(); block_atomic_max(fm[d], a.out + 3 + d); __syncthreads(); }
}

Range valid_range(const vdn_ctx *c, int fdir)
{
    const Geo &g = c->geo;
    return mk_range(0, g.n[0] - 1 + (fdir == 0), 0, g.n[1] - 1 + (fdir == 1), 0, g.n[2] - 1 + (fdir == 2));
}

} // namespace

// host box (flat on the device, Fortran order with its own ghost width) <-> region array: vdn_ctx.cu box_copy
__global__ void k_box_copy(BoxCopyArgs a)
{
    const long per = (long)a.n[0] * a.n[1] * a.n[2], tot = per * a.ncomp;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int comp = (int)(t / per); const long r = t - (long)comp * per;
        const int i = (int)(r % a.n[0]), j = (int)((r / a.n[0]) % a.n[1]), k = (int)(r / ((long)a.n[0] * a.n[1]));
        const long h = (a.hofs[0] + i) + (long)a.hext[0] * ((a.hofs[1] + j) + (long)a.hext[1] * ((a.hofs[2] + k) + (long)a.hext[2] * comp));
        const long d = (long)comp * a.cs + (a.dofs[0] + i) + (long)a.dext0 * ((a.dofs[1] + j) + (long)a.dext1 * (a.dofs[2] + k));
        if (a.upload) a.base[d] = a.stage[h]; else a.stage[h] = a.base[d];
    }
}
void st_box_copy(const BoxCopyArgs &a, cudaStream_t stream)
{
    const long tot = (long)a.n[0] * a.n[1] * a.n[2] * a.ncomp;
    if (tot <= 0) return;
    k_box_copy<<<(unsigned)std::min<long>(148 * 16, (tot + 255) / 256), 256, 0, stream>>>(a);
    VDN_CUDA(cudaGetLastError());
}

void st_setval(vdn_ctx *c, int field, double val)
{
    if (field >= VDN_UMAC_X && field <= VDN_UMAC_Z) ++c->umac_epoch;
    DField &f = c->f[field];
    LaunchScope ls(c, "setval", (double)f.bytes);
    SetArgs a; a.p = f.base; a.n = (long)(f.bytes / 8); a.v = val;
    k_setval<<<1184, 256, 0, c->stream>>>(a);
    VDN_CUDA(cudaGetLastError());
}

// estdt (estdt.f90:15-87) on the resident UOLD / SOLD / GP / EXT_VEL_FORCE: the six maxima are reduced over the whole region at once (the
// reference reduces per box and takes the minimum dt; the limits depend on the maxima only, so the result is the same number), all-reduced
// over the ranks (parallel_reduce MPI_MIN of dt == MAX of the maxima), and the driver's fallback / cflfac / growth limit applied on the host.
double st_estdt(vdn_ctx *c, double dtold, double cflfac, double max_dt_growth)
{
    const int dim = c->dim;
    {
        LaunchScope ls(c, "estdt", (double)c->ncells() * 8.0 * (3 * dim + 1));
        VDN_CUDA(cudaMemsetAsync(c->d_red, 0, 6 * 8, c->stream));
        EstArgs a; a.r = valid_range(c, -1); a.dim = dim;
        a.u = c->f[VDN_UOLD].view(); a.s = c->f[VDN_SOLD].view(); a.gp = c->f[VDN_GP].view(); a.f = c->f[VDN_EXT_VEL_FORCE].view(); a.out = c->d_red;
        k_estdt<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
    VDN_CUDA(cudaMemcpyAsync(c->h_pin, c->d_red, 6 * 8, cudaMemcpyDeviceToHost, c->stream));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
    double m[6];
    for (int q = 0; q < 6; ++q) m[q] = c->h_pin[q];
    for (int q = 0; q < 6; ++q) m[q] = comm_allreduce_max(c, m[q]);
    const double eps = (double)1.0e-8f;           // a single-precision literal in the reference (estdt.f90:104,147)
    const double dt_start = 1.e20;
    double dt = dt_start;
    for (int d = 0; d < dim; ++d) if (m[d] > eps) dt = fmin(dt, c->geo.h[d] / m[d]);
    for (int d = 0; d < dim; ++d) if (m[3 + d] > eps) dt = fmin(dt, sqrt(2.0 * c->geo.h[d] / m[3 + d]));
    if (dt == dt_start) { dt = fmin(c->geo.h[0], c->geo.h[1]); if (dim == 3) dt = fmin(dt, c->geo.h[2]); }
    dt = dt * cflfac;
    if (dtold > 0.0) dt = fmin(dt, max_dt_growth * dtold);
    return dt;
}

// dst = src for two fields of one layout (the driver's uold <- unew, sold <- snew, varden.f90:321-324), ghost cells included
void st_field_copy(vdn_ctx *c, int dst, int src)
{
    DField &d = c->f[dst], &s = c->f[src];
    VDN_REQUIRE(d.base && s.base && d.bytes == s.bytes && d.ng == s.ng && d.nc == s.nc && d.fdir == s.fdir, "vdn_field_copy: the two fields have different layouts");
    if (dst >= VDN_UMAC_X && dst <= VDN_UMAC_Z) ++c->umac_epoch;
    LaunchScope ls(c, "field_copy", 2.0 * (double)d.bytes);
    VDN_CUDA(cudaMemcpyAsync(d.base, s.base, d.bytes, cudaMemcpyDeviceToDevice, c->stream));
}

// ---- visc_solve / diff_scalar_solve (viscsolve.f90) around the Helmholtz multigrid of vdn_mg.cu ----
// mkrhs_2d / mkrhs_3d (viscsolve.f90:193-299 for a velocity component, :464-513 for a scalar), the initial guess phi = the current field,
// alpha = rho (or 1), and the boundary data: a Dirichlet face whose ghost cell holds the boundary VALUE (multifab_physbc EXT_DIR) enters the
// stencil_order-2 operator as  beta (3 phi_0 - phi_1/3 - 8/3 phi_b) / h^2 ; the homogeneous part lives in the operator, the 8/3 beta phi_b / h^2
// part is folded into the right-hand side here, once.
namespace {
struct HelmFillArgs {
    Range r; int dim, comp, visc, cn;            // visc: velocity component (mac_rhs gradient term, alpha = rho); cn: Crank-Nicolson (+ mu * lap)
    View u, lap, rho, mrhs;
    double mu, visc_mu_dt, dx[3], h2[3];
    int dirbc[3][2], n[3];
    double *phi, *rhs, *alpha; long off, sy, sz;
};
__global__ void k_helm_fill(HelmFillArgs a)
{
    const int i = a.r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.z;
    if (i > a.r.hi[0] || j > a.r.hi[1]) return;
    const int ix[3] = { i, j, k };
    const double u = a.u(i, j, k, a.comp);
    const double rho = a.visc ? a.rho(i, j, k, 0) : 1.0;
    double rh = a.visc ? u * rho : u;
    if (a.cn) rh = rh + a.mu * a.lap(i, j, k, a.comp);
    if (a.visc) {
        const int e0 = a.comp == 0, e1 = a.comp == 1, e2 = a.comp == 2;
        rh = rh + (1.0 / 3.0) * a.visc_mu_dt * (a.mrhs(i + e0, j + e1, k + e2, 0) - a.mrhs(i - e0, j - e1, k - e2, 0)) / a.dx[a.comp];
    }
    for (int d = 0; d < a.dim; ++d) {
        const int o0 = d == 0, o1 = d == 1, o2 = d == 2;
        if (ix[d] == 0 && a.dirbc[d][0])            rh = rh + ((8.0 / 3.0) * a.mu * a.h2[d]) * a.u(i - o0, j - o1, k - o2, a.comp);
        if (ix[d] == a.n[d] - 1 && a.dirbc[d][1])   rh = rh + ((8.0 / 3.0) * a.mu * a.h2[d]) * a.u(i + o0, j + o1, k + o2, a.comp);
    }
    const long c = a.off + i + a.sy * j + a.sz * k;
    a.rhs[c] = rh; a.phi[c] = u; a.alpha[c] = rho;
}
struct HelmStoreArgs { Range r; int comp; View u; const double *phi; long off, sy, sz; };
__global__ void k_helm_store(HelmStoreArgs a)
{
    const int i = a.r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.z;
    if (i > a.r.hi[0] || j > a.r.hi[1]) return;
    a.u(i, j, k, a.comp) = a.phi[a.off + i + a.sy * j + a.sz * k];
}

// ell_bc_level_build (define_bc_tower.f90:254-340): elliptic boundary type of a velocity component (comp < dim) or a scalar on a domain face
int helm_mode(int phys, int comp_is_vel, int comp, int d)
{
    switch (phys) {
    case BC_PERIODIC:     return VDN_MODE_WRAP;
    case BC_SLIP_WALL:
    case BC_SYMMETRY:     return (comp_is_vel && comp == d) ? VDN_MODE_DIR : VDN_MODE_NEU;
    case BC_NO_SLIP_WALL: return comp_is_vel ? VDN_MODE_DIR : VDN_MODE_NEU;
    case BC_INLET:        return VDN_MODE_DIR;
    case BC_OUTLET:       return VDN_MODE_NEU;
    default: throw VdnError("Helmholtz solve: unsupported boundary type on a domain face");
    }
}

int helm_component(vdn_ctx *c, int field, int comp, bool visc, double mu, int diffusion_type, int *ncycles, double *resnorm)
{
    const int dim = c->dim;
    VDN_REQUIRE(diffusion_type == 1 || diffusion_type == 2, "diffusion_type must be 1 (Crank-Nicolson) or 2 (backward Euler)");
    HelmLev0 L;
    mg_helm_level0(c, &L);
    int mode[3][2];
    HelmFillArgs a;
    a.r = valid_range(c, -1); a.dim = dim; a.comp = comp; a.visc = visc ? 1 : 0; a.cn = diffusion_type == 1 ? 1 : 0;
    a.u = c->f[field].view(); a.lap = c->f[VDN_LAPU].view(); a.rho = c->f[VDN_RHOHALF].view(); a.mrhs = c->f[VDN_MAC_RHS].view();
    a.mu = mu; a.visc_mu_dt = diffusion_type == 1 ? 2.0 * mu : mu;
    for (int d = 0; d < 3; ++d) {
        a.dx[d] = c->geo.h[d < dim ? d : 0]; a.h2[d] = 1.0 / (a.dx[d] * a.dx[d]); a.n[d] = c->geo.n[d];
        for (int s = 0; s < 2; ++s) {
            mode[d][s] = d < dim ? helm_mode(c->dom_bc[d][s], visc ? 1 : 0, comp, d) : VDN_MODE_NEU;
            a.dirbc[d][s] = mode[d][s] == VDN_MODE_DIR;
        }
    }
    a.phi = L.phi; a.rhs = L.rhs; a.alpha = L.alpha; a.off = L.off; a.sy = L.sy; a.sz = L.sz;
    {
        LaunchScope ls(c, "helm_mkrhs", (double)c->ncells() * 8.0 * 7.0, 1 + dim);
        for (int d = 0; d < dim; ++d) {         // beta = mu on every face (viscsolve.f90:59-61)
            SetArgs sa; sa.p = L.b[d]; sa.n = L.ntot; sa.v = mu;
            k_setval<<<1184, 256, 0, c->stream>>>(sa);
        }
        k_helm_fill<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
    // rel_solver_eps = 1.d-12, abs_solver_eps = -1 (viscsolve.f90:88-89)
    const int rc = st_helm_solve(c, mode, 1.0e-12, -1.0, ncycles, resnorm);
    {
        LaunchScope ls(c, "helm_store", (double)c->ncells() * 16.0);
        HelmStoreArgs sa; sa.r = a.r; sa.comp = comp; sa.u = a.u; sa.phi = L.phi; sa.off = L.off; sa.sy = L.sy; sa.sz = L.sz;
        k_helm_store<<<grid3(sa.r, BLK), BLK, 0, c->stream>>>(sa);
        VDN_CUDA(cudaGetLastError());
    }
    return rc;
}
} // namespace

// visc_solve (viscsolve.f90:19-146): one Helmholtz solve per velocity component on UNEW with alpha = RHOHALF, beta = mu, LAPU (Crank-Nicolson)
// and MAC_RHS; then ml_restrict_and_fill(unew) (:105).  mu is (1/2) dt visc_coef (Crank-Nicolson) or dt visc_coef (velocity_advance.f90:105-111).
int st_visc_solve(vdn_ctx *c, double mu, int diffusion_type, int *ncycles, double *resnorm)
{
    VDN_REQUIRE(diffusion_type != 1 || c->lapu_set, "visc_solve with diffusion_type = 1 needs the LAPU field (vdn_field_upload)");
    int rc = 0, cyc = 0, ctot = 0; double res = 0.0, rmax = 0.0;
    for (int d = 0; d < c->dim; ++d) {
        rc |= helm_component(c, VDN_UNEW, d, true, mu, diffusion_type, &cyc, &res);
        ctot += cyc; rmax = fmax(rmax, res);
    }
    st_fill_boundary(c, VDN_UNEW);
    st_physbc(c, VDN_UNEW, 0, false);
    if (ncycles) *ncycles = ctot;
    if (resnorm) *resnorm = rmax;
    return rc;
}

// diff_scalar_solve (viscsolve.f90:310-423): alpha = 1, beta = mu on component icomp of SNEW, then fill_boundary + physbc of that component
// (:379-382; bc_comp = dm + icomp).  diffusion_type = 1 would need the explicit term laps, for which the context has no field yet.
int st_diff_scalar_solve(vdn_ctx *c, double mu, int icomp, int diffusion_type, int *ncycles, double *resnorm)
{
    VDN_REQUIRE(icomp >= 0 && icomp < c->prm.nscal, "scalar component out of range");
    VDN_REQUIRE(diffusion_type == 2, "diff_scalar_solve: only diffusion_type = 2 (backward Euler) in this version (no LAPS field yet)");
    const int rc = helm_component(c, VDN_SNEW, icomp, false, mu, diffusion_type, ncycles, resnorm);
    st_fill_boundary(c, VDN_SNEW);
    st_physbc(c, VDN_SNEW, c->dim, false);
    return rc;
}

double st_absmax_valid(vdn_ctx *c, int field)
{
    DField &f = c->f[field];
    LaunchScope ls(c, "norm_inf", (double)c->ncells() * 8.0 * f.nc);
    VDN_CUDA(cudaMemsetAsync(c->d_red, 0, 8, c->stream));
    AmaxArgs a; a.r = valid_range(c, f.fdir); a.ncomp = f.nc; a.v = f.view(); a.out = c->d_red;
    k_absmax<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
    VDN_CUDA(cudaMemcpyAsync(c->h_pin, c->d_red, 8, cudaMemcpyDeviceToHost, c->stream));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
    return comm_allreduce_max(c, c->h_pin[0]);
}

void st_update(vdn_ctx *c, int is_vel, double dt)
{
    const int dim = c->dim;
    UpdArgs a; a.r = valid_range(c, -1); a.dim = dim; a.is_vel = is_vel;
    a.ncomp = is_vel ? dim : c->prm.nscal;
    a.cons_mask = is_vel ? 0 : 1;                                   // scalar_advance.f90:54-57: density only
    a.sold = c->f[is_vel ? VDN_UOLD : VDN_SOLD].view(); a.snew = c->f[is_vel ? VDN_UNEW : VDN_SNEW].view();
    a.force = c->f[is_vel ? VDN_VEL_FORCE : VDN_SCAL_FORCE].view();
    for (int d = 0; d < 3; ++d) {
        const int dd = d < dim ? d : 0;
        a.mac[d] = c->f[VDN_UMAC_X + dd].view();
        a.sedge[d] = c->f[(is_vel ? VDN_UEDGE_X : VDN_SEDGE_X) + dd].view();
        a.flux[d] = c->f[(is_vel ? VDN_UEDGE_X : VDN_SFLUX_X) + dd].view();
        a.dx[d] = c->geo.h[dd];
    }
    a.dt = dt;
    {
        // velocity 168 B/cell, scalars 120 B/cell (3-D)
        double bpc = is_vel ? 8.0 * (dim + dim + dim * dim + dim + dim) : 8.0 * (2 + dim + dim + dim + 2 + 2);
        LaunchScope ls(c, is_vel ? "update_vel" : "update_scal", (double)c->ncells() * bpc);
        // the flux arrays carry one comp (density): comps >= 1 are never conservative, so their index stays inside sedge
        auto go = [&](auto D, auto NC, auto VEL) {
            k_update<decltype(D)::value, decltype(NC)::value, decltype(VEL)::value><<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        };
        using std::integral_constant;
        const int nc = a.ncomp;
        VDN_REQUIRE(nc >= 1 && nc <= 8, "update: component count out of range");
        auto by_nc = [&](auto D, auto VEL) {
            switch (nc) {
            case 1: go(D, integral_constant<int, 1>(), VEL); break; case 2: go(D, integral_constant<int, 2>(), VEL); break;
            case 3: go(D, integral_constant<int, 3>(), VEL); break; case 4: go(D, integral_constant<int, 4>(), VEL); break;
            case 5: go(D, integral_constant<int, 5>(), VEL); break; case 6: go(D, integral_constant<int, 6>(), VEL); break;
            case 7: go(D, integral_constant<int, 7>(), VEL); break; default: go(D, integral_constant<int, 8>(), VEL); break;
            }
        };
        if (dim == 3) { if (is_vel) by_nc(integral_constant<int, 3>(), std::true_type()); else by_nc(integral_constant<int, 3>(), std::false_type()); }
        else          { if (is_vel) by_nc(integral_constant<int, 2>(), std::true_type()); else by_nc(integral_constant<int, 2>(), std::false_type()); }
        VDN_CUDA(cudaGetLastError());
    }
    // ml_restrict_and_fill (update.f90:103-107)
    const int fld = is_vel ? VDN_UNEW : VDN_SNEW;
    st_fill_boundary(c, fld);
    st_physbc(c, fld, is_vel ? 0 : dim, false);
}

void st_mkvelforce(vdn_ctx *c, int rho_field, double visc_fac)
{
    VDN_REQUIRE(rho_field == VDN_SOLD || rho_field == VDN_RHOHALF, "mkvelforce: rho_field must be SOLD or RHOHALF");
    VfArgs a; a.r = valid_range(c, -1); a.dim = c->dim; a.boussinesq = c->prm.boussinesq;
    a.vf = c->f[VDN_VEL_FORCE].view(); a.ext = c->f[VDN_EXT_VEL_FORCE].view(); a.gp = c->f[VDN_GP].view();
    a.s = c->f[rho_field].view(); a.nscal_s = c->f[rho_field].nc; a.lapu = c->f[VDN_LAPU].view();
    a.visc = c->prm.visc_coef * visc_fac;
    {
        LaunchScope ls(c, "mkvelforce", (double)c->ncells() * 8.0 * (3 * c->dim + 1 + c->dim));
        k_mkvelforce<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
    st_fill_boundary(c, VDN_VEL_FORCE);
    st_physbc(c, VDN_VEL_FORCE, c->dim + c->prm.nscal + 1, true);      // bcomp = extrap_comp, same_boundary (mkforce.f90:75-76)
}

void st_mkscalforce(vdn_ctx *c, double diff_fac)
{
    (void)diff_fac;
    VDN_REQUIRE(c->prm.diff_coef == 0.0, "mkscalforce: diff_coef > 0 (explicit diffusive term) is outside the device path");
    SfArgs a; a.r = valid_range(c, -1); a.nscal = c->prm.nscal; a.sf = c->f[VDN_SCAL_FORCE].view(); a.ext = c->f[VDN_EXT_SCAL_FORCE].view();
    {
        LaunchScope ls(c, "mkscalforce", (double)c->ncells() * 8.0 * (2 * c->prm.nscal - 1));
        k_mkscalforce<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
    st_fill_boundary(c, VDN_SCAL_FORCE);
    st_physbc(c, VDN_SCAL_FORCE, c->dim + c->prm.nscal + 1, true);
}

void st_make_at_halftime(vdn_ctx *c)
{
    HtArgs a; a.r = valid_range(c, -1); a.rh = c->f[VDN_RHOHALF].view(); a.ro = c->f[VDN_SOLD].view(); a.rn = c->f[VDN_SNEW].view();
    {
        LaunchScope ls(c, "make_at_halftime", (double)c->ncells() * 24.0);
        k_halftime<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
    st_fill_boundary(c, VDN_RHOHALF);
    st_physbc(c, VDN_RHOHALF, c->dim, false);                           // bcomp = dm + in_comp (make_at_halftime.f90:64-65)
}

double st_divumac(vdn_ctx *c, bool want_norm)
{
    DivArgs a; a.r = valid_range(c, -1); a.dim = c->dim;
    for (int d = 0; d < 3; ++d) { const int dd = d < c->dim ? d : 0; a.mac[d] = c->f[VDN_UMAC_X + dd].view(); a.dxinv[d] = 1.0 / c->geo.h[dd]; }
    a.mrhs = c->f[VDN_MAC_RHS].view(); a.rh = c->f[VDN_RH].view();
    a.nrm = want_norm ? c->d_red : nullptr;
    if (want_norm) VDN_CUDA(cudaMemsetAsync(c->d_red, 0, 8, c->stream));
    {
        LaunchScope ls(c, "divumac", (double)c->ncells() * 8.0 * (c->dim + 2));      // a6: 40 B/cell in 3-D
        if (c->dim == 3) k_divumac<3><<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        else             k_divumac<2><<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
    if (!want_norm) return 0.0;
    VDN_CUDA(cudaMemcpyAsync(c->h_pin, c->d_red, 8, cudaMemcpyDeviceToHost, c->stream));
    VDN_CUDA(cudaStreamSynchronize(c->stream));
    return comm_allreduce_max(c, c->h_pin[0]);
}

void st_mk_mac_coeffs(vdn_ctx *c)
{
    LaunchScope ls(c, "mk_mac_coeffs", (double)c->ncells() * 8.0 * (1 + c->dim), c->dim);   // a7: 32 B/cell
    for (int d = 0; d < c->dim; ++d) {
        CoefArgs a; a.r = valid_range(c, d); a.d = d; a.rho = c->f[VDN_SOLD].view(); a.beta = c->f[VDN_BETA_X + d].view();
        k_mk_mac_coeffs<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
    }
    VDN_CUDA(cudaGetLastError());
}

void st_mkumac(vdn_ctx *c)
{
    // the solver leaves phi's periodic / rank-boundary ghost cells stale (it wraps by index); the box-boundary faces need them
    st_fill_boundary(c, VDN_PHI);
    {
        LaunchScope ls(c, "mkumac", (double)c->ncells() * 8.0 * (1 + 3 * c->dim), c->dim);    // a9: 80 B/cell
        for (int d = 0; d < c->dim; ++d) {
            UmacArgs a; a.r = valid_range(c, d); a.d = d; a.n = c->geo.n[d];
            a.bclo = c->ell_bc[d][0]; a.bchi = c->ell_bc[d][1];
            a.mac = c->f[VDN_UMAC_X + d].view(); a.phi = c->f[VDN_PHI].view(); a.beta = c->f[VDN_BETA_X + d].view();
            a.dx = c->geo.h[d];
            k_mkumac<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        }
        VDN_CUDA(cudaGetLastError());
    }
    ++c->umac_epoch;
    for (int d = 0; d < c->dim; ++d) st_fill_boundary(c, VDN_UMAC_X + d);       // macproject.f90:491-493
}

// multifab_fill_boundary on the merged region.  Directions split across ranks: one exchange fills the ghost cells of all of them (faces,
// edges and corners of the neighbour ranks, valid range along the other directions).  Periodic directions this rank owns alone then wrap in
// order x, y, z, each over the ghosted range of the split directions and of the wrapped directions before it, so that edge and corner ghost
// cells receive the diagonal images.
void st_fill_boundary(vdn_ctx *c, int field)
{
    ctx_require_comm(c);
    DField &f = c->f[field];
    if (f.ng == 0) return;
    const Geo &g = c->geo;
    const int *pg = comm_pgrid_or_null(c);
    if (pg) comm_exchange_field(c, field);
    for (int d = 0; d < c->dim; ++d) {
        if (!c->wrap[d]) continue;
        WrapArgs a; a.d = d; a.n = g.n[d]; a.ng = f.ng; a.nodal = (f.fdir == d); a.ncomp = f.nc; a.v = f.view();
        int lo[3], hi[3];
        for (int t = 0; t < 3; ++t) {
            const bool filled = t < c->dim && t != d && ((pg && pg[t] > 1) || (c->wrap[t] && t < d));
            if (t >= c->dim) { lo[t] = 0; hi[t] = 0; }
            else if (filled) { lo[t] = -f.ng; hi[t] = g.n[t] - 1 + (f.fdir == t) + f.ng; }       // already filled directions: full
            else             { lo[t] = 0;     hi[t] = g.n[t] - 1 + (f.fdir == t); }               // not (yet) filled: valid only
        }
        lo[d] = 1; hi[d] = f.ng;
        a.r = mk_range(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
        LaunchScope ls(c, "fill_boundary", 0.0);
        if (d == 0 && f.ng <= 4) k_wrap_x<<<dim3(cdiv(hi[1] - lo[1] + 1, 64), hi[2] - lo[2] + 1), BLK, 0, c->stream>>>(a);
        else k_wrap<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
        VDN_CUDA(cudaGetLastError());
    }
}

// multifab_physbc (multifab_physbc.f90:17-61): per component, x faces, then y over the full x range, then z.
void st_physbc(vdn_ctx *c, int field, int bccomp, bool same_boundary)
{
    DField &f = c->f[field];
    VDN_REQUIRE(f.fdir < 0, "physbc applies to cell-centred fields");
    if (f.ng == 0) return;
    const Geo &g = c->geo;
    const int dim = c->dim;
    for (int comp = 0; comp < f.nc; ++comp) {
        const int bcc = same_boundary ? bccomp : bccomp + comp;
        const int (*bc)[2] = c->adv_bc[bcc];
        bool any = false;
        for (int d = 0; d < dim; ++d) for (int s = 0; s < 2; ++s) if (bc[d][s] != BC_INTERIOR && bc[d][s] != BC_PERIODIC) any = true;
        if (!any) continue;
        // EXT_DIR constant slot: icomp = bcc+1; 2-D: 1,2 vel, 3 rho, 4 trac; 3-D: 1..3, 4, 5
        const int icomp = bcc + 1;
        int slot;
        if (dim == 2) slot = (icomp == 1) ? 0 : (icomp == 2) ? 1 : (icomp == 3) ? 3 : (icomp == 4) ? 4 : -1;
        else          slot = (icomp >= 1 && icomp <= 5) ? icomp - 1 : -1;
        for (int d = 0; d < dim; ++d) {
            if ((bc[d][0] == BC_INTERIOR || bc[d][0] == BC_PERIODIC) && (bc[d][1] == BC_INTERIOR || bc[d][1] == BC_PERIODIC)) continue;
            PbcArgs a; a.d = d; a.n = g.n[d]; a.ng = f.ng; a.v = f.view().comp(comp);
            for (int s = 0; s < 2; ++s) {
                a.bc[s] = bc[d][s];
                a.val[s] = slot >= 0 ? c->prm.bc_val[slot][d][s] : 0.0;
                if (a.bc[s] == BC_EXT_DIR && slot < 0) a.bc[s] = BC_INTERIOR;      // no branch in the reference => untouched
            }
            int lo[3], hi[3];
            for (int t = 0; t < 3; ++t) {
                if (t >= dim) { lo[t] = 0; hi[t] = 0; a.full_lo[t] = 0; a.full_hi[t] = 0; continue; }
                lo[t] = -f.ng; hi[t] = g.n[t] - 1 + f.ng;
                if (t < d) { a.full_lo[t] = lo[t]; a.full_hi[t] = hi[t]; }
                else {   // t > d: skip ghost rows on non-interior sides (ngylo.. logic, multifab_physbc.f90:254-276)
                    const bool il = (bc[t][0] == BC_INTERIOR || bc[t][0] == BC_PERIODIC);
                    const bool ih = (bc[t][1] == BC_INTERIOR || bc[t][1] == BC_PERIODIC);
                    a.full_lo[t] = il ? lo[t] : 0; a.full_hi[t] = ih ? hi[t] : g.n[t] - 1;
                }
            }
            lo[d] = 0; hi[d] = 0; a.full_lo[d] = 0; a.full_hi[d] = 0;
            a.r = mk_range(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
            LaunchScope ls(c, "physbc", 0.0);
            k_physbc<<<grid3(a.r, BLK), BLK, 0, c->stream>>>(a);
            VDN_CUDA(cudaGetLastError());
        }
    }
}
