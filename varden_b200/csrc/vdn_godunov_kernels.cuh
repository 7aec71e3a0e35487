// vdn_godunov_kernels.cuh -- device code of the Godunov edge-state predictor (slopes, velpred, mkflux) and the stage
// orchestration, written against an abstract launcher so that the SAME source runs under nvcc (vdn_godunov.cu) and,
// as plain C++ with tests/emu/cuda_emu.h, in the CPU-only test tier.
#pragma once


constexpr double HALF = 0.5, ZERO = 0.0, ONE = 1.0, TWO = 2.0;

// ------------------------------------------------------------------------------------------
// slopes (slope.f90).  s points at cell m along a direction with stride st; n = region cells along it.
// bclo/bchi: adv_bc is EXT_DIR or HOEXTRAP on that region face.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void slope_parts(const double *s, int st, double &cen, double &lim, double &flag, double &fromm)
{
    const double sm = s[-st], s0 = s[0], sp = s[st];
    cen = HALF * (sp - sm);
    const double dmn = TWO * (s0 - sm), dpls = TWO * (sp - s0);
    lim = fmin(fabs(dmn), fabs(dpls));
    lim = (dpls * dmn > ZERO) ? lim : ZERO;
    flag = copysign(ONE, cen);
    fromm = flag * fmin(lim, fabs(cen));
}
// one-sided 4th-order slope at the first interior cell next to a lo face (slope.f90:247-254); s points at that cell
__device__ __forceinline__ double slope4_lo(const double *s, int st)
{
    const double two3rd = 2.0 / 3.0, tenth = 0.1;
    double del = (-(16.0 / 15.0)) * s[-st] + HALF * s[0] + two3rd * s[st] - tenth * s[2 * st];
    double dmn = TWO * (s[0] - s[-st]), dpls = TWO * (s[st] - s[0]);
    double slim = fmin(fabs(dpls), fabs(dmn));
    slim = (dpls * dmn > ZERO) ? slim : ZERO;
    return copysign(ONE, del) * fmin(slim, fabs(del));
}
__device__ __forceinline__ double slope4_hi(const double *s, int st)   // slope.f90:268-275
{
    const double two3rd = 2.0 / 3.0, tenth = 0.1;
    double del = -((-(16.0 / 15.0)) * s[st] + HALF * s[0] + two3rd * s[-st] - tenth * s[-2 * st]);
    double dmn = TWO * (s[0] - s[-st]), dpls = TWO * (s[st] - s[0]);
    double slim = fmin(fabs(dpls), fabs(dmn));
    slim = (dpls * dmn > ZERO) ? slim : ZERO;
    return copysign(ONE, del) * fmin(slim, fabs(del));
}
__device__ __forceinline__ double slope2_lo(const double *s, int st)   // slope.f90:193-200
{
    double del = (s[st] + 3.0 * s[0] - 4.0 * s[-st]) * (1.0 / 3.0);
    double dpls = TWO * (s[st] - s[0]), dmn = TWO * (s[0] - s[-st]);
    double slim = fmin(fabs(dpls), fabs(dmn));
    slim = (dpls * dmn > ZERO) ? slim : ZERO;
    return copysign(ONE, del) * fmin(slim, fabs(del));
}
__device__ __forceinline__ double slope2_hi(const double *s, int st)   // slope.f90:207-214
{
    double del = -(s[-st] + 3.0 * s[0] - 4.0 * s[st]) * (1.0 / 3.0);
    double dpls = TWO * (s[0] - s[-st]), dmn = TWO * (s[st] - s[0]);
    double slim = fmin(fabs(dpls), fabs(dmn));
    slim = (dpls * dmn > ZERO) ? slim : ZERO;
    return copysign(ONE, del) * fmin(slim, fabs(del));
}
__device__ double slope_at(const double *s, int st, int m, int n, bool bclo, bool bchi, int order)
{
    if (order == 0) return ZERO;
    if ((bclo && m == -1) || (bchi && m == n)) return ZERO;
    if (order == 2) {
        if (bclo && m == 0) return slope2_lo(s, st);
        if (bchi && m == n - 1) return slope2_hi(s, st);
        double del = HALF * (s[st] - s[-st]);
        double dpls = TWO * (s[st] - s[0]), dmn = TWO * (s[0] - s[-st]);
        double slim = fmin(fabs(dpls), fabs(dmn));
        slim = (dpls * dmn > ZERO) ? slim : ZERO;
        return copysign(ONE, del) * fmin(slim, fabs(del));
    }
    if (bclo && m == 0) return slope4_lo(s, st);
    if (bchi && m == n - 1) return slope4_hi(s, st);
    const double two3rd = 2.0 / 3.0, sixth = 1.0 / 6.0;
    double cen, lim, flag, fr, c2, l2, f2, frp, frm;
    slope_parts(s, st, cen, lim, flag, fr);
    slope_parts(s + st, st, c2, l2, f2, frp);
    slope_parts(s - st, st, c2, l2, f2, frm);
    if (bclo && m - 1 == 0) frm = slope4_lo(s - st, st);          // revised fromm(is), slope.f90:257
    if (bchi && m + 1 == n - 1) frp = slope4_hi(s + st, st);      // revised fromm(ie), slope.f90:278
    double ds = TWO * two3rd * cen - sixth * (frp + frm);
    return flag * fmin(fabs(ds), lim);
}

struct SlopeArgs { Geo g; View s; View out[3]; int ncomp; int order; int bc[3][3][2]; Range r; };

// slopes of up to 3 comps in all directions on cells -1..n
__global__ void k_slopes(SlopeArgs a)
{
    const int i = a.r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = a.r.lo[2] + blockIdx.z;
    if (i > a.r.hi[0] || j > a.r.hi[1]) return;
    const int ix[3] = { i, j, k };
    for (int c = 0; c < a.ncomp; ++c) {
        const double *p = &a.s(i, j, k, c);
        for (int d = 0; d < a.g.dim; ++d) {
            const bool bl = a.bc[c][d][0] == BC_EXT_DIR || a.bc[c][d][0] == BC_HOEXTRAP;
            const bool bh = a.bc[c][d][1] == BC_EXT_DIR || a.bc[c][d][1] == BC_HOEXTRAP;
            a.out[d](i, j, k, c) = slope_at(p, a.s.st(d), ix[d], a.g.n[d], bl, bh, a.order);
        }
    }
}

// ------------------------------------------------------------------------------------------
// shared Riemann / upwind / BC helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double riemann_n(double l, double r, double eps)      // velpred.f90:2084-2088
{
    double uavg = HALF * (l + r);
    bool test = ((l <= ZERO && r >= ZERO) || (fabs(l + r) < eps));
    double v = (uavg > ZERO) ? l : r;
    return test ? ZERO : v;
}
__device__ __forceinline__ double upwind_t(double l, double r, double un, double eps)   // velpred.f90:2091-2093
{
    double v = (un > ZERO) ? l : r;
    double uavg = HALF * (l + r);
    return (fabs(un) < eps) ? uavg : v;
}
__device__ __forceinline__ double upw(double l, double r, double um, double eps)        // mkflux.f90:1520-1522
{
    double v = (um > ZERO) ? l : r;
    double savg = HALF * (l + r);
    return (fabs(um) > eps) ? v : savg;
}
template <int NC>
__device__ __forceinline__ void bc_normal(double (&ul)[NC], double (&ur)[NC], int d, int side, int bc, const double (&ug)[NC], bool hi_outlet_min)
{
    if (bc == BC_INLET) {
#pragma unroll
        for (int c = 0; c < NC; ++c) { ul[c] = ug[c]; ur[c] = ug[c]; }
    } else if (bc == BC_SLIP_WALL) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c == d) { ul[c] = ZERO; ur[c] = ZERO; }
            else if (side == 0) ul[c] = ur[c]; else ur[c] = ul[c];
        }
    } else if (bc == BC_NO_SLIP_WALL) {
#pragma unroll
        for (int c = 0; c < NC; ++c) { ul[c] = ZERO; ur[c] = ZERO; }
    } else if (bc == BC_OUTLET) {
        if (side == 0) {
#pragma unroll
            for (int c = 0; c < NC; ++c) { if (c == d) ur[c] = fmin(ur[c], ZERO); ul[c] = ur[c]; }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) { if (c == d) ul[c] = hi_outlet_min ? fmin(ul[c], ZERO) : fmax(ul[c], ZERO); ur[c] = ul[c]; }
        }
    }
}
__device__ __forceinline__ void bc_trans(double &l, double &r, int side, int bc, double ug)     // velpred.f90:2202-2221
{
    if (bc == BC_INLET) { l = ug; r = ug; }
    else if (bc == BC_SLIP_WALL || bc == BC_OUTLET) { if (side == 0) l = r; else r = l; }
    else if (bc == BC_NO_SLIP_WALL) { l = ZERO; r = ZERO; }
}
__device__ __forceinline__ double bc_mac(double v, double ml, double mr, int side, int bc, double ug)  // velpred.f90:2644-2659
{
    if (bc == BC_SLIP_WALL || bc == BC_NO_SLIP_WALL) return ZERO;
    if (bc == BC_INLET) return ug;
    if (bc == BC_OUTLET) return side == 0 ? fmin(mr, ZERO) : fmax(ml, ZERO);
    return v;
}
__device__ __forceinline__ void bc_pair(double &l, double &r, int d, int side, int bc, int is_vel, int comp, double sg)  // mkflux.f90:1463-1515
{
    if (bc == BC_INLET) { l = sg; r = sg; }
    else if (bc == BC_SLIP_WALL) {
        if (is_vel && comp == d) { l = ZERO; r = ZERO; }
        else if (side == 0) l = r; else r = l;
    } else if (bc == BC_NO_SLIP_WALL) {
        if (is_vel) { l = ZERO; r = ZERO; }
        else if (side == 0) l = r; else r = l;
    } else if (bc == BC_OUTLET) {
        if (is_vel && comp == d) {
            if (side == 0) { l = fmin(r, ZERO); r = fmin(r, ZERO); }
            else           { l = fmax(l, ZERO); r = fmax(l, ZERO); }
        } else if (side == 0) l = r; else r = l;
    }
}
__device__ __forceinline__ double bc_edge(double v, double el, double er, int d, int side, int bc, int is_vel, int comp, double sg)  // mkflux.f90:2356-2397
{
    const double in = (side == 0) ? er : el;
    if (bc == BC_INLET) return sg;
    if (bc == BC_SLIP_WALL) return (is_vel && comp == d) ? ZERO : in;
    if (bc == BC_NO_SLIP_WALL) return is_vel ? ZERO : in;
    if (bc == BC_OUTLET) {
        if (is_vel && comp == d) return (side == 0) ? fmin(er, ZERO) : fmax(el, ZERO);
        return in;
    }
    return v;
}

#define THREAD_IJK(r)                                                            \
    const int i = (r).lo[0] + blockIdx.x * blockDim.x + threadIdx.x;             \
    const int j = (r).lo[1] + blockIdx.y * blockDim.y + threadIdx.y;             \
    const int k = (r).lo[2] + blockIdx.z;                                        \
    if (i > (r).hi[0] || j > (r).hi[1]) return;                                  \
    const int ix[3] = { i, j, k }; (void)ix;

// ------------------------------------------------------------------------------------------
// velpred, staged form: every stage is a per-point device function (one face of direction D at (i,j,k)), one direction per
// launch, intermediates in the S-layout scratch arena.  This is the 2-D path (velpred_2d / mkflux_2d; config 1 is tiny); 3-D
// runs the plane-marching kernels of vdn_godunov_march.cuh.  The templates stay DIM-generic so that the CPU test tier can run
// the staged 3-D form as an independent second implementation against the marching kernels.
// ------------------------------------------------------------------------------------------
struct VpArgs {
    Geo g; Range r;
    View u, force;
    View sl[3];                     // slopes along D, DIM comps
    View ul[3], ur[3], uimh[3];     // per direction, DIM comps each
    View X[3][3];                   // transverse-corrected states X[D][T] (3-D)
    View out[3];                    // umac_D (field views)
    const double *eps;
    double dt; int use_minion;
    int order; int sbc[3][3][2];    // slope order and adv_bc[comp][d][side] for in-register slopes
};

// normal predictor + Riemann/upwind: velpred.f90:2019-2099 (x), 2105-2185 (y), 2283-2367 (z); 2-D :258-322, :330-396
template <int DIM, int D>
__device__ __forceinline__ void vp_normal_pt(const VpArgs &a, int i, int j, int k)
{
    const int ix[3] = { i, j, k };
    const double dt2 = HALF * a.dt, h = a.g.h[D];
    const int su = a.u.st(D);
    const double *uR = &a.u(i, j, k), *uL = uR - su;
    double slL[DIM], slR[DIM];
    {
        const int ss = a.sl[D].st(D);
        const double *sR = &a.sl[D](i, j, k), *sL = sR - ss;
#pragma unroll
        for (int c = 0; c < DIM; ++c) { slL[c] = sL[a.sl[D].cs * c]; slR[c] = sR[a.sl[D].cs * c]; }
    }
    const double unL = uL[a.u.cs * D], unR = uR[a.u.cs * D];
    // operation-order quirks: 3-D x,z: dt2*max(0,u)/h ; 3-D y left: dt2*max(0,u/h) ; 2-D: both inside
    double cl, cr;
    if (DIM == 2)             { cl = HALF - dt2 * fmax(ZERO, unL / h); cr = HALF + dt2 * fmin(ZERO, unR / h); }
    else if (D == 1)          { cl = HALF - dt2 * fmax(ZERO, unL / h); cr = HALF + dt2 * fmin(ZERO, unR) / h; }
    else                      { cl = HALF - dt2 * fmax(ZERO, unL) / h; cr = HALF + dt2 * fmin(ZERO, unR) / h; }
    double ul[DIM], ur[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        ul[c] = uL[a.u.cs * c] + cl * slL[c];
        ur[c] = uR[a.u.cs * c] - cr * slR[c];
    }
    if (a.use_minion) {
        const int sf = a.force.st(D);
        const double *fR = &a.force(i, j, k), *fL = fR - sf;
#pragma unroll
        for (int c = 0; c < DIM; ++c) { ul[c] = ul[c] + dt2 * fL[a.force.cs * c]; ur[c] = ur[c] + dt2 * fR[a.force.cs * c]; }
    }
    if (ix[D] == 0) {
        double ug[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) ug[c] = uL[a.u.cs * c];
        bc_normal<DIM>(ul, ur, D, 0, a.g.pbc[D][0], ug, false);
    }
    if (ix[D] == a.g.n[D]) {
        double ug[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) ug[c] = uR[a.u.cs * c];
        bc_normal<DIM>(ul, ur, D, 1, a.g.pbc[D][1], ug, DIM == 3 && D == 0 /* velpred.f90:2075 */);
    }
    const double eps = a.eps[a.g.box(i, j, k)];
    const double un = riemann_n(ul[D], ur[D], eps);
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        a.ul[D](i, j, k, c) = ul[c];
        a.ur[D](i, j, k, c) = ur[c];
        a.uimh[D](i, j, k, c) = (c == D) ? un : upwind_t(ul[c], ur[c], un, eps);
    }
}
template <int DIM, int D>
__global__ void k_vp_normal(VpArgs a)
{
    THREAD_IJK(a.r)
    vp_normal_pt<DIM, D>(a, i, j, k);
}

// transverse-corrected tangential state of comp C = 3-D-T on D faces, corrected by T (3-D only):
// wimhxy :2191, wimhyx :2236, vimhzx :2425, uimhzy :2474, vimhxz :2527, uimhyz :2572
template <int D, int T>
__device__ __forceinline__ void vp_trans_pt(const VpArgs &a, int i, int j, int k)
{
    const int ix[3] = { i, j, k };
    constexpr int C = 3 - D - T;
    const View ulD = a.ul[D].comp(C), urD = a.ur[D].comp(C);
    const View uimhD_n = a.uimh[D].comp(D), uimhT_n = a.uimh[T].comp(T), uimhT_c = a.uimh[T].comp(C);
    const double dt6 = a.dt / 6.0, hT = a.g.h[T];
    const int sD = uimhT_n.st(D), sT = uimhT_n.st(T);
    // R cell = (i,j,k), L cell = R - e_D; T-faces of a cell: lo = cell index, hi = cell index + e_T
    const double *nR = &uimhT_n(i, j, k), *cR = &uimhT_c(i, j, k);
    const double *nL = nR - sD, *cL = cR - sD;
    double l = ulD(i, j, k) - (dt6 / hT) * (nL[sT] + nL[0]) * (cL[sT] - cL[0]);
    double r = urD(i, j, k) - (dt6 / hT) * (nR[sT] + nR[0]) * (cR[sT] - cR[0]);
    if (ix[D] == 0)        bc_trans(l, r, 0, a.g.pbc[D][0], (&a.u(i, j, k, C))[-a.u.st(D)]);
    if (ix[D] == a.g.n[D]) bc_trans(l, r, 1, a.g.pbc[D][1], a.u(i, j, k, C));
    const double eps = a.eps[a.g.box(i, j, k)];
    a.X[D][T](i, j, k) = upwind_t(l, r, uimhD_n(i, j, k), eps);
}
template <int D, int T>
__global__ void k_vp_trans(VpArgs a)
{
    THREAD_IJK(a.r)
    vp_trans_pt<D, T>(a, i, j, k);
}

// final MAC velocity: umac :2617-2659, vmac :2665-2707, wmac :2373-2419 ; 2-D :402-444, :454-496
template <int DIM, int D>
__device__ __forceinline__ void vp_final_pt(const VpArgs &a, int i, int j, int k)
{
    const int ix[3] = { i, j, k };
    constexpr int T1 = (D == 0) ? 1 : 0;
    constexpr int T2 = (D == 2) ? 1 : 2;
    const View ulD = a.ul[D].comp(D), urD = a.ur[D].comp(D);
    const View n1 = a.uimh[T1].comp(T1), n2 = a.uimh[T2 < DIM ? T2 : T1].comp(T2 < DIM ? T2 : T1);
    const View x1 = DIM == 3 ? a.X[T1][T2] : a.uimh[T1].comp(D);
    const View x2 = DIM == 3 ? a.X[T2][T1] : x1;
    const double dt2 = HALF * a.dt, dt4 = a.dt / 4.0;
    double ml, mr;
    {
        const int sD = n1.st(D), s1 = n1.st(T1);
        const double *nR = &n1(i, j, k), *xR = &x1(i, j, k);
        const double *nL = nR - sD, *xL = xR - sD;
        ml = ulD(i, j, k) - (dt4 / a.g.h[T1]) * (nL[s1] + nL[0]) * (xL[s1] - xL[0]);
        mr = urD(i, j, k) - (dt4 / a.g.h[T1]) * (nR[s1] + nR[0]) * (xR[s1] - xR[0]);
    }
    if (DIM == 3) {
        const int sD = n2.st(D), s2 = n2.st(T2);
        const double *nR = &n2(i, j, k), *xR = &x2(i, j, k);
        const double *nL = nR - sD, *xL = xR - sD;
        ml = ml - (dt4 / a.g.h[T2]) * (nL[s2] + nL[0]) * (xL[s2] - xL[0]);
        mr = mr - (dt4 / a.g.h[T2]) * (nR[s2] + nR[0]) * (xR[s2] - xR[0]);
    }
    if (!a.use_minion) {
        const double *fR = &a.force(i, j, k, D);
        ml = ml + dt2 * fR[-a.force.st(D)];
        mr = mr + dt2 * fR[0];
    }
    const double eps = a.eps[a.g.box(i, j, k)];
    double v = riemann_n(ml, mr, eps);
    if (ix[D] == 0)        v = bc_mac(v, ml, mr, 0, a.g.pbc[D][0], (&a.u(i, j, k, D))[-a.u.st(D)]);
    if (ix[D] == a.g.n[D]) v = bc_mac(v, ml, mr, 1, a.g.pbc[D][1], a.u(i, j, k, D));
    a.out[D](i, j, k) = v;
}
template <int DIM, int D>
__global__ void k_vp_final(VpArgs a)
{
    THREAD_IJK(a.r)
    vp_final_pt<DIM, D>(a, i, j, k);
}

// ------------------------------------------------------------------------------------------
// mkflux, staged form (one component per pass)
// ------------------------------------------------------------------------------------------
struct MfArgs {
    Geo g; Range r;
    View s;                         // comp already selected
    View sl[3];                     // slope along D of this comp
    View mac[3];                    // MAC velocities
    View force, mac_rhs;
    View l[3], rr[3], simh[3];      // 1-D extrapolated L/R states and their upwinded value, per direction
    View X[3][3];                   // transverse-once states
    View sedge[3], flux[3];         // outputs (comp selected)
    const double *eps;
    double dt; int use_minion, is_vel, comp, cons, use_rhs;
    int order; int sbc[3][2];       // slope order and adv_bc[d][side] of this comp for in-register slopes
};
// 1-D extrapolation + BC + upwind: mkflux.f90:1443-1524 (x), 1530-1611 (y), 1779-1864 (z)
template <int D>
__device__ __forceinline__ void mf_normal_pt(const MfArgs &a, int i, int j, int k)
{
    const int ix[3] = { i, j, k };
    const double dt2 = HALF * a.dt, h = a.g.h[D];
    const double *sR = &a.s(i, j, k), *sL = sR - a.s.st(D);
    double pLv, pRv;
    {
        const double *pR = &a.sl[D](i, j, k), *pL = pR - a.sl[D].st(D);
        pLv = pL[0]; pRv = pR[0];
    }
    const double um = a.mac[D](i, j, k);
    double l = sL[0] + (HALF - dt2 * um / h) * pLv;
    double r = sR[0] - (HALF + dt2 * um / h) * pRv;
    if (a.use_minion) {
        const double *fR = &a.force(i, j, k);
        l = l + dt2 * fR[-a.force.st(D)]; r = r + dt2 * fR[0];
        if (a.cons && a.use_rhs) {
            const double *dR = &a.mac_rhs(i, j, k);
            l = l - dt2 * sL[0] * dR[-a.mac_rhs.st(D)]; r = r - dt2 * sR[0] * dR[0];
        }
    }
    if (ix[D] == 0)        bc_pair(l, r, D, 0, a.g.pbc[D][0], a.is_vel, a.comp, sL[0]);
    if (ix[D] == a.g.n[D]) bc_pair(l, r, D, 1, a.g.pbc[D][1], a.is_vel, a.comp, sR[0]);
    const double eps = a.eps[a.g.box(i, j, k)];
    a.l[D](i, j, k) = l; a.rr[D](i, j, k) = r;
    a.simh[D](i, j, k) = upw(l, r, um, eps);
}
template <int D>
__global__ void k_mf_normal(MfArgs a)
{
    THREAD_IJK(a.r)
    mf_normal_pt<D>(a, i, j, k);
}

// transverse-once states: simhxy :1617, simhyx :1697, simhzx :1978, simhzy :2062, simhxz :2150, simhyz :2230
template <int D, int T>
__device__ __forceinline__ void mf_trans_pt(const MfArgs &a, int i, int j, int k)
{
    const int ix[3] = { i, j, k };
    const double dt3 = a.dt / 3.0, dt6 = a.dt / 6.0, hT = a.g.h[T];
    const double *qR = &a.simh[T](i, j, k), *qL = qR - a.simh[T].st(D);
    const double *mR = &a.mac[T](i, j, k), *mL = mR - a.mac[T].st(D);
    const int sq = a.simh[T].st(T), sm = a.mac[T].st(T);
    double l, r;
    if (a.cons) {
        l = a.l[D](i, j, k) - (dt3 / hT) * (qL[sq] * mL[sm] - qL[0] * mL[0]);
        r = a.rr[D](i, j, k) - (dt3 / hT) * (qR[sq] * mR[sm] - qR[0] * mR[0]);
    } else {
        l = a.l[D](i, j, k) - (dt6 / hT) * (mL[sm] + mL[0]) * (qL[sq] - qL[0]);
        r = a.rr[D](i, j, k) - (dt6 / hT) * (mR[sm] + mR[0]) * (qR[sq] - qR[0]);
    }
    const double *sR = &a.s(i, j, k);
    if (ix[D] == 0)        bc_pair(l, r, D, 0, a.g.pbc[D][0], a.is_vel, a.comp, sR[-a.s.st(D)]);
    if (ix[D] == a.g.n[D]) bc_pair(l, r, D, 1, a.g.pbc[D][1], a.is_vel, a.comp, sR[0]);
    const double eps = a.eps[a.g.box(i, j, k)];
    a.X[D][T](i, j, k) = upw(l, r, a.mac[D](i, j, k), eps);
}
template <int D, int T>
__global__ void k_mf_trans(MfArgs a)
{
    THREAD_IJK(a.r)
    mf_trans_pt<D, T>(a, i, j, k);
}

// final edge state + flux: sedgex :2310-2408, sedgey :2414-2512, sedgez :1870-1972 ; 2-D :470-566, :572-666
template <int DIM, int D>
__device__ __forceinline__ void mf_final_pt(const MfArgs &a, int i, int j, int k)
{
    const int ix[3] = { i, j, k };
    constexpr int T1 = (D == 0) ? 1 : 0;
    constexpr int T2 = (D == 2) ? 1 : 2;
    const View x1 = DIM == 3 ? a.X[T1][T2] : a.simh[T1];
    const View mac1 = a.mac[T1];
    const double dt2 = HALF * a.dt, dt4 = a.dt / 4.0;
    const double h1 = a.g.h[T1], h2 = a.g.h[T2];
    const double *sR = &a.s(i, j, k), *sL = sR - a.s.st(D);
    const double *x1R = &x1(i, j, k), *x1L = x1R - x1.st(D);
    const double *m1R = &mac1(i, j, k), *m1L = m1R - mac1.st(D);
    const int sx1 = x1.st(T1), sm1 = mac1.st(T1);
    double el = a.l[D](i, j, k), er = a.rr[D](i, j, k);
    if (DIM == 3) {
        const View x2 = a.X[T2][T1], mac2 = a.mac[T2];
        const double *x2R = &x2(i, j, k), *x2L = x2R - x2.st(D);
        const double *m2R = &mac2(i, j, k), *m2L = m2R - mac2.st(D);
        const int sx2 = x2.st(T2), sm2 = mac2.st(T2);
        if (a.cons) {
            el = el - (dt2 / h1) * (x1L[sx1] * m1L[sm1] - x1L[0] * m1L[0])
                    - (dt2 / h2) * (x2L[sx2] * m2L[sm2] - x2L[0] * m2L[0])
                    + (dt2 / h1) * sL[0] * (m1L[sm1] - m1L[0])
                    + (dt2 / h2) * sL[0] * (m2L[sm2] - m2L[0]);
            er = er - (dt2 / h1) * (x1R[sx1] * m1R[sm1] - x1R[0] * m1R[0])
                    - (dt2 / h2) * (x2R[sx2] * m2R[sm2] - x2R[0] * m2R[0])
                    + (dt2 / h1) * sR[0] * (m1R[sm1] - m1R[0])
                    + (dt2 / h2) * sR[0] * (m2R[sm2] - m2R[0]);
        } else {
            el = el - (dt4 / h1) * (m1L[sm1] + m1L[0]) * (x1L[sx1] - x1L[0])
                    - (dt4 / h2) * (m2L[sm2] + m2L[0]) * (x2L[sx2] - x2L[0]);
            er = er - (dt4 / h1) * (m1R[sm1] + m1R[0]) * (x1R[sx1] - x1R[0])
                    - (dt4 / h2) * (m2R[sm2] + m2R[0]) * (x2R[sx2] - x2R[0]);
        }
    } else {
        if (a.cons) {
            el = el - (dt2 / h1) * (x1L[sx1] * m1L[sm1] - x1L[0] * m1L[0]) + (dt2 / h1) * sL[0] * (m1L[sm1] - m1L[0]);
            er = er - (dt2 / h1) * (x1R[sx1] * m1R[sm1] - x1R[0] * m1R[0]) + (dt2 / h1) * sR[0] * (m1R[sm1] - m1R[0]);
        } else {
            el = el - (dt4 / h1) * (m1L[sm1] + m1L[0]) * (x1L[sx1] - x1L[0]);
            er = er - (dt4 / h1) * (m1R[sm1] + m1R[0]) * (x1R[sx1] - x1R[0]);
        }
    }
    if (!a.use_minion) {
        const double *fR = &a.force(i, j, k);
        el = el + dt2 * fR[-a.force.st(D)]; er = er + dt2 * fR[0];
        if (a.cons && a.use_rhs) {
            const double *dR = &a.mac_rhs(i, j, k);
            el = el - dt2 * sL[0] * dR[-a.mac_rhs.st(D)]; er = er - dt2 * sR[0] * dR[0];
        }
    }
    const double um = a.mac[D](i, j, k);
    const double eps = a.eps[a.g.box(i, j, k)];
    double v = upw(el, er, um, eps);
    if (ix[D] == 0)        v = bc_edge(v, el, er, D, 0, a.g.pbc[D][0], a.is_vel, a.comp, sL[0]);
    if (ix[D] == a.g.n[D]) v = bc_edge(v, el, er, D, 1, a.g.pbc[D][1], a.is_vel, a.comp, sR[0]);
    a.sedge[D](i, j, k) = v;
    if (a.cons) a.flux[D](i, j, k) = v * um;
}
template <int DIM, int D>
__global__ void k_mf_final(MfArgs a)
{
    THREAD_IJK(a.r)
    mf_final_pt<DIM, D>(a, i, j, k);
}


// ------------------------------------------------------------------------------------------
// stage orchestration.  L is a launcher:  L.scope(name, alg_bytes, nlaunch) -> RAII profiling bracket,
// L(kernel, range, args) launches kernel over the index range with 64x4x1 thread blocks.
// ------------------------------------------------------------------------------------------

template <int DIM, class L>
void velpred_stages(L &launch, VpArgs a)
{
    const Geo &g = a.g;
    const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    const int zl = DIM == 3 ? -1 : 0, zh = DIM == 3 ? n2 : 0;       // grown z range
    const double cells = (double)n0 * n1 * n2;
    // slopes on cells -1..n (velpred.f90:1848-1852)
    {
        auto ls = launch.scope("vp_slopes", cells * 8.0 * (DIM + DIM * DIM), 1);
        SlopeArgs sa; sa.g = g; sa.s = a.u; sa.ncomp = DIM; sa.order = a.order;
        for (int d = 0; d < 3; ++d) sa.out[d] = a.sl[d];
        for (int cc = 0; cc < 3; ++cc) for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) sa.bc[cc][d][s] = a.sbc[cc][d][s];
        sa.r = mk_range(-1, n0, -1, n1, zl, zh);
        launch(k_slopes, sa.r, sa);
    }
    // normal predictors
    {
        auto ls = launch.scope("vp_normal", cells * 8.0 * DIM * (2 * DIM + 3 * DIM), DIM);
        a.r = mk_range(0, n0, -1, n1, zl, zh); launch(k_vp_normal<DIM, 0>, a.r, a);
        a.r = mk_range(-1, n0, 0, n1, zl, zh); launch(k_vp_normal<DIM, 1>, a.r, a);
        if (DIM == 3) { a.r = mk_range(-1, n0, -1, n1, 0, n2); launch(k_vp_normal<3, 2>, a.r, a); }
    }
    // transverse states (3-D), index ranges from the reference pseudo-code velpred.f90:1986-2004
    if (DIM == 3) {
        auto ls = launch.scope("vp_trans", cells * 8.0 * 6 * 6, 6);
        a.r = mk_range(0, n0, 0, n1 - 1, -1, n2); launch(k_vp_trans<0, 1>, a.r, a);         // wimhxy (is:ie+1, js:je, ks-1:ke+1)
        a.r = mk_range(0, n0 - 1, 0, n1, -1, n2); launch(k_vp_trans<1, 0>, a.r, a);         // wimhyx (is:ie, js:je+1, ks-1:ke+1)
        a.r = mk_range(0, n0 - 1, -1, n1, 0, n2); launch(k_vp_trans<2, 0>, a.r, a);         // vimhzx (is:ie, js-1:je+1, ks:ke+1)
        a.r = mk_range(-1, n0, 0, n1 - 1, 0, n2); launch(k_vp_trans<2, 1>, a.r, a);         // uimhzy (is-1:ie+1, js:je, ks:ke+1)
        a.r = mk_range(0, n0, -1, n1, 0, n2 - 1); launch(k_vp_trans<0, 2>, a.r, a);         // vimhxz (is:ie+1, js-1:je+1, ks:ke)
        a.r = mk_range(-1, n0, 0, n1, 0, n2 - 1); launch(k_vp_trans<1, 2>, a.r, a);         // uimhyz (is-1:ie+1, js:je+1, ks:ke)
    }
    // final MAC velocities on the valid faces
    {
        auto ls = launch.scope("vp_final", cells * 8.0 * DIM * (DIM == 3 ? 8 : 6), DIM);
        a.r = mk_range(0, n0, 0, n1 - 1, 0, n2 - 1); launch(k_vp_final<DIM, 0>, a.r, a);
        a.r = mk_range(0, n0 - 1, 0, n1, 0, n2 - 1); launch(k_vp_final<DIM, 1>, a.r, a);
        if (DIM == 3) { a.r = mk_range(0, n0 - 1, 0, n1 - 1, 0, n2); launch(k_vp_final<3, 2>, a.r, a); }
    }
}

// one component of mkflux (a.s, a.force, a.sedge, a.flux, a.comp, a.cons, a.sbc select it)
template <int DIM, class L>
void mkflux_stages(L &launch, MfArgs a)
{
    const Geo &g = a.g;
    const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    const int zl = DIM == 3 ? -1 : 0, zh = DIM == 3 ? n2 : 0;
    const double cells = (double)n0 * n1 * n2;
    {
        auto ls = launch.scope("mf_slopes", cells * 8.0 * (1 + DIM), 1);
        SlopeArgs sa; sa.g = g; sa.s = a.s; sa.ncomp = 1; sa.order = a.order;
        for (int d = 0; d < 3; ++d) sa.out[d] = a.sl[d];
        for (int d = 0; d < 3; ++d) for (int sd = 0; sd < 2; ++sd) sa.bc[0][d][sd] = a.sbc[d][sd];
        sa.r = mk_range(-1, n0, -1, n1, zl, zh);
        launch(k_slopes, sa.r, sa);
    }
    {
        auto ls = launch.scope("mf_normal", cells * 8.0 * DIM * 6, DIM);
        a.r = mk_range(0, n0, -1, n1, zl, zh); launch(k_mf_normal<0>, a.r, a);
        a.r = mk_range(-1, n0, 0, n1, zl, zh); launch(k_mf_normal<1>, a.r, a);
        if (DIM == 3) { a.r = mk_range(-1, n0, -1, n1, 0, n2); launch(k_mf_normal<2>, a.r, a); }
    }
    if (DIM == 3) {
        auto ls = launch.scope("mf_trans", cells * 8.0 * 6 * 7, 6);
        a.r = mk_range(0, n0, 0, n1 - 1, -1, n2); launch(k_mf_trans<0, 1>, a.r, a);         // simhxy
        a.r = mk_range(0, n0 - 1, 0, n1, -1, n2); launch(k_mf_trans<1, 0>, a.r, a);         // simhyx
        a.r = mk_range(0, n0 - 1, -1, n1, 0, n2); launch(k_mf_trans<2, 0>, a.r, a);         // simhzx
        a.r = mk_range(-1, n0, 0, n1 - 1, 0, n2); launch(k_mf_trans<2, 1>, a.r, a);         // simhzy
        a.r = mk_range(0, n0, -1, n1, 0, n2 - 1); launch(k_mf_trans<0, 2>, a.r, a);         // simhxz
        a.r = mk_range(-1, n0, 0, n1, 0, n2 - 1); launch(k_mf_trans<1, 2>, a.r, a);         // simhyz
    }
    {
        auto ls = launch.scope("mf_final", cells * 8.0 * DIM * (DIM == 3 ? 11 : 8), DIM);
        a.r = mk_range(0, n0, 0, n1 - 1, 0, n2 - 1); launch(k_mf_final<DIM, 0>, a.r, a);
        a.r = mk_range(0, n0 - 1, 0, n1, 0, n2 - 1); launch(k_mf_final<DIM, 1>, a.r, a);
        if (DIM == 3) { a.r = mk_range(0, n0 - 1, 0, n1 - 1, 0, n2); launch(k_mf_final<3, 2>, a.r, a); }
    }
}
