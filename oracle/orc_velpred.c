/*
 * oracle/orc_velpred.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of src/velpred.f90: velpred_3d (:1776-2765, the production
 * routine selected when use_godunov_debug=.false.) and velpred_2d (:125-524).
 * The reference cycles two k-planes of scratch; here every intermediate is a
 * full array (same values, same operation order), indexed by the absolute k.
 */
#include "orc_common.h"

void orc_slope(const V *s, V *sl, const int *lo, const int *hi, int dim, int dir, int ncomp,
               const int *adv_bc, int order);

/* normal Riemann problem, velpred.f90:2084-2088 */
static inline double riemann_n(double l, double r, double eps)
{
    double uavg = HALF*(l + r);
    int test = ((l <= ZERO && r >= ZERO) || (fabs(l + r) < eps));
    double v = (uavg > ZERO) ? l : r;
    return test ? ZERO : v;
}
/* upwind a transverse quantity by the normal velocity un, velpred.f90:2091-2093 */
static inline double upwind_t(double l, double r, double un, double eps)
{
    double v = (un > ZERO) ? l : r;
    double uavg = HALF*(l + r);
    return (fabs(un) < eps) ? uavg : v;
}

/* BC override on the normal-predictor pair (all comps), velpred.f90:2044-2079 (x), 2130-2165 (y), 2308-2347 (z).
 * d = face direction, side 0=lo 1=hi, ug = u in the ghost cell beyond the face.
 * hi_outlet_min reproduces the reference's production 3-D hi-x OUTLET quirk (min instead of max, :2075). */
static inline void bc_normal(double *ul, double *ur, int nc, int d, int side, int bc, const double *ug, int hi_outlet_min)
{
    if (bc == BC_INLET) {
        for (int c = 0; c < nc; ++c) { ul[c] = ug[c]; ur[c] = ug[c]; }
    } else if (bc == BC_SLIP_WALL) {
        ul[d] = ZERO; ur[d] = ZERO;
        for (int c = 0; c < nc; ++c) if (c != d) { if (side == 0) ul[c] = ur[c]; else ur[c] = ul[c]; }
    } else if (bc == BC_NO_SLIP_WALL) {
        for (int c = 0; c < nc; ++c) { ul[c] = ZERO; ur[c] = ZERO; }
    } else if (bc == BC_OUTLET) {
        if (side == 0) {
            ur[d] = dmin(ur[d], ZERO);
            for (int c = 0; c < nc; ++c) ul[c] = ur[c];
        } else {
            ul[d] = hi_outlet_min ? dmin(ul[d], ZERO) : dmax(ul[d], ZERO);
            for (int c = 0; c < nc; ++c) ur[c] = ul[c];
        }
    }
}
/* BC override on a transverse-corrected pair, velpred.f90:2202-2221 */
static inline void bc_trans(double *l, double *r, int side, int bc, double ug)
{
    if (bc == BC_INLET) { *l = ug; *r = ug; }
    else if (bc == BC_SLIP_WALL || bc == BC_OUTLET) { if (side == 0) *l = *r; else *r = *l; }
    else if (bc == BC_NO_SLIP_WALL) { *l = ZERO; *r = ZERO; }
}
/* final MAC face value BC, velpred.f90:2644-2659 */
static inline double bc_mac(double v, double ml, double mr, int side, int bc, double ug)
{
    if (bc == BC_SLIP_WALL || bc == BC_NO_SLIP_WALL) return ZERO;
    if (bc == BC_INLET)  return ug;
    if (bc == BC_OUTLET) return side == 0 ? dmin(mr, ZERO) : dmax(ml, ZERO);
    return v;
}

/* eps = 1e-8 * max|u| over the valid region of this box (velpred.f90:1965-1980, :215-227) */
double orc_velpred_eps(const V *u, const int *lo, const int *hi, int dim)
{
    int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    double umax = fabs(AT(*u, lo[0], lo[1], k0, 0));
    for (int k = k0; k <= k1; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i)
                for (int c = 0; c < dim; ++c) umax = dmax(umax, fabs(AT(*u, i, j, k, c)));
    return (umax == 0.0) ? 1.0e-8 : 1.0e-8*umax;
}

/*
 * phys_bc[d][side]; adv_bc[comp][d][side] (comps 0..2 = velocity).
 * u: ng_u ghosts, 3 comps; umac/vmac/wmac: face arrays, ng_m ghosts; force: ng_f ghosts, 3 comps.
 */
void orc_velpred_3d(const double *u_, double *umac_, double *vmac_, double *wmac_, const double *force_,
                    const int *lo, const int *hi, const double *dx, double dt,
                    const int *phys_bc, const int *adv_bc, int ng_u, int ng_m, int ng_f,
                    int use_minion, int slope_order)
{
    const int is = lo[0], ie = hi[0], js = lo[1], je = hi[1], ks = lo[2], ke = hi[2];
    V u     = v_box((double*)u_, lo, hi, ng_u, -1, 3, 3);
    V force = v_box((double*)force_, lo, hi, ng_f, -1, 3, 3);
    V umac  = v_box(umac_, lo, hi, ng_m, 0, 1, 3);
    V vmac  = v_box(vmac_, lo, hi, ng_m, 1, 1, 3);
    V wmac  = v_box(wmac_, lo, hi, ng_m, 2, 1, 3);
#define PB(d,s) phys_bc[(d)*2+(s)]

    V slopex = v_alloc(is-1,ie+1, js-1,je+1, ks-1,ke+1, 3);
    V slopey = v_alloc(is-1,ie+1, js-1,je+1, ks-1,ke+1, 3);
    V slopez = v_alloc(is-1,ie+1, js-1,je+1, ks-1,ke+1, 3);
    orc_slope(&u, &slopex, lo, hi, 3, 0, 3, adv_bc, slope_order);
    orc_slope(&u, &slopey, lo, hi, 3, 1, 3, adv_bc, slope_order);
    orc_slope(&u, &slopez, lo, hi, 3, 2, 3, adv_bc, slope_order);

    /* extents follow the allocate statements, velpred.f90:1865-1946, third extent made full */
    V ulx = v_alloc(is,ie+1, js-1,je+1, ks-1,ke+1, 3), urx = v_alloc(is,ie+1, js-1,je+1, ks-1,ke+1, 3), uimhx = v_alloc(is,ie+1, js-1,je+1, ks-1,ke+1, 3);
    V uly = v_alloc(is-1,ie+1, js,je+1, ks-1,ke+1, 3), ury = v_alloc(is-1,ie+1, js,je+1, ks-1,ke+1, 3), uimhy = v_alloc(is-1,ie+1, js,je+1, ks-1,ke+1, 3);
    V ulz = v_alloc(is-1,ie+1, js-1,je+1, ks,ke+1, 3), urz = v_alloc(is-1,ie+1, js-1,je+1, ks,ke+1, 3), uimhz = v_alloc(is-1,ie+1, js-1,je+1, ks,ke+1, 3);
    V uimhyz = v_alloc(is-1,ie+1, js,je+1, ks,ke, 1);
    V uimhzy = v_alloc(is-1,ie+1, js,je, ks,ke+1, 1);
    V vimhxz = v_alloc(is,ie+1, js-1,je+1, ks,ke, 1);
    V vimhzx = v_alloc(is,ie, js-1,je+1, ks,ke+1, 1);
    V wimhxy = v_alloc(is,ie+1, js,je, ks-1,ke+1, 1);
    V wimhyx = v_alloc(is,ie, js,je+1, ks-1,ke+1, 1);

    const double dt2 = HALF*dt, dt4 = dt/4.0, dt6 = dt/6.0;
    const double hx = dx[0], hy = dx[1], hz = dx[2];
    const double eps = orc_velpred_eps(&u, lo, hi, 3);

    /* 1. uimhx (is:ie+1, js-1:je+1, k), velpred.f90:2019-2099 */
    #pragma omp parallel for collapse(2)
    for (int k = ks-1; k <= ke+1; ++k)
    for (int j = js-1; j <= je+1; ++j)
    for (int i = is; i <= ie+1; ++i) {
        double ul[3], ur[3];
        for (int c = 0; c < 3; ++c) {
            ul[c] = AT(u,i-1,j,k,c) + (HALF - dt2*dmax(ZERO,AT(u,i-1,j,k,0))/hx)*AT(slopex,i-1,j,k,c);
            ur[c] = AT(u,i  ,j,k,c) - (HALF + dt2*dmin(ZERO,AT(u,i  ,j,k,0))/hx)*AT(slopex,i  ,j,k,c);
            if (use_minion) { ul[c] = ul[c] + dt2*AT(force,i-1,j,k,c); ur[c] = ur[c] + dt2*AT(force,i,j,k,c); }
        }
        if (i == is)   { double ug[3] = { AT(u,is-1,j,k,0), AT(u,is-1,j,k,1), AT(u,is-1,j,k,2) }; bc_normal(ul,ur,3,0,0,PB(0,0),ug,0); }
        if (i == ie+1) { double ug[3] = { AT(u,ie+1,j,k,0), AT(u,ie+1,j,k,1), AT(u,ie+1,j,k,2) }; bc_normal(ul,ur,3,0,1,PB(0,1),ug,1 /* quirk :2075 */); }
        for (int c = 0; c < 3; ++c) { AT(ulx,i,j,k,c) = ul[c]; AT(urx,i,j,k,c) = ur[c]; }
        double un = riemann_n(ul[0], ur[0], eps);
        AT(uimhx,i,j,k,0) = un;
        AT(uimhx,i,j,k,1) = upwind_t(ul[1], ur[1], un, eps);
        AT(uimhx,i,j,k,2) = upwind_t(ul[2], ur[2], un, eps);
    }

    /* 2. uimhy (is-1:ie+1, js:je+1, k), velpred.f90:2105-2185 */
    #pragma omp parallel for collapse(2)
    for (int k = ks-1; k <= ke+1; ++k)
    for (int j = js; j <= je+1; ++j)
    for (int i = is-1; i <= ie+1; ++i) {
        double ul[3], ur[3];
        for (int c = 0; c < 3; ++c) {
            ul[c] = AT(u,i,j-1,k,c) + (HALF - dt2*dmax(ZERO,AT(u,i,j-1,k,1)/hy))*AT(slopey,i,j-1,k,c);   /* note max(0,u/hy) :2108 */
            ur[c] = AT(u,i,j  ,k,c) - (HALF + dt2*dmin(ZERO,AT(u,i,j  ,k,1))/hy)*AT(slopey,i,j  ,k,c);
            if (use_minion) { ul[c] = ul[c] + dt2*AT(force,i,j-1,k,c); ur[c] = ur[c] + dt2*AT(force,i,j,k,c); }
        }
        if (j == js)   { double ug[3] = { AT(u,i,js-1,k,0), AT(u,i,js-1,k,1), AT(u,i,js-1,k,2) }; bc_normal(ul,ur,3,1,0,PB(1,0),ug,0); }
        if (j == je+1) { double ug[3] = { AT(u,i,je+1,k,0), AT(u,i,je+1,k,1), AT(u,i,je+1,k,2) }; bc_normal(ul,ur,3,1,1,PB(1,1),ug,0); }
        for (int c = 0; c < 3; ++c) { AT(uly,i,j,k,c) = ul[c]; AT(ury,i,j,k,c) = ur[c]; }
        double un = riemann_n(ul[1], ur[1], eps);
        AT(uimhy,i,j,k,1) = un;
        AT(uimhy,i,j,k,0) = upwind_t(ul[0], ur[0], un, eps);
        AT(uimhy,i,j,k,2) = upwind_t(ul[2], ur[2], un, eps);
    }

    /* 5. uimhz (is-1:ie+1, js-1:je+1, k), k = ks..ke+1, velpred.f90:2283-2367 */
    #pragma omp parallel for collapse(2)
    for (int k = ks; k <= ke+1; ++k)
    for (int j = js-1; j <= je+1; ++j)
    for (int i = is-1; i <= ie+1; ++i) {
        double ul[3], ur[3];
        for (int c = 0; c < 3; ++c) {
            ul[c] = AT(u,i,j,k-1,c) + (HALF - dt2*dmax(ZERO,AT(u,i,j,k-1,2))/hz)*AT(slopez,i,j,k-1,c);
            ur[c] = AT(u,i,j,k  ,c) - (HALF + dt2*dmin(ZERO,AT(u,i,j,k  ,2))/hz)*AT(slopez,i,j,k  ,c);
            if (use_minion) { ul[c] = ul[c] + dt2*AT(force,i,j,k-1,c); ur[c] = ur[c] + dt2*AT(force,i,j,k,c); }
        }
        if (k == ks)   { double ug[3] = { AT(u,i,j,ks-1,0), AT(u,i,j,ks-1,1), AT(u,i,j,ks-1,2) }; bc_normal(ul,ur,3,2,0,PB(2,0),ug,0); }
        if (k == ke+1) { double ug[3] = { AT(u,i,j,ke+1,0), AT(u,i,j,ke+1,1), AT(u,i,j,ke+1,2) }; bc_normal(ul,ur,3,2,1,PB(2,1),ug,0); }
        for (int c = 0; c < 3; ++c) { AT(ulz,i,j,k,c) = ul[c]; AT(urz,i,j,k,c) = ur[c]; }
        double un = riemann_n(ul[2], ur[2], eps);
        AT(uimhz,i,j,k,2) = un;
        AT(uimhz,i,j,k,0) = upwind_t(ul[0], ur[0], un, eps);
        AT(uimhz,i,j,k,1) = upwind_t(ul[1], ur[1], un, eps);
    }

    /* 3. wimhxy (is:ie+1, js:je, k), velpred.f90:2191-2230 */
    #pragma omp parallel for collapse(2)
    for (int k = ks-1; k <= ke+1; ++k)
    for (int j = js; j <= je; ++j)
    for (int i = is; i <= ie+1; ++i) {
        double l = AT(ulx,i,j,k,2) - (dt6/hy)*(AT(uimhy,i-1,j+1,k,1)+AT(uimhy,i-1,j,k,1))*(AT(uimhy,i-1,j+1,k,2)-AT(uimhy,i-1,j,k,2));
        double r = AT(urx,i,j,k,2) - (dt6/hy)*(AT(uimhy,i  ,j+1,k,1)+AT(uimhy,i  ,j,k,1))*(AT(uimhy,i  ,j+1,k,2)-AT(uimhy,i  ,j,k,2));
        if (i == is)   bc_trans(&l,&r,0,PB(0,0),AT(u,is-1,j,k,2));
        if (i == ie+1) bc_trans(&l,&r,1,PB(0,1),AT(u,ie+1,j,k,2));
        AT(wimhxy,i,j,k,0) = upwind_t(l, r, AT(uimhx,i,j,k,0), eps);
    }
    /* 4. wimhyx (is:ie, js:je+1, k), velpred.f90:2236-2275 */
    #pragma omp parallel for collapse(2)
    for (int k = ks-1; k <= ke+1; ++k)
    for (int j = js; j <= je+1; ++j)
    for (int i = is; i <= ie; ++i) {
        double l = AT(uly,i,j,k,2) - (dt6/hx)*(AT(uimhx,i+1,j-1,k,0)+AT(uimhx,i,j-1,k,0))*(AT(uimhx,i+1,j-1,k,2)-AT(uimhx,i,j-1,k,2));
        double r = AT(ury,i,j,k,2) - (dt6/hx)*(AT(uimhx,i+1,j  ,k,0)+AT(uimhx,i,j  ,k,0))*(AT(uimhx,i+1,j  ,k,2)-AT(uimhx,i,j  ,k,2));
        if (j == js)   bc_trans(&l,&r,0,PB(1,0),AT(u,i,js-1,k,2));
        if (j == je+1) bc_trans(&l,&r,1,PB(1,1),AT(u,i,je+1,k,2));
        AT(wimhyx,i,j,k,0) = upwind_t(l, r, AT(uimhy,i,j,k,1), eps);
    }
    /* 7. vimhzx (is:ie, js-1:je+1, k), k = ks..ke+1, velpred.f90:2425-2468 (kp == k-1, kc == k) */
    #pragma omp parallel for collapse(2)
    for (int k = ks; k <= ke+1; ++k)
    for (int j = js-1; j <= je+1; ++j)
    for (int i = is; i <= ie; ++i) {
        double l = AT(ulz,i,j,k,1) - (dt6/hx)*(AT(uimhx,i+1,j,k-1,0)+AT(uimhx,i,j,k-1,0))*(AT(uimhx,i+1,j,k-1,1)-AT(uimhx,i,j,k-1,1));
        double r = AT(urz,i,j,k,1) - (dt6/hx)*(AT(uimhx,i+1,j,k  ,0)+AT(uimhx,i,j,k  ,0))*(AT(uimhx,i+1,j,k  ,1)-AT(uimhx,i,j,k  ,1));
        if (k == ks)   bc_trans(&l,&r,0,PB(2,0),AT(u,i,j,ks-1,1));
        if (k == ke+1) bc_trans(&l,&r,1,PB(2,1),AT(u,i,j,ke+1,1));
        AT(vimhzx,i,j,k,0) = upwind_t(l, r, AT(uimhz,i,j,k,2), eps);
    }
    /* 8. uimhzy (is-1:ie+1, js:je, k), velpred.f90:2474-2517 */
    #pragma omp parallel for collapse(2)
    for (int k = ks; k <= ke+1; ++k)
    for (int j = js; j <= je; ++j)
    for (int i = is-1; i <= ie+1; ++i) {
        double l = AT(ulz,i,j,k,0) - (dt6/hy)*(AT(uimhy,i,j+1,k-1,1)+AT(uimhy,i,j,k-1,1))*(AT(uimhy,i,j+1,k-1,0)-AT(uimhy,i,j,k-1,0));
        double r = AT(urz,i,j,k,0) - (dt6/hy)*(AT(uimhy,i,j+1,k  ,1)+AT(uimhy,i,j,k  ,1))*(AT(uimhy,i,j+1,k  ,0)-AT(uimhy,i,j,k  ,0));
        if (k == ks)   bc_trans(&l,&r,0,PB(2,0),AT(u,i,j,ks-1,0));
        if (k == ke+1) bc_trans(&l,&r,1,PB(2,1),AT(u,i,j,ke+1,0));
        AT(uimhzy,i,j,k,0) = upwind_t(l, r, AT(uimhz,i,j,k,2), eps);
    }
    /* 9. vimhxz (is:ie+1, js-1:je+1, k-1) -> index kk = k-1 in ks..ke, velpred.f90:2527-2566 (kc == kk+1, kp == kk) */
    #pragma omp parallel for collapse(2)
    for (int kk = ks; kk <= ke; ++kk)
    for (int j = js-1; j <= je+1; ++j)
    for (int i = is; i <= ie+1; ++i) {
        double l = AT(ulx,i,j,kk,1) - (dt6/hz)*(AT(uimhz,i-1,j,kk+1,2)+AT(uimhz,i-1,j,kk,2))*(AT(uimhz,i-1,j,kk+1,1)-AT(uimhz,i-1,j,kk,1));
        double r = AT(urx,i,j,kk,1) - (dt6/hz)*(AT(uimhz,i  ,j,kk+1,2)+AT(uimhz,i  ,j,kk,2))*(AT(uimhz,i  ,j,kk+1,1)-AT(uimhz,i  ,j,kk,1));
        if (i == is)   bc_trans(&l,&r,0,PB(0,0),AT(u,is-1,j,kk,1));
        if (i == ie+1) bc_trans(&l,&r,1,PB(0,1),AT(u,ie+1,j,kk,1));
        AT(vimhxz,i,j,kk,0) = upwind_t(l, r, AT(uimhx,i,j,kk,0), eps);
    }
    /* 10. uimhyz (is-1:ie+1, js:je+1, k-1), velpred.f90:2572-2611 */
    #pragma omp parallel for collapse(2)
    for (int kk = ks; kk <= ke; ++kk)
    for (int j = js; j <= je+1; ++j)
    for (int i = is-1; i <= ie+1; ++i) {
        double l = AT(uly,i,j,kk,0) - (dt6/hz)*(AT(uimhz,i,j-1,kk+1,2)+AT(uimhz,i,j-1,kk,2))*(AT(uimhz,i,j-1,kk+1,0)-AT(uimhz,i,j-1,kk,0));
        double r = AT(ury,i,j,kk,0) - (dt6/hz)*(AT(uimhz,i,j  ,kk+1,2)+AT(uimhz,i,j  ,kk,2))*(AT(uimhz,i,j  ,kk+1,0)-AT(uimhz,i,j  ,kk,0));
        if (j == js)   bc_trans(&l,&r,0,PB(1,0),AT(u,i,js-1,kk,0));
        if (j == je+1) bc_trans(&l,&r,1,PB(1,1),AT(u,i,je+1,kk,0));
        AT(uimhyz,i,j,kk,0) = upwind_t(l, r, AT(uimhy,i,j,kk,1), eps);
    }

    /* 6. wmac (is:ie, js:je, k), k = ks..ke+1, velpred.f90:2373-2419 */
    #pragma omp parallel for collapse(2)
    for (int k = ks; k <= ke+1; ++k)
    for (int j = js; j <= je; ++j)
    for (int i = is; i <= ie; ++i) {
        double ml = AT(ulz,i,j,k,2)
            - (dt4/hx)*(AT(uimhx,i+1,j,k-1,0)+AT(uimhx,i,j,k-1,0))*(AT(wimhxy,i+1,j,k-1,0)-AT(wimhxy,i,j,k-1,0))
            - (dt4/hy)*(AT(uimhy,i,j+1,k-1,1)+AT(uimhy,i,j,k-1,1))*(AT(wimhyx,i,j+1,k-1,0)-AT(wimhyx,i,j,k-1,0));
        double mr = AT(urz,i,j,k,2)
            - (dt4/hx)*(AT(uimhx,i+1,j,k  ,0)+AT(uimhx,i,j,k  ,0))*(AT(wimhxy,i+1,j,k  ,0)-AT(wimhxy,i,j,k  ,0))
            - (dt4/hy)*(AT(uimhy,i,j+1,k  ,1)+AT(uimhy,i,j,k  ,1))*(AT(wimhyx,i,j+1,k  ,0)-AT(wimhyx,i,j,k  ,0));
        if (!use_minion) { ml = ml + dt2*AT(force,i,j,k-1,2); mr = mr + dt2*AT(force,i,j,k,2); }
        double v = riemann_n(ml, mr, eps);
        if (k == ks)   v = bc_mac(v, ml, mr, 0, PB(2,0), AT(u,i,j,ks-1,2));
        if (k == ke+1) v = bc_mac(v, ml, mr, 1, PB(2,1), AT(u,i,j,ke+1,2));
        AT(wmac,i,j,k,0) = v;
    }
    /* 11. umac (is:ie+1, js:je, k-1), velpred.f90:2617-2659 */
    #pragma omp parallel for collapse(2)
    for (int kk = ks; kk <= ke; ++kk)
    for (int j = js; j <= je; ++j)
    for (int i = is; i <= ie+1; ++i) {
        double ml = AT(ulx,i,j,kk,0)
            - (dt4/hy)*(AT(uimhy,i-1,j+1,kk,1)+AT(uimhy,i-1,j,kk,1))*(AT(uimhyz,i-1,j+1,kk,0)-AT(uimhyz,i-1,j,kk,0))
            - (dt4/hz)*(AT(uimhz,i-1,j,kk+1,2)+AT(uimhz,i-1,j,kk,2))*(AT(uimhzy,i-1,j,kk+1,0)-AT(uimhzy,i-1,j,kk,0));
        double mr = AT(urx,i,j,kk,0)
            - (dt4/hy)*(AT(uimhy,i  ,j+1,kk,1)+AT(uimhy,i  ,j,kk,1))*(AT(uimhyz,i  ,j+1,kk,0)-AT(uimhyz,i  ,j,kk,0))
            - (dt4/hz)*(AT(uimhz,i  ,j,kk+1,2)+AT(uimhz,i  ,j,kk,2))*(AT(uimhzy,i  ,j,kk+1,0)-AT(uimhzy,i  ,j,kk,0));
        if (!use_minion) { ml = ml + dt2*AT(force,i-1,j,kk,0); mr = mr + dt2*AT(force,i,j,kk,0); }
        double v = riemann_n(ml, mr, eps);
        if (i == is)   v = bc_mac(v, ml, mr, 0, PB(0,0), AT(u,is-1,j,kk,0));
        if (i == ie+1) v = bc_mac(v, ml, mr, 1, PB(0,1), AT(u,ie+1,j,kk,0));
        AT(umac,i,j,kk,0) = v;
    }
    /* 12. vmac (is:ie, js:je+1, k-1), velpred.f90:2665-2707 */
    #pragma omp parallel for collapse(2)
    for (int kk = ks; kk <= ke; ++kk)
    for (int j = js; j <= je+1; ++j)
    for (int i = is; i <= ie; ++i) {
        double ml = AT(uly,i,j,kk,1)
            - (dt4/hx)*(AT(uimhx,i+1,j-1,kk,0)+AT(uimhx,i,j-1,kk,0))*(AT(vimhxz,i+1,j-1,kk,0)-AT(vimhxz,i,j-1,kk,0))
            - (dt4/hz)*(AT(uimhz,i,j-1,kk+1,2)+AT(uimhz,i,j-1,kk,2))*(AT(vimhzx,i,j-1,kk+1,0)-AT(vimhzx,i,j-1,kk,0));
        double mr = AT(ury,i,j,kk,1)
            - (dt4/hx)*(AT(uimhx,i+1,j  ,kk,0)+AT(uimhx,i,j  ,kk,0))*(AT(vimhxz,i+1,j  ,kk,0)-AT(vimhxz,i,j  ,kk,0))
            - (dt4/hz)*(AT(uimhz,i,j  ,kk+1,2)+AT(uimhz,i,j  ,kk,2))*(AT(vimhzx,i,j  ,kk+1,0)-AT(vimhzx,i,j  ,kk,0));
        if (!use_minion) { ml = ml + dt2*AT(force,i,j-1,kk,1); mr = mr + dt2*AT(force,i,j,kk,1); }
        double v = riemann_n(ml, mr, eps);
        if (j == js)   v = bc_mac(v, ml, mr, 0, PB(1,0), AT(u,i,js-1,kk,1));
        if (j == je+1) v = bc_mac(v, ml, mr, 1, PB(1,1), AT(u,i,je+1,kk,1));
        AT(vmac,i,j,kk,0) = v;
    }

    v_free(&slopex); v_free(&slopey); v_free(&slopez);
    v_free(&ulx); v_free(&urx); v_free(&uimhx); v_free(&uly); v_free(&ury); v_free(&uimhy);
    v_free(&ulz); v_free(&urz); v_free(&uimhz);
    v_free(&uimhyz); v_free(&uimhzy); v_free(&vimhxz); v_free(&vimhzx); v_free(&wimhxy); v_free(&wimhyx);
#undef PB
}

/* velpred_2d, velpred.f90:125-524.  Arrays have a unit third extent. */
void orc_velpred_2d(const double *u_, double *umac_, double *vmac_, const double *force_,
                    const int *lo, const int *hi, const double *dx, double dt,
                    const int *phys_bc, const int *adv_bc, int ng_u, int ng_m, int ng_f,
                    int use_minion, int slope_order)
{
    const int is = lo[0], ie = hi[0], js = lo[1], je = hi[1];
    V u     = v_box((double*)u_, lo, hi, ng_u, -1, 2, 2);
    V force = v_box((double*)force_, lo, hi, ng_f, -1, 2, 2);
    V umac  = v_box(umac_, lo, hi, ng_m, 0, 1, 2);
    V vmac  = v_box(vmac_, lo, hi, ng_m, 1, 1, 2);
#define PB(d,s) phys_bc[(d)*2+(s)]
    V slopex = v_alloc(is-1,ie+1, js-1,je+1, 0,0, 2);
    V slopey = v_alloc(is-1,ie+1, js-1,je+1, 0,0, 2);
    orc_slope(&u, &slopex, lo, hi, 2, 0, 2, adv_bc, slope_order);
    orc_slope(&u, &slopey, lo, hi, 2, 1, 2, adv_bc, slope_order);

    V ulx = v_alloc(is,ie+1, js-1,je+1, 0,0, 2), urx = v_alloc(is,ie+1, js-1,je+1, 0,0, 2), uimhx = v_alloc(is,ie+1, js-1,je+1, 0,0, 2);
    V uly = v_alloc(is-1,ie+1, js,je+1, 0,0, 2), ury = v_alloc(is-1,ie+1, js,je+1, 0,0, 2), uimhy = v_alloc(is-1,ie+1, js,je+1, 0,0, 2);

    const double dt2 = HALF*dt, dt4 = dt/4.0;
    const double hx = dx[0], hy = dx[1];
    const double eps = orc_velpred_eps(&u, lo, hi, 2);

    /* 1. uimhx, velpred.f90:258-322 (note max(0,u/hx) and min(0,u/hx) forms) */
    for (int j = js-1; j <= je+1; ++j)
    for (int i = is; i <= ie+1; ++i) {
        double ul[2], ur[2];
        for (int c = 0; c < 2; ++c) {
            ul[c] = AT(u,i-1,j,0,c) + (HALF - dt2*dmax(ZERO,AT(u,i-1,j,0,0)/hx))*AT(slopex,i-1,j,0,c);
            ur[c] = AT(u,i  ,j,0,c) - (HALF + dt2*dmin(ZERO,AT(u,i  ,j,0,0)/hx))*AT(slopex,i  ,j,0,c);
            if (use_minion) { ul[c] = ul[c] + dt2*AT(force,i-1,j,0,c); ur[c] = ur[c] + dt2*AT(force,i,j,0,c); }
        }
        if (i == is)   { double ug[2] = { AT(u,is-1,j,0,0), AT(u,is-1,j,0,1) }; bc_normal(ul,ur,2,0,0,PB(0,0),ug,0); }
        if (i == ie+1) { double ug[2] = { AT(u,ie+1,j,0,0), AT(u,ie+1,j,0,1) }; bc_normal(ul,ur,2,0,1,PB(0,1),ug,0); }
        for (int c = 0; c < 2; ++c) { AT(ulx,i,j,0,c) = ul[c]; AT(urx,i,j,0,c) = ur[c]; }
        double un = riemann_n(ul[0], ur[0], eps);
        AT(uimhx,i,j,0,0) = un;
        AT(uimhx,i,j,0,1) = upwind_t(ul[1], ur[1], un, eps);
    }
    /* 2. uimhy, velpred.f90:330-396 */
    for (int j = js; j <= je+1; ++j)
    for (int i = is-1; i <= ie+1; ++i) {
        double ul[2], ur[2];
        for (int c = 0; c < 2; ++c) {
            ul[c] = AT(u,i,j-1,0,c) + (HALF - dt2*dmax(ZERO,AT(u,i,j-1,0,1)/hy))*AT(slopey,i,j-1,0,c);
            ur[c] = AT(u,i,j  ,0,c) - (HALF + dt2*dmin(ZERO,AT(u,i,j  ,0,1)/hy))*AT(slopey,i,j  ,0,c);
            if (use_minion) { ul[c] = ul[c] + dt2*AT(force,i,j-1,0,c); ur[c] = ur[c] + dt2*AT(force,i,j,0,c); }
        }
        if (j == js)   { double ug[2] = { AT(u,i,js-1,0,0), AT(u,i,js-1,0,1) }; bc_normal(ul,ur,2,1,0,PB(1,0),ug,0); }
        if (j == je+1) { double ug[2] = { AT(u,i,je+1,0,0), AT(u,i,je+1,0,1) }; bc_normal(ul,ur,2,1,1,PB(1,1),ug,0); }
        for (int c = 0; c < 2; ++c) { AT(uly,i,j,0,c) = ul[c]; AT(ury,i,j,0,c) = ur[c]; }
        double un = riemann_n(ul[1], ur[1], eps);
        AT(uimhy,i,j,0,1) = un;
        AT(uimhy,i,j,0,0) = upwind_t(ul[0], ur[0], un, eps);
    }
    /* 3. vmac(is:ie, j), velpred.f90:402-444 (jp == j-1, jc == j) */
    for (int j = js; j <= je+1; ++j)
    for (int i = is; i <= ie; ++i) {
        double ml = AT(uly,i,j,0,1) - (dt4/hx)*(AT(uimhx,i+1,j-1,0,0)+AT(uimhx,i,j-1,0,0))*(AT(uimhx,i+1,j-1,0,1)-AT(uimhx,i,j-1,0,1));
        double mr = AT(ury,i,j,0,1) - (dt4/hx)*(AT(uimhx,i+1,j  ,0,0)+AT(uimhx,i,j  ,0,0))*(AT(uimhx,i+1,j  ,0,1)-AT(uimhx,i,j  ,0,1));
        if (!use_minion) { ml = ml + dt2*AT(force,i,j-1,0,1); mr = mr + dt2*AT(force,i,j,0,1); }
        double v = riemann_n(ml, mr, eps);
        if (j == js)   v = bc_mac(v, ml, mr, 0, PB(1,0), AT(u,i,js-1,0,1));
        if (j == je+1) v = bc_mac(v, ml, mr, 1, PB(1,1), AT(u,i,je+1,0,1));
        AT(vmac,i,j,0,0) = v;
    }
    /* 4. umac(is:ie+1, j-1), velpred.f90:454-496 (jc == jj+1, jp == jj) */
    for (int jj = js; jj <= je; ++jj)
    for (int i = is; i <= ie+1; ++i) {
        double ml = AT(ulx,i,jj,0,0) - (dt4/hy)*(AT(uimhy,i-1,jj+1,0,1)+AT(uimhy,i-1,jj,0,1))*(AT(uimhy,i-1,jj+1,0,0)-AT(uimhy,i-1,jj,0,0));
        double mr = AT(urx,i,jj,0,0) - (dt4/hy)*(AT(uimhy,i  ,jj+1,0,1)+AT(uimhy,i  ,jj,0,1))*(AT(uimhy,i  ,jj+1,0,0)-AT(uimhy,i  ,jj,0,0));
        if (!use_minion) { ml = ml + dt2*AT(force,i-1,jj,0,0); mr = mr + dt2*AT(force,i,jj,0,0); }
        double v = riemann_n(ml, mr, eps);
        if (i == is)   v = bc_mac(v, ml, mr, 0, PB(0,0), AT(u,is-1,jj,0,0));
        if (i == ie+1) v = bc_mac(v, ml, mr, 1, PB(0,1), AT(u,ie+1,jj,0,0));
        AT(umac,i,jj,0,0) = v;
    }
    v_free(&slopex); v_free(&slopey);
    v_free(&ulx); v_free(&urx); v_free(&uimhx); v_free(&uly); v_free(&ury); v_free(&uimhy);
#undef PB
}
