/*
 * oracle/orc_mkflux.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of src/mkflux.f90: mkflux_3d (:1186-2567, production routine)
 * and mkflux_2d (:152-691).  Intermediates are full arrays indexed by absolute k
 * (the reference cycles two planes kc/kp); values and operation order are the same.
 */
#include "orc_common.h"

void orc_slope(const V *s, V *sl, const int *lo, const int *hi, int dim, int dir, int ncomp,
               const int *adv_bc, int order);

/* upwind on the sign of the given MAC velocity, mkflux.f90:1520-1522 */
static inline double upw(double l, double r, double um, double eps)
{
    double v = (um > ZERO) ? l : r;
    double savg = HALF*(l + r);
    return (fabs(um) > eps) ? v : savg;
}
/* BC override of an (L,R) pair on a face in direction d, mkflux.f90:1463-1515 */
static inline void bc_pair(double *l, double *r, int d, int side, int bc, int is_vel, int comp, double sg)
{
    if (bc == BC_INLET) { *l = sg; *r = sg; }
    else if (bc == BC_SLIP_WALL) {
        if (is_vel && comp == d) { *l = ZERO; *r = ZERO; }
        else { if (side == 0) *l = *r; else *r = *l; }
    } else if (bc == BC_NO_SLIP_WALL) {
        if (is_vel) { *l = ZERO; *r = ZERO; }
        else { if (side == 0) *l = *r; else *r = *l; }
    } else if (bc == BC_OUTLET) {
        if (is_vel && comp == d) {
            if (side == 0) { *l = dmin(*r, ZERO); *r = dmin(*r, ZERO); }
            else           { *l = dmax(*l, ZERO); *r = dmax(*l, ZERO); }
        } else { if (side == 0) *l = *r; else *r = *l; }
    }
}
/* BC override of the final edge state, mkflux.f90:2356-2397 */
static inline double bc_edge(double v, double el, double er, int d, int side, int bc, int is_vel, int comp, double sg)
{
    double in = (side == 0) ? er : el;
    if (bc == BC_INLET) return sg;
    if (bc == BC_SLIP_WALL)    return (is_vel && comp == d) ? ZERO : in;
    if (bc == BC_NO_SLIP_WALL) return is_vel ? ZERO : in;
    if (bc == BC_OUTLET) {
        if (is_vel && comp == d) return (side == 0) ? dmin(er, ZERO) : dmax(el, ZERO);
        return in;
    }
    return v;
}

/* eps = 1e-8 * max |umac,vmac,wmac| over this box's valid faces, mkflux.f90:1374-1401 (3-D), :249-265 (2-D) */
double orc_mkflux_eps(const V *umac, const V *vmac, const V *wmac, const int *lo, const int *hi, int dim)
{
    int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    double umax = fabs(AT(*umac, lo[0], lo[1], k0, 0));
    for (int k = k0; k <= k1; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]+1; ++i)
        umax = dmax(umax, fabs(AT(*umac,i,j,k,0)));
    for (int k = k0; k <= k1; ++k) for (int j = lo[1]; j <= hi[1]+1; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
        umax = dmax(umax, fabs(AT(*vmac,i,j,k,0)));
    if (dim == 3)
        for (int k = k0; k <= k1+1; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
            umax = dmax(umax, fabs(AT(*wmac,i,j,k,0)));
    return (umax == 0.0) ? 1.0e-8 : 1.0e-8*umax;
}

/*
 * adv_bc: [ncomp][3][2] table for the comps of s (i.e. already offset by bccomp, mkflux.f90:95).
 * is_cons[ncomp].
 */
void orc_mkflux_3d(const double *s_, double *sedgex_, double *sedgey_, double *sedgez_,
                   double *fluxx_, double *fluxy_, double *fluxz_,
                   const double *umac_, const double *vmac_, const double *wmac_,
                   const double *force_, const double *mac_rhs_,
                   const int *lo, const int *hi, const double *dx, double dt, int is_vel,
                   const int *phys_bc, const int *adv_bc,
                   int ng_s, int ng_e, int ng_f, int ng_u, int ng_o, int ng_m,
                   const int *is_cons, int ncomp, int use_minion, int slope_order)
{
    const int is = lo[0], ie = hi[0], js = lo[1], je = hi[1], ks = lo[2], ke = hi[2];
    V s      = v_box((double*)s_, lo, hi, ng_s, -1, ncomp, 3);
    V sedgex = v_box(sedgex_, lo, hi, ng_e, 0, ncomp, 3);
    V sedgey = v_box(sedgey_, lo, hi, ng_e, 1, ncomp, 3);
    V sedgez = v_box(sedgez_, lo, hi, ng_e, 2, ncomp, 3);
    V fluxx  = v_box(fluxx_, lo, hi, ng_f, 0, ncomp, 3);
    V fluxy  = v_box(fluxy_, lo, hi, ng_f, 1, ncomp, 3);
    V fluxz  = v_box(fluxz_, lo, hi, ng_f, 2, ncomp, 3);
    V umac   = v_box((double*)umac_, lo, hi, ng_u, 0, 1, 3);
    V vmac   = v_box((double*)vmac_, lo, hi, ng_u, 1, 1, 3);
    V wmac   = v_box((double*)wmac_, lo, hi, ng_u, 2, 1, 3);
    V force  = v_box((double*)force_, lo, hi, ng_o, -1, ncomp, 3);
    V mac_rhs= v_box((double*)mac_rhs_, lo, hi, ng_m, -1, 1, 3);
#define PB(d,sd) phys_bc[(d)*2+(sd)]

    V slopex = v_alloc(is-1,ie+1, js-1,je+1, ks-1,ke+1, ncomp);
    V slopey = v_alloc(is-1,ie+1, js-1,je+1, ks-1,ke+1, ncomp);
    V slopez = v_alloc(is-1,ie+1, js-1,je+1, ks-1,ke+1, ncomp);
    orc_slope(&s, &slopex, lo, hi, 3, 0, ncomp, adv_bc, slope_order);
    orc_slope(&s, &slopey, lo, hi, 3, 1, ncomp, adv_bc, slope_order);
    orc_slope(&s, &slopez, lo, hi, 3, 2, ncomp, adv_bc, slope_order);

    /* extents from the allocate statements mkflux.f90:1273-1354 */
    V slx = v_alloc(is,ie+1, js-1,je+1, ks-1,ke+1, 1), srx = v_alloc(is,ie+1, js-1,je+1, ks-1,ke+1, 1), simhx = v_alloc(is,ie+1, js-1,je+1, ks-1,ke+1, 1);
    V sly = v_alloc(is-1,ie+1, js,je+1, ks-1,ke+1, 1), sry = v_alloc(is-1,ie+1, js,je+1, ks-1,ke+1, 1), simhy = v_alloc(is-1,ie+1, js,je+1, ks-1,ke+1, 1);
    V slz = v_alloc(is-1,ie+1, js-1,je+1, ks,ke+1, 1), srz = v_alloc(is-1,ie+1, js-1,je+1, ks,ke+1, 1), simhz = v_alloc(is-1,ie+1, js-1,je+1, ks,ke+1, 1);
    V simhxy = v_alloc(is,ie+1, js,je, ks-1,ke+1, 1);
    V simhxz = v_alloc(is,ie+1, js-1,je+1, ks,ke, 1);
    V simhyx = v_alloc(is,ie, js,je+1, ks-1,ke+1, 1);
    V simhyz = v_alloc(is-1,ie+1, js,je+1, ks,ke, 1);
    V simhzx = v_alloc(is,ie, js-1,je+1, ks,ke+1, 1);
    V simhzy = v_alloc(is-1,ie+1, js,je, ks,ke+1, 1);

    const double dt2 = HALF*dt, dt3 = dt/3.0, dt4 = dt/4.0, dt6 = dt/6.0;
    const double hx = dx[0], hy = dx[1], hz = dx[2];
    const double eps = orc_mkflux_eps(&umac, &vmac, &wmac, lo, hi, 3);

    for (int comp = 0; comp < ncomp; ++comp) {
        const int cons = is_cons[comp];

        /* 1. simhx (is:ie+1, js-1:je+1, k), mkflux.f90:1443-1524 */
        #pragma omp parallel for collapse(2)
        for (int k = ks-1; k <= ke+1; ++k)
        for (int j = js-1; j <= je+1; ++j)
        for (int i = is; i <= ie+1; ++i) {
            double l = AT(s,i-1,j,k,comp) + (HALF - dt2*AT(umac,i,j,k,0)/hx)*AT(slopex,i-1,j,k,comp);
            double r = AT(s,i  ,j,k,comp) - (HALF + dt2*AT(umac,i,j,k,0)/hx)*AT(slopex,i  ,j,k,comp);
            if (use_minion) { l = l + dt2*AT(force,i-1,j,k,comp); r = r + dt2*AT(force,i,j,k,comp); }
            if (use_minion && cons) { l = l - dt2*AT(s,i-1,j,k,comp)*AT(mac_rhs,i-1,j,k,0); r = r - dt2*AT(s,i,j,k,comp)*AT(mac_rhs,i,j,k,0); }
            if (i == is)   bc_pair(&l,&r,0,0,PB(0,0),is_vel,comp,AT(s,is-1,j,k,comp));
            if (i == ie+1) bc_pair(&l,&r,0,1,PB(0,1),is_vel,comp,AT(s,ie+1,j,k,comp));
            AT(slx,i,j,k,0) = l; AT(srx,i,j,k,0) = r;
            AT(simhx,i,j,k,0) = upw(l, r, AT(umac,i,j,k,0), eps);
        }
        /* 2. simhy (is-1:ie+1, js:je+1, k), mkflux.f90:1530-1611 */
        #pragma omp parallel for collapse(2)
        for (int k = ks-1; k <= ke+1; ++k)
        for (int j = js; j <= je+1; ++j)
        for (int i = is-1; i <= ie+1; ++i) {
            double l = AT(s,i,j-1,k,comp) + (HALF - dt2*AT(vmac,i,j,k,0)/hy)*AT(slopey,i,j-1,k,comp);
            double r = AT(s,i,j  ,k,comp) - (HALF + dt2*AT(vmac,i,j,k,0)/hy)*AT(slopey,i,j  ,k,comp);
            if (use_minion) { l = l + dt2*AT(force,i,j-1,k,comp); r = r + dt2*AT(force,i,j,k,comp); }
            if (use_minion && cons) { l = l - dt2*AT(s,i,j-1,k,comp)*AT(mac_rhs,i,j-1,k,0); r = r - dt2*AT(s,i,j,k,comp)*AT(mac_rhs,i,j,k,0); }
            if (j == js)   bc_pair(&l,&r,1,0,PB(1,0),is_vel,comp,AT(s,i,js-1,k,comp));
            if (j == je+1) bc_pair(&l,&r,1,1,PB(1,1),is_vel,comp,AT(s,i,je+1,k,comp));
            AT(sly,i,j,k,0) = l; AT(sry,i,j,k,0) = r;
            AT(simhy,i,j,k,0) = upw(l, r, AT(vmac,i,j,k,0), eps);
        }
        /* 5. simhz (is-1:ie+1, js-1:je+1, k), k=ks..ke+1, mkflux.f90:1779-1864 */
        #pragma omp parallel for collapse(2)
        for (int k = ks; k <= ke+1; ++k)
        for (int j = js-1; j <= je+1; ++j)
        for (int i = is-1; i <= ie+1; ++i) {
            double l = AT(s,i,j,k-1,comp) + (HALF - dt2*AT(wmac,i,j,k,0)/hz)*AT(slopez,i,j,k-1,comp);
            double r = AT(s,i,j,k  ,comp) - (HALF + dt2*AT(wmac,i,j,k,0)/hz)*AT(slopez,i,j,k  ,comp);
            if (use_minion) { l = l + dt2*AT(force,i,j,k-1,comp); r = r + dt2*AT(force,i,j,k,comp); }
            if (use_minion && cons) { l = l - dt2*AT(s,i,j,k-1,comp)*AT(mac_rhs,i,j,k-1,0); r = r - dt2*AT(s,i,j,k,comp)*AT(mac_rhs,i,j,k,0); }
            if (k == ks)   bc_pair(&l,&r,2,0,PB(2,0),is_vel,comp,AT(s,i,j,ks-1,comp));
            if (k == ke+1) bc_pair(&l,&r,2,1,PB(2,1),is_vel,comp,AT(s,i,j,ke+1,comp));
            AT(slz,i,j,k,0) = l; AT(srz,i,j,k,0) = r;
            AT(simhz,i,j,k,0) = upw(l, r, AT(wmac,i,j,k,0), eps);
        }

        /* 3. simhxy (is:ie+1, js:je, k), mkflux.f90:1617-1691 */
        #pragma omp parallel for collapse(2)
        for (int k = ks-1; k <= ke+1; ++k)
        for (int j = js; j <= je; ++j)
        for (int i = is; i <= ie+1; ++i) {
            double l, r;
            if (cons) {
                l = AT(slx,i,j,k,0) - (dt3/hy)*(AT(simhy,i-1,j+1,k,0)*AT(vmac,i-1,j+1,k,0) - AT(simhy,i-1,j,k,0)*AT(vmac,i-1,j,k,0));
                r = AT(srx,i,j,k,0) - (dt3/hy)*(AT(simhy,i  ,j+1,k,0)*AT(vmac,i  ,j+1,k,0) - AT(simhy,i  ,j,k,0)*AT(vmac,i  ,j,k,0));
            } else {
                l = AT(slx,i,j,k,0) - (dt6/hy)*(AT(vmac,i-1,j+1,k,0)+AT(vmac,i-1,j,k,0))*(AT(simhy,i-1,j+1,k,0)-AT(simhy,i-1,j,k,0));
                r = AT(srx,i,j,k,0) - (dt6/hy)*(AT(vmac,i  ,j+1,k,0)+AT(vmac,i  ,j,k,0))*(AT(simhy,i  ,j+1,k,0)-AT(simhy,i  ,j,k,0));
            }
            if (i == is)   bc_pair(&l,&r,0,0,PB(0,0),is_vel,comp,AT(s,is-1,j,k,comp));
            if (i == ie+1) bc_pair(&l,&r,0,1,PB(0,1),is_vel,comp,AT(s,ie+1,j,k,comp));
            AT(simhxy,i,j,k,0) = upw(l, r, AT(umac,i,j,k,0), eps);
        }
        /* 4. simhyx (is:ie, js:je+1, k), mkflux.f90:1697-1771 */
        #pragma omp parallel for collapse(2)
        for (int k = ks-1; k <= ke+1; ++k)
        for (int j = js; j <= je+1; ++j)
        for (int i = is; i <= ie; ++i) {
            double l, r;
            if (cons) {
                l = AT(sly,i,j,k,0) - (dt3/hx)*(AT(simhx,i+1,j-1,k,0)*AT(umac,i+1,j-1,k,0) - AT(simhx,i,j-1,k,0)*AT(umac,i,j-1,k,0));
                r = AT(sry,i,j,k,0) - (dt3/hx)*(AT(simhx,i+1,j  ,k,0)*AT(umac,i+1,j  ,k,0) - AT(simhx,i,j  ,k,0)*AT(umac,i,j  ,k,0));
            } else {
                l = AT(sly,i,j,k,0) - (dt6/hx)*(AT(umac,i+1,j-1,k,0)+AT(umac,i,j-1,k,0))*(AT(simhx,i+1,j-1,k,0)-AT(simhx,i,j-1,k,0));
                r = AT(sry,i,j,k,0) - (dt6/hx)*(AT(umac,i+1,j  ,k,0)+AT(umac,i,j  ,k,0))*(AT(simhx,i+1,j  ,k,0)-AT(simhx,i,j  ,k,0));
            }
            if (j == js)   bc_pair(&l,&r,1,0,PB(1,0),is_vel,comp,AT(s,i,js-1,k,comp));
            if (j == je+1) bc_pair(&l,&r,1,1,PB(1,1),is_vel,comp,AT(s,i,je+1,k,comp));
            AT(simhyx,i,j,k,0) = upw(l, r, AT(vmac,i,j,k,0), eps);
        }
        /* 7. simhzx (is:ie, js-1:je+1, k), k=ks..ke+1, mkflux.f90:1978-2056 (kp == k-1) */
        #pragma omp parallel for collapse(2)
        for (int k = ks; k <= ke+1; ++k)
        for (int j = js-1; j <= je+1; ++j)
        for (int i = is; i <= ie; ++i) {
            double l, r;
            if (cons) {
                l = AT(slz,i,j,k,0) - (dt3/hx)*(AT(simhx,i+1,j,k-1,0)*AT(umac,i+1,j,k-1,0) - AT(simhx,i,j,k-1,0)*AT(umac,i,j,k-1,0));
                r = AT(srz,i,j,k,0) - (dt3/hx)*(AT(simhx,i+1,j,k  ,0)*AT(umac,i+1,j,k  ,0) - AT(simhx,i,j,k  ,0)*AT(umac,i,j,k  ,0));
            } else {
                l = AT(slz,i,j,k,0) - (dt6/hx)*(AT(umac,i+1,j,k-1,0)+AT(umac,i,j,k-1,0))*(AT(simhx,i+1,j,k-1,0)-AT(simhx,i,j,k-1,0));
                r = AT(srz,i,j,k,0) - (dt6/hx)*(AT(umac,i+1,j,k  ,0)+AT(umac,i,j,k  ,0))*(AT(simhx,i+1,j,k  ,0)-AT(simhx,i,j,k  ,0));
            }
            if (k == ks)   bc_pair(&l,&r,2,0,PB(2,0),is_vel,comp,AT(s,i,j,ks-1,comp));
            if (k == ke+1) bc_pair(&l,&r,2,1,PB(2,1),is_vel,comp,AT(s,i,j,ke+1,comp));
            AT(simhzx,i,j,k,0) = upw(l, r, AT(wmac,i,j,k,0), eps);
        }
        /* 8. simhzy (is-1:ie+1, js:je, k), mkflux.f90:2062-2140 */
        #pragma omp parallel for collapse(2)
        for (int k = ks; k <= ke+1; ++k)
        for (int j = js; j <= je; ++j)
        for (int i = is-1; i <= ie+1; ++i) {
            double l, r;
            if (cons) {
                l = AT(slz,i,j,k,0) - (dt3/hy)*(AT(simhy,i,j+1,k-1,0)*AT(vmac,i,j+1,k-1,0) - AT(simhy,i,j,k-1,0)*AT(vmac,i,j,k-1,0));
                r = AT(srz,i,j,k,0) - (dt3/hy)*(AT(simhy,i,j+1,k  ,0)*AT(vmac,i,j+1,k  ,0) - AT(simhy,i,j,k  ,0)*AT(vmac,i,j,k  ,0));
            } else {
                l = AT(slz,i,j,k,0) - (dt6/hy)*(AT(vmac,i,j+1,k-1,0)+AT(vmac,i,j,k-1,0))*(AT(simhy,i,j+1,k-1,0)-AT(simhy,i,j,k-1,0));
                r = AT(srz,i,j,k,0) - (dt6/hy)*(AT(vmac,i,j+1,k  ,0)+AT(vmac,i,j,k  ,0))*(AT(simhy,i,j+1,k  ,0)-AT(simhy,i,j,k  ,0));
            }
            if (k == ks)   bc_pair(&l,&r,2,0,PB(2,0),is_vel,comp,AT(s,i,j,ks-1,comp));
            if (k == ke+1) bc_pair(&l,&r,2,1,PB(2,1),is_vel,comp,AT(s,i,j,ke+1,comp));
            AT(simhzy,i,j,k,0) = upw(l, r, AT(wmac,i,j,k,0), eps);
        }
        /* 9. simhxz (is:ie+1, js-1:je+1, kk), kk=ks..ke, mkflux.f90:2150-2222 (kc == kk+1, kp == kk; k == kk+1) */
        #pragma omp parallel for collapse(2)
        for (int kk = ks; kk <= ke; ++kk)
        for (int j = js-1; j <= je+1; ++j)
        for (int i = is; i <= ie+1; ++i) {
            double l, r;
            if (cons) {
                l = AT(slx,i,j,kk,0) - (dt3/hz)*(AT(simhz,i-1,j,kk+1,0)*AT(wmac,i-1,j,kk+1,0) - AT(simhz,i-1,j,kk,0)*AT(wmac,i-1,j,kk,0));
                r = AT(srx,i,j,kk,0) - (dt3/hz)*(AT(simhz,i  ,j,kk+1,0)*AT(wmac,i  ,j,kk+1,0) - AT(simhz,i  ,j,kk,0)*AT(wmac,i  ,j,kk,0));
            } else {
                l = AT(slx,i,j,kk,0) - (dt6/hz)*(AT(wmac,i-1,j,kk+1,0)+AT(wmac,i-1,j,kk,0))*(AT(simhz,i-1,j,kk+1,0)-AT(simhz,i-1,j,kk,0));
                r = AT(srx,i,j,kk,0) - (dt6/hz)*(AT(wmac,i  ,j,kk+1,0)+AT(wmac,i  ,j,kk,0))*(AT(simhz,i  ,j,kk+1,0)-AT(simhz,i  ,j,kk,0));
            }
            if (i == is)   bc_pair(&l,&r,0,0,PB(0,0),is_vel,comp,AT(s,is-1,j,kk,comp));
            if (i == ie+1) bc_pair(&l,&r,0,1,PB(0,1),is_vel,comp,AT(s,ie+1,j,kk,comp));
            AT(simhxz,i,j,kk,0) = upw(l, r, AT(umac,i,j,kk,0), eps);
        }
        /* 10. simhyz (is-1:ie+1, js:je+1, kk), mkflux.f90:2230-2304 */
        #pragma omp parallel for collapse(2)
        for (int kk = ks; kk <= ke; ++kk)
        for (int j = js; j <= je+1; ++j)
        for (int i = is-1; i <= ie+1; ++i) {
            double l, r;
            if (cons) {
                l = AT(sly,i,j,kk,0) - (dt3/hz)*(AT(simhz,i,j-1,kk+1,0)*AT(wmac,i,j-1,kk+1,0) - AT(simhz,i,j-1,kk,0)*AT(wmac,i,j-1,kk,0));
                r = AT(sry,i,j,kk,0) - (dt3/hz)*(AT(simhz,i,j  ,kk+1,0)*AT(wmac,i,j  ,kk+1,0) - AT(simhz,i,j  ,kk,0)*AT(wmac,i,j  ,kk,0));
            } else {
                l = AT(sly,i,j,kk,0) - (dt6/hz)*(AT(wmac,i,j-1,kk+1,0)+AT(wmac,i,j-1,kk,0))*(AT(simhz,i,j-1,kk+1,0)-AT(simhz,i,j-1,kk,0));
                r = AT(sry,i,j,kk,0) - (dt6/hz)*(AT(wmac,i,j  ,kk+1,0)+AT(wmac,i,j  ,kk,0))*(AT(simhz,i,j  ,kk+1,0)-AT(simhz,i,j  ,kk,0));
            }
            if (j == js)   bc_pair(&l,&r,1,0,PB(1,0),is_vel,comp,AT(s,i,js-1,kk,comp));
            if (j == je+1) bc_pair(&l,&r,1,1,PB(1,1),is_vel,comp,AT(s,i,je+1,kk,comp));
            AT(simhyz,i,j,kk,0) = upw(l, r, AT(vmac,i,j,kk,0), eps);
        }

        /* 6. sedgez (is:ie, js:je, k), k=ks..ke+1, mkflux.f90:1870-1972 */
        #pragma omp parallel for collapse(2)
        for (int k = ks; k <= ke+1; ++k)
        for (int j = js; j <= je; ++j)
        for (int i = is; i <= ie; ++i) {
            double el, er;
            if (cons) {
                el = AT(slz,i,j,k,0)
                    - (dt2/hx)*(AT(simhxy,i+1,j,k-1,0)*AT(umac,i+1,j,k-1,0) - AT(simhxy,i,j,k-1,0)*AT(umac,i,j,k-1,0))
                    - (dt2/hy)*(AT(simhyx,i,j+1,k-1,0)*AT(vmac,i,j+1,k-1,0) - AT(simhyx,i,j,k-1,0)*AT(vmac,i,j,k-1,0))
                    + (dt2/hx)*AT(s,i,j,k-1,comp)*(AT(umac,i+1,j,k-1,0)-AT(umac,i,j,k-1,0))
                    + (dt2/hy)*AT(s,i,j,k-1,comp)*(AT(vmac,i,j+1,k-1,0)-AT(vmac,i,j,k-1,0));
                er = AT(srz,i,j,k,0)
                    - (dt2/hx)*(AT(simhxy,i+1,j,k,0)*AT(umac,i+1,j,k,0) - AT(simhxy,i,j,k,0)*AT(umac,i,j,k,0))
                    - (dt2/hy)*(AT(simhyx,i,j+1,k,0)*AT(vmac,i,j+1,k,0) - AT(simhyx,i,j,k,0)*AT(vmac,i,j,k,0))
                    + (dt2/hx)*AT(s,i,j,k,comp)*(AT(umac,i+1,j,k,0)-AT(umac,i,j,k,0))
                    + (dt2/hy)*AT(s,i,j,k,comp)*(AT(vmac,i,j+1,k,0)-AT(vmac,i,j,k,0));
            } else {
                el = AT(slz,i,j,k,0)
                    - (dt4/hx)*(AT(umac,i+1,j,k-1,0)+AT(umac,i,j,k-1,0))*(AT(simhxy,i+1,j,k-1,0)-AT(simhxy,i,j,k-1,0))
                    - (dt4/hy)*(AT(vmac,i,j+1,k-1,0)+AT(vmac,i,j,k-1,0))*(AT(simhyx,i,j+1,k-1,0)-AT(simhyx,i,j,k-1,0));
                er = AT(srz,i,j,k,0)
                    - (dt4/hx)*(AT(umac,i+1,j,k,0)+AT(umac,i,j,k,0))*(AT(simhxy,i+1,j,k,0)-AT(simhxy,i,j,k,0))
                    - (dt4/hy)*(AT(vmac,i,j+1,k,0)+AT(vmac,i,j,k,0))*(AT(simhyx,i,j+1,k,0)-AT(simhyx,i,j,k,0));
            }
            if (!use_minion) { el = el + dt2*AT(force,i,j,k-1,comp); er = er + dt2*AT(force,i,j,k,comp); }
            if (!use_minion && cons) { el = el - dt2*AT(s,i,j,k-1,comp)*AT(mac_rhs,i,j,k-1,0); er = er - dt2*AT(s,i,j,k,comp)*AT(mac_rhs,i,j,k,0); }
            double v = upw(el, er, AT(wmac,i,j,k,0), eps);
            if (k == ks)   v = bc_edge(v, el, er, 2, 0, PB(2,0), is_vel, comp, AT(s,i,j,ks-1,comp));
            if (k == ke+1) v = bc_edge(v, el, er, 2, 1, PB(2,1), is_vel, comp, AT(s,i,j,ke+1,comp));
            AT(sedgez,i,j,k,comp) = v;
            if (cons) AT(fluxz,i,j,k,comp) = v*AT(wmac,i,j,k,0);
        }
        /* 11. sedgex (is:ie+1, js:je, kk), mkflux.f90:2310-2408 */
        #pragma omp parallel for collapse(2)
        for (int kk = ks; kk <= ke; ++kk)
        for (int j = js; j <= je; ++j)
        for (int i = is; i <= ie+1; ++i) {
            double el, er;
            if (cons) {
                el = AT(slx,i,j,kk,0)
                    - (dt2/hy)*(AT(simhyz,i-1,j+1,kk,0)*AT(vmac,i-1,j+1,kk,0) - AT(simhyz,i-1,j,kk,0)*AT(vmac,i-1,j,kk,0))
                    - (dt2/hz)*(AT(simhzy,i-1,j,kk+1,0)*AT(wmac,i-1,j,kk+1,0) - AT(simhzy,i-1,j,kk,0)*AT(wmac,i-1,j,kk,0))
                    + (dt2/hy)*AT(s,i-1,j,kk,comp)*(AT(vmac,i-1,j+1,kk,0)-AT(vmac,i-1,j,kk,0))
                    + (dt2/hz)*AT(s,i-1,j,kk,comp)*(AT(wmac,i-1,j,kk+1,0)-AT(wmac,i-1,j,kk,0));
                er = AT(srx,i,j,kk,0)
                    - (dt2/hy)*(AT(simhyz,i,j+1,kk,0)*AT(vmac,i,j+1,kk,0) - AT(simhyz,i,j,kk,0)*AT(vmac,i,j,kk,0))
                    - (dt2/hz)*(AT(simhzy,i,j,kk+1,0)*AT(wmac,i,j,kk+1,0) - AT(simhzy,i,j,kk,0)*AT(wmac,i,j,kk,0))
                    + (dt2/hy)*AT(s,i,j,kk,comp)*(AT(vmac,i,j+1,kk,0)-AT(vmac,i,j,kk,0))
                    + (dt2/hz)*AT(s,i,j,kk,comp)*(AT(wmac,i,j,kk+1,0)-AT(wmac,i,j,kk,0));
            } else {
                el = AT(slx,i,j,kk,0)
                    - (dt4/hy)*(AT(vmac,i-1,j+1,kk,0)+AT(vmac,i-1,j,kk,0))*(AT(simhyz,i-1,j+1,kk,0)-AT(simhyz,i-1,j,kk,0))
                    - (dt4/hz)*(AT(wmac,i-1,j,kk+1,0)+AT(wmac,i-1,j,kk,0))*(AT(simhzy,i-1,j,kk+1,0)-AT(simhzy,i-1,j,kk,0));
                er = AT(srx,i,j,kk,0)
                    - (dt4/hy)*(AT(vmac,i,j+1,kk,0)+AT(vmac,i,j,kk,0))*(AT(simhyz,i,j+1,kk,0)-AT(simhyz,i,j,kk,0))
                    - (dt4/hz)*(AT(wmac,i,j,kk+1,0)+AT(wmac,i,j,kk,0))*(AT(simhzy,i,j,kk+1,0)-AT(simhzy,i,j,kk,0));
            }
            if (!use_minion) { el = el + dt2*AT(force,i-1,j,kk,comp); er = er + dt2*AT(force,i,j,kk,comp); }
            if (!use_minion && cons) { el = el - dt2*AT(s,i-1,j,kk,comp)*AT(mac_rhs,i-1,j,kk,0); er = er - dt2*AT(s,i,j,kk,comp)*AT(mac_rhs,i,j,kk,0); }
            double v = upw(el, er, AT(umac,i,j,kk,0), eps);
            if (i == is)   v = bc_edge(v, el, er, 0, 0, PB(0,0), is_vel, comp, AT(s,is-1,j,kk,comp));
            if (i == ie+1) v = bc_edge(v, el, er, 0, 1, PB(0,1), is_vel, comp, AT(s,ie+1,j,kk,comp));
            AT(sedgex,i,j,kk,comp) = v;
            if (cons) AT(fluxx,i,j,kk,comp) = v*AT(umac,i,j,kk,0);
        }
        /* 12. sedgey (is:ie, js:je+1, kk), mkflux.f90:2414-2512 */
        #pragma omp parallel for collapse(2)
        for (int kk = ks; kk <= ke; ++kk)
        for (int j = js; j <= je+1; ++j)
        for (int i = is; i <= ie; ++i) {
            double el, er;
            if (cons) {
                el = AT(sly,i,j,kk,0)
                    - (dt2/hx)*(AT(simhxz,i+1,j-1,kk,0)*AT(umac,i+1,j-1,kk,0) - AT(simhxz,i,j-1,kk,0)*AT(umac,i,j-1,kk,0))
                    - (dt2/hz)*(AT(simhzx,i,j-1,kk+1,0)*AT(wmac,i,j-1,kk+1,0) - AT(simhzx,i,j-1,kk,0)*AT(wmac,i,j-1,kk,0))
                    + (dt2/hx)*AT(s,i,j-1,kk,comp)*(AT(umac,i+1,j-1,kk,0)-AT(umac,i,j-1,kk,0))
                    + (dt2/hz)*AT(s,i,j-1,kk,comp)*(AT(wmac,i,j-1,kk+1,0)-AT(wmac,i,j-1,kk,0));
                er = AT(sry,i,j,kk,0)
                    - (dt2/hx)*(AT(simhxz,i+1,j,kk,0)*AT(umac,i+1,j,kk,0) - AT(simhxz,i,j,kk,0)*AT(umac,i,j,kk,0))
                    - (dt2/hz)*(AT(simhzx,i,j,kk+1,0)*AT(wmac,i,j,kk+1,0) - AT(simhzx,i,j,kk,0)*AT(wmac,i,j,kk,0))
                    + (dt2/hx)*AT(s,i,j,kk,comp)*(AT(umac,i+1,j,kk,0)-AT(umac,i,j,kk,0))
                    + (dt2/hz)*AT(s,i,j,kk,comp)*(AT(wmac,i,j,kk+1,0)-AT(wmac,i,j,kk,0));
            } else {
                el = AT(sly,i,j,kk,0)
                    - (dt4/hx)*(AT(umac,i+1,j-1,kk,0)+AT(umac,i,j-1,kk,0))*(AT(simhxz,i+1,j-1,kk,0)-AT(simhxz,i,j-1,kk,0))
                    - (dt4/hz)*(AT(wmac,i,j-1,kk+1,0)+AT(wmac,i,j-1,kk,0))*(AT(simhzx,i,j-1,kk+1,0)-AT(simhzx,i,j-1,kk,0));
                er = AT(sry,i,j,kk,0)
                    - (dt4/hx)*(AT(umac,i+1,j,kk,0)+AT(umac,i,j,kk,0))*(AT(simhxz,i+1,j,kk,0)-AT(simhxz,i,j,kk,0))
                    - (dt4/hz)*(AT(wmac,i,j,kk+1,0)+AT(wmac,i,j,kk,0))*(AT(simhzx,i,j,kk+1,0)-AT(simhzx,i,j,kk,0));
            }
            if (!use_minion) { el = el + dt2*AT(force,i,j-1,kk,comp); er = er + dt2*AT(force,i,j,kk,comp); }
            if (!use_minion && cons) { el = el - dt2*AT(s,i,j-1,kk,comp)*AT(mac_rhs,i,j-1,kk,0); er = er - dt2*AT(s,i,j,kk,comp)*AT(mac_rhs,i,j,kk,0); }
            double v = upw(el, er, AT(vmac,i,j,kk,0), eps);
            if (j == js)   v = bc_edge(v, el, er, 1, 0, PB(1,0), is_vel, comp, AT(s,i,js-1,kk,comp));
            if (j == je+1) v = bc_edge(v, el, er, 1, 1, PB(1,1), is_vel, comp, AT(s,i,je+1,kk,comp));
            AT(sedgey,i,j,kk,comp) = v;
            if (cons) AT(fluxy,i,j,kk,comp) = v*AT(vmac,i,j,kk,0);
        }
    }

    v_free(&slopex); v_free(&slopey); v_free(&slopez);
    v_free(&slx); v_free(&srx); v_free(&simhx); v_free(&sly); v_free(&sry); v_free(&simhy);
    v_free(&slz); v_free(&srz); v_free(&simhz);
    v_free(&simhxy); v_free(&simhxz); v_free(&simhyx); v_free(&simhyz); v_free(&simhzx); v_free(&simhzy);
#undef PB
}

/* mkflux_2d, mkflux.f90:152-691 */
void orc_mkflux_2d(const double *s_, double *sedgex_, double *sedgey_, double *fluxx_, double *fluxy_,
                   const double *umac_, const double *vmac_, const double *force_, const double *mac_rhs_,
                   const int *lo, const int *hi, const double *dx, double dt, int is_vel,
                   const int *phys_bc, const int *adv_bc,
                   int ng_s, int ng_e, int ng_f, int ng_u, int ng_o, int ng_m,
                   const int *is_cons, int ncomp, int use_minion, int slope_order)
{
    const int is = lo[0], ie = hi[0], js = lo[1], je = hi[1];
    V s      = v_box((double*)s_, lo, hi, ng_s, -1, ncomp, 2);
    V sedgex = v_box(sedgex_, lo, hi, ng_e, 0, ncomp, 2);
    V sedgey = v_box(sedgey_, lo, hi, ng_e, 1, ncomp, 2);
    V fluxx  = v_box(fluxx_, lo, hi, ng_f, 0, ncomp, 2);
    V fluxy  = v_box(fluxy_, lo, hi, ng_f, 1, ncomp, 2);
    V umac   = v_box((double*)umac_, lo, hi, ng_u, 0, 1, 2);
    V vmac   = v_box((double*)vmac_, lo, hi, ng_u, 1, 1, 2);
    V force  = v_box((double*)force_, lo, hi, ng_o, -1, ncomp, 2);
    V mac_rhs= v_box((double*)mac_rhs_, lo, hi, ng_m, -1, 1, 2);
#define PB(d,sd) phys_bc[(d)*2+(sd)]
    V slopex = v_alloc(is-1,ie+1, js-1,je+1, 0,0, ncomp);
    V slopey = v_alloc(is-1,ie+1, js-1,je+1, 0,0, ncomp);
    orc_slope(&s, &slopex, lo, hi, 2, 0, ncomp, adv_bc, slope_order);
    orc_slope(&s, &slopey, lo, hi, 2, 1, ncomp, adv_bc, slope_order);
    V slx = v_alloc(is,ie+1, js-1,je+1, 0,0, 1), srx = v_alloc(is,ie+1, js-1,je+1, 0,0, 1), simhx = v_alloc(is,ie+1, js-1,je+1, 0,0, 1);
    V sly = v_alloc(is-1,ie+1, js,je+1, 0,0, 1), sry = v_alloc(is-1,ie+1, js,je+1, 0,0, 1), simhy = v_alloc(is-1,ie+1, js,je+1, 0,0, 1);

    const double dt2 = HALF*dt, dt4 = dt/4.0;
    const double hx = dx[0], hy = dx[1];
    const double eps = orc_mkflux_eps(&umac, &vmac, &vmac, lo, hi, 2);

    for (int comp = 0; comp < ncomp; ++comp) {
        const int cons = is_cons[comp];
        for (int j = js-1; j <= je+1; ++j)
        for (int i = is; i <= ie+1; ++i) {
            double l = AT(s,i-1,j,0,comp) + (HALF - dt2*AT(umac,i,j,0,0)/hx)*AT(slopex,i-1,j,0,comp);
            double r = AT(s,i  ,j,0,comp) - (HALF + dt2*AT(umac,i,j,0,0)/hx)*AT(slopex,i  ,j,0,comp);
            if (use_minion) { l = l + dt2*AT(force,i-1,j,0,comp); r = r + dt2*AT(force,i,j,0,comp); }
            if (use_minion && cons) { l = l - dt2*AT(s,i-1,j,0,comp)*AT(mac_rhs,i-1,j,0,0); r = r - dt2*AT(s,i,j,0,comp)*AT(mac_rhs,i,j,0,0); }
            if (i == is)   bc_pair(&l,&r,0,0,PB(0,0),is_vel,comp,AT(s,is-1,j,0,comp));
            if (i == ie+1) bc_pair(&l,&r,0,1,PB(0,1),is_vel,comp,AT(s,ie+1,j,0,comp));
            AT(slx,i,j,0,0) = l; AT(srx,i,j,0,0) = r;
            AT(simhx,i,j,0,0) = upw(l, r, AT(umac,i,j,0,0), eps);
        }
        for (int j = js; j <= je+1; ++j)
        for (int i = is-1; i <= ie+1; ++i) {
            double l = AT(s,i,j-1,0,comp) + (HALF - dt2*AT(vmac,i,j,0,0)/hy)*AT(slopey,i,j-1,0,comp);
            double r = AT(s,i,j  ,0,comp) - (HALF + dt2*AT(vmac,i,j,0,0)/hy)*AT(slopey,i,j  ,0,comp);
            if (use_minion) { l = l + dt2*AT(force,i,j-1,0,comp); r = r + dt2*AT(force,i,j,0,comp); }
            if (use_minion && cons) { l = l - dt2*AT(s,i,j-1,0,comp)*AT(mac_rhs,i,j-1,0,0); r = r - dt2*AT(s,i,j,0,comp)*AT(mac_rhs,i,j,0,0); }
            if (j == js)   bc_pair(&l,&r,1,0,PB(1,0),is_vel,comp,AT(s,i,js-1,0,comp));
            if (j == je+1) bc_pair(&l,&r,1,1,PB(1,1),is_vel,comp,AT(s,i,je+1,0,comp));
            AT(sly,i,j,0,0) = l; AT(sry,i,j,0,0) = r;
            AT(simhy,i,j,0,0) = upw(l, r, AT(vmac,i,j,0,0), eps);
        }
        /* sedgey (is:ie, j), mkflux.f90:470-566 */
        for (int j = js; j <= je+1; ++j)
        for (int i = is; i <= ie; ++i) {
            double el, er;
            if (cons) {
                el = AT(sly,i,j,0,0)
                    - (dt2/hx)*(AT(simhx,i+1,j-1,0,0)*AT(umac,i+1,j-1,0,0) - AT(simhx,i,j-1,0,0)*AT(umac,i,j-1,0,0))
                    + (dt2/hx)*AT(s,i,j-1,0,comp)*(AT(umac,i+1,j-1,0,0)-AT(umac,i,j-1,0,0));
                er = AT(sry,i,j,0,0)
                    - (dt2/hx)*(AT(simhx,i+1,j,0,0)*AT(umac,i+1,j,0,0) - AT(simhx,i,j,0,0)*AT(umac,i,j,0,0))
                    + (dt2/hx)*AT(s,i,j,0,comp)*(AT(umac,i+1,j,0,0)-AT(umac,i,j,0,0));
            } else {
                el = AT(sly,i,j,0,0) - (dt4/hx)*(AT(umac,i+1,j-1,0,0)+AT(umac,i,j-1,0,0))*(AT(simhx,i+1,j-1,0,0)-AT(simhx,i,j-1,0,0));
                er = AT(sry,i,j,0,0) - (dt4/hx)*(AT(umac,i+1,j  ,0,0)+AT(umac,i,j  ,0,0))*(AT(simhx,i+1,j  ,0,0)-AT(simhx,i,j  ,0,0));
            }
            if (!use_minion) { el = el + dt2*AT(force,i,j-1,0,comp); er = er + dt2*AT(force,i,j,0,comp); }
            if (!use_minion && cons) { el = el - dt2*AT(s,i,j-1,0,comp)*AT(mac_rhs,i,j-1,0,0); er = er - dt2*AT(s,i,j,0,comp)*AT(mac_rhs,i,j,0,0); }
            double v = upw(el, er, AT(vmac,i,j,0,0), eps);
            if (j == js)   v = bc_edge(v, el, er, 1, 0, PB(1,0), is_vel, comp, AT(s,i,js-1,0,comp));
            if (j == je+1) v = bc_edge(v, el, er, 1, 1, PB(1,1), is_vel, comp, AT(s,i,je+1,0,comp));
            AT(sedgey,i,j,0,comp) = v;
            if (cons) AT(fluxy,i,j,0,comp) = v*AT(vmac,i,j,0,0);
        }
        /* sedgex (is:ie+1, jj), mkflux.f90:572-666 */
        for (int jj = js; jj <= je; ++jj)
        for (int i = is; i <= ie+1; ++i) {
            double el, er;
            if (cons) {
                el = AT(slx,i,jj,0,0)
                    - (dt2/hy)*(AT(simhy,i-1,jj+1,0,0)*AT(vmac,i-1,jj+1,0,0) - AT(simhy,i-1,jj,0,0)*AT(vmac,i-1,jj,0,0))
                    + (dt2/hy)*AT(s,i-1,jj,0,comp)*(AT(vmac,i-1,jj+1,0,0)-AT(vmac,i-1,jj,0,0));
                er = AT(srx,i,jj,0,0)
                    - (dt2/hy)*(AT(simhy,i,jj+1,0,0)*AT(vmac,i,jj+1,0,0) - AT(simhy,i,jj,0,0)*AT(vmac,i,jj,0,0))
                    + (dt2/hy)*AT(s,i,jj,0,comp)*(AT(vmac,i,jj+1,0,0)-AT(vmac,i,jj,0,0));
            } else {
                el = AT(slx,i,jj,0,0) - (dt4/hy)*(AT(vmac,i-1,jj+1,0,0)+AT(vmac,i-1,jj,0,0))*(AT(simhy,i-1,jj+1,0,0)-AT(simhy,i-1,jj,0,0));
                er = AT(srx,i,jj,0,0) - (dt4/hy)*(AT(vmac,i  ,jj+1,0,0)+AT(vmac,i  ,jj,0,0))*(AT(simhy,i  ,jj+1,0,0)-AT(simhy,i  ,jj,0,0));
            }
            if (!use_minion) { el = el + dt2*AT(force,i-1,jj,0,comp); er = er + dt2*AT(force,i,jj,0,comp); }
            if (!use_minion && cons) { el = el - dt2*AT(s,i-1,jj,0,comp)*AT(mac_rhs,i-1,jj,0,0); er = er - dt2*AT(s,i,jj,0,comp)*AT(mac_rhs,i,jj,0,0); }
            double v = upw(el, er, AT(umac,i,jj,0,0), eps);
            if (i == is)   v = bc_edge(v, el, er, 0, 0, PB(0,0), is_vel, comp, AT(s,is-1,jj,0,comp));
            if (i == ie+1) v = bc_edge(v, el, er, 0, 1, PB(0,1), is_vel, comp, AT(s,ie+1,jj,0,comp));
            AT(sedgex,i,jj,0,comp) = v;
            if (cons) AT(fluxx,i,jj,0,comp) = v*AT(umac,i,jj,0,0);
        }
    }
    v_free(&slopex); v_free(&slopey);
    v_free(&slx); v_free(&srx); v_free(&simhx); v_free(&sly); v_free(&sry); v_free(&simhy);
#undef PB
}
