/*
 * oracle/orc_slope.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of src/slope.f90 (slopex_2d :148, slopey_2d :291, slopez_3d :437).
 * The three reference routines are the same arithmetic along different axes, so
 * one axis-generic routine restates them; `dir` selects the axis.
 */
#include "orc_common.h"

/* One pencil along `dir` through cell (i0,j0,k0) (the index along dir is ignored).
 * s has >= 3 ghost cells along dir (order 4) ; sl is written on lo-1..hi+1 along dir.
 * bclo/bchi: adv_bc(dir,1,comp), adv_bc(dir,2,comp). */
static void slope_pencil(const V *s, V *sl, int comp, int dir, int i0, int j0, int k0,
                         int lo, int hi, int bclo, int bchi, int order)
{
    /* pointer to the pencil element whose index along dir equals 0, and the stride along dir */
    int ix[3] = { i0, j0, k0 };
    ix[dir] = s->l[dir];
    const long sst = (dir == 0) ? 1 : (dir == 1 ? s->n[0] : s->n[0]*s->n[1]);
    const double *sp = &AT(*s, ix[0], ix[1], ix[2], comp) - (long)s->l[dir]*sst;
    ix[dir] = sl->l[dir];
    const long lst = (dir == 0) ? 1 : (dir == 1 ? sl->n[0] : sl->n[0]*sl->n[1]);
    double *lp = &AT(*sl, ix[0], ix[1], ix[2], comp) - (long)sl->l[dir]*lst;
#define S_(m)  (sp[(long)(m)*sst])
#define SL_(m) (lp[(long)(m)*lst])
    const int is = lo, ie = hi;
    double del, slim, sflag, dpls, dmn, ds;

    if (order == 0) {                                    /* slope.f90:172-173 */
        for (int i = is-1; i <= ie+1; ++i) SL_(i) = ZERO;
        return;
    }
    if (order == 2) {                                    /* slope.f90:178-218 */
        for (int i = is-1; i <= ie+1; ++i) {
            del  = HALF*(S_(i+1) - S_(i-1));
            dpls = TWO*(S_(i+1) - S_(i));
            dmn  = TWO*(S_(i) - S_(i-1));
            slim = dmin(fabs(dpls), fabs(dmn));
            slim = (dpls*dmn > ZERO) ? slim : ZERO;
            sflag = copysign(ONE, del);
            SL_(i) = sflag*dmin(slim, fabs(del));
        }
        if (bclo == BC_EXT_DIR || bclo == BC_HOEXTRAP) {
            SL_(is-1) = ZERO;
            del  = (S_(is+1) + 3.0*S_(is) - 4.0*S_(is-1)) * (1.0/3.0);
            dpls = TWO*(S_(is+1) - S_(is));
            dmn  = TWO*(S_(is) - S_(is-1));
            slim = dmin(fabs(dpls), fabs(dmn));
            slim = (dpls*dmn > ZERO) ? slim : ZERO;
            sflag = copysign(ONE, del);
            SL_(is) = sflag*dmin(slim, fabs(del));
        }
        if (bchi == BC_EXT_DIR || bchi == BC_HOEXTRAP) {
            SL_(ie+1) = ZERO;
            del  = -(S_(ie-1) + 3.0*S_(ie) - 4.0*S_(ie+1)) * (1.0/3.0);
            dpls = TWO*(S_(ie) - S_(ie-1));
            dmn  = TWO*(S_(ie+1) - S_(ie));
            slim = dmin(fabs(dpls), fabs(dmn));
            slim = (dpls*dmn > ZERO) ? slim : ZERO;
            sflag = copysign(ONE, del);
            SL_(ie) = sflag*dmin(slim, fabs(del));
        }
        return;
    }

    /* 4th order: slope.f90:223-285 */
    const int n = ie - is + 5;                            /* is-2 .. ie+2 */
    double *scr = (double*)malloc(sizeof(double)*4*(size_t)n);
    double *cen = scr, *lim = scr+n, *flg = scr+2*n, *frm = scr+3*n;
#define X(a,i) a[(i)-(is-2)]
    const double two3rd = 2.0/3.0, sixth = 1.0/6.0, tenth = 0.1, sixteen = 16.0, fifteen = 15.0;
    for (int i = is-2; i <= ie+2; ++i) {
        X(cen,i) = HALF*(S_(i+1) - S_(i-1));
        dmn  = TWO*(S_(i) - S_(i-1));
        dpls = TWO*(S_(i+1) - S_(i));
        X(lim,i) = dmin(fabs(dmn), fabs(dpls));
        X(lim,i) = (dpls*dmn > ZERO) ? X(lim,i) : ZERO;
        X(flg,i) = copysign(ONE, X(cen,i));
        X(frm,i) = X(flg,i)*dmin(X(lim,i), fabs(X(cen,i)));
    }
    for (int i = is-1; i <= ie+1; ++i) {
        ds = TWO * two3rd * X(cen,i) - sixth * (X(frm,i+1) + X(frm,i-1));
        SL_(i) = X(flg,i)*dmin(fabs(ds), X(lim,i));
    }
    if (bclo == BC_EXT_DIR || bclo == BC_HOEXTRAP) {
        SL_(is-1) = ZERO;
        del  = -sixteen/fifteen*S_(is-1) + HALF*S_(is) + two3rd*S_(is+1) - tenth*S_(is+2);
        dmn  = TWO*(S_(is) - S_(is-1));
        dpls = TWO*(S_(is+1) - S_(is));
        slim = dmin(fabs(dpls), fabs(dmn));
        slim = (dpls*dmn > ZERO) ? slim : ZERO;
        sflag = copysign(ONE, del);
        SL_(is) = sflag*dmin(slim, fabs(del));
        /* recompute is+1 with the revised fromm(is) */
        X(frm,is) = SL_(is);
        ds = TWO * two3rd * X(cen,is+1) - sixth * (X(frm,is+2) + X(frm,is));
        SL_(is+1) = X(flg,is+1)*dmin(fabs(ds), X(lim,is+1));
    }
    if (bchi == BC_EXT_DIR || bchi == BC_HOEXTRAP) {
        SL_(ie+1) = ZERO;
        del  = -( -sixteen/fifteen*S_(ie+1) + HALF*S_(ie) + two3rd*S_(ie-1) - tenth*S_(ie-2) );
        dmn  = TWO*(S_(ie) - S_(ie-1));
        dpls = TWO*(S_(ie+1) - S_(ie));
        slim = dmin(fabs(dpls), fabs(dmn));
        slim = (dpls*dmn > ZERO) ? slim : ZERO;
        sflag = copysign(ONE, del);
        SL_(ie) = sflag*dmin(slim, fabs(del));
        X(frm,ie) = SL_(ie);
        ds = TWO * two3rd * X(cen,ie-1) - sixth * (X(frm,ie-2) + X(frm,ie));
        SL_(ie-1) = X(flg,ie-1)*dmin(fabs(ds), X(lim,ie-1));
    }
    free(scr);
#undef X
#undef S_
#undef SL_
}

/* Slopes along `dir` for all comps on the box grown by one in every direction
 * (the range the 3-D callers build plane by plane: velpred.f90:1848-1852).
 * adv_bc is the C table [comp][dir][side] (comp slowest) starting at the first comp of s. */
void orc_slope(const V *s, V *sl, const int *lo, const int *hi, int dim, int dir, int ncomp,
               const int *adv_bc /* [ncomp][3][2] */, int order)
{
    int glo[3], ghi[3];
    for (int d = 0; d < 3; ++d) {
        if (d < dim) { glo[d] = lo[d]-1; ghi[d] = hi[d]+1; } else { glo[d] = 0; ghi[d] = 0; }
    }
    for (int comp = 0; comp < ncomp; ++comp) {
        int bclo = adv_bc[(comp*3+dir)*2+0];
        int bchi = adv_bc[(comp*3+dir)*2+1];
        int a = (dir+1)%3, b = (dir+2)%3;
        int ix[3];
        #pragma omp parallel for collapse(2) private(ix)
        for (int q = glo[b]; q <= ghi[b]; ++q)
            for (int p = glo[a]; p <= ghi[a]; ++p) {
                ix[a] = p; ix[b] = q; ix[dir] = 0;
                slope_pencil(s, sl, comp, dir, ix[0], ix[1], ix[2], lo[dir], hi[dir], bclo, bchi, order);
            }
    }
}
