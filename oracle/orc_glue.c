/*
 * oracle/orc_glue.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the streaming / glue routines on the hot path:
 *   update_3d/_2d          src/update.f90:186-278, :113-184
 *   physbc_3d/_2d          src/multifab_physbc.f90:238-561, :64-236
 *   divumac_3d/_2d         src/macproject.f90:250-278, :227-248   (+ rh = mac_rhs - rh, :189-196)
 *   mk_mac_coeffs_3d/_2d   src/macproject.f90:361-401, :338-359
 *   mkumac_3d/_2d          src/macproject.f90:578-645, :533-576   (box-boundary faces: see note)
 *   mkvelforce_3d/_2d      src/mkforce.f90:144-236, :82-142  (valid cells only; ghosts come from fill+FOEXTRAP, :75-76)
 *   mkscalforce_3d/_2d     src/mkforce.f90:333-402, :290-331
 *   make_at_halftime       src/make_at_halftime.f90:80-115
 */
#include "orc_common.h"

void orc_update(const double *sold_, const double *umac_, const double *vmac_, const double *wmac_,
                const double *sedgex_, const double *sedgey_, const double *sedgez_,
                const double *fluxx_, const double *fluxy_, const double *fluxz_,
                const double *force_, double *snew_,
                const int *lo, const int *hi, int dim, int ng_s, int ng_u, int ng_e, int ng_f, int ng_o,
                const double *dx, double dt, int is_vel, const int *is_cons, int ncomp)
{
    V sold = v_box((double*)sold_, lo, hi, ng_s, -1, ncomp, dim);
    V snew = v_box(snew_, lo, hi, ng_s, -1, ncomp, dim);
    V umac = v_box((double*)umac_, lo, hi, ng_u, 0, 1, dim);
    V vmac = v_box((double*)vmac_, lo, hi, ng_u, 1, 1, dim);
    V wmac = v_box((double*)(dim == 3 ? wmac_ : vmac_), lo, hi, ng_u, 2, 1, dim);
    V sex = v_box((double*)sedgex_, lo, hi, ng_e, 0, ncomp, dim);
    V sey = v_box((double*)sedgey_, lo, hi, ng_e, 1, ncomp, dim);
    V sez = v_box((double*)(dim == 3 ? sedgez_ : sedgey_), lo, hi, ng_e, 2, ncomp, dim);
    V fx = v_box((double*)fluxx_, lo, hi, ng_f, 0, ncomp, dim);
    V fy = v_box((double*)fluxy_, lo, hi, ng_f, 1, ncomp, dim);
    V fz = v_box((double*)(dim == 3 ? fluxz_ : fluxy_), lo, hi, ng_f, 2, ncomp, dim);
    V force = v_box((double*)force_, lo, hi, ng_o, -1, ncomp, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;

    for (int comp = 0; comp < ncomp; ++comp) {
        const int conservative = (!is_vel) && is_cons[comp];
        #pragma omp parallel for
        for (int k = k0; k <= k1; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
            double adv;
            if (conservative) {
                if (dim == 3)
                    adv = (AT(fx,i+1,j,k,comp)-AT(fx,i,j,k,comp))/dx[0]
                        + (AT(fy,i,j+1,k,comp)-AT(fy,i,j,k,comp))/dx[1]
                        + (AT(fz,i,j,k+1,comp)-AT(fz,i,j,k,comp))/dx[2];
                else
                    adv = (AT(fx,i+1,j,k,comp)-AT(fx,i,j,k,comp))/dx[0]
                        + (AT(fy,i,j+1,k,comp)-AT(fy,i,j,k,comp))/dx[1];
            } else {
                double ubar = HALF*(AT(umac,i,j,k,0) + AT(umac,i+1,j,k,0));
                double vbar = HALF*(AT(vmac,i,j,k,0) + AT(vmac,i,j+1,k,0));
                if (dim == 3) {
                    double wbar = HALF*(AT(wmac,i,j,k,0) + AT(wmac,i,j,k+1,0));
                    adv = ubar*(AT(sex,i+1,j,k,comp) - AT(sex,i,j,k,comp))/dx[0]
                        + vbar*(AT(sey,i,j+1,k,comp) - AT(sey,i,j,k,comp))/dx[1]
                        + wbar*(AT(sez,i,j,k+1,comp) - AT(sez,i,j,k,comp))/dx[2];
                } else {
                    adv = ubar*(AT(sex,i+1,j,k,comp) - AT(sex,i,j,k,comp))/dx[0]
                        + vbar*(AT(sey,i,j+1,k,comp) - AT(sey,i,j,k,comp))/dx[1];
                }
            }
            AT(snew,i,j,k,comp) = AT(sold,i,j,k,comp) - dt*adv + dt*AT(force,i,j,k,comp);
        }
    }
}

/* physbc on ONE component. bc[d][side] = adv_bc for this comp on this box; icomp is the 1-based bc component
 * used to pick the EXT_DIR constant (1..dm velocity, dm+1 rho, dm+2 tracer); bcval[5][3][2] = u_bc,v_bc,w_bc,rho_bc,trac_bc.
 * Sweep order and ranges exactly as multifab_physbc.f90:254-561 (x faces, then y over full x, then z over full x,y). */
void orc_physbc(double *s_, const int *lo, const int *hi, int dim, int ng, const int *bc, int icomp, const double *bcval)
{
    if (ng == 0) return;
    V s = v_box(s_, lo, hi, ng, -1, 1, dim);
#define BCV(d,sd) bc[(d)*2+(sd)]
    int glo[3], ghi[3], nglo[3], nghi[3];
    for (int d = 0; d < 3; ++d) {
        if (d < dim) { glo[d] = lo[d]-ng; ghi[d] = hi[d]+ng; } else { glo[d] = 0; ghi[d] = 0; }
        nglo[d] = (d < dim && BCV(d,0) == BC_INTERIOR) ? ng : 0;
        nghi[d] = (d < dim && BCV(d,1) == BC_INTERIOR) ? ng : 0;
    }
    /* the EXT_DIR constant: 2-D uses icomp 1,2 vel, 3 rho, 4 trac (multifab_physbc.f90:98-101); 3-D 1..3,4,5 */
    int slot;
    if (dim == 2) slot = (icomp == 1) ? 0 : (icomp == 2) ? 1 : (icomp == 3) ? 3 : (icomp == 4) ? 4 : -1;
    else          slot = (icomp >= 1 && icomp <= 5) ? icomp-1 : -1;

    for (int d = 0; d < dim; ++d) {
        /* transverse ranges: directions below d use the full ghosted range, directions above d skip
         * ghost rows on non-interior sides (only x-sweep skips y,z; y-sweep skips z; z skips none) */
        int rlo[3], rhi[3];
        for (int t = 0; t < 3; ++t) {
            if (t >= dim) { rlo[t] = 0; rhi[t] = 0; }
            else if (t < d) { rlo[t] = glo[t]; rhi[t] = ghi[t]; }
            else if (t > d) { rlo[t] = lo[t]-nglo[t]; rhi[t] = hi[t]+nghi[t]; }
        }
        for (int side = 0; side < 2; ++side) {
            const int b = BCV(d,side);
            if (b == BC_INTERIOR) continue;
            /* EXT_DIR assigns over the full ghosted transverse range (:283-287) */
            int tlo[3], thi[3];
            for (int t = 0; t < 3; ++t) {
                if (t == d) continue;
                if (b == BC_EXT_DIR && t < dim) { tlo[t] = glo[t]; thi[t] = ghi[t]; }
                else { tlo[t] = rlo[t]; thi[t] = rhi[t]; }
            }
            const int a = (d+1)%3, c = (d+2)%3;
            for (int q = tlo[c]; q <= thi[c]; ++q)
            for (int p = tlo[a]; p <= thi[a]; ++p) {
                int ix[3]; ix[a] = p; ix[c] = q;
                const int e  = side == 0 ? lo[d] : hi[d];   /* first interior cell */
                const int sg = side == 0 ? -1 : +1;          /* outward direction   */
#define SS(m) (ix[d] = (m), &AT(s, ix[0], ix[1], ix[2], 0))
                if (b == BC_EXT_DIR) {
                    if (slot >= 0) { double v = bcval[(slot*3+d)*2+side]; for (int g = 1; g <= ng; ++g) *SS(e+sg*g) = v; }
                } else if (b == BC_FOEXTRAP) {
                    double v = *SS(e); for (int g = 1; g <= ng; ++g) *SS(e+sg*g) = v;
                } else if (b == BC_HOEXTRAP) {
                    double s0 = *SS(e), s1 = *SS(e-sg), s2 = *SS(e-2*sg);
                    double v = ( 15.0*s0 - 10.0*s1 + 3.0*s2 ) * 0.125;
                    for (int g = 1; g <= ng; ++g) *SS(e+sg*g) = v;
                } else if (b == BC_REFLECT_EVEN) {
                    for (int g = 1; g <= ng; ++g) { double v = *SS(e-sg*(g-1)); *SS(e+sg*g) = v; }
                } else if (b == BC_REFLECT_ODD) {
                    for (int g = 1; g <= ng; ++g) { double v = *SS(e-sg*(g-1)); *SS(e+sg*g) = -v; }
                }
#undef SS
            }
        }
    }
#undef BCV
}

/* rh = mac_rhs - div(umac), macproject.f90:189-196 + :250-278 */
void orc_divumac(const double *umac_, const double *vmac_, const double *wmac_, int ng_um,
                 const double *mac_rhs_, int ng_m, double *rh_, int ng_rh,
                 const double *dx, const int *lo, const int *hi, int dim)
{
    V umac = v_box((double*)umac_, lo, hi, ng_um, 0, 1, dim);
    V vmac = v_box((double*)vmac_, lo, hi, ng_um, 1, 1, dim);
    V wmac = v_box((double*)(dim == 3 ? wmac_ : vmac_), lo, hi, ng_um, 2, 1, dim);
    V mrhs = v_box((double*)mac_rhs_, lo, hi, ng_m, -1, 1, dim);
    V rh   = v_box(rh_, lo, hi, ng_rh, -1, 1, dim);
    double dxinv[3] = { 1.0/dx[0], 1.0/dx[1], dim == 3 ? 1.0/dx[2] : 0.0 };
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    #pragma omp parallel for
    for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
    for (int i = lo[0]; i <= hi[0]; ++i) {
        double d;
        if (dim == 3)
            d = (AT(umac,i+1,j,k,0) - AT(umac,i,j,k,0)) * dxinv[0] +
                (AT(vmac,i,j+1,k,0) - AT(vmac,i,j,k,0)) * dxinv[1] +
                (AT(wmac,i,j,k+1,0) - AT(wmac,i,j,k,0)) * dxinv[2];
        else
            d = (AT(umac,i+1,j,k,0) - AT(umac,i,j,k,0)) * dxinv[0] +
                (AT(vmac,i,j+1,k,0) - AT(vmac,i,j,k,0)) * dxinv[1];
        /* multifab_mult_mult_s(rh,-1) then multifab_plus_plus(rh,mac_rhs) */
        AT(rh,i,j,k,0) = d * (-ONE) + AT(mrhs,i,j,k,0);
    }
}

/* beta_d(face) = 2/(rho_i + rho_{i-e_d}) on all faces of the box incl. its boundary, macproject.f90:361-401 */
void orc_mk_mac_coeffs(double *bx_, double *by_, double *bz_, int ng_b, const double *rho_, int ng_r,
                       const int *lo, const int *hi, int dim)
{
    V bx = v_box(bx_, lo, hi, ng_b, 0, 1, dim);
    V by = v_box(by_, lo, hi, ng_b, 1, 1, dim);
    V bz = v_box(dim == 3 ? bz_ : by_, lo, hi, ng_b, 2, 1, dim);
    V rho = v_box((double*)rho_, lo, hi, ng_r, -1, 1, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    for (int k = k0; k <= k1; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]+1; ++i)
        AT(bx,i,j,k,0) = TWO / (AT(rho,i,j,k,0) + AT(rho,i-1,j,k,0));
    for (int k = k0; k <= k1; ++k) for (int j = lo[1]; j <= hi[1]+1; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
        AT(by,i,j,k,0) = TWO / (AT(rho,i,j,k,0) + AT(rho,i,j-1,k,0));
    if (dim == 3)
        for (int k = k0; k <= k1+1; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
            AT(bz,i,j,k,0) = TWO / (AT(rho,i,j,k,0) + AT(rho,i,j,k-1,0));
}

/*
 * umac -= beta * grad(phi).  Interior faces exactly as macproject.f90:610-613.
 * Box-boundary faces: the reference takes them from fine_flx, which F_MG (absent) fills with the
 * stencil flux; here the same quantity is formed from phi's ghost cell (already filled by the
 * caller with neighbour / periodic data) or from the elliptic BC: Neumann => 0, Dirichlet
 * (stencil_order 2, wall value 0) => one-sided gradient (3 phi_0 - phi_1/3)/h.   [SURVEY Q7]
 * ell_bc[d][side] in {ELL_INT, ELL_PER, ELL_NEU, ELL_DIR}.
 */
void orc_mkumac(double *umac_, double *vmac_, double *wmac_, int ng_um, const double *phi_, int ng_p,
                const double *bx_, const double *by_, const double *bz_, int ng_b,
                const int *lo, const int *hi, int dim, const double *dx, const int *ell_bc)
{
    V um[3]; V be[3];
    double *up[3] = { umac_, vmac_, dim == 3 ? wmac_ : vmac_ };
    const double *bp[3] = { bx_, by_, dim == 3 ? bz_ : by_ };
    for (int d = 0; d < 3; ++d) { um[d] = v_box(up[d], lo, hi, ng_um, d, 1, dim); be[d] = v_box((double*)bp[d], lo, hi, ng_b, d, 1, dim); }
    V phi = v_box((double*)phi_, lo, hi, ng_p, -1, 1, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    for (int d = 0; d < dim; ++d) {
        int e[3] = { d == 0, d == 1, d == 2 };
        for (int k = k0; k <= k1 + e[2]; ++k)
        for (int j = lo[1]; j <= hi[1] + e[1]; ++j)
        for (int i = lo[0]; i <= hi[0] + e[0]; ++i) {
            int ix[3] = { i, j, k };
            double g;
            if (ix[d] == lo[d] && ell_bc[d*2+0] == ELL_NEU) continue;
            if (ix[d] == hi[d]+1 && ell_bc[d*2+1] == ELL_NEU) continue;
            if (ix[d] == lo[d] && ell_bc[d*2+0] == ELL_DIR)
                g = (3.0*AT(phi,i,j,k,0) - AT(phi,i+e[0],j+e[1],k+e[2],0)/3.0) / dx[d];
            else if (ix[d] == hi[d]+1 && ell_bc[d*2+1] == ELL_DIR)
                g = -(3.0*AT(phi,i-e[0],j-e[1],k-e[2],0) - AT(phi,i-2*e[0],j-2*e[1],k-2*e[2],0)/3.0) / dx[d];
            else
                g = (AT(phi,i,j,k,0) - AT(phi,i-e[0],j-e[1],k-e[2],0)) / dx[d];
            AT(um[d],i,j,k,0) = AT(um[d],i,j,k,0) - AT(be[d],i,j,k,0)*g;
        }
    }
}

/* valid cells of vel_force, mkforce.f90:160-185 (3-D) / :98-118 (2-D).  s comp 0 = rho, comp 1 = tracer. */
void orc_mkvelforce(double *vf_, const double *ext_, const double *gp_, const double *s_, const double *lapu_,
                    int ng_f, int ng_e, int ng_g, int ng_s, int ng_l, int ncomp_s,
                    double visc_fac, double visc_coef, int boussinesq, const int *lo, const int *hi, int dim)
{
    V vf = v_box(vf_, lo, hi, ng_f, -1, dim, dim);
    V ext = v_box((double*)ext_, lo, hi, ng_e, -1, dim, dim);
    V gp = v_box((double*)gp_, lo, hi, ng_g, -1, dim, dim);
    V s = v_box((double*)s_, lo, hi, ng_s, -1, ncomp_s, dim);
    V lapu = v_box((double*)lapu_, lo, hi, ng_l, -1, dim, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    #pragma omp parallel for
    for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
    for (int i = lo[0]; i <= hi[0]; ++i)
        for (int c = 0; c < dim; ++c) {
            /* rhohalf (dm comps in the reference, only comp 1 filled, advance_timestep.f90:70,114) reads as 0 in comp 2 */
            double tr = ncomp_s > 1 ? AT(s,i,j,k,1) : 0.0;
            double f = boussinesq == 1 ? tr * AT(ext,i,j,k,c) : AT(ext,i,j,k,c);
            double ll = visc_coef * visc_fac * AT(lapu,i,j,k,c);
            AT(vf,i,j,k,c) = f + (ll - AT(gp,i,j,k,c)) / AT(s,i,j,k,0);
        }
}

/* valid cells of scal_force, mkforce.f90:349-356: comp 0 (density) = 0, comps >=1 = ext + diff_coef*diff_fac*laps */
void orc_mkscalforce(double *sf_, const double *ext_, const double *laps_, int ng_f, int ng_e, int ng_l,
                     int nscal, double diff_fac, double diff_coef, const int *lo, const int *hi, int dim)
{
    V sf = v_box(sf_, lo, hi, ng_f, -1, nscal, dim);
    V ext = v_box((double*)ext_, lo, hi, ng_e, -1, nscal, dim);
    V laps = v_box((double*)laps_, lo, hi, ng_l, -1, nscal, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
    for (int i = lo[0]; i <= hi[0]; ++i) {
        AT(sf,i,j,k,0) = 0.0;
        for (int c = 1; c < nscal; ++c) {
            double ll = diff_coef * diff_fac * AT(laps,i,j,k,c);
            AT(sf,i,j,k,c) = AT(ext,i,j,k,c) + ll;
        }
    }
}

/* rhohalf = 0.5*(rhoold + rhonew) on valid cells, make_at_halftime.f90:107-113 */
void orc_make_at_halftime(double *rh_, const double *ro_, const double *rn_, const int *lo, const int *hi, int dim,
                          int ng_h, int ng_o)
{
    V rh = v_box(rh_, lo, hi, ng_h, -1, 1, dim);
    V ro = v_box((double*)ro_, lo, hi, ng_o, -1, 1, dim);
    V rn = v_box((double*)rn_, lo, hi, ng_o, -1, 1, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
    for (int i = lo[0]; i <= hi[0]; ++i)
        AT(rh,i,j,k,0) = HALF * (AT(ro,i,j,k,0) + AT(rn,i,j,k,0));
}

/* estdt_3d / estdt_2d  (src/estdt.f90:131-181, :89-129): per-box time-step estimate.  dt (in/out) is lowered by the advective limit
 * dx/max|u_d| and the forcing limit sqrt(2 dx / max|gp_d/rho - f_d|) of every direction whose maximum exceeds eps = 1.0e-8 (a
 * single-precision literal in the reference).  s is the first scalar (density) only. */
void orc_estdt(const double *vel_, int ng_u, const double *s_, int ng_s, const double *gp_, int ng_g, const double *f_, int ng_f,
               const int *lo, const int *hi, int dim, const double *dx, double *dt)
{
    V vel = v_box((double*)vel_, lo, hi, ng_u, -1, dim, dim);
    V s = v_box((double*)s_, lo, hi, ng_s, -1, 1, dim);
    V gp = v_box((double*)gp_, lo, hi, ng_g, -1, dim, dim);
    V f = v_box((double*)f_, lo, hi, ng_f, -1, dim, dim);
    const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
    const double eps = (double)1.0e-8f;
    double um[3] = { 0.0, 0.0, 0.0 }, fm[3] = { 0.0, 0.0, 0.0 };
    for (int k = k0; k <= k1; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
    for (int i = lo[0]; i <= hi[0]; ++i)
        for (int d = 0; d < dim; ++d) {
            um[d] = dmax(um[d], fabs(AT(vel, i, j, k, d)));
            fm[d] = dmax(fm[d], fabs(AT(gp, i, j, k, d) / AT(s, i, j, k, 0) - AT(f, i, j, k, d)));
        }
    /* the reference applies all velocity limits first, then the forcing limits (min is order-independent) */
    for (int d = 0; d < dim; ++d) if (um[d] > eps) *dt = dmin(*dt, dx[d] / um[d]);
    for (int d = 0; d < dim; ++d) if (fm[d] > eps) *dt = dmin(*dt, sqrt(2.0 * dx[d] / fm[d]));
}

/* estdt (src/estdt.f90:15-87): min over the boxes (and, in the reference, over the MPI ranks), the fallback min(dx) when nothing limits,
 * the CFL factor and the growth limit.  u / s / gp / f: one array per box. */
double orc_estdt_mf(const double *const *u, int ng_u, const double *const *s, int ng_s, const double *const *gp, int ng_g,
                    const double *const *f, int ng_f, const int *boxes_lo, const int *boxes_hi, int nboxes, int dim, const double *dx,
                    double dtold, double cflfac, double max_dt_growth)
{
    const double dt_start = 1.e20;
    double dt = 1.e20;
    for (int b = 0; b < nboxes; ++b) {
        double dt_grid = 1.e20;
        orc_estdt(u[b], ng_u, s[b], ng_s, gp[b], ng_g, f[b], ng_f, boxes_lo + 3 * b, boxes_hi + 3 * b, dim, dx, &dt_grid);
        dt = dmin(dt_grid, dt);
    }
    if (dt == dt_start) {
        dt = dmin(dx[0], dx[1]);
        if (dim == 3) dt = dmin(dt, dx[2]);
    }
    dt = dt * cflfac;
    if (dtold > 0.0) dt = dmin(dt, max_dt_growth * dtold);
    return dt;
}
