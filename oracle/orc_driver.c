/*
 * oracle/orc_driver.c -- TEST INFRASTRUCTURE ONLY.
 * Multi-box orchestration of the oracle kernels, mirroring the reference call
 * sequence for nlevs == 1:
 *   advance_timestep.f90:66-124, advance_premac.f90:44-51, macproject.f90:20-133,
 *   scalar_advance.f90:54-119, velocity_advance.f90:51-93, make_at_halftime.f90:38-65,
 *   and the FBoxLib pieces those call: multifab_fill_boundary and
 *   ml_restrict_and_fill (nlevs==1: fill_boundary + multifab_physbc).
 * A "multifab" here is an array of per-box pointers, each box a Fortran-ordered
 * array with ng ghost cells (face arrays have hi+1 in their direction).
 */
#include "orc_common.h"

typedef struct {
    int dim, nscal, slope_order, use_minion, boussinesq, stencil_order;
    double visc_coef, diff_coef;
    double bcval[5][3][2];      /* u_bc, v_bc, w_bc, rho_bc, trac_bc  [d][side] */
    double mg_rel_eps;          /* macproject.f90:92 -> 1e-10 */
    double mg_bottom_eps;       /* mac_multigrid.f90:56 -> 1e-3 */
    int mg_max_cycles, mg_nu1, mg_nu2, mg_verbose;
} orc_params;

typedef struct {
    int nboxes;
    const int *blo, *bhi;       /* [nboxes][3] */
    int dlo[3], dhi[3];         /* domain */
    int phys_bc[3][2];          /* domain physical BCs (PERIODIC = -1) */
    double dx[3];
} orc_geom;

/* ---- prototypes of the per-box kernels ---- */
void orc_velpred_3d(const double*, double*, double*, double*, const double*, const int*, const int*, const double*, double,
                    const int*, const int*, int, int, int, int, int);
void orc_velpred_2d(const double*, double*, double*, const double*, const int*, const int*, const double*, double,
                    const int*, const int*, int, int, int, int, int);
void orc_mkflux_3d(const double*, double*, double*, double*, double*, double*, double*, const double*, const double*, const double*,
                   const double*, const double*, const int*, const int*, const double*, double, int, const int*, const int*,
                   int, int, int, int, int, int, const int*, int, int, int);
void orc_mkflux_2d(const double*, double*, double*, double*, double*, const double*, const double*, const double*, const double*,
                   const int*, const int*, const double*, double, int, const int*, const int*,
                   int, int, int, int, int, int, const int*, int, int, int);
void orc_update(const double*, const double*, const double*, const double*, const double*, const double*, const double*,
                const double*, const double*, const double*, const double*, double*, const int*, const int*, int,
                int, int, int, int, int, const double*, double, int, const int*, int);
void orc_physbc(double*, const int*, const int*, int, int, const int*, int, const double*);
void orc_divumac(const double*, const double*, const double*, int, const double*, int, double*, int, const double*, const int*, const int*, int);
void orc_mk_mac_coeffs(double*, double*, double*, int, const double*, int, const int*, const int*, int);
void orc_mkumac(double*, double*, double*, int, const double*, int, const double*, const double*, const double*, int,
                const int*, const int*, int, const double*, const int*);
void orc_mkvelforce(double*, const double*, const double*, const double*, const double*, int, int, int, int, int, int,
                    double, double, int, const int*, const int*, int);
void orc_mkscalforce(double*, const double*, const double*, int, int, int, int, double, double, const int*, const int*, int);
void orc_make_at_halftime(double*, const double*, const double*, const int*, const int*, int, int, int);
int orc_mg_solve(int, const int*, const double*, const int*, const double*, const double*, const double*, const double*,
                 double*, double, int, int, int, double, int, double*);

/* ---- BC tables (define_bc_tower.f90:129-340) ---- */
/* per-box phys_bc: the domain BC where the box touches the domain boundary, INTERIOR elsewhere (:140-154) */
void orc_box_phys_bc(const orc_geom *g, int ib, int *pb /* [3][2] */)
{
    for (int d = 0; d < 3; ++d) {
        pb[d*2+0] = (g->blo[ib*3+d] == g->dlo[d]) ? g->phys_bc[d][0] : BC_INTERIOR;
        pb[d*2+1] = (g->bhi[ib*3+d] == g->dhi[d]) ? g->phys_bc[d][1] : BC_INTERIOR;
    }
}
/* adv_bc[comp][d][side], comps: 0..dm-1 vel, dm..dm+nscal-1 scalars, dm+nscal press, dm+nscal+1 extrap (:158-252) */
void orc_adv_bc(const int *pb, int dm, int nscal, int *adv /* [dm+nscal+2][3][2] */)
{
    const int ncomp = dm + nscal + 2, press = dm + nscal, extrap = press + 1;
    for (int q = 0; q < ncomp*6; ++q) adv[q] = BC_INTERIOR;
#define ADV(c,d,s) adv[((c)*3+(d))*2+(s)]
    for (int d = 0; d < dm; ++d) for (int s = 0; s < 2; ++s) {
        const int p = pb[d*2+s];
        if (p == BC_SLIP_WALL) {
            for (int c = 0; c < dm; ++c) ADV(c,d,s) = BC_HOEXTRAP;
            ADV(d,d,s) = BC_EXT_DIR;
            for (int ns = 0; ns < nscal; ++ns) ADV(dm+ns,d,s) = BC_HOEXTRAP;
            ADV(press,d,s) = BC_FOEXTRAP; ADV(extrap,d,s) = BC_FOEXTRAP;
        } else if (p == BC_NO_SLIP_WALL) {
            for (int c = 0; c < dm; ++c) ADV(c,d,s) = BC_EXT_DIR;
            for (int ns = 0; ns < nscal; ++ns) ADV(dm+ns,d,s) = BC_HOEXTRAP;
            ADV(press,d,s) = BC_FOEXTRAP; ADV(extrap,d,s) = BC_FOEXTRAP;
        } else if (p == BC_INLET) {
            for (int c = 0; c < dm; ++c) ADV(c,d,s) = BC_EXT_DIR;
            for (int ns = 0; ns < nscal; ++ns) ADV(dm+ns,d,s) = BC_EXT_DIR;
            ADV(press,d,s) = BC_FOEXTRAP; ADV(extrap,d,s) = BC_FOEXTRAP;
        } else if (p == BC_OUTLET) {
            for (int c = 0; c < dm; ++c) ADV(c,d,s) = BC_FOEXTRAP;
            for (int ns = 0; ns < nscal; ++ns) ADV(dm+ns,d,s) = BC_FOEXTRAP;
            ADV(press,d,s) = BC_EXT_DIR; ADV(extrap,d,s) = BC_FOEXTRAP;
        } else if (p == BC_SYMMETRY) {
            for (int c = 0; c < dm; ++c) ADV(c,d,s) = BC_REFLECT_EVEN;
            ADV(d,d,s) = BC_REFLECT_ODD;
            for (int ns = 0; ns < nscal; ++ns) ADV(dm+ns,d,s) = BC_REFLECT_EVEN;
            ADV(press,d,s) = BC_EXT_DIR; ADV(extrap,d,s) = BC_REFLECT_EVEN;
        }
    }
#undef ADV
}
/* elliptic BC of the pressure component (:291-334) */
void orc_ell_bc_press(const int *pb, int dm, int *ell /* [3][2] */)
{
    for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) {
        const int p = pb[d*2+s];
        int e = ELL_INT;
        if (d < dm) {
            if (p == BC_SLIP_WALL || p == BC_NO_SLIP_WALL || p == BC_INLET || p == BC_SYMMETRY) e = ELL_NEU;
            else if (p == BC_OUTLET) e = ELL_DIR;
            else if (p == BC_PERIODIC) e = ELL_PER;
        }
        ell[d*2+s] = e;
    }
}

/* ---- multifab_fill_boundary: copy valid -> ghost between boxes incl. periodic images ---- */
void orc_fill_boundary(const orc_geom *g, int dim, double **mf, int ng, int ncomp, int face_dir)
{
    if (ng == 0) return;
    int per[3], dlen[3];
    for (int d = 0; d < 3; ++d) { per[d] = (d < dim && g->phys_bc[d][0] == BC_PERIODIC); dlen[d] = g->dhi[d]-g->dlo[d]+1; }
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        V dst = v_box(mf[ib], lo, hi, ng, face_dir, ncomp, dim);
        int vlo[3], vhi[3];            /* valid index range of this box */
        for (int d = 0; d < 3; ++d) { vlo[d] = d < dim ? lo[d] : 0; vhi[d] = d < dim ? hi[d] + (d == face_dir) : 0; }
        for (int k = dst.l[2]; k < dst.l[2]+dst.n[2]; ++k)
        for (int j = dst.l[1]; j < dst.l[1]+dst.n[1]; ++j)
        for (int i = dst.l[0]; i < dst.l[0]+dst.n[0]; ++i) {
            int ix[3] = { i, j, k };
            int inside = 1;
            for (int d = 0; d < dim; ++d) if (ix[d] < vlo[d] || ix[d] > vhi[d]) inside = 0;
            if (inside) continue;
            /* search a source box (and periodic shift) whose valid region contains this point */
            int found = 0;
            for (int sx = -per[0]; sx <= per[0] && !found; ++sx)
            for (int sy = -per[1]; sy <= per[1] && !found; ++sy)
            for (int sz = -per[2]; sz <= per[2] && !found; ++sz) {
                int sh[3] = { sx*dlen[0], sy*dlen[1], sz*dlen[2] };
                int p[3] = { ix[0]+sh[0], ix[1]+sh[1], ix[2]+sh[2] };
                for (int jb = 0; jb < g->nboxes && !found; ++jb) {
                    const int *slo = &g->blo[jb*3], *shi = &g->bhi[jb*3];
                    int in = 1;
                    for (int d = 0; d < dim; ++d) if (p[d] < slo[d] || p[d] > shi[d] + (d == face_dir)) in = 0;
                    if (!in) continue;
                    V src = v_box(mf[jb], slo, shi, ng, face_dir, ncomp, dim);
                    for (int c = 0; c < ncomp; ++c) AT(dst,i,j,k,c) = AT(src,p[0],p[1],p[2],c);
                    found = 1;
                }
            }
        }
    }
}

/* ml_restrict_and_fill for nlevs == 1: fill_boundary then multifab_physbc; same_boundary => every comp uses bccomp */
void orc_fill_and_physbc(const orc_geom *g, const orc_params *P, double **mf, int ng, int ncomp_total,
                         int scomp, int bccomp /* 0-based into adv_bc */, int nc, int same_boundary)
{
    const int dim = P->dim;
    orc_fill_boundary(g, dim, mf, ng, ncomp_total, -1);
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        int pb[6]; orc_box_phys_bc(g, ib, pb);
        int adv[(3+8+2)*6]; orc_adv_bc(pb, dim, P->nscal, adv);
        V v = v_box(mf[ib], lo, hi, ng, -1, ncomp_total, dim);
        for (int c = 0; c < nc; ++c) {
            int bcc = same_boundary ? bccomp : bccomp + c;
            orc_physbc(v.p + v.cs*(scomp+c), lo, hi, dim, ng, &adv[bcc*6], bcc+1, &P->bcval[0][0][0]);
        }
    }
}

static long box_size(const int *lo, const int *hi, int ng, int face_dir, int dim)
{
    long n = 1;
    for (int d = 0; d < dim; ++d) n *= (hi[d]-lo[d]+1 + 2*ng + (d == face_dir));
    return n;
}
static double **mf_alloc(const orc_geom *g, int dim, int ng, int ncomp, int face_dir, double val)
{
    double **mf = (double**)malloc(sizeof(double*)*g->nboxes);
    for (int ib = 0; ib < g->nboxes; ++ib) {
        long n = box_size(&g->blo[ib*3], &g->bhi[ib*3], ng, face_dir, dim)*ncomp;
        mf[ib] = (double*)malloc(sizeof(double)*n);
        for (long q = 0; q < n; ++q) mf[ib][q] = val;
    }
    return mf;
}
static void mf_free(const orc_geom *g, double **mf) { for (int ib = 0; ib < g->nboxes; ++ib) free(mf[ib]); free(mf); }

/* mkvelforce + ghost fill with the extrap component (mkforce.f90:18-80) */
void orc_mkvelforce_mf(const orc_geom *g, const orc_params *P, double **vel_force, double **ext, double **gp,
                       double **s, int ng_s, int ncomp_s, double **lapu, double visc_fac)
{
    const int dim = P->dim;
    for (int ib = 0; ib < g->nboxes; ++ib) {
        long n = box_size(&g->blo[ib*3], &g->bhi[ib*3], 1, -1, dim)*dim;
        for (long q = 0; q < n; ++q) vel_force[ib][q] = 0.0;
        orc_mkvelforce(vel_force[ib], ext[ib], gp[ib], s[ib], lapu[ib], 1, 1, 1, ng_s, 0, ncomp_s,
                       visc_fac, P->visc_coef, P->boussinesq, &g->blo[ib*3], &g->bhi[ib*3], dim);
    }
    orc_fill_and_physbc(g, P, vel_force, 1, dim, 0, dim + P->nscal + 1, dim, 1);
}
void orc_mkscalforce_mf(const orc_geom *g, const orc_params *P, double **scal_force, double **ext, double **laps, double diff_fac)
{
    const int dim = P->dim;
    for (int ib = 0; ib < g->nboxes; ++ib) {
        long n = box_size(&g->blo[ib*3], &g->bhi[ib*3], 1, -1, dim)*P->nscal;
        for (long q = 0; q < n; ++q) scal_force[ib][q] = 0.0;
        orc_mkscalforce(scal_force[ib], ext[ib], laps[ib], 1, 1, 0, P->nscal, diff_fac, P->diff_coef, &g->blo[ib*3], &g->bhi[ib*3], dim);
    }
    orc_fill_and_physbc(g, P, scal_force, 1, P->nscal, 0, dim + P->nscal + 1, P->nscal, 1);
}

/* velpred driver (velpred.f90:16-123): per box kernel, then fill_boundary(umac(d)) */
void orc_velpred_mf(const orc_geom *g, const orc_params *P, double **u, double **umac[3], double **force, double dt)
{
    const int dim = P->dim;
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        int pb[6]; orc_box_phys_bc(g, ib, pb);
        int adv[(3+8+2)*6]; orc_adv_bc(pb, dim, P->nscal, adv);
        if (dim == 3) orc_velpred_3d(u[ib], umac[0][ib], umac[1][ib], umac[2][ib], force[ib], lo, hi, g->dx, dt, pb, adv, 3, 1, 1, P->use_minion, P->slope_order);
        else          orc_velpred_2d(u[ib], umac[0][ib], umac[1][ib], force[ib], lo, hi, g->dx, dt, pb, adv, 3, 1, 1, P->use_minion, P->slope_order);
    }
    for (int d = 0; d < dim; ++d) orc_fill_boundary(g, dim, umac[d], 1, 1, d);
}

/* mkflux driver (mkflux.f90:16-150) */
void orc_mkflux_mf(const orc_geom *g, const orc_params *P, double **sold, int ncomp, double **sedge[3], double **flux[3],
                   double **umac[3], double **force, double **mac_rhs, double dt, int is_vel, const int *is_cons)
{
    const int dim = P->dim;
    const int bccomp = is_vel ? 0 : dim;
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        int pb[6]; orc_box_phys_bc(g, ib, pb);
        int adv[(3+8+2)*6]; orc_adv_bc(pb, dim, P->nscal, adv);
        if (dim == 3)
            orc_mkflux_3d(sold[ib], sedge[0][ib], sedge[1][ib], sedge[2][ib], flux[0][ib], flux[1][ib], flux[2][ib],
                          umac[0][ib], umac[1][ib], umac[2][ib], force[ib], mac_rhs[ib], lo, hi, g->dx, dt, is_vel,
                          pb, &adv[bccomp*6], 3, 0, 0, 1, 1, 1, is_cons, ncomp, P->use_minion, P->slope_order);
        else
            orc_mkflux_2d(sold[ib], sedge[0][ib], sedge[1][ib], flux[0][ib], flux[1][ib],
                          umac[0][ib], umac[1][ib], force[ib], mac_rhs[ib], lo, hi, g->dx, dt, is_vel,
                          pb, &adv[bccomp*6], 3, 0, 0, 1, 1, 1, is_cons, ncomp, P->use_minion, P->slope_order);
    }
}

/* update driver (update.f90:16-111) */
void orc_update_mf(const orc_geom *g, const orc_params *P, double **sold, int ncomp, double **umac[3], double **sedge[3], double **flux[3],
                   double **force, double **snew, double dt, int is_vel, const int *is_cons)
{
    const int dim = P->dim;
    for (int ib = 0; ib < g->nboxes; ++ib)
        orc_update(sold[ib], umac[0][ib], umac[1][ib], dim == 3 ? umac[2][ib] : NULL, sedge[0][ib], sedge[1][ib], dim == 3 ? sedge[2][ib] : NULL,
                   flux[0][ib], flux[1][ib], dim == 3 ? flux[2][ib] : NULL, force[ib], snew[ib], &g->blo[ib*3], &g->bhi[ib*3], dim,
                   3, 1, 0, 0, 1, g->dx, dt, is_vel, is_cons, ncomp);
    orc_fill_and_physbc(g, P, snew, 3, ncomp, 0, is_vel ? 0 : dim, ncomp, 0);
}

/*
 * macproject (macproject.f90:20-133) for nlevs == 1.  phi_out (optional) receives per-box phi with 1 ghost.
 * rel_eps <= 0 selects the reference's 1e-10.  Returns V-cycle count.
 */
int orc_macproject_mf(const orc_geom *g, const orc_params *P, double **umac[3], double **rho /* s, ng 3, comp 0 */, int ncomp_s,
                      double **mac_rhs, double **phi_out, double rel_eps, double *resnorm)
{
    const int dim = P->dim;
    (void)ncomp_s;
    double **rh = mf_alloc(g, dim, 0, 1, -1, 0.0);
    double **phi = mf_alloc(g, dim, 1, 1, -1, 0.0);
    double **beta[3] = { mf_alloc(g, dim, 0, 1, 0, 0.0), mf_alloc(g, dim, 0, 1, 1, 0.0), dim == 3 ? mf_alloc(g, dim, 0, 1, 2, 0.0) : NULL };
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        orc_divumac(umac[0][ib], umac[1][ib], dim == 3 ? umac[2][ib] : NULL, 1, mac_rhs[ib], 1, rh[ib], 0, g->dx, lo, hi, dim);
        orc_mk_mac_coeffs(beta[0][ib], beta[1][ib], dim == 3 ? beta[2][ib] : NULL, 0, rho[ib], 3, lo, hi, dim);
    }
    /* gather to merged-domain arrays */
    int n[3] = { g->dhi[0]-g->dlo[0]+1, g->dhi[1]-g->dlo[1]+1, dim == 3 ? g->dhi[2]-g->dlo[2]+1 : 1 };
    int zlo[3] = { 0, 0, 0 }, zhi[3] = { n[0]-1, n[1]-1, n[2]-1 };
    double *RH = (double*)calloc((long)n[0]*n[1]*n[2], 8);
    double *B[3] = { NULL, NULL, NULL };
    V vRH = v_box(RH, zlo, zhi, 0, -1, 1, dim);
    V vB[3];
    for (int d = 0; d < dim; ++d) {
        B[d] = (double*)calloc(box_size(zlo, zhi, 0, d, dim), 8);
        vB[d] = v_box(B[d], zlo, zhi, 0, d, 1, dim);
    }
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        V r = v_box(rh[ib], lo, hi, 0, -1, 1, dim);
        const int k0 = dim == 3 ? lo[2] : 0, k1 = dim == 3 ? hi[2] : 0;
        for (int k = k0; k <= k1; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
            AT(vRH, i-g->dlo[0], j-g->dlo[1], dim == 3 ? k-g->dlo[2] : 0, 0) = AT(r,i,j,k,0);
        for (int d = 0; d < dim; ++d) {
            V b = v_box(beta[d][ib], lo, hi, 0, d, 1, dim);
            int e[3] = { d == 0, d == 1, d == 2 };
            for (int k = k0; k <= k1 + e[2]; ++k) for (int j = lo[1]; j <= hi[1]+e[1]; ++j) for (int i = lo[0]; i <= hi[0]+e[0]; ++i)
                AT(vB[d], i-g->dlo[0], j-g->dlo[1], dim == 3 ? k-g->dlo[2] : 0, 0) = AT(b,i,j,k,0);
        }
    }
    int dpb[6]; for (int d = 0; d < 3; ++d) { dpb[d*2] = g->phys_bc[d][0]; dpb[d*2+1] = g->phys_bc[d][1]; }
    int dell[6]; orc_ell_bc_press(dpb, dim, dell);
    long npad = (long)(n[0]+2)*(n[1]+2)*(dim == 3 ? n[2]+2 : 1);
    double *PHI = (double*)calloc(npad, 8);
    int cycles = orc_mg_solve(dim, n, g->dx, dell, RH, B[0], B[1], B[2], PHI,
                              rel_eps > 0 ? rel_eps : P->mg_rel_eps, P->mg_max_cycles, P->mg_nu1, P->mg_nu2,
                              P->mg_bottom_eps, P->mg_verbose, resnorm);
    /* scatter phi (with its ghost ring) back to the boxes */
    int plo[3] = { 0, 0, 0 };
    V vP = v_box(PHI, plo, zhi, 1, -1, 1, dim);
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        V p = v_box(phi[ib], lo, hi, 1, -1, 1, dim);
        const int k0 = dim == 3 ? lo[2]-1 : 0, k1 = dim == 3 ? hi[2]+1 : 0;
        for (int k = k0; k <= k1; ++k) for (int j = lo[1]-1; j <= hi[1]+1; ++j) for (int i = lo[0]-1; i <= hi[0]+1; ++i)
            AT(p,i,j,k,0) = AT(vP, i-g->dlo[0], j-g->dlo[1], dim == 3 ? k-g->dlo[2] : 0, 0);
    }
    for (int ib = 0; ib < g->nboxes; ++ib) {
        const int *lo = &g->blo[ib*3], *hi = &g->bhi[ib*3];
        int pb[6]; orc_box_phys_bc(g, ib, pb);
        int ell[6]; orc_ell_bc_press(pb, dim, ell);
        orc_mkumac(umac[0][ib], umac[1][ib], dim == 3 ? umac[2][ib] : NULL, 1, phi[ib], 1,
                   beta[0][ib], beta[1][ib], dim == 3 ? beta[2][ib] : NULL, 0, lo, hi, dim, g->dx, ell);
    }
    for (int d = 0; d < dim; ++d) orc_fill_boundary(g, dim, umac[d], 1, 1, d);
    if (phi_out)
        for (int ib = 0; ib < g->nboxes; ++ib)
            memcpy(phi_out[ib], phi[ib], 8*box_size(&g->blo[ib*3], &g->bhi[ib*3], 1, -1, dim));
    free(RH); free(PHI); for (int d = 0; d < dim; ++d) free(B[d]);
    mf_free(g, rh); mf_free(g, phi); for (int d = 0; d < dim; ++d) mf_free(g, beta[d]);
    return cycles;
}

/*
 * One pass of the hot path: advance_timestep.f90:66-124 (everything between the entry state and hgproject).
 * Inputs (per box): uold(ng3,dm) sold(ng3,nscal) gp(ng1,dm) ext_vel_force(ng1,dm) ext_scal_force(ng1,nscal) lapu(ng0,dm).
 * Outputs: unew(ng3,dm) snew(ng3,nscal) rhohalf(ng1, 1 comp) umac[d](ng1) (projected); optional phi(ng1).
 */
int orc_advance_mf(const orc_geom *g, const orc_params *P, double **uold, double **sold, double **gp,
                   double **ext_vel_force, double **ext_scal_force, double **lapu,
                   double **unew, double **snew, double **rhohalf, double **umac_x, double **umac_y, double **umac_z,
                   double **phi_out, double dt, double mac_rel_eps, double *mac_resnorm)
{
    const int dim = P->dim, nscal = P->nscal;
    double **umac[3] = { umac_x, umac_y, umac_z };
    double **mac_rhs = mf_alloc(g, dim, 1, 1, -1, 0.0);
    double **vel_force = mf_alloc(g, dim, 1, dim, -1, 0.0);
    double **scal_force = mf_alloc(g, dim, 1, nscal, -1, 0.0);
    double **laps = mf_alloc(g, dim, 0, nscal, -1, 0.0);
    double **divu = mf_alloc(g, dim, 1, 1, -1, 0.0);
    double **sedge[3], **sflux[3], **uedge[3], **uflux[3];
    for (int d = 0; d < 3; ++d) {
        sedge[d] = d < dim ? mf_alloc(g, dim, 0, nscal, d, 0.0) : NULL; sflux[d] = d < dim ? mf_alloc(g, dim, 0, nscal, d, 0.0) : NULL;
        uedge[d] = d < dim ? mf_alloc(g, dim, 0, dim, d, 0.0) : NULL;   uflux[d] = d < dim ? mf_alloc(g, dim, 0, dim, d, 0.0) : NULL;
    }
    /* umac = 1.d20 (advance_timestep.f90:76-77) */
    for (int d = 0; d < dim; ++d) for (int ib = 0; ib < g->nboxes; ++ib) {
        long n = box_size(&g->blo[ib*3], &g->bhi[ib*3], 1, d, dim);
        for (long q = 0; q < n; ++q) umac[d][ib][q] = 1.0e20;
    }
    /* advance_premac */
    orc_mkvelforce_mf(g, P, vel_force, ext_vel_force, gp, sold, 3, nscal, lapu, 1.0);
    orc_velpred_mf(g, P, uold, umac, vel_force, dt);
    /* macproject */
    int cycles = orc_macproject_mf(g, P, umac, sold, nscal, mac_rhs, phi_out, mac_rel_eps, mac_resnorm);
    /* scalar_advance */
    int is_cons_s[8]; is_cons_s[0] = 1; for (int c = 1; c < 8; ++c) is_cons_s[c] = 0;
    orc_mkscalforce_mf(g, P, scal_force, ext_scal_force, laps, 1.0);
    orc_mkflux_mf(g, P, sold, nscal, sedge, sflux, umac, scal_force, divu, dt, 0, is_cons_s);
    orc_mkscalforce_mf(g, P, scal_force, ext_scal_force, laps, 0.0);
    orc_update_mf(g, P, sold, nscal, umac, sedge, sflux, scal_force, snew, dt, 0, is_cons_s);
    /* make_at_halftime(rhohalf, sold, snew, 1, 1) + fill with the density BC (bcomp = dm+in_comp) */
    for (int ib = 0; ib < g->nboxes; ++ib)
        orc_make_at_halftime(rhohalf[ib], sold[ib], snew[ib], &g->blo[ib*3], &g->bhi[ib*3], dim, 1, 3);
    orc_fill_and_physbc(g, P, rhohalf, 1, 1, 0, dim, 1, 0);
    /* velocity_advance */
    int is_cons_v[3] = { 0, 0, 0 };
    orc_mkvelforce_mf(g, P, vel_force, ext_vel_force, gp, sold, 3, nscal, lapu, 1.0);
    orc_mkflux_mf(g, P, uold, dim, uedge, uflux, umac, vel_force, mac_rhs, dt, 1, is_cons_v);
    orc_mkvelforce_mf(g, P, vel_force, ext_vel_force, gp, rhohalf, 1, 1, lapu, 0.0);
    orc_update_mf(g, P, uold, dim, umac, uedge, uflux, vel_force, unew, dt, 1, is_cons_v);

    mf_free(g, mac_rhs); mf_free(g, vel_force); mf_free(g, scal_force); mf_free(g, laps); mf_free(g, divu);
    for (int d = 0; d < dim; ++d) { mf_free(g, sedge[d]); mf_free(g, sflux[d]); mf_free(g, uedge[d]); mf_free(g, uflux[d]); }
    return cycles;
}

/* thread control for the timing runs (bench.py): torchrun exports OMP_NUM_THREADS=1, which libgomp reads at load time */
#ifdef _OPENMP
#include <omp.h>
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_get_threads(void) { return omp_get_max_threads(); }
#else
void orc_set_threads(int n) { (void)n; }
int orc_get_threads(void) { return 1; }
#endif
