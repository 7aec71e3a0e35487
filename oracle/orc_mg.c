/*
 * oracle/orc_mg.c -- TEST INFRASTRUCTURE ONLY.  ** parity unpinned **
 *
 * Cell-centred variable-coefficient multigrid for  -div(beta grad phi) = rh,
 * the solve VARDEN reaches through mac_multigrid.f90:53-62 -> ml_cc_solve.
 * ml_cc_solve lives in FBoxLib / AMReX F_MG (third party, NOT in /root/reference,
 * version unpinned: exec/test/GNUmakefile:12), so this file restates the published
 * algorithm (V-cycle, red-black Gauss-Seidel nu1=nu2=2, cell-average restriction,
 * piecewise-constant prolongation, BiCGStab bottom solve with eps 1e-3, stencil_order 2
 * Dirichlet ghost, stop when |r|_inf <= eps*|rh|_inf) rather than the F_MG source.
 * It is validated in tests/ against a direct sparse solve of the same discrete operator.
 *
 * Layout: one array per level covering the whole (merged) domain, padded by one
 * cell on every side: index (i+1) + (nx+2)*((j+1) + (ny+2)*(k+1)); 2-D uses nz=1
 * and no z faces.  beta_d[idx] is the coefficient on the LOW face of cell idx in d.
 */
#include "orc_common.h"

typedef struct {
    int n[3];            /* cells */
    long s[3];           /* strides */
    long ntot;
    double h[3];
    double *phi, *rhs, *res, *b[3];
    double *alpha;       /* Helmholtz solves (alpha - div beta grad): cell coefficient, NULL for the MAC projection */
} mglev;

typedef struct {
    int dim, nlev;
    int bc[3][2];        /* ELL_PER / ELL_NEU / ELL_DIR on the domain faces */
    int singular;
    mglev *L;
} mgtower;

#define IDX(L,i,j,k) ((long)((i)+1) + (L)->s[1]*(long)((j)+1) + (L)->s[2]*(long)((k)+1))

static void lev_alloc(mglev *L, int nx, int ny, int nz, const double *h, int dim)
{
    L->n[0] = nx; L->n[1] = ny; L->n[2] = nz;
    L->s[0] = 1; L->s[1] = nx+2; L->s[2] = (long)(nx+2)*(ny+2);
    L->ntot = L->s[2]*(nz+2);
    for (int d = 0; d < 3; ++d) L->h[d] = h[d];
    L->phi = (double*)calloc(L->ntot, sizeof(double));
    L->rhs = (double*)calloc(L->ntot, sizeof(double));
    L->res = (double*)calloc(L->ntot, sizeof(double));
    for (int d = 0; d < 3; ++d) L->b[d] = (d < dim) ? (double*)calloc(L->ntot, sizeof(double)) : NULL;
    L->alpha = NULL;
}
static void lev_free(mglev *L)
{
    free(L->phi); free(L->rhs); free(L->res); free(L->alpha);
    for (int d = 0; d < 3; ++d) free(L->b[d]);
}

/* fill periodic ghost cells of x (physical BCs are synthesised inside the operator) */
static void fill_periodic(const mgtower *T, const mglev *L, double *x)
{
    const int nx = L->n[0], ny = L->n[1], nz = L->n[2];
    if (T->bc[0][0] == ELL_PER)
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) {
            x[IDX(L,-1,j,k)] = x[IDX(L,nx-1,j,k)]; x[IDX(L,nx,j,k)] = x[IDX(L,0,j,k)];
        }
    if (T->bc[1][0] == ELL_PER)
        for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) {
            x[IDX(L,i,-1,k)] = x[IDX(L,i,ny-1,k)]; x[IDX(L,i,ny,k)] = x[IDX(L,i,0,k)];
        }
    if (T->dim == 3 && T->bc[2][0] == ELL_PER)
        for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            x[IDX(L,i,j,-1)] = x[IDX(L,i,j,nz-1)]; x[IDX(L,i,j,nz)] = x[IDX(L,i,j,0)];
        }
}

/* (A x)_c and diag_c for cell c=(i,j,k); ghost cells of x must hold periodic images */
static inline void op_cell(const mgtower *T, const mglev *L, const double *x, int i, int j, int k, double *Ax, double *diag)
{
    const long c = IDX(L,i,j,k);
    const int ix[3] = { i, j, k };
    double a = 0.0, dg = 0.0;
    if (L->alpha) { a = L->alpha[c]*x[c]; dg = L->alpha[c]; }
    for (int d = 0; d < T->dim; ++d) {
        const long st = L->s[d];
        const double h2 = 1.0/(L->h[d]*L->h[d]);
        const double blo = L->b[d][c], bhi = L->b[d][c+st];
        /* low face */
        if (ix[d] == 0 && T->bc[d][0] == ELL_NEU) { /* no flux */ }
        else if (ix[d] == 0 && T->bc[d][0] == ELL_DIR) { a += blo*(3.0*x[c] - x[c+st]/3.0)*h2; dg += 3.0*blo*h2; }
        else { a += blo*(x[c] - x[c-st])*h2; dg += blo*h2; }
        /* high face */
        if (ix[d] == L->n[d]-1 && T->bc[d][1] == ELL_NEU) { }
        else if (ix[d] == L->n[d]-1 && T->bc[d][1] == ELL_DIR) { a += bhi*(3.0*x[c] - x[c-st]/3.0)*h2; dg += 3.0*bhi*h2; }
        else { a += bhi*(x[c] - x[c+st])*h2; dg += bhi*h2; }
    }
    *Ax = a; *diag = dg;
}

static void gsrb(const mgtower *T, mglev *L, int sweeps)
{
    for (int sw = 0; sw < sweeps; ++sw)
        for (int color = 0; color < 2; ++color) {
            fill_periodic(T, L, L->phi);
            #pragma omp parallel for
            for (int k = 0; k < L->n[2]; ++k)
            for (int j = 0; j < L->n[1]; ++j)
            for (int i = (j + k + color) & 1; i < L->n[0]; i += 2) {
                double Ax, dg; op_cell(T, L, L->phi, i, j, k, &Ax, &dg);
                if (dg != 0.0) L->phi[IDX(L,i,j,k)] += (L->rhs[IDX(L,i,j,k)] - Ax)/dg;
            }
        }
}

static double residual(const mgtower *T, mglev *L)
{
    double nrm = 0.0;
    fill_periodic(T, L, L->phi);
    #pragma omp parallel for reduction(max:nrm)
    for (int k = 0; k < L->n[2]; ++k)
    for (int j = 0; j < L->n[1]; ++j)
    for (int i = 0; i < L->n[0]; ++i) {
        double Ax, dg; op_cell(T, L, L->phi, i, j, k, &Ax, &dg);
        double r = L->rhs[IDX(L,i,j,k)] - Ax;
        L->res[IDX(L,i,j,k)] = r;
        if (fabs(r) > nrm) nrm = fabs(r);
    }
    return nrm;
}

static void restrict_res(const mgtower *T, const mglev *F, mglev *C)
{
    const int rz = T->dim == 3 ? 2 : 1;
    const double w = 1.0/(4.0*rz);
    for (int k = 0; k < C->n[2]; ++k)
    for (int j = 0; j < C->n[1]; ++j)
    for (int i = 0; i < C->n[0]; ++i) {
        double s = 0.0;
        for (int kk = 0; kk < rz; ++kk) for (int jj = 0; jj < 2; ++jj) for (int ii = 0; ii < 2; ++ii)
            s += F->res[IDX(F, 2*i+ii, 2*j+jj, rz*k+kk)];
        C->rhs[IDX(C,i,j,k)] = s*w;
    }
}
static void prolong_add(const mgtower *T, mglev *F, const mglev *C)
{
    const int rz = T->dim == 3 ? 2 : 1;
    for (int k = 0; k < F->n[2]; ++k)
    for (int j = 0; j < F->n[1]; ++j)
    for (int i = 0; i < F->n[0]; ++i)
        F->phi[IDX(F,i,j,k)] += C->phi[IDX(C, i/2, j/2, k/rz)];
}

static double dotp(const mglev *L, const double *a, const double *b)
{
    double s = 0.0;
    for (int k = 0; k < L->n[2]; ++k) for (int j = 0; j < L->n[1]; ++j) for (int i = 0; i < L->n[0]; ++i)
        s += a[IDX(L,i,j,k)]*b[IDX(L,i,j,k)];
    return s;
}
static void apply(const mgtower *T, const mglev *L, double *x, double *y)
{
    fill_periodic(T, L, x);
    for (int k = 0; k < L->n[2]; ++k) for (int j = 0; j < L->n[1]; ++j) for (int i = 0; i < L->n[0]; ++i) {
        double Ax, dg; op_cell(T, L, x, i, j, k, &Ax, &dg); y[IDX(L,i,j,k)] = Ax;
    }
}
static void sub_mean(const mglev *L, double *x)
{
    double s = 0.0; long n = (long)L->n[0]*L->n[1]*L->n[2];
    for (int k = 0; k < L->n[2]; ++k) for (int j = 0; j < L->n[1]; ++j) for (int i = 0; i < L->n[0]; ++i) s += x[IDX(L,i,j,k)];
    s /= (double)n;
    for (int k = 0; k < L->n[2]; ++k) for (int j = 0; j < L->n[1]; ++j) for (int i = 0; i < L->n[0]; ++i) x[IDX(L,i,j,k)] -= s;
}

/* BiCGStab on the coarsest level, phi starts at 0; relative tolerance eps on |r|_2. */
static void bottom_bicgstab(const mgtower *T, mglev *L, double eps, int maxit)
{
    const long n = L->ntot;
    double *r = (double*)calloc(n, 8), *rh = (double*)calloc(n, 8), *p = (double*)calloc(n, 8),
           *v = (double*)calloc(n, 8), *s = (double*)calloc(n, 8), *t = (double*)calloc(n, 8);
    if (T->singular) sub_mean(L, L->rhs);
    memset(L->phi, 0, n*8);
    memcpy(r, L->rhs, n*8); memcpy(rh, r, n*8);
    double rho = 1, alpha = 1, omega = 1;
    const double bnorm = sqrt(dotp(L, r, r));
    if (bnorm > 0.0)
    for (int it = 0; it < maxit; ++it) {
        double rho1 = dotp(L, rh, r);
        if (rho1 == 0.0) break;
        if (it == 0) memcpy(p, r, n*8);
        else {
            double beta = (rho1/rho)*(alpha/omega);
            for (long q = 0; q < n; ++q) p[q] = r[q] + beta*(p[q] - omega*v[q]);
        }
        apply(T, L, p, v);
        double den = dotp(L, rh, v);
        if (den == 0.0) break;
        alpha = rho1/den;
        for (long q = 0; q < n; ++q) s[q] = r[q] - alpha*v[q];
        if (sqrt(dotp(L, s, s)) <= eps*bnorm) { for (long q = 0; q < n; ++q) L->phi[q] += alpha*p[q]; break; }
        apply(T, L, s, t);
        double tt = dotp(L, t, t);
        if (tt == 0.0) { for (long q = 0; q < n; ++q) L->phi[q] += alpha*p[q]; break; }
        omega = dotp(L, t, s)/tt;
        for (long q = 0; q < n; ++q) { L->phi[q] += alpha*p[q] + omega*s[q]; r[q] = s[q] - omega*t[q]; }
        rho = rho1;
        if (sqrt(dotp(L, r, r)) <= eps*bnorm) break;
        if (omega == 0.0) break;
    }
    if (T->singular) sub_mean(L, L->phi);
    free(r); free(rh); free(p); free(v); free(s); free(t);
}

static void vcycle(const mgtower *T, int l, int nu1, int nu2, double bottom_eps)
{
    mglev *L = &T->L[l];
    if (l == T->nlev-1) {
        long nc = (long)L->n[0]*L->n[1]*L->n[2];
        if (nc == 1) {   /* a single cell: exact */
            double Ax, dg; L->phi[IDX(L,0,0,0)] = 0.0; op_cell(T, L, L->phi, 0, 0, 0, &Ax, &dg);
            L->phi[IDX(L,0,0,0)] = (dg != 0.0 && !T->singular) ? L->rhs[IDX(L,0,0,0)]/dg : 0.0;
        } else bottom_bicgstab(T, L, bottom_eps, 200);
        return;
    }
    gsrb(T, L, nu1);
    residual(T, L);
    mglev *C = &T->L[l+1];
    restrict_res(T, L, C);
    memset(C->phi, 0, C->ntot*8);
    vcycle(T, l+1, nu1, nu2, bottom_eps);
    prolong_add(T, L, C);
    gsrb(T, L, nu2);
}

/*
 * Solve on the merged domain.  rh: nx*ny*nz (no ghosts, Fortran order); beta_d: face arrays
 * (nx+1)*ny*nz etc.; phi out: (nx+2)*(ny+2)*(nz+2) padded (2-D: (nx+2)*(ny+2)), ghost cells
 * hold periodic images (physical-boundary ghosts are left 0).  Returns the number of V-cycles;
 * *resnorm = final |r|_inf / |rh|_inf.
 */
/* alpha, phi0: NULL, or nx*ny*nz cell arrays (no ghosts) -- the Helmholtz form (alpha - div beta grad) phi = rh of viscsolve.f90
 * (visc_solve: alpha = rho, diff_scalar_solve: alpha = 1) and its initial guess (the current field).  Dirichlet DATA is not handled here:
 * the caller folds 8/3 beta phi_b / h^2 into rh (see orc_helm_rhs in orc_glue.c). */
int orc_mg_solve_ex(int dim, const int *n, const double *h, const int *ell_bc /* [3][2] */,
                    const double *rh, const double *bx, const double *by, const double *bz, const double *alpha, const double *phi0,
                    double *phi, double rel_eps, int max_cycles, int nu1, int nu2, double bottom_eps,
                    int verbose, double *resnorm)
{
    mgtower T; T.dim = dim;
    for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) T.bc[d][s] = (d < dim) ? ell_bc[d*2+s] : ELL_NEU;
    T.singular = alpha ? 0 : 1;
    for (int d = 0; d < dim; ++d) for (int s = 0; s < 2; ++s) if (T.bc[d][s] == ELL_DIR) T.singular = 0;

    int nn[3] = { n[0], n[1], dim == 3 ? n[2] : 1 };
    int nlev = 1;
    { int m[3] = { nn[0], nn[1], nn[2] };
      for (;;) { int ok = 1; for (int d = 0; d < dim; ++d) if (m[d] % 2 != 0 || m[d]/2 < 2) ok = 0;
                 if (!ok) break;
                 for (int d = 0; d < dim; ++d) m[d] /= 2;
                 ++nlev; } }
    T.nlev = nlev;
    T.L = (mglev*)calloc(nlev, sizeof(mglev));
    { int m[3] = { nn[0], nn[1], nn[2] }; double hh[3] = { h[0], h[1], dim == 3 ? h[2] : 1.0 };
      for (int l = 0; l < nlev; ++l) { lev_alloc(&T.L[l], m[0], m[1], m[2], hh, dim);
          for (int d = 0; d < dim; ++d) { m[d] /= 2; hh[d] *= 2.0; } } }

    mglev *F = &T.L[0];
    const double *bsrc[3] = { bx, by, bz };
    for (int k = 0; k < nn[2]; ++k) for (int j = 0; j < nn[1]; ++j) for (int i = 0; i < nn[0]; ++i)
        F->rhs[IDX(F,i,j,k)] = rh[(long)i + (long)nn[0]*(j + (long)nn[1]*k)];
    for (int d = 0; d < dim; ++d) {
        int e[3] = { d == 0, d == 1, d == 2 };
        long m0 = nn[0]+e[0], m1 = nn[1]+e[1];
        for (int k = 0; k < nn[2]+e[2]; ++k) for (int j = 0; j < nn[1]+e[1]; ++j) for (int i = 0; i < nn[0]+e[0]; ++i)
            F->b[d][IDX(F,i,j,k)] = bsrc[d][(long)i + m0*(j + m1*k)];
    }
    if (phi0)
        for (int k = 0; k < nn[2]; ++k) for (int j = 0; j < nn[1]; ++j) for (int i = 0; i < nn[0]; ++i)
            F->phi[IDX(F,i,j,k)] = phi0[(long)i + (long)nn[0]*(j + (long)nn[1]*k)];
    if (alpha) {
        for (int l = 0; l < nlev; ++l) T.L[l].alpha = (double*)calloc(T.L[l].ntot, sizeof(double));
        for (int k = 0; k < nn[2]; ++k) for (int j = 0; j < nn[1]; ++j) for (int i = 0; i < nn[0]; ++i)
            F->alpha[IDX(F,i,j,k)] = alpha[(long)i + (long)nn[0]*(j + (long)nn[1]*k)];
        for (int l = 1; l < nlev; ++l) {        /* coarse alpha = average of the fine cells it covers */
            mglev *f = &T.L[l-1], *c = &T.L[l];
            const int rz = dim == 3 ? 2 : 1;
            for (int k = 0; k < c->n[2]; ++k) for (int j = 0; j < c->n[1]; ++j) for (int i = 0; i < c->n[0]; ++i) {
                double s = 0.0;
                for (int kk = 0; kk < rz; ++kk) for (int jj = 0; jj < 2; ++jj) for (int ii = 0; ii < 2; ++ii)
                    s += f->alpha[IDX(f, 2*i+ii, 2*j+jj, rz*k+kk)];
                c->alpha[IDX(c,i,j,k)] = s/(4.0*rz);
            }
        }
    }
    /* coarsen coefficients: arithmetic average of the fine faces covering the coarse face */
    for (int l = 1; l < nlev; ++l) {
        mglev *f = &T.L[l-1], *c = &T.L[l];
        const int rz = dim == 3 ? 2 : 1;
        for (int d = 0; d < dim; ++d) {
            int e[3] = { d == 0, d == 1, d == 2 };
            for (int k = 0; k < c->n[2]+e[2]; ++k) for (int j = 0; j < c->n[1]+e[1]; ++j) for (int i = 0; i < c->n[0]+e[0]; ++i) {
                double s = 0.0; int cnt = 0;
                for (int kk = 0; kk < (d == 2 ? 1 : rz); ++kk)
                for (int jj = 0; jj < (d == 1 ? 1 : 2); ++jj)
                for (int ii = 0; ii < (d == 0 ? 1 : 2); ++ii) { s += f->b[d][IDX(f, 2*i+ii, 2*j+jj, rz*k+kk)]; ++cnt; }
                c->b[d][IDX(c,i,j,k)] = s/cnt;
            }
        }
    }

    double bnorm = 0.0;
    for (int k = 0; k < nn[2]; ++k) for (int j = 0; j < nn[1]; ++j) for (int i = 0; i < nn[0]; ++i)
        bnorm = dmax(bnorm, fabs(F->rhs[IDX(F,i,j,k)]));
    int cycles = 0;
    double rn = residual(&T, F);
    if (verbose) printf("orc_mg: levels %d, |rh|=%g, initial |r|=%g\n", nlev, bnorm, rn);
    while (bnorm > 0.0 && rn > rel_eps*bnorm && cycles < max_cycles) {
        vcycle(&T, 0, nu1, nu2, bottom_eps);
        rn = residual(&T, F);
        ++cycles;
        if (verbose) printf("orc_mg: cycle %d |r|/|rh| = %g\n", cycles, rn/bnorm);
    }
    if (resnorm) *resnorm = bnorm > 0.0 ? rn/bnorm : 0.0;
    fill_periodic(&T, F, F->phi);
    const long np = (dim == 3) ? F->ntot : F->s[2];
    if (dim == 3) memcpy(phi, F->phi, np*8);
    else          memcpy(phi, F->phi + F->s[2], np*8);   /* the k=0 plane */
    for (int l = 0; l < nlev; ++l) lev_free(&T.L[l]);
    free(T.L);
    return cycles;
}

int orc_mg_solve(int dim, const int *n, const double *h, const int *ell_bc /* [3][2] */,
                 const double *rh, const double *bx, const double *by, const double *bz,
                 double *phi, double rel_eps, int max_cycles, int nu1, int nu2, double bottom_eps,
                 int verbose, double *resnorm)
{
    return orc_mg_solve_ex(dim, n, h, ell_bc, rh, bx, by, bz, NULL, NULL, phi, rel_eps, max_cycles, nu1, nu2, bottom_eps, verbose, resnorm);
}
