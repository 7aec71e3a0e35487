/*
 * tests/mock_driver.c -- test infrastructure: a compiled C stand-in for VARDEN's Fortran driver.
 *
 * The Fortran side cannot be compiled here (no Fortran compiler, no FBoxLib), so this program plays it: it owns "multifabs"
 * (one Fortran-layout host array a(lo-ng:hi+ng, ..., ncomp) per box, exactly what dataptr(mf,i) addresses) and replays
 * advance_timestep.f90:95-124 through the SAME-NAMED procedures of fortran/vdn_modules.f90 --
 *     velpred(u, umac, force)                      velpred.f90:16        (called from advance_premac.f90:51)
 *     macproject(umac, rho, mac_rhs)               macproject.f90:20     (advance_timestep.f90:100)
 *     mkflux(sold, sedge, flux, umac, force, ...)  mkflux.f90:16-17      (scalar_advance.f90:102, velocity_advance.f90:76)
 *     update(sold, umac, sedge, flux, force, snew) update.f90:16-17      (scalar_advance.f90:118, velocity_advance.f90:92)
 * -- each of which, like its Fortran twin, copies its input multifabs to the device through the C ABI (include/vdn.h), runs
 * the stage and copies its outputs back; the host arrays are authoritative between the calls (the device copies of a
 * procedure's inputs are overwritten with a poison value before it uploads them).  mkforce / make_at_halftime are not among
 * the replaced modules; they run through their own ABI entry points with the same copy discipline.
 *
 * usage: mock_driver <libvdn.so> <deck.bin> <out.bin>   (deck written and output checked by tests/test_mock_driver.py
 * against the golden fixtures of the reference's own routines, tests/golden/).  Not ctypes: plain C against the ABI.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/vdn.h"

#define MAXB 64
typedef struct { int ng, nc, fdir, field; double *p[MAXB]; size_t n[MAXB]; } mfab;   /* fdir: -1 cell, 0..2 face */

static int dim, nboxes, nscal;
static int blo[MAXB][3], bhi[MAXB][3];
static vdn_ctx *ctx;

/* ABI entry points, bound at run time */
#define FN(ret, name, args) static ret (*p_##name) args;
FN(void, vdn_params_default, (vdn_params *))
FN(int, vdn_ctx_create, (const vdn_params *, int, int, const int *, const int *, const int *, const int *, const int *, const double *, int, vdn_ctx **))
FN(void, vdn_ctx_destroy, (vdn_ctx *))
FN(const char *, vdn_last_error, (const vdn_ctx *))
FN(int, vdn_field_upload, (vdn_ctx *, int, int, const double *, int, int))
FN(int, vdn_field_download, (vdn_ctx *, int, int, double *, int, int))
FN(int, vdn_field_setval, (vdn_ctx *, int, double))
FN(int, vdn_mkvelforce, (vdn_ctx *, int, double))
FN(int, vdn_mkscalforce, (vdn_ctx *, double))
FN(int, vdn_velpred, (vdn_ctx *, double))
FN(int, vdn_macproject, (vdn_ctx *, double, double, int *, double *))
FN(int, vdn_mkflux, (vdn_ctx *, int, double))
FN(int, vdn_update, (vdn_ctx *, int, double))
FN(int, vdn_make_at_halftime, (vdn_ctx *))

static void die(const char *what) { fprintf(stderr, "mock_driver: %s: %s\n", what, ctx ? p_vdn_last_error(ctx) : p_vdn_last_error(NULL)); exit(1); }
#define CHK(call) do { if ((call) != 0) die(#call); } while (0)

static mfab mf_build(int ng, int nc, int fdir, int field)            /* multifab_build / multifab_build_edge */
{
    mfab m; memset(&m, 0, sizeof m); m.ng = ng; m.nc = nc; m.fdir = fdir; m.field = field;
    for (int b = 0; b < nboxes; ++b) {
        size_t n = nc;
        for (int d = 0; d < dim; ++d) n *= (size_t)(bhi[b][d] - blo[b][d] + 1 + 2 * ng + (d == fdir));
        m.n[b] = n; m.p[b] = (double *)calloc(n, sizeof(double));
    }
    return m;
}
static void mf_read(mfab *m, FILE *f) { for (int b = 0; b < nboxes; ++b) if (fread(m->p[b], 8, m->n[b], f) != m->n[b]) { fprintf(stderr, "short deck\n"); exit(1); } }
static void mf_write(const mfab *m, FILE *f) { for (int b = 0; b < nboxes; ++b) fwrite(m->p[b], 8, m->n[b], f); }
/* vdn_put / vdn_get of fortran/vdn_iso_c.f90; nc < m->nc copies the leading components only (the density flux) */
static void put_nc(const mfab *m, int field, int nc)
{
    CHK(p_vdn_field_setval(ctx, field, -7.0e77));                      /* the device copy is NOT trusted between procedures */
    for (int b = 0; b < nboxes; ++b) CHK(p_vdn_field_upload(ctx, field, b, m->p[b], m->ng, nc));
}
static void put(const mfab *m, int field) { put_nc(m, field, m->nc); }
static void get_nc(mfab *m, int field, int nc) { for (int b = 0; b < nboxes; ++b) CHK(p_vdn_field_download(ctx, field, b, m->p[b], m->ng, nc)); }
static void get(mfab *m, int field) { get_nc(m, field, m->nc); }

/* ---- the same-named procedures (bodies = fortran/vdn_modules.f90) ---- */
static void velpred(const mfab *u, mfab *umac, const mfab *force, double dt)
{
    put(u, VDN_UOLD); put(force, VDN_VEL_FORCE);
    CHK(p_vdn_velpred(ctx, dt));
    for (int d = 0; d < dim; ++d) get(&umac[d], VDN_UMAC_X + d);
}
static void macproject(mfab *umac, const mfab *rho, const mfab *mac_rhs, int *ncyc, double *res)
{
    put(rho, VDN_SOLD); put(mac_rhs, VDN_MAC_RHS);
    for (int d = 0; d < dim; ++d) put(&umac[d], VDN_UMAC_X + d);
    int rc = p_vdn_macproject(ctx, 1.0e-10, -1.0, ncyc, res);
    if (rc != 0 && rc != 2) die("vdn_macproject");
    for (int d = 0; d < dim; ++d) get(&umac[d], VDN_UMAC_X + d);
}
static void mkflux(const mfab *sold, mfab *sedge, mfab *flux, const mfab *umac, const mfab *force, const mfab *mac_rhs, double dt, int is_vel)
{
    put(sold, is_vel ? VDN_UOLD : VDN_SOLD); put(force, is_vel ? VDN_VEL_FORCE : VDN_SCAL_FORCE); put(mac_rhs, VDN_MAC_RHS);
    for (int d = 0; d < dim; ++d) put(&umac[d], VDN_UMAC_X + d);
    CHK(p_vdn_mkflux(ctx, is_vel, dt));
    for (int d = 0; d < dim; ++d) {
        get(&sedge[d], (is_vel ? VDN_UEDGE_X : VDN_SEDGE_X) + d);
        if (!is_vel) get_nc(&flux[d], VDN_SFLUX_X + d, 1);
    }
}
static void update(const mfab *sold, const mfab *umac, const mfab *sedge, const mfab *flux, const mfab *force, mfab *snew, double dt, int is_vel)
{
    put(sold, is_vel ? VDN_UOLD : VDN_SOLD); put(force, is_vel ? VDN_VEL_FORCE : VDN_SCAL_FORCE);
    for (int d = 0; d < dim; ++d) {
        put(&umac[d], VDN_UMAC_X + d);
        put(&sedge[d], (is_vel ? VDN_UEDGE_X : VDN_SEDGE_X) + d);
        if (!is_vel) put_nc(&flux[d], VDN_SFLUX_X + d, 1);
    }
    CHK(p_vdn_update(ctx, is_vel, dt));
    get(snew, is_vel ? VDN_UNEW : VDN_SNEW);
}
/* not replaced by vdn_modules.f90 (they stay the reference's mkforce.f90 / make_at_halftime.f90); here through their own entry points */
static void mkvelforce(mfab *vel_force, const mfab *ext, const mfab *gp, const mfab *rho, int rho_field, double visc_fac)
{
    put(ext, VDN_EXT_VEL_FORCE); put(gp, VDN_GP);
    put_nc(rho, rho_field, rho_field == VDN_RHOHALF ? 1 : rho->nc);   /* rhohalf has dm comps, only the first is meaningful */
    CHK(p_vdn_mkvelforce(ctx, rho_field, visc_fac));
    get(vel_force, VDN_VEL_FORCE);
}
static void mkscalforce(mfab *scal_force, const mfab *ext, double diff_fac)
{
    put(ext, VDN_EXT_SCAL_FORCE);
    CHK(p_vdn_mkscalforce(ctx, diff_fac));
    get(scal_force, VDN_SCAL_FORCE);
}
static void make_at_halftime(mfab *rhohalf, const mfab *sold, const mfab *snew)
{
    put(sold, VDN_SOLD); put(snew, VDN_SNEW);
    CHK(p_vdn_make_at_halftime(ctx));
    get_nc(rhohalf, VDN_RHOHALF, 1);
}

int main(int argc, char **argv)
{
    if (argc != 4) { fprintf(stderr, "usage: mock_driver libvdn.so deck.bin out.bin\n"); return 2; }
    void *h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "mock_driver: %s\n", dlerror()); return 2; }
#define BIND(name) do { *(void **)(&p_##name) = dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; } } while (0)
    BIND(vdn_params_default); BIND(vdn_ctx_create); BIND(vdn_ctx_destroy); BIND(vdn_last_error); BIND(vdn_field_upload);
    BIND(vdn_field_download); BIND(vdn_field_setval); BIND(vdn_mkvelforce); BIND(vdn_mkscalforce); BIND(vdn_velpred);
    BIND(vdn_macproject); BIND(vdn_mkflux); BIND(vdn_update); BIND(vdn_make_at_halftime);

    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 2; }
    int hdr[6], dlo[3], dhi[3], pbc[6], use_given_umac;
    double dx[3], dt, bcval[30];
    if (fread(hdr, 4, 6, f) != 6) return 2;
    dim = hdr[0]; nboxes = hdr[1]; nscal = hdr[2];
    if (nboxes > MAXB) return 2;
    int lo[MAXB * 3], hi[MAXB * 3];
    if (fread(lo, 4, 3 * nboxes, f) != (size_t)(3 * nboxes) || fread(hi, 4, 3 * nboxes, f) != (size_t)(3 * nboxes)) return 2;
    for (int b = 0; b < nboxes; ++b) for (int d = 0; d < 3; ++d) { blo[b][d] = lo[3 * b + d]; bhi[b][d] = hi[3 * b + d]; }
    if (fread(dlo, 4, 3, f) != 3 || fread(dhi, 4, 3, f) != 3 || fread(pbc, 4, 6, f) != 6 || fread(dx, 8, 3, f) != 3 ||
        fread(&dt, 8, 1, f) != 1 || fread(bcval, 8, 30, f) != 30 || fread(&use_given_umac, 4, 1, f) != 1) return 2;

    vdn_params prm; p_vdn_params_default(&prm);
    prm.nscal = nscal; prm.slope_order = hdr[3]; prm.use_minion = hdr[4]; prm.boussinesq = hdr[5];
    memcpy(prm.bc_val, bcval, sizeof bcval);
    if (p_vdn_ctx_create(&prm, dim, nboxes, lo, hi, dlo, dhi, pbc, dx, 0, &ctx) != 0) die("vdn_ctx_create");

    /* the multifabs of advance_timestep.f90:26-80 */
    mfab uold = mf_build(3, dim, -1, VDN_UOLD), sold = mf_build(3, nscal, -1, VDN_SOLD), gp = mf_build(1, dim, -1, VDN_GP);
    mfab evf = mf_build(1, dim, -1, VDN_EXT_VEL_FORCE), esf = mf_build(1, nscal, -1, VDN_EXT_SCAL_FORCE);
    mfab unew = mf_build(3, dim, -1, VDN_UNEW), snew = mf_build(3, nscal, -1, VDN_SNEW), rhohalf = mf_build(1, dim, -1, VDN_RHOHALF);
    mfab mac_rhs = mf_build(1, 1, -1, VDN_MAC_RHS), divu = mf_build(1, 1, -1, VDN_MAC_RHS);
    mfab vel_force = mf_build(1, dim, -1, VDN_VEL_FORCE), scal_force = mf_build(1, nscal, -1, VDN_SCAL_FORCE);
    mfab umac[3], umac_given[3], sedge[3], sflux[3], uedge[3], uflux[3];
    for (int d = 0; d < dim; ++d) {
        umac[d] = mf_build(1, 1, d, VDN_UMAC_X + d); umac_given[d] = mf_build(1, 1, d, VDN_UMAC_X + d);
        sedge[d] = mf_build(0, nscal, d, VDN_SEDGE_X + d); sflux[d] = mf_build(0, nscal, d, VDN_SFLUX_X + d);
        uedge[d] = mf_build(0, dim, d, VDN_UEDGE_X + d); uflux[d] = mf_build(0, dim, d, VDN_UEDGE_X + d);
        for (int b = 0; b < nboxes; ++b) for (size_t q = 0; q < umac[d].n[b]; ++q) umac[d].p[b][q] = 1.0e20;     /* advance_timestep.f90:77 */
    }
    mf_read(&uold, f); mf_read(&sold, f); mf_read(&gp, f); mf_read(&evf, f); mf_read(&esf, f);
    for (int d = 0; d < dim; ++d) mf_read(&umac_given[d], f);
    fclose(f);

    FILE *o = fopen(argv[3], "wb");
    if (!o) { perror(argv[3]); return 2; }
    int ncyc = 0; double res = 0.0;
    /* advance_premac (advance_premac.f90:44-51) */
    mkvelforce(&vel_force, &evf, &gp, &sold, VDN_SOLD, 1.0);           mf_write(&vel_force, o);
    velpred(&uold, umac, &vel_force, dt);                              for (int d = 0; d < dim; ++d) mf_write(&umac[d], o);
    /* macproject (advance_timestep.f90:100) */
    macproject(umac, &sold, &mac_rhs, &ncyc, &res);                    for (int d = 0; d < dim; ++d) mf_write(&umac[d], o);
    /* the golden fixtures pin the downstream stages on a stored projected umac (MAC solves agree to the solver tolerance only) */
    mfab *um = use_given_umac ? umac_given : umac;
    /* scalar_advance (scalar_advance.f90:96-119) */
    mkscalforce(&scal_force, &esf, 1.0);
    mkflux(&sold, sedge, sflux, um, &scal_force, &divu, dt, 0);        for (int d = 0; d < dim; ++d) { mf_write(&sedge[d], o); mf_write(&sflux[d], o); }
    mkscalforce(&scal_force, &esf, 0.0);
    update(&sold, um, sedge, sflux, &scal_force, &snew, dt, 0);        mf_write(&snew, o);
    /* make_at_halftime (advance_timestep.f90:114) */
    make_at_halftime(&rhohalf, &sold, &snew);                          mf_write(&rhohalf, o);
    /* velocity_advance (velocity_advance.f90:70-93) */
    mkvelforce(&vel_force, &evf, &gp, &sold, VDN_SOLD, 1.0);
    mkflux(&uold, uedge, uflux, um, &vel_force, &mac_rhs, dt, 1);      for (int d = 0; d < dim; ++d) mf_write(&uedge[d], o);
    mkvelforce(&vel_force, &evf, &gp, &rhohalf, VDN_RHOHALF, 0.0);     mf_write(&vel_force, o);
    update(&uold, um, uedge, uflux, &vel_force, &unew, dt, 1);         mf_write(&unew, o);
    fwrite(&ncyc, 4, 1, o); fwrite(&res, 8, 1, o);
    fclose(o);
    p_vdn_ctx_destroy(ctx);
    printf("mock_driver: %d-D, %d boxes, MAC V-cycles %d, |r|/|rh| %.2e\n", dim, nboxes, ncyc, res);
    return 0;
}
