/*
 * oracle/orc_common.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the VARDEN hot path).
 *
 * This directory is the *checker*: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it.  The product
 * (varden_b200/libvdn.so) never links or loads anything from oracle/.
 *
 * Parity status: PINNED for the Godunov / update / physbc / force / macproject-glue arithmetic -- oracle/f2c.py
 * transpiles the reference's own per-box Fortran routines to C (oracle/_ref/, built where /root/reference is
 * mounted) and tests/test_ref_pin.py demands bit-identical results stage by stage; tests/golden/*.npz carry the
 * same reference outputs to machines without the reference tree (tests/test_golden.py).
 * The multigrid (FBoxLib F_MG, third party, absent, version unpinned) and multifab_fill_boundary (FBoxLib) are
 * "parity unpinned": restated from the published algorithm / documented semantics; the MG is checked against a
 * direct sparse solve of the same discrete operator.
 *
 * Array convention everywhere: Fortran column-major boxes with ghost cells,
 *   a(lo1-ng:hi1+ng, lo2-ng:hi2+ng, lo3-ng:hi3+ng, ncomp), i fastest,
 * face arrays in direction d have one extra point (hi_d+1) in d.
 */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* FBoxLib bc_module codes (F_BaseLib/bc.f90; used at define_bc_tower.f90:158-340) */
#define BC_PERIODIC     (-1)
#define BC_INTERIOR     0
#define BC_INLET        11
#define BC_OUTLET       12
#define BC_SYMMETRY     13
#define BC_SLIP_WALL    14
#define BC_NO_SLIP_WALL 15
#define BC_REFLECT_ODD  20
#define BC_REFLECT_EVEN 21
#define BC_FOEXTRAP     22
#define BC_EXT_DIR      23
#define BC_HOEXTRAP     24
/* elliptic */
#define ELL_PER (-1)
#define ELL_INT 0
#define ELL_DIR 1
#define ELL_NEU 2

#define HALF 0.5
#define ZERO 0.0
#define ONE  1.0
#define TWO  2.0

/* A 4-D view with explicit lower bounds (Fortran style). */
typedef struct {
    double *p;
    int  l[3];   /* lower bound of each index */
    long n[3];   /* extent of each index      */
    long cs;     /* component stride          */
    int  nc;
} V;

#define AT(v,i,j,k,c) ((v).p[ ((long)(i)-(v).l[0]) + (v).n[0]*( ((long)(j)-(v).l[1]) + (v).n[1]*((long)(k)-(v).l[2]) ) + (v).cs*(long)(c) ])

static inline V v_wrap(double *p, int l0, int h0, int l1, int h1, int l2, int h2, int nc)
{
    V v; v.p = p;
    v.l[0] = l0; v.l[1] = l1; v.l[2] = l2;
    v.n[0] = h0 - l0 + 1; v.n[1] = h1 - l1 + 1; v.n[2] = h2 - l2 + 1;
    v.cs = v.n[0]*v.n[1]*v.n[2]; v.nc = nc;
    return v;
}
/* cell box lo:hi grown by ng; face_dir in {0,1,2} adds one point on the hi side of that dir; -1 = cell centred.
 * dim==2 arrays have a unit third extent (k index fixed at lo[2], which callers pass as 0). */
static inline V v_box(double *p, const int *lo, const int *hi, int ng, int face_dir, int nc, int dim)
{
    int l[3], h[3];
    for (int d = 0; d < 3; ++d) {
        if (d < dim) { l[d] = lo[d]-ng; h[d] = hi[d]+ng + (d == face_dir ? 1 : 0); }
        else         { l[d] = 0; h[d] = 0; }
    }
    return v_wrap(p, l[0],h[0], l[1],h[1], l[2],h[2], nc);
}
static inline V v_alloc(int l0, int h0, int l1, int h1, int l2, int h2, int nc)
{
    long n = (long)(h0-l0+1)*(h1-l1+1)*(h2-l2+1)*nc;
    double *p = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    return v_wrap(p, l0,h0, l1,h1, l2,h2, nc);
}
static inline void v_free(V *v) { free(v->p); v->p = NULL; }

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

#endif
