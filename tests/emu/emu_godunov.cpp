// emu_godunov.cpp -- test infrastructure: runs the real Godunov kernel source and stage orchestration
// (varden_b200/csrc/vdn_godunov_kernels.cuh) on the CPU for tests/test_emu_godunov.py.  One box = the whole region.
#define VDN_EMU 1
#include "cuda_emu.h"
#include "../../varden_b200/csrc/vdn_common.cuh"
#include "../../varden_b200/csrc/vdn_godunov_kernels.cuh"
#ifndef MARCH_HEADER            // development: -DMARCH_HEADER='"/path/to/a/working/copy.cuh"'
#define MARCH_HEADER "../../varden_b200/csrc/vdn_godunov_march.cuh"
#endif
#include MARCH_HEADER

namespace {
struct NoScope { };
struct EmuLauncher {
    NoScope scope(const char *, double, int) { return NoScope(); }
    template <class A> void operator()(void (*k)(A), const Range &r, const A &a) { emu_launch_seq(k, grid3(r, dim3(64, 4, 1)), dim3(64, 4, 1), a); }
};
// view of a host array a(-ng:n0+ng-1(+face), ..., ncomp) pointing at cell (0,0,0)
View mkview(double *base, const int *n, int ng, int fdir)
{
    View v;
    const long e0 = n[0] + 2 * ng + (fdir == 0), e1 = n[1] + 2 * ng + (fdir == 1), e2 = n[2] + 2 * ng + (fdir == 2);
    v.sy = e0; v.sz = e0 * e1; v.cs = e0 * e1 * e2;
    v.p = base + ng + v.sy * ng + v.sz * ng;
    return v;
}
Geo mkgeo(const int *n, const int *pbc, const double *h)
{
    Geo g; memset(&g, 0, sizeof g);
    g.dim = 3;
    for (int d = 0; d < 3; ++d) { g.n[d] = n[d]; g.h[d] = h[d]; g.pbc[d][0] = pbc[2 * d]; g.pbc[d][1] = pbc[2 * d + 1]; g.nb[d] = 1; g.cut[d][0] = 0; g.cut[d][1] = n[d]; }
    return g;
}
// launcher of the plane-marching kernels: one OS thread per CUDA thread, warp shuffles and shared memory emulated
struct MarchLauncher {
    NoScope scope(const char *, double, int) { return NoScope(); }
    template <class A> void run(void (*k)(A), dim3 grid, dim3 block, size_t smem, const A &a) { emu_launch2(k, grid, block, smem, a); }
    template <class A> int slots(void (*)(A), int, size_t) { return 1; }
};
struct Scratch {
    std::vector<double> buf; long sn, sy, sz, off;
    Scratch(const int *n, int nslots) {
        sy = n[0] + 2; sz = sy * (n[1] + 2); sn = sz * (n[2] + 2); off = 1 + sy + sz;
        buf.assign((size_t)sn * nslots, std::nan(""));
    }
    View S(int q) { View v; v.sy = sy; v.sz = sz; v.cs = sn; v.p = buf.data() + (long)q * sn + off; return v; }
};
}

// adv_bc: [3 comps][3 dirs][2 sides]
extern "C" int emu_velpred(int fused, const int *n, const int *pbc, const int *adv_bc, int order, int use_minion, double dt, const double *h,
                           double eps, double *u, double *force, double *umac0, double *umac1, double *umac2)
{
    Scratch sc(n, 36);
    VpArgs a; a.g = mkgeo(n, pbc, h); a.u = mkview(u, n, 3, -1); a.force = mkview(force, n, 1, -1);
    a.eps = &eps; a.dt = dt; a.use_minion = use_minion; a.order = order;
    for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) a.sbc[c][d][s] = adv_bc[(c * 3 + d) * 2 + s];
    double *um[3] = { umac0, umac1, umac2 };
    for (int d = 0; d < 3; ++d) {
        a.sl[d] = sc.S(0 + 3 * d); a.ul[d] = sc.S(9 + 3 * d); a.ur[d] = sc.S(18 + 3 * d); a.uimh[d] = sc.S(27 + 3 * d);
        a.out[d] = mkview(um[d], n, 1, d);
        for (int t = 0; t < 3; ++t) a.X[d][t] = sc.S(0 + d * 3 + t);
    }
    EmuLauncher L;
    velpred_stages<3>(L, a);
    return 0;
}

// one component; sbc: [3 dirs][2 sides]
extern "C" int emu_mkflux(int fused, const int *n, const int *pbc, const int *sbc, int order, int use_minion, int is_vel, int comp, int cons, int use_rhs,
                          double dt, const double *h, double eps, double *s, double *mac0, double *mac1, double *mac2, double *force, double *mac_rhs,
                          double *sedge0, double *sedge1, double *sedge2, double *flux0, double *flux1, double *flux2)
{
    Scratch sc(n, 21);
    MfArgs a; a.g = mkgeo(n, pbc, h); a.s = mkview(s, n, 3, -1); a.force = mkview(force, n, 1, -1); a.mac_rhs = mkview(mac_rhs, n, 1, -1);
    a.eps = &eps; a.dt = dt; a.use_minion = use_minion; a.is_vel = is_vel; a.comp = comp; a.cons = cons; a.use_rhs = use_rhs; a.order = order;
    for (int d = 0; d < 3; ++d) for (int sd = 0; sd < 2; ++sd) a.sbc[d][sd] = sbc[2 * d + sd];
    double *mac[3] = { mac0, mac1, mac2 }, *se[3] = { sedge0, sedge1, sedge2 }, *fl[3] = { flux0, flux1, flux2 };
    for (int d = 0; d < 3; ++d) {
        a.mac[d] = mkview(mac[d], n, 1, d);
        a.sl[d] = sc.S(d); a.l[d] = sc.S(3 + d); a.rr[d] = sc.S(6 + d); a.simh[d] = sc.S(9 + d);
        for (int t = 0; t < 3; ++t) a.X[d][t] = sc.S(12 + d * 3 + t);
        a.sedge[d] = mkview(se[d], n, 0, d); a.flux[d] = mkview(fl[d], n, 0, d);
    }
    EmuLauncher L;
    mkflux_stages<3>(L, a);
    return 0;
}

// plane-marching mkflux: all ncomp components of s (ng 3), force (ng 1), sedge (ng 0, ncomp comps), flux (ng 0, comp 0 only);
// adv_bc: [ncomp][3 dirs][2 sides]; slots = resident-CTA count the z-chunking is planned for
extern "C" int emu_mkflux_march(const int *n, const int *pbc, const int *adv_bc, int order, int use_minion, int is_vel, int ncomp,
                                double dt, const double *h, double eps, int slots, double *s, double *mac0, double *mac1, double *mac2, double *force,
                                double *sedge0, double *sedge1, double *sedge2, double *flux0, double *flux1, double *flux2)
{
    Geo g = mkgeo(n, pbc, h);
    View sv = mkview(s, n, 3, -1), fv = mkview(force, n, 1, -1);
    double *mac[3] = { mac0, mac1, mac2 }, *se[3] = { sedge0, sedge1, sedge2 }, *fl[3] = { flux0, flux1, flux2 };
    View mv[3], ev[3], xv[3];
    for (int d = 0; d < 3; ++d) { mv[d] = mkview(mac[d], n, 1, d); ev[d] = mkview(se[d], n, 0, d); xv[d] = mkview(fl[d], n, 0, d); }
    int bc[8][3][2];
    for (int c = 0; c < ncomp; ++c) for (int d = 0; d < 3; ++d) for (int sd = 0; sd < 2; ++sd) bc[c][d][sd] = adv_bc[(c * 3 + d) * 2 + sd];
    MarchLauncher L;
    march::mkflux_march(L, g, sv, fv, mv, ev, xv, &eps, dt, is_vel, ncomp, order, use_minion, bc, slots);
    return 0;
}

extern "C" int emu_velpred_march(const int *n, const int *pbc, const int *adv_bc, int order, int use_minion, double dt, const double *h,
                                 double eps, int slots, double *u, double *force, double *umac0, double *umac1, double *umac2)
{
    Geo g = mkgeo(n, pbc, h);
    View uv = mkview(u, n, 3, -1), fv = mkview(force, n, 1, -1);
    double *um[3] = { umac0, umac1, umac2 };
    View ov[3];
    for (int d = 0; d < 3; ++d) ov[d] = mkview(um[d], n, 1, d);
    int bc[3][3][2];
    for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) for (int sd = 0; sd < 2; ++sd) bc[c][d][sd] = adv_bc[(c * 3 + d) * 2 + sd];
    MarchLauncher L;
    march::velpred_march(L, g, uv, fv, ov, &eps, dt, order, use_minion, bc, slots);
    return 0;
}
