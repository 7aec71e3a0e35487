// emu_wave.cpp -- test infrastructure: CPU emulation of k_sweep3 (the fused red-black multigrid smoother, vdn_mg_fused.cuh) for
// tests/test_emu_wave.py.  cfg 0..4 are the PRODUCTION tile shapes of the launcher (vdn_mg.cu: sweep3_get), cfg 5 a small one.
#define VDN_EMU 1
#include "cuda_emu.h"
double sm[1 << 17];
#ifndef FUSED_HEADER            // development: -DFUSED_HEADER='"/path/to/a/working/copy.cuh"'
#define FUSED_HEADER "../../varden_b200/csrc/vdn_mg_fused.cuh"
#endif
#include FUSED_HEADER

// peer_delta: null, or 27 byte distances (process-grid offset (ox,oy,oz) at (ox+1)+3(oy+1)+9(oz+1)) from this "rank's" out / crhs / czero
// arrays to the same arrays of the neighbour "ranks" (the test lays the three arrays of every rank out in one buffer, like the symmetric heap):
// the peer-memory mode, in which the kernel also stores what it writes near a shared face into the neighbours' ghost layers
extern "C" int emu_sweep3_p2p(int nsw, int pre, int post, int cfg, const int *n, const int *mode, int par0, const double *h2,
                              const double *rhs, const double *b0, const double *b1, const double *b2,
                              const double *in, double *out, const double *cphi, double *crhs, double *czero, double *nrm, int zchunk, int pad,
                              const long *peer_delta, const double *dinv)
{
    if (nsw != 1) return 1;
    WaveArgs a;
    memset(&a, 0, sizeof a);
    if (peer_delta) { a.p2p = 1; for (int q = 0; q < 27; ++q) { a.peer_delta[q] = peer_delta[q]; if (peer_delta[q] != 0) a.peer_mask |= 1u << q; } }
    for (int d = 0; d < 3; ++d) { a.n[d] = n[d]; a.h2[d] = h2[d]; a.mode[d][0] = mode[2 * d]; a.mode[d][1] = mode[2 * d + 1]; }
    a.s1 = n[0] + 2 * pad; a.s2 = (long)(n[0] + 2 * pad) * (n[1] + 2 * pad); a.off = pad * (1 + a.s1 + a.s2); a.par0 = par0;
    a.rhs = rhs; a.b0 = b0; a.b1 = b1; a.b2 = b2; a.dinv = dinv; a.in = in; a.out = out;
    a.cphi = cphi; a.crhs = crhs; a.czero = czero;
    a.cs1 = n[0] / 2 + 2 * pad; a.cs2 = (long)(n[0] / 2 + 2 * pad) * (n[1] / 2 + 2 * pad); a.coff = pad * (1 + a.cs1 + a.cs2);
    a.nrm = nrm; a.zchunk = zchunk;
#define G3(PRE, POST, C, TX, TY) if (pre == PRE && post == POST && cfg == C) { \
        const dim3 gr((n[0] + TX - 1) / TX, (n[1] + TY - 1) / TY, (n[2] + zchunk - 1) / zchunk); \
        if (a.p2p) emu_launch(k_sweep3<PRE, POST, TX, TY, true>, gr, Sweep3Cfg<PRE, POST, TX, TY>::NT, a); \
        else       emu_launch(k_sweep3<PRE, POST, TX, TY, false>, gr, Sweep3Cfg<PRE, POST, TX, TY>::NT, a); \
        return 0; }
#define G36(C, TX, TY) G3(0, 0, C, TX, TY) G3(0, 2, C, TX, TY) G3(0, 3, C, TX, TY) G3(1, 0, C, TX, TY) G3(1, 2, C, TX, TY) G3(1, 3, C, TX, TY)
    G36(0, 32, 32) G36(1, 64, 16) G36(2, 32, 16) G36(3, 64, 14) G36(4, 32, 24) G36(5, 16, 8)
    return 1;
}
extern "C" int emu_sweep3(int nsw, int pre, int post, int cfg, const int *n, const int *mode, int par0, const double *h2,
                          const double *rhs, const double *b0, const double *b1, const double *b2,
                          const double *in, double *out, const double *cphi, double *crhs, double *czero, double *nrm, int zchunk, int pad, const double *dinv)
{
    return emu_sweep3_p2p(nsw, pre, post, cfg, n, mode, par0, h2, rhs, b0, b1, b2, in, out, cphi, crhs, czero, nrm, zchunk, pad, nullptr, dinv);
}
