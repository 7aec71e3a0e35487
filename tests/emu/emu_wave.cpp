// emu_wave.cpp -- test infrastructure: CPU emulation of k_wave (the fused wavefront multigrid smoother) for tests/test_emu_wave.py
#define VDN_EMU 1
#include "cuda_emu.h"
double sm[1 << 17];
#include "../../varden_b200/csrc/vdn_mg_wave.cuh"

extern "C" int emu_wave(int nsw, int pre, int post, int cfg, const int *n, const int *mode, int par0, const double *h2,
                        const double *rhs, const double *b0, const double *b1, const double *b2,
                        const double *in, double *out, const double *cphi, double *crhs, double *czero, double *nrm, int zchunk, int pad)
{
    WaveArgs a;
    for (int d = 0; d < 3; ++d) { a.n[d] = n[d]; a.h2[d] = h2[d]; a.mode[d][0] = mode[2 * d]; a.mode[d][1] = mode[2 * d + 1]; }
    a.s1 = n[0] + 2 * pad; a.s2 = (long)(n[0] + 2 * pad) * (n[1] + 2 * pad); a.off = pad * (1 + a.s1 + a.s2); a.par0 = par0;
    a.rhs = rhs; a.b0 = b0; a.b1 = b1; a.b2 = b2; a.in = in; a.out = out;
    a.cphi = cphi; a.crhs = crhs; a.czero = czero;
    a.cs1 = n[0] / 2 + 2 * pad; a.cs2 = (long)(n[0] / 2 + 2 * pad) * (n[1] / 2 + 2 * pad); a.coff = pad * (1 + a.cs1 + a.cs2);
    a.nrm = nrm; a.zchunk = zchunk;
#define GO(NSW, PRE, POST) if (nsw == NSW && pre == PRE && post == POST) { \
        if (cfg == 0) { emu_launch(k_wave<NSW, PRE, POST, 32, 16, 512, 2>, dim3((n[0] + 31) / 32, (n[1] + 15) / 16, (n[2] + zchunk - 1) / zchunk), 512, a); return 0; } \
        else          { emu_launch(k_wave<NSW, PRE, POST, 32, 8, 256, 2>, dim3((n[0] + 31) / 32, (n[1] + 7) / 8, (n[2] + zchunk - 1) / zchunk), 256, a); return 0; } }
    GO(1, 0, 0) GO(1, 0, 2) GO(1, 0, 3) GO(1, 1, 0) GO(1, 1, 2) GO(1, 1, 3)
    return 1;
}

// ---- k_sweep (vdn_mg_sweep.cuh): same arguments, tile sizes small enough for one OS thread per CUDA thread ----
#include "../../varden_b200/csrc/vdn_mg_sweep.cuh"
extern "C" int emu_sweep(int nsw, int pre, int post, int cfg, const int *n, const int *mode, int par0, const double *h2,
                         const double *rhs, const double *b0, const double *b1, const double *b2,
                         const double *in, double *out, const double *cphi, double *crhs, double *czero, double *nrm, int zchunk, int pad)
{
    WaveArgs a;
    for (int d = 0; d < 3; ++d) { a.n[d] = n[d]; a.h2[d] = h2[d]; a.mode[d][0] = mode[2 * d]; a.mode[d][1] = mode[2 * d + 1]; }
    a.s1 = n[0] + 2 * pad; a.s2 = (long)(n[0] + 2 * pad) * (n[1] + 2 * pad); a.off = pad * (1 + a.s1 + a.s2); a.par0 = par0;
    a.rhs = rhs; a.b0 = b0; a.b1 = b1; a.b2 = b2; a.in = in; a.out = out;
    a.cphi = cphi; a.crhs = crhs; a.czero = czero;
    a.cs1 = n[0] / 2 + 2 * pad; a.cs2 = (long)(n[0] / 2 + 2 * pad) * (n[1] / 2 + 2 * pad); a.coff = pad * (1 + a.cs1 + a.cs2);
    a.nrm = nrm; a.zchunk = zchunk;
#define GS(NSW, PRE, POST, C, TX, TY) if (nsw == NSW && pre == PRE && post == POST && cfg == C) { \
        emu_launch(k_sweep<NSW, PRE, POST, TX, TY>, dim3((n[0] + TX - 1) / TX, (n[1] + TY - 1) / TY, (n[2] + zchunk - 1) / zchunk), \
                   SweepCfg<NSW, PRE, POST, TX, TY>::NT, a); return 0; }
#define GS6(NSW, C, TX, TY) GS(NSW, 0, 0, C, TX, TY) GS(NSW, 0, 2, C, TX, TY) GS(NSW, 0, 3, C, TX, TY) GS(NSW, 1, 0, C, TX, TY) GS(NSW, 1, 2, C, TX, TY) GS(NSW, 1, 3, C, TX, TY)
    GS6(1, 0, 16, 8) GS6(1, 1, 32, 16) GS6(2, 0, 16, 8) GS6(2, 1, 32, 16)
    return 1;
}

// ---- k_sweep2 (vdn_mg_sweep2.cuh): operator data staged through shared memory ----
#include "../../varden_b200/csrc/vdn_mg_sweep2.cuh"
extern "C" int emu_sweep2(int nsw, int pre, int post, int cfg, const int *n, const int *mode, int par0, const double *h2,
                          const double *rhs, const double *b0, const double *b1, const double *b2,
                          const double *in, double *out, const double *cphi, double *crhs, double *czero, double *nrm, int zchunk, int pad)
{
    if (nsw != 1) return 1;
    WaveArgs a;
    for (int d = 0; d < 3; ++d) { a.n[d] = n[d]; a.h2[d] = h2[d]; a.mode[d][0] = mode[2 * d]; a.mode[d][1] = mode[2 * d + 1]; }
    a.s1 = n[0] + 2 * pad; a.s2 = (long)(n[0] + 2 * pad) * (n[1] + 2 * pad); a.off = pad * (1 + a.s1 + a.s2); a.par0 = par0;
    a.rhs = rhs; a.b0 = b0; a.b1 = b1; a.b2 = b2; a.in = in; a.out = out;
    a.cphi = cphi; a.crhs = crhs; a.czero = czero;
    a.cs1 = n[0] / 2 + 2 * pad; a.cs2 = (long)(n[0] / 2 + 2 * pad) * (n[1] / 2 + 2 * pad); a.coff = pad * (1 + a.cs1 + a.cs2);
    a.nrm = nrm; a.zchunk = zchunk;
#define G2(PRE, POST, C, TX, TY) if (pre == PRE && post == POST && cfg == C) { \
        emu_launch(k_sweep2<PRE, POST, TX, TY>, dim3((n[0] + TX - 1) / TX, (n[1] + TY - 1) / TY, (n[2] + zchunk - 1) / zchunk), \
                   Sweep2Cfg<PRE, POST, TX, TY>::NT, a); return 0; }
#define G26(C, TX, TY) G2(0, 0, C, TX, TY) G2(0, 2, C, TX, TY) G2(0, 3, C, TX, TY) G2(1, 0, C, TX, TY) G2(1, 2, C, TX, TY) G2(1, 3, C, TX, TY)
    G26(0, 16, 8) G26(1, 32, 16)
    return 1;
}

// ---- k_sweep3 (vdn_mg_sweep3.cuh): one column of cell pairs per thread, register-pipelined operator data ----
#include "../../varden_b200/csrc/vdn_mg_sweep3.cuh"
extern "C" int emu_sweep3(int nsw, int pre, int post, int cfg, const int *n, const int *mode, int par0, const double *h2,
                          const double *rhs, const double *b0, const double *b1, const double *b2,
                          const double *in, double *out, const double *cphi, double *crhs, double *czero, double *nrm, int zchunk, int pad)
{
    if (nsw != 1) return 1;
    WaveArgs a;
    for (int d = 0; d < 3; ++d) { a.n[d] = n[d]; a.h2[d] = h2[d]; a.mode[d][0] = mode[2 * d]; a.mode[d][1] = mode[2 * d + 1]; }
    a.s1 = n[0] + 2 * pad; a.s2 = (long)(n[0] + 2 * pad) * (n[1] + 2 * pad); a.off = pad * (1 + a.s1 + a.s2); a.par0 = par0;
    a.rhs = rhs; a.b0 = b0; a.b1 = b1; a.b2 = b2; a.in = in; a.out = out;
    a.cphi = cphi; a.crhs = crhs; a.czero = czero;
    a.cs1 = n[0] / 2 + 2 * pad; a.cs2 = (long)(n[0] / 2 + 2 * pad) * (n[1] / 2 + 2 * pad); a.coff = pad * (1 + a.cs1 + a.cs2);
    a.nrm = nrm; a.zchunk = zchunk;
#define G3(PRE, POST, C, TX, TY) if (pre == PRE && post == POST && cfg == C) { \
        emu_launch(k_sweep3<PRE, POST, TX, TY>, dim3((n[0] + TX - 1) / TX, (n[1] + TY - 1) / TY, (n[2] + zchunk - 1) / zchunk), \
                   Sweep3Cfg<PRE, POST, TX, TY>::NT, a); return 0; }
#define G36(C, TX, TY) G3(0, 0, C, TX, TY) G3(0, 2, C, TX, TY) G3(0, 3, C, TX, TY) G3(1, 0, C, TX, TY) G3(1, 2, C, TX, TY) G3(1, 3, C, TX, TY)
    G36(0, 16, 8) G36(1, 32, 16)
    return 1;
}
