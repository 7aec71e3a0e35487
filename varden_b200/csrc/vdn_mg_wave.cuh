// vdn_mg_wave.cuh -- the fused wavefront multigrid smoother kernel (included by vdn_mg.cu; also compiled as plain C++
// by tests/emu/ where one OS thread plays each CUDA thread, so the kernel logic is checked on CPU-only machines).
#pragma once

enum : int { M_GHOST = 0, M_NEU = 1, M_DIR = 2, M_WRAP = 3 };

// ------------------------------------------------------------------------------------------
// Fused wavefront smoother (3-D, rank-local levels).
//
// One launch applies NSW full red-black Gauss-Seidel sweeps (S = 2*NSW colour stages) to the whole level while the
// grid streams through shared memory ONCE: a CTA owns an (TX x TY) column of cells plus a halo of H cells, marches
// along z and keeps a ring of NP phi planes in shared memory.  In iteration t stage s relaxes its colour on plane
// t-1-s (so every stage sees exactly the neighbour values the plain sweep order would give it), the plane that has
// passed all stages is written to the OUTPUT array (ping-pong: halos of neighbouring CTAs still read the input), and
//   PRE  = 1 : the load adds the piecewise-constant prolongation of the coarse correction (k_prolong fused),
//   POST = 2 : the residual of the finished plane is averaged 2x2x2 into the coarse right-hand side (k_residual +
//              k_restrict fused; the fine residual is never stored) and the coarse phi is zeroed,
//   POST = 3 : the inf-norm of the residual is reduced (the convergence test of the V-cycle, one atomic per CTA).
// Halo cells are relaxed redundantly (region of stage s = core grown by E+S-1-s); periodic directions wrap by index.
// HBM traffic per launch: phi in + out, rhs, 1/diag, 3 face-coefficient arrays  ~ 56 B/cell for S colour stages
// (+ residual), against 48 B/cell for EVERY colour stage of the plain kernels.
// ------------------------------------------------------------------------------------------
struct WaveArgs {
    int n[3]; long s1, s2, off;
    double h2[3]; int mode[3][2]; int par0;
    const double *rhs, *dgi, *b0, *b1, *b2;
    const double *in; double *out;
    const double *cphi; double *crhs, *czero; long cs1, cs2, coff;   // coarse level (PRE / POST == 2)
    double *nrm;
    int zchunk;
};

template <int H>
__device__ __forceinline__ int wave_wrap(int g, int n, bool wrap)
{
    if (g < 0) return (wrap && g >= -H) ? g + n : -1;
    if (g >= n) return (wrap && g < n + H) ? g - n : -1;
    return g;
}

// A*phi at one cell, phi taken from the shared-memory planes (same face formulas as cell_op)
__device__ __forceinline__ double wave_dir(double blo, double bhi, double h2, double p0, double pm, double pp,
                                           bool atlo, bool athi, int mlo, int mhi)
{
    double a = 0.0;
    if (atlo && mlo == M_NEU) { }
    else if (atlo && mlo == M_DIR) a += blo * (3.0 * p0 - pp * (1.0 / 3.0)) * h2;
    else a += blo * (p0 - pm) * h2;
    if (athi && mhi == M_NEU) { }
    else if (athi && mhi == M_DIR) a += bhi * (3.0 * p0 - pm * (1.0 / 3.0)) * h2;
    else a += bhi * (p0 - pp) * h2;
    return a;
}

template <int NSW, int PRE, int POST, int TX, int TY, int NT>
__global__ void __launch_bounds__(NT) k_wave(const WaveArgs a)
{
    constexpr int S = 2 * NSW, E = POST ? 1 : 0, H = S + E, W = TX + 2 * H, HH = TY + 2 * H, NP = S + 4;
    constexpr int PLANE = W * HH, NL = (PLANE + NT - 1) / NT;
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk, z1 = min(z0 + a.zchunk, a.n[2]);
    const int gx0 = x0 - H, gy0 = y0 - H;
    const bool wrx = a.mode[0][0] == M_WRAP, wry = a.mode[1][0] == M_WRAP, wrz = a.mode[2][0] == M_WRAP;
    const int n0 = a.n[0], n1 = a.n[1], n2 = a.n[2];

    // per-thread load slots of a plane: global offset (x,y part) or -1
    int lofs[NL], cofs[NL];
#pragma unroll
    for (int m = 0; m < NL; ++m) {
        const int q = tid + m * NT;
        lofs[m] = -1; cofs[m] = 0;
        if (q < PLANE) {
            const int ly = q / W, lx = q - ly * W;
            const int wx = wave_wrap<H>(gx0 + lx, n0, wrx), wy = wave_wrap<H>(gy0 + ly, n1, wry);
            if (wx >= 0 && wy >= 0) {
                lofs[m] = (int)(wx + a.s1 * wy);
                if (PRE) cofs[m] = (int)((wx >> 1) + a.cs1 * (wy >> 1));
            }
        }
    }
    auto slot = [&](int p) { return ((p + 16 * NP) % NP) * PLANE; };

    double pre[NL];
    double nmax = 0.0;
    constexpr int NQ = (TX / 2) * (TY / 2), NI = (NQ + NT - 1) / NT;
    double acc[NI];
#pragma unroll
    for (int m = 0; m < NI; ++m) acc[m] = 0.0;

    for (int t = z0 - H - 1; t <= z1 + S; ++t) {
        // ---- issue the loads of plane t+1 (consumed at the end of the iteration) ----
        const int pn = t + 1;
        const int wzn = (pn >= z0 - H && pn <= z1 - 1 + H) ? wave_wrap<H>(pn, n2, wrz) : -1;
#pragma unroll
        for (int m = 0; m < NL; ++m) {
            double v = 0.0;
            if (wzn >= 0 && lofs[m] >= 0) {
                v = a.in[a.off + lofs[m] + a.s2 * wzn];
                if (PRE) v += a.cphi[a.coff + cofs[m] + a.cs2 * (wzn >> 1)];
            }
            pre[m] = v;
        }
        // ---- colour stages ----
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int RS = E + S - 1 - s;                      // region = core grown by RS
            const int p = t - 1 - s;
            const int wz = (p >= z0 - RS && p <= z1 - 1 + RS) ? wave_wrap<H>(p, n2, wrz) : -1;
            if (wz >= 0) {
                const int Ws = TX + 2 * RS, Hs = TY + 2 * RS, hw = Ws / 2, ls = H - RS;
                const int color = s & 1;
                double *P0 = sm + slot(p);
                const double *PM = sm + slot(p - 1), *PP = sm + slot(p + 1);
                const bool zlo = p == 0, zhi = p == n2 - 1;
                for (int q = tid; q < hw * Hs; q += NT) {
                    const int yy = q / hw, xh = q - yy * hw;
                    const int ly = ls + yy, gy = gy0 + ly;
                    const int lx = ls + 2 * xh + ((color ^ (gx0 + ls + gy + p + a.par0)) & 1);
                    const int gx = gx0 + lx;
                    const int wx = wave_wrap<H>(gx, n0, wrx), wy = wave_wrap<H>(gy, n1, wry);
                    if (wx < 0 || wy < 0) continue;
                    const long c = a.off + wx + a.s1 * wy + a.s2 * wz;
                    const double dgi = __ldg(a.dgi + c);
                    const double rhs = __ldg(a.rhs + c);
                    const double bxl = __ldg(a.b0 + c), bxh = __ldg(a.b0 + c + 1);
                    const double byl = __ldg(a.b1 + c), byh = __ldg(a.b1 + c + a.s1);
                    const double bzl = __ldg(a.b2 + c), bzh = __ldg(a.b2 + c + a.s2);
                    const int id = ly * W + lx;
                    const double p0 = P0[id];
                    double ax = wave_dir(bxl, bxh, a.h2[0], p0, P0[id - 1], P0[id + 1], gx == 0, gx == n0 - 1, a.mode[0][0], a.mode[0][1]);
                    ax += wave_dir(byl, byh, a.h2[1], p0, P0[id - W], P0[id + W], gy == 0, gy == n1 - 1, a.mode[1][0], a.mode[1][1]);
                    ax += wave_dir(bzl, bzh, a.h2[2], p0, PM[id], PP[id], zlo, zhi, a.mode[2][0], a.mode[2][1]);
                    if (dgi != 0.0) P0[id] = p0 + (rhs - ax) * dgi;
                }
            }
            __syncthreads();
        }
        // ---- finished plane: write out (+ residual -> restriction / norm) ----
        {
            const int r = t - S - 1;
            if (r >= z0 && r < z1) {
                const double *P0 = sm + slot(r);
                const double *PM = sm + slot(r - 1), *PP = sm + slot(r + 1);
                const bool zlo = r == 0, zhi = r == n2 - 1;
                auto resid = [&](int lx, int ly, int gx, int gy, long c) {
                    const int id = ly * W + lx;
                    const double p0 = P0[id];
                    double ax = wave_dir(__ldg(a.b0 + c), __ldg(a.b0 + c + 1), a.h2[0], p0, P0[id - 1], P0[id + 1], gx == 0, gx == n0 - 1, a.mode[0][0], a.mode[0][1]);
                    ax += wave_dir(__ldg(a.b1 + c), __ldg(a.b1 + c + a.s1), a.h2[1], p0, P0[id - W], P0[id + W], gy == 0, gy == n1 - 1, a.mode[1][0], a.mode[1][1]);
                    ax += wave_dir(__ldg(a.b2 + c), __ldg(a.b2 + c + a.s2), a.h2[2], p0, PM[id], PP[id], zlo, zhi, a.mode[2][0], a.mode[2][1]);
                    return __ldg(a.rhs + c) - ax;
                };
                if (POST == 2) {
#pragma unroll
                    for (int m = 0; m < NI; ++m) {
                        const int q = tid + m * NT;
                        if (q < NQ) {
                            const int by = q / (TX / 2), bx = q - by * (TX / 2);
                            const int gx = x0 + 2 * bx, gy = y0 + 2 * by;
                            if (gx < n0 && gy < n1) {
                                const int lx = H + 2 * bx, ly = H + 2 * by;
                                const long c = a.off + gx + a.s1 * gy + a.s2 * r;
                                const int id = ly * W + lx;
                                a.out[c] = P0[id]; a.out[c + 1] = P0[id + 1];
                                a.out[c + a.s1] = P0[id + W]; a.out[c + a.s1 + 1] = P0[id + W + 1];
                                double s4 = resid(lx, ly, gx, gy, c) + resid(lx + 1, ly, gx + 1, gy, c + 1);
                                s4 += resid(lx, ly + 1, gx, gy + 1, c + a.s1);
                                s4 += resid(lx + 1, ly + 1, gx + 1, gy + 1, c + a.s1 + 1);
                                if ((r & 1) == 0) acc[m] = s4;
                                else {
                                    const long cc = a.coff + (gx >> 1) + a.cs1 * (gy >> 1) + a.cs2 * (r >> 1);
                                    a.crhs[cc] = (acc[m] + s4) * 0.125;
                                    a.czero[cc] = 0.0;
                                }
                            }
                        }
                    }
                } else {
                    for (int q = tid; q < TX * TY; q += NT) {
                        const int yy = q / TX, xx = q - yy * TX;
                        const int gx = x0 + xx, gy = y0 + yy;
                        if (gx < n0 && gy < n1) {
                            const long c = a.off + gx + a.s1 * gy + a.s2 * r;
                            a.out[c] = P0[(H + yy) * W + H + xx];
                            if (POST == 3) nmax = fmax(nmax, fabs(resid(H + xx, H + yy, gx, gy, c)));
                        }
                    }
                }
            }
        }
        // ---- park the prefetched plane ----
        {
            double *PN = sm + slot(pn);
#pragma unroll
            for (int m = 0; m < NL; ++m) { const int q = tid + m * NT; if (q < PLANE) PN[q] = pre[m]; }
        }
        __syncthreads();
    }
    if (POST == 3) block_atomic_max(nmax, a.nrm);
}

