// vdn_mg_wave.cuh -- the fused wavefront multigrid smoother kernel (included by vdn_mg.cu; also compiled as plain C++
// by tests/emu/ where one OS thread plays each CUDA thread, so the kernel logic is checked on CPU-only machines).
#pragma once

enum : int { M_GHOST = 0, M_NEU = 1, M_DIR = 2, M_WRAP = 3 };

// ------------------------------------------------------------------------------------------
// Fused wavefront smoother (3-D, rank-local levels).
//
// One launch applies NSW full red-black Gauss-Seidel sweeps (S = 2*NSW colour stages) to the whole level while the
// grid streams through shared memory ONCE.  A CTA owns a (TX x TY) column of cells plus a halo of H cells and marches
// along z.  Rings of planes of phi AND of the operator data (rhs, the three face-coefficient arrays) live in shared
// memory, filled PF planes ahead with cp.async (LDGSTS), so every colour stage reads only shared memory:
//   iteration t:  plane t has landed | stage s relaxes its colour on plane t-1-s (each stage therefore sees exactly
//   the neighbour values the plain sweep order gives it) | the plane that passed all stages, t-S-1, is written to the
//   OUTPUT array (ping-pong: halos of neighbouring CTAs still read the input) | plane t+PF is requested.
//   PRE  = 1 : the prolongation of the coarse correction is added as a plane lands (k_prolong fused),
//   POST = 2 : the residual of the finished plane is averaged 2x2x2 into the coarse right-hand side (k_residual +
//              k_restrict fused; the fine residual is never stored) and the coarse phi is zeroed,
//   POST = 3 : the inf-norm of the residual is reduced (the convergence test of the V-cycle, one atomic per CTA).
// Halo cells are relaxed redundantly (region of stage s = core grown by E+S-1-s); periodic directions wrap by index.
// HBM traffic per launch: phi in + out, rhs, 3 face-coefficient arrays = 48 B/cell for S colour stages + residual +
// transfer operator, against 48 B/cell for EVERY colour stage (and again for the residual) of the plain kernels.
// ------------------------------------------------------------------------------------------
struct WaveArgs {
    int n[3]; long s1, s2, off;
    double h2[3]; int mode[3][2]; int par0;
    const double *rhs, *b0, *b1, *b2;
    const double *in; double *out;
    const double *cphi; double *crhs, *czero; long cs1, cs2, coff;   // coarse level (PRE / POST == 2)
    double *nrm;
    int zchunk;
};

#ifdef VDN_EMU
__device__ __forceinline__ void wave_cp8(double *dst, const double *src) { *dst = *src; }
__device__ __forceinline__ void wave_commit() { }
template <int N> __device__ __forceinline__ void wave_wait() { }
#else
__device__ __forceinline__ void wave_cp8(double *dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void wave_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void wave_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
#endif

// index of a tile cell in the level arrays, or WAVE_NONE.  Cells that take part in the relaxation: the level's own cells,
// periodic images (M_WRAP, wrap by index) and neighbour-rank cells held in the level's ghost layers (M_GHOST) ...
constexpr int WAVE_NONE = -(1 << 28);
template <int H>
__device__ __forceinline__ int wave_idx(int g, int n, int mlo, int mhi)
{
    if (g < 0) { if (g < -H) return WAVE_NONE; return mlo == M_WRAP ? g + n : (mlo == M_GHOST ? g : WAVE_NONE); }
    if (g >= n) { if (g >= n + H) return WAVE_NONE; return mhi == M_WRAP ? g - n : (mhi == M_GHOST ? g : WAVE_NONE); }
    return g;
}
// ... and cells whose data is LOADED: additionally index n next to a physical boundary, which holds the coefficient of the
// high boundary face (the padded level layout stores face n at cell index n)
template <int H>
__device__ __forceinline__ int wave_idx_ld(int g, int n, int mlo, int mhi)
{
    if (g == n && mhi != M_WRAP && mhi != M_GHOST) return n;
    return wave_idx<H>(g, n, mlo, mhi);
}

// A*phi contribution and diagonal of one direction (same face formulas as cell_op in vdn_mg.cu)
__device__ __forceinline__ void wave_dir(double blo, double bhi, double h2, double p0, double pm, double pp,
                                         bool atlo, bool athi, int mlo, int mhi, double &a, double &g)
{
    if (atlo && mlo == M_NEU) { }
    else if (atlo && mlo == M_DIR) { a += blo * (3.0 * p0 - pp * (1.0 / 3.0)) * h2; g += 3.0 * blo * h2; }
    else { a += blo * (p0 - pm) * h2; g += blo * h2; }
    if (athi && mhi == M_NEU) { }
    else if (athi && mhi == M_DIR) { a += bhi * (3.0 * p0 - pm * (1.0 / 3.0)) * h2; g += 3.0 * bhi * h2; }
    else { a += bhi * (p0 - pp) * h2; g += bhi * h2; }
}

template <int NSW, int PRE, int POST, int TX, int TY, int NT, int PF>
struct WaveCfg {
    static constexpr int S = 2 * NSW, E = POST ? 1 : 0, H = S + E, W = TX + 2 * H, HH = TY + 2 * H;
    static constexpr int PLANE = W * HH;
    static constexpr int NPP = PF + S + 3;          // phi ring: planes t-S-2 .. t+PF
    static constexpr int NPC = PF + S + 2;          // operator rings: planes t-S-1 .. t+PF
    static constexpr int NL = (PLANE + NT - 1) / NT;
    static constexpr size_t SMEM = sizeof(double) * PLANE * (NPP + 4 * NPC);
};

template <int NSW, int PRE, int POST, int TX, int TY, int NT, int PF>
__global__ void __launch_bounds__(NT) k_wave(const WaveArgs a)
{
    using C = WaveCfg<NSW, PRE, POST, TX, TY, NT, PF>;
    constexpr int S = C::S, E = C::E, H = C::H, W = C::W, PLANE = C::PLANE, NPP = C::NPP, NPC = C::NPC, NL = C::NL;
    extern __shared__ double sm[];
    double *const sP = sm;                          // phi ring
    double *const sR = sP + NPP * PLANE;            // rhs ring
    double *const sX = sR + NPC * PLANE;            // beta_x, beta_y, beta_z rings
    double *const sY = sX + NPC * PLANE;
    double *const sZ = sY + NPC * PLANE;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk, z1 = min(z0 + a.zchunk, a.n[2]);
    const int gx0 = x0 - H, gy0 = y0 - H;
    const int mx0 = a.mode[0][0], mx1 = a.mode[0][1], my0 = a.mode[1][0], my1 = a.mode[1][1], mz0 = a.mode[2][0], mz1 = a.mode[2][1];
    const int n0 = a.n[0], n1 = a.n[1], n2 = a.n[2];

    // per-thread load slots of a plane: global offset of the (x,y) part, or -1
    int lofs[NL], cofs[NL];
#pragma unroll
    for (int m = 0; m < NL; ++m) {
        const int q = tid + m * NT;
        lofs[m] = WAVE_NONE; cofs[m] = WAVE_NONE;
        if (q < PLANE) {
            const int ly = q / W, lx = q - ly * W;
            const int wx = wave_idx_ld<H>(gx0 + lx, n0, mx0, mx1), wy = wave_idx_ld<H>(gy0 + ly, n1, my0, my1);
            if (wx != WAVE_NONE && wy != WAVE_NONE) {
                lofs[m] = (int)(wx + a.s1 * wy);
                // coarse cell under a fine cell (ghost cells: floor division, the coarse level carries ghost layers too)
                if (PRE && wave_idx<H>(gx0 + lx, n0, mx0, mx1) != WAVE_NONE && wave_idx<H>(gy0 + ly, n1, my0, my1) != WAVE_NONE)
                    cofs[m] = (int)((wx >> 1) + a.cs1 * (wy >> 1));
            }
        }
    }
    auto pslot = [&](int p) { return ((p + 16 * NPP) % NPP) * PLANE; };
    auto cslot = [&](int p) { return ((p + 16 * NPC) % NPC) * PLANE; };
    // request plane p (phi + operator data) with cp.async; always commits one group
    auto request = [&](int p) {
        const int wz = (p >= z0 - H && p <= z1 - 1 + H) ? wave_idx_ld<H>(p, n2, mz0, mz1) : WAVE_NONE;
        if (wz != WAVE_NONE) {
            const int ps = pslot(p), cs = cslot(p);
#pragma unroll
            for (int m = 0; m < NL; ++m)
                if (lofs[m] != WAVE_NONE) {
                    const int q = tid + m * NT;
                    const long c = a.off + lofs[m] + a.s2 * wz;
                    wave_cp8(sP + ps + q, a.in + c);
                    wave_cp8(sR + cs + q, a.rhs + c);
                    wave_cp8(sX + cs + q, a.b0 + c);
                    wave_cp8(sY + cs + q, a.b1 + c);
                    wave_cp8(sZ + cs + q, a.b2 + c);
                }
        }
        wave_commit();
    };

    double nmax = 0.0;
    constexpr int NQ = (TX / 2) * (TY / 2), NI = (NQ + NT - 1) / NT;
    double acc[NI];
#pragma unroll
    for (int m = 0; m < NI; ++m) acc[m] = 0.0;

    // relaxation / residual operands of the tile cell (lx, ly) of plane p
    auto apply = [&](int p, int lx, int ly, int gx, int gy, double &ax, double &dg, double &p0, double &rhs) {
        const int id = ly * W + lx;
        const double *P0 = sP + pslot(p), *PM = sP + pslot(p - 1), *PP = sP + pslot(p + 1);
        const int c0 = cslot(p), c1 = cslot(p + 1);
        p0 = P0[id]; rhs = sR[c0 + id];
        ax = 0.0; dg = 0.0;
        wave_dir(sX[c0 + id], sX[c0 + id + 1], a.h2[0], p0, P0[id - 1], P0[id + 1], gx == 0, gx == n0 - 1, a.mode[0][0], a.mode[0][1], ax, dg);
        wave_dir(sY[c0 + id], sY[c0 + id + W], a.h2[1], p0, P0[id - W], P0[id + W], gy == 0, gy == n1 - 1, a.mode[1][0], a.mode[1][1], ax, dg);
        wave_dir(sZ[c0 + id], sZ[c1 + id], a.h2[2], p0, PM[id], PP[id], p == 0, p == n2 - 1, a.mode[2][0], a.mode[2][1], ax, dg);
    };

    const int tfirst = z0 - H;
#pragma unroll
    for (int q = 0; q < PF; ++q) request(tfirst + q);

    for (int t = tfirst; t <= z1 + S; ++t) {
        // ---- plane t has landed (at most PF-1 younger requests may still be in flight) ----
        wave_wait<PF - 1>();
        if (PRE) {
            const int wz = (t >= z0 - H && t <= z1 - 1 + H) ? wave_idx<H>(t, n2, mz0, mz1) : WAVE_NONE;
            if (wz != WAVE_NONE) {
                const int ps = pslot(t);
#pragma unroll
                for (int m = 0; m < NL; ++m)
                    if (cofs[m] != WAVE_NONE) sP[ps + tid + m * NT] += __ldg(a.cphi + a.coff + cofs[m] + a.cs2 * (wz >> 1));
            }
        }
        __syncthreads();
        request(t + PF);
        // ---- colour stages ----
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int RS = E + S - 1 - s;                      // region = core grown by RS
            const int p = t - 1 - s;
            const int wz = (p >= z0 - RS && p <= z1 - 1 + RS) ? wave_idx<H>(p, n2, mz0, mz1) : WAVE_NONE;
            if (wz != WAVE_NONE) {
                const int Ws = TX + 2 * RS, Hs = TY + 2 * RS, hw = Ws / 2, ls = H - RS;
                const int color = s & 1;
                double *P0 = sP + pslot(p);
                for (int q = tid; q < hw * Hs; q += NT) {
                    const int yy = q / hw, xh = q - yy * hw;
                    const int ly = ls + yy, gy = gy0 + ly;
                    const int lx = ls + 2 * xh + ((color ^ (gx0 + ls + gy + p + a.par0)) & 1);
                    const int gx = gx0 + lx;
                    if (wave_idx<H>(gx, n0, mx0, mx1) == WAVE_NONE || wave_idx<H>(gy, n1, my0, my1) == WAVE_NONE) continue;
                    double ax, dg, p0, rhs;
                    apply(p, lx, ly, gx, gy, ax, dg, p0, rhs);
                    if (dg != 0.0) P0[ly * W + lx] = p0 + (rhs - ax) / dg;
                }
            }
            __syncthreads();
        }
        // ---- finished plane: write out (+ residual -> restriction / norm) ----
        {
            const int r = t - S - 1;
            if (r >= z0 && r < z1) {
                const double *P0 = sP + pslot(r);
                auto resid = [&](int lx, int ly, int gx, int gy) {
                    double ax, dg, p0, rhs;
                    apply(r, lx, ly, gx, gy, ax, dg, p0, rhs);
                    return rhs - ax;
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
                                double s4 = resid(lx, ly, gx, gy) + resid(lx + 1, ly, gx + 1, gy);
                                s4 += resid(lx, ly + 1, gx, gy + 1);
                                s4 += resid(lx + 1, ly + 1, gx + 1, gy + 1);
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
                            a.out[a.off + gx + a.s1 * gy + a.s2 * r] = P0[(H + yy) * W + H + xx];
                            if (POST == 3) nmax = fmax(nmax, fabs(resid(H + xx, H + yy, gx, gy)));
                        }
                    }
                }
            }
        }
        // no barrier here: the next request (issued after the next iteration's barrier) is the first writer of a slot
        // this iteration still reads
    }
    wave_wait<0>();
    if (POST == 3) block_atomic_max(nmax, a.nrm);
}
