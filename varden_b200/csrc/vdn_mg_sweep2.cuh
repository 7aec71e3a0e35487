// vdn_mg_sweep2.cuh -- k_sweep2: k_sweep (vdn_mg_sweep.cuh: 2x2 column blocks per thread, every colour stage of a step in
// program order, ONE barrier per plane) with the operator data staged through shared memory by cp.async one plane ahead.
//
// Why (ncu of k_sweep, profiles/r01_ncu_full_ksweep_v3.txt): 700 warp instructions per thread and plane, most of them 64-bit
// address arithmetic for 28 scalar global loads; the loads of a step sit on its critical path (9500 cycles per plane
// against 4000 at the HBM rate), and half of all shared-memory wavefronts were bank conflicts of the 16-byte-strided
// pair layout.  Here
//   * rhs and the three face-coefficient arrays of plane t+1 (b2: t+2) are copied global -> shared with cp.async (LDGSTS,
//     no registers, no wait) at the top of step t and are first read in step t+1, after the step's barrier;
//     every value is fetched from L2/HBM once per launch and then read twice (once by each colour) from shared memory;
//   * all shared-memory planes are colour-split: the even and the odd columns of a row are stored in two separate halves,
//     so the active cells of a warp (every other column) are contiguous 8-byte words -- no bank conflicts -- and phi,
//     rhs and the coefficients of a cell share ONE index;
//   * the residual stage (red cells of plane t-2) still loads its operator data directly: its ring slots are gone by then
//     and a fourth set would not fit (the lines are L2-resident, they were streamed two steps earlier).
// One GSRB sweep (S = 2 colour stages) per launch.  Shared memory: phi ring 4 (5 with the residual stage) planes,
// rhs/b0/b1 rings 3 planes each, b2 ring 4 planes: 17 (18) planes of (TX+2H) x (TY+2H) doubles.
#pragma once
#include "vdn_mg_sweep.cuh"

template <int PRE, int POST, int TX, int TY>
struct Sweep2Cfg {
    static constexpr int S = 2, E = POST ? 1 : 0, H = S + E, O = H & 1;
    static constexpr int RX = TX + 2 * H, RY = TY + 2 * H;      // loaded region = core grown by H
    static constexpr int RXH = RX / 2, HALF = RY * RXH, PLANE = RX * RY;
    static constexpr int BX = (RX + 2 * O) / 2, BY = (RY + 2 * O) / 2;   // 2x2 blocks aligned to even GLOBAL indices
    static constexpr int NPL = S + 2 + E, NC3 = 3, NC4 = 4;
    static constexpr size_t SMEM = sizeof(double) * PLANE * (NPL + 3 * NC3 + NC4);
    static constexpr int NT = ((BX * BY + 31) / 32) * 32;
    static_assert(TX % 2 == 0 && TY % 2 == 0, "even tiles");
};

template <int PRE, int POST, int TX, int TY>
__global__ void __launch_bounds__(Sweep2Cfg<PRE, POST, TX, TY>::NT, 1) k_sweep2(const WaveArgs a)
{
    using C = Sweep2Cfg<PRE, POST, TX, TY>;
    constexpr int S = C::S, E = C::E, H = C::H, O = C::O, RX = C::RX, RY = C::RY, RXH = C::RXH, HALF = C::HALF, PLANE = C::PLANE;
    constexpr int NPL = C::NPL, NC3 = C::NC3, NC4 = C::NC4;
    extern __shared__ double sm[];
    double *const sP = sm;                          // phi ring
    double *const sR = sP + NPL * PLANE;            // rhs ring
    double *const sX = sR + NC3 * PLANE;            // b0 (x faces)
    double *const sY = sX + NC3 * PLANE;            // b1 (y faces)
    double *const sZ = sY + NC3 * PLANE;            // b2 (z faces), NC4 planes
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk, z1 = min(z0 + a.zchunk, a.n[2]);
    const int n0 = a.n[0], n1 = a.n[1], n2 = a.n[2];
    const int mx0 = a.mode[0][0], mx1 = a.mode[0][1], my0 = a.mode[1][0], my1 = a.mode[1][1], mz0 = a.mode[2][0], mz1 = a.mode[2][1];
    const bool have = tid < C::BX * C::BY;
    const int by = have ? tid / C::BX : 0, bx = have ? tid - by * C::BX : 0;
    const int u0 = 2 * bx - O, v0 = 2 * by - O;                 // region coordinates of the block's (0,0) cell
    const int gx0 = x0 - H + u0, gy0 = y0 - H + v0;             // unwrapped global coordinates (even)

    // per cell (row r, column q): 0 = nothing, 1 = operator data only (face index n next to a physical boundary),
    // 2 + d = phi too, may run the stages s < d (d = distance to the edge of the loaded region)
    int lev[2][2];
    int wxq[2], wyq[2];
    bool inx[2], iny[2], phx[2], phy[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        inx[q] = have && u0 + q >= 0 && u0 + q < RX; iny[q] = have && v0 + q >= 0 && v0 + q < RY;
        wxq[q] = wave_idx_ld<H>(gx0 + q, n0, mx0, mx1); wyq[q] = wave_idx_ld<H>(gy0 + q, n1, my0, my1);
        phx[q] = wave_idx<H>(gx0 + q, n0, mx0, mx1) != WAVE_NONE; phy[q] = wave_idx<H>(gy0 + q, n1, my0, my1) != WAVE_NONE;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int u = u0 + q, v = v0 + r;
            const int d = min(min(u, RX - 1 - u), min(v, RY - 1 - v));
            lev[r][q] = !(inx[q] && iny[r] && wxq[q] != WAVE_NONE && wyq[r] != WAVE_NONE) ? 0 : (phx[q] && phy[r]) ? 2 + min(d, H) : 1;
        }
    const int wxb = wxq[0] != WAVE_NONE ? wxq[0] : wxq[1] - 1, wyb = wyq[0] != WAVE_NONE ? wyq[0] : wyq[1] - 1;
    const long gofs = a.off + wxb + a.s1 * (long)wyb;            // + s2 * plane
    const long cofs = PRE ? a.coff + (wxb >> 1) + a.cs1 * (long)(wyb >> 1) : 0;
    const bool anyc = (lev[0][0] | lev[0][1] | lev[1][0] | lev[1][1]) != 0;
    // colour-split index of cell (r, q): half (u & 1), row v, position u >> 1
    int si[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) si[r][q] = ((q ^ O) ? HALF : 0) + (v0 + r) * RXH + bx - ((O && q == 0) ? 1 : 0);
    bool bnd = false;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        bnd = bnd || (gx0 + q == 0 && (mx0 == M_NEU || mx0 == M_DIR)) || (gx0 + q == n0 - 1 && (mx1 == M_NEU || mx1 == M_DIR));
        bnd = bnd || (gy0 + q == 0 && (my0 == M_NEU || my0 == M_DIR)) || (gy0 + q == n1 - 1 && (my1 == M_NEU || my1 == M_DIR));
    }
    const bool core = have && u0 >= H && u0 < H + TX && v0 >= H && v0 < H + TY && gx0 < n0 && gy0 < n1;

    auto zld = [&](int p) { return (p >= z0 - H && p <= z1 - 1 + H) ? wave_idx_ld<H>(p, n2, mz0, mz1) : WAVE_NONE; };
    auto zph = [&](int p) { return (p >= z0 - H && p <= z1 - 1 + H) ? wave_idx<H>(p, n2, mz0, mz1) : WAVE_NONE; };
    auto ring = [](int p, int n) { return (p + 64 * n) % n; };

    // ---- operator planes: cp.async, one commit group per step.  w3 / w4: array plane index (or WAVE_NONE) of the rhs/b0/b1
    // plane and of the b2 plane, o3 / o4: their ring offsets ----
    auto request = [&](int w3, int o3, int w4, int o4) {
        if (anyc) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (lev[r][q]) {
                        const long g = gofs + a.s1 * r + q;
                        if (w3 != WAVE_NONE) {
                            const long g3 = g + a.s2 * (long)w3;
                            wave_cp8(sR + o3 + si[r][q], a.rhs + g3); wave_cp8(sX + o3 + si[r][q], a.b0 + g3); wave_cp8(sY + o3 + si[r][q], a.b1 + g3);
                        }
                        if (w4 != WAVE_NONE) wave_cp8(sZ + o4 + si[r][q], a.b2 + g + a.s2 * (long)w4);
                    }
        }
        wave_commit();
    };

    // ---- phi prefetch registers ----
    double pf[2][2], pc = 0.0;
    auto fetch = [&](int wz) {                   // wz: array plane index of a plane whose phi exists, or WAVE_NONE
#pragma unroll
        for (int r = 0; r < 2; ++r) { pf[r][0] = 0.0; pf[r][1] = 0.0; }
        pc = 0.0;
        if (wz != WAVE_NONE && anyc) {
            const double *src = a.in + gofs + a.s2 * (long)wz;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (lev[r][0] >= 2 && lev[r][1] >= 2) { const sweep_d2 v = *reinterpret_cast<const sweep_d2 *>(src + a.s1 * r); pf[r][0] = v.x; pf[r][1] = v.y; }
                else if (lev[r][0] >= 2) pf[r][0] = src[a.s1 * r];
                else if (lev[r][1] >= 2) pf[r][1] = src[a.s1 * r + 1];
            }
            if (PRE) pc = __ldg(a.cphi + cofs + a.cs2 * (long)(wz >> 1));
        }
    };
    auto stash = [&](int oP) {
        double *dst = sP + oP;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (inx[q] && iny[r]) dst[si[r][q]] = pf[r][q] + ((PRE && lev[r][q] >= 2) ? pc : 0.0);
    };

    // operands of tile cell `id` (colour-split index) of the plane at ring offset o0 (o0m / o0p: the planes below / above);
    // hf = the cell sits in the odd-column half; p = plane number (z-boundary flags)
    auto apply = [&](int p, int o0, int o0m, int o0p, int id, bool hf, int gx, int gy, bool general, const SweepCoef &c, double &ax, double &dg, double &p0) {
        const double *P0 = sP + o0 + id;
        const int xo = hf ? -HALF : HALF - 1;                   // the x-neighbours live in the other half
        p0 = P0[0];
        const double xm = P0[xo], xp = P0[xo + 1], ym = P0[-RXH], yp = P0[RXH], zm = sP[o0m + id], zp = sP[o0p + id];
        if (!general) {
            ax = (c.xl * (p0 - xm) + c.xh * (p0 - xp)) * a.h2[0] + (c.yl * (p0 - ym) + c.yh * (p0 - yp)) * a.h2[1]
               + (c.zl * (p0 - zm) + c.zh * (p0 - zp)) * a.h2[2];
            dg = (c.xl + c.xh) * a.h2[0] + (c.yl + c.yh) * a.h2[1] + (c.zl + c.zh) * a.h2[2];
        } else {
            ax = 0.0; dg = 0.0;
            wave_dir(c.xl, c.xh, a.h2[0], p0, xm, xp, gx == 0, gx == n0 - 1, mx0, mx1, ax, dg);
            wave_dir(c.yl, c.yh, a.h2[1], p0, ym, yp, gy == 0, gy == n1 - 1, my0, my1, ax, dg);
            wave_dir(c.zl, c.zh, a.h2[2], p0, zm, zp, p == 0, p == n2 - 1, mz0, mz1, ax, dg);
        }
    };
    auto staged = [&](int o3, int o4, int o4p, int id, bool hf, SweepCoef &c) {  // operator data from the rings (o4p: b2 of the plane above)
        const int xo = hf ? -HALF : HALF - 1;
        c.rhs = sR[o3 + id];
        c.xl = sX[o3 + id]; c.xh = sX[o3 + id + xo + 1];
        c.yl = sY[o3 + id]; c.yh = sY[o3 + id + RXH];
        c.zl = sZ[o4 + id]; c.zh = sZ[o4p + id];
    };

    double nmax = 0.0, acc = 0.0;
    const int tfirst = z0 - H, tlast = z1 + S - 2 + E;
    // Ring offsets as rotating registers (no modulo in the loop): at step t, oP[k] holds plane t+1-k of the phi ring,
    // o3[k] plane t+1-k of the rhs/b0/b1 rings, o4[k] plane t+2-k of the b2 ring; ex[k] = phi of plane t+2-k exists.
    int oP[NPL], o3[NC3], o4[NC4];
    bool ex[S + 3 + E];
#pragma unroll
    for (int k = 0; k < NPL; ++k) oP[k] = ring(tfirst - k, NPL) * PLANE;        // state "step tfirst-1"; rotated at the top of the loop
#pragma unroll
    for (int k = 0; k < NC3; ++k) o3[k] = ring(tfirst - k, NC3) * PLANE;
#pragma unroll
    for (int k = 0; k < NC4; ++k) o4[k] = ring(tfirst + 1 - k, NC4) * PLANE;
#pragma unroll
    for (int k = 0; k < S + 3 + E; ++k) ex[k] = zph(tfirst + 1 - k) != WAVE_NONE;
    // prologue: operator planes of the first step, phi of the first two planes
    request(WAVE_NONE, 0, zld(tfirst), o4[1]);
    request(zld(tfirst), o3[0], zld(tfirst + 1), o4[0]);
    fetch(zph(tfirst)); stash(oP[0]); fetch(zph(tfirst + 1));
    wave_wait<0>();
    __syncthreads();
    int wl1 = zld(tfirst + 1);                                   // array plane index (operator data) of plane t+1 at the top of step t

    for (int t = tfirst; t <= tlast; ++t) {
        const int sel = (t + a.par0) & 1;                        // active column of row 0 (row 1: the other one)
        // rotate the rings: the slot of the oldest plane receives the newest one
        { const int x = oP[NPL - 1];
#pragma unroll
          for (int k = NPL - 1; k > 0; --k) oP[k] = oP[k - 1];
          oP[0] = x; }
        { const int x = o3[NC3 - 1];
#pragma unroll
          for (int k = NC3 - 1; k > 0; --k) o3[k] = o3[k - 1];
          o3[0] = x; }
        { const int x = o4[NC4 - 1];
#pragma unroll
          for (int k = NC4 - 1; k > 0; --k) o4[k] = o4[k - 1];
          o4[0] = x; }
        const int wl2 = zld(t + 2), wp2 = zph(t + 2);
#pragma unroll
        for (int k = S + 2 + E; k > 0; --k) ex[k] = ex[k - 1];
        ex[0] = wp2 != WAVE_NONE;
        request(wl1, o3[0], wl2, o4[0]);                         // lands during this step, first read after the barrier
        wl1 = wl2;
        stash(oP[0]);                                            // plane t+1 (fetched during the previous step)
        fetch(wp2);                                              // plane t+2
        // ---- colour stages: stage s relaxes the active columns of plane t-s ----
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int p = t - s, RS = E + S - 1 - s;
            const bool zok = p >= z0 - RS && p <= z1 - 1 + RS && ex[s + 2];
            const bool zb = (p == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (p == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int q = sel ^ r;
                const int lv = q ? lev[r][1] : lev[r][0];
                if (zok && lv > s + 2) {
                    const int id = q ? si[r][1] : si[r][0];
                    const bool hf = (q ^ O) != 0;
                    SweepCoef c; staged(o3[s + 1], o4[s + 2], o4[s + 1], id, hf, c);
                    double ax, dg, p0;
                    apply(p, oP[s + 1], oP[s + 2], oP[s], id, hf, gx0 + q, gy0 + r, bnd || zb, c, ax, dg, p0);
                    if (dg != 0.0) sP[oP[s + 1] + id] = p0 + (c.rhs - ax) / dg;
                }
            }
        }
        // ---- plane t-S+1 has passed every stage: write it out ----
        {
            const int r1 = t - S + 1;
            if (core && r1 >= z0 && r1 < z1) {
                const double *P0 = sP + oP[S];
                double *dst = a.out + gofs + a.s2 * (long)r1;
#pragma unroll
                for (int r = 0; r < 2; ++r) { sweep_d2 v; v.x = P0[si[r][0]]; v.y = P0[si[r][1]]; *reinterpret_cast<sweep_d2 *>(dst + a.s1 * r) = v; }
            }
        }
        // ---- residual of plane t-S: red cells only (the black ones were just relaxed), operator data straight from L2 ----
        if (POST) {
            const int r0 = t - S;
            if (core && r0 >= z0 && r0 < z1) {
                const bool zb = (r0 == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (r0 == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
                double s2 = 0.0;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int q = sel ^ r, id = q ? si[r][1] : si[r][0];
                    SweepCoef c; sweep_load(c, a, gofs + a.s1 * r + q + a.s2 * (long)r0);
                    double ax, dg, p0;
                    apply(r0, oP[S + 1], oP[(S + 2) % NPL], oP[S], id, (q ^ O) != 0, gx0 + q, gy0 + r, bnd || zb, c, ax, dg, p0);
                    const double res = c.rhs - ax;
                    s2 += res;
                    if (POST == 3) nmax = fmax(nmax, fabs(res));
                }
                if (POST == 2) {
                    if ((r0 & 1) == 0) acc = s2;
                    else {
                        const long cc = a.coff + (gx0 >> 1) + a.cs1 * (long)(gy0 >> 1) + a.cs2 * (long)(r0 >> 1);
                        a.crhs[cc] = (acc + s2) * 0.125;
                        a.czero[cc] = 0.0;
                    }
                }
            }
        }
        wave_wait<0>();
        __syncthreads();
    }
    if (POST == 3) block_atomic_max(nmax, a.nrm);
}
