// vdn_mg_sweep.cuh -- k_sweep: the fused red-black multigrid smoother, second generation (included by vdn_mg.cu; also
// compiled as plain C++ by tests/emu/ so the kernel logic runs in the CPU-only test tier).
//
// Same contract as k_wave (vdn_mg_wave.cuh): one launch applies S = 2*NSW colour stages of Gauss-Seidel to a whole level
// while phi streams through shared memory once (input array -> output array, ping-pong), with
//   PRE  = 1 : prolongation of the coarse correction added as a plane lands,
//   POST = 2 : residual of the finished plane averaged 2x2x2 into the coarse right-hand side, coarse phi zeroed,
//   POST = 3 : inf-norm of the residual reduced (one atomic per CTA).
// What changed, from the ncu capture of k_wave (profiles/r01_ncu_full_wave_v2.txt: 128 CTAs, 16 warps/SM, 3 barriers per
// plane, 31 planes of shared memory, issue slots 23-37 % busy):
//   * a thread owns a 2x2 block of cell columns for the whole march.  Of its four cells in a plane two are red and two
//     black, and (the block origin is even) WHICH column of each row is active depends only on (t + par0) & 1.  In step t
//     the thread runs ALL stages on its active columns: stage s relaxes plane t-s.  The z-neighbours a stage needs were
//     written by this same thread (program order), the x/y-neighbours were written by other threads one step earlier,
//     so ONE __syncthreads per plane is enough and no thread idles on the "wrong colour";
//   * only phi lives in shared memory (S+2 or S+3 planes).  rhs and the face coefficients are read straight from global
//     memory where they are used -- all loads of a step are issued before the first use, the lines of the next step are
//     requested with prefetch.global.L2 one step ahead -- so a CTA needs ~80-110 KB instead of 207 KB;
//   * after the last (black) stage the residual of the black cells is zero up to round-off (a cell that was just relaxed
//     satisfies its equation exactly when its neighbours do not change any more), so the residual stage only visits the
//     red cells: it is one more pass over the active columns, S planes behind the front;
//   * phi moves as 16-byte pairs (MG_PAD = 4 keeps every row of an even-sized level 32-byte aligned).
// HBM traffic per launch: phi in + out, rhs, 3 face-coefficient arrays = 48 B/cell (x tile halo) for S colour stages +
// residual + transfer operator, against 48 B/cell for EVERY colour stage of the plain kernels.
#pragma once
#include "vdn_mg_wave.cuh"

#ifdef VDN_EMU
struct sweep_d2 { double x, y; };
__device__ __forceinline__ void sweep_pf(const double *) { }
#else
typedef double2 sweep_d2;
__device__ __forceinline__ void sweep_pf(const double *p) { asm volatile("prefetch.global.L2 [%0];\n" :: "l"(p)); }
#endif

template <int NSW, int PRE, int POST, int TX, int TY, int MINB = 1>
struct SweepCfg {
    static constexpr int S = 2 * NSW, E = POST ? 1 : 0, H = S + E, HE = (H + 1) & ~1;
    static constexpr int X = TX + 2 * HE, Y = TY + 2 * HE;      // shared-memory tile: 2x2 blocks aligned to even global indices
    static constexpr int BX = X / 2, BY = Y / 2;
    static constexpr int PLANE = X * Y;
    static constexpr int NPL = S + 2 + E;                        // phi ring: planes t-S-E .. t+1
    static constexpr size_t SMEM = sizeof(double) * PLANE * NPL;
    static constexpr int NT = ((BX * BY + 31) / 32) * 32;       // one thread per 2x2 block of columns
    static_assert(TX % 2 == 0 && TY % 2 == 0, "even tiles");
};

// relaxation operands of one cell: phi neighbours from the shared-memory ring, operator data in registers
struct SweepCoef { double rhs, xl, xh, yl, yh, zl, zh; };

__device__ __forceinline__ void sweep_load(SweepCoef &c, const WaveArgs &a, long g)
{
    c.rhs = __ldg(a.rhs + g);
    c.xl = __ldg(a.b0 + g); c.xh = __ldg(a.b0 + g + 1);
    c.yl = __ldg(a.b1 + g); c.yh = __ldg(a.b1 + g + a.s1);
    c.zl = __ldg(a.b2 + g); c.zh = __ldg(a.b2 + g + a.s2);
}
__device__ __forceinline__ void sweep_prefetch(const WaveArgs &a, long g)
{
    sweep_pf(a.rhs + g); sweep_pf(a.b0 + g); sweep_pf(a.b1 + g); sweep_pf(a.b2 + g + a.s2);     // b2 of plane p came with plane p-1
}

template <int NSW, int PRE, int POST, int TX, int TY, int MINB = 1>
__global__ void __launch_bounds__(SweepCfg<NSW, PRE, POST, TX, TY>::NT, MINB) k_sweep(const WaveArgs a)
{
    using C = SweepCfg<NSW, PRE, POST, TX, TY>;
    constexpr int S = C::S, E = C::E, H = C::H, HE = C::HE, X = C::X, PLANE = C::PLANE, NPL = C::NPL;
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk, z1 = min(z0 + a.zchunk, a.n[2]);
    const int n0 = a.n[0], n1 = a.n[1], n2 = a.n[2];
    const int mx0 = a.mode[0][0], mx1 = a.mode[0][1], my0 = a.mode[1][0], my1 = a.mode[1][1], mz0 = a.mode[2][0], mz1 = a.mode[2][1];
    const bool have = tid < C::BX * C::BY;
    const int by = have ? tid / C::BX : 0, bx = have ? tid - by * C::BX : 0;
    const int lx0 = 2 * bx, ly0 = 2 * by;                       // tile coordinates of the block's (0,0) cell
    const int gx0 = x0 - HE + lx0, gy0 = y0 - HE + ly0;         // unwrapped global coordinates (even)
    const int sid = ly0 * X + lx0;                              // shared-memory index of the block's (0,0) cell

    // per cell (row r, column c) of the block: number of stages it may run (0: not loaded at all).  A cell takes part in
    // stage s when it is at least s+1 cells inside the loaded region (core grown by H).
    int wx[2], wy[2], depth[2][2];
    bool ldx[2], ldy[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ux = lx0 + q - (HE - H), uy = ly0 + q - (HE - H);       // position inside the H-grown region
        wx[q] = wave_idx<H>(gx0 + q, n0, mx0, mx1); wy[q] = wave_idx<H>(gy0 + q, n1, my0, my1);
        ldx[q] = have && wx[q] != WAVE_NONE && ux >= 0 && ux < TX + 2 * H;
        ldy[q] = have && wy[q] != WAVE_NONE && uy >= 0 && uy < TY + 2 * H;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int ux = lx0 + q - (HE - H), uy = ly0 + r - (HE - H);
            const int d = min(min(ux, TX + 2 * H - 1 - ux), min(uy, TY + 2 * H - 1 - uy));
            depth[r][q] = (ldx[q] && ldy[r]) ? min(max(d, 0), H) : 0;
        }
    // both columns / rows of a block exist together except on the masked outer ring (HE > H)
    const int wxb = ldx[0] ? wx[0] : wx[1] - 1, wyb = ldy[0] ? wy[0] : wy[1] - 1;
    const long gofs = a.off + wxb + a.s1 * (long)wyb;            // + s2 * plane
    const long cofs = PRE ? a.coff + (wxb >> 1) + a.cs1 * (long)(wyb >> 1) : 0;
    const bool anyld = (ldx[0] || ldx[1]) && (ldy[0] || ldy[1]);
    // does any cell of the block touch a physical boundary in x or y (Neumann / Dirichlet face formulas)?
    bool bnd = false;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        bnd = bnd || (gx0 + q == 0 && (mx0 == M_NEU || mx0 == M_DIR)) || (gx0 + q == n0 - 1 && (mx1 == M_NEU || mx1 == M_DIR));
        bnd = bnd || (gy0 + q == 0 && (my0 == M_NEU || my0 == M_DIR)) || (gy0 + q == n1 - 1 && (my1 == M_NEU || my1 == M_DIR));
    }
    const bool core = have && lx0 >= HE && lx0 < HE + TX && ly0 >= HE && ly0 < HE + TY && gx0 < n0 && gy0 < n1;

    auto zidx = [&](int p) { return (p >= z0 - H && p <= z1 - 1 + H) ? wave_idx<H>(p, n2, mz0, mz1) : WAVE_NONE; };
    auto slot = [&](int p) { return ((p + 64 * NPL) % NPL) * PLANE; };

    // ---- phi prefetch registers: the block's four cells of one plane (+ the coarse correction under them) ----
    double pf[2][2], pc = 0.0;
    auto fetch = [&](int p) {
        const int wz = zidx(p);
#pragma unroll
        for (int r = 0; r < 2; ++r) { pf[r][0] = 0.0; pf[r][1] = 0.0; }
        pc = 0.0;
        if (wz != WAVE_NONE && anyld) {
            const double *src = a.in + gofs + a.s2 * (long)wz;
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (ldy[r]) {
                    if (ldx[0] && ldx[1]) { const sweep_d2 v = *reinterpret_cast<const sweep_d2 *>(src + a.s1 * r); pf[r][0] = v.x; pf[r][1] = v.y; }
                    else if (ldx[0]) pf[r][0] = src[a.s1 * r];
                    else if (ldx[1]) pf[r][1] = src[a.s1 * r + 1];
                }
            if (PRE) pc = __ldg(a.cphi + cofs + a.cs2 * (long)(wz >> 1));
        }
    };
    auto stash = [&](int p) {                                    // prefetched plane p -> its ring slot
        if (!have) return;
        double *dst = sm + slot(p) + sid;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            sweep_d2 v; v.x = pf[r][0] + (PRE && ldx[0] && ldy[r] ? pc : 0.0); v.y = pf[r][1] + (PRE && ldx[1] && ldy[r] ? pc : 0.0);
            *reinterpret_cast<sweep_d2 *>(dst + X * r) = v;
        }
    };

    // A*phi and the diagonal at tile cell `id` of plane p (z-boundary flags from p, x/y flags from the unwrapped indices)
    auto apply = [&](int p, int id, int gx, int gy, const SweepCoef &c, bool general, double &ax, double &dg, double &p0) {
        const double *P0 = sm + slot(p), *PM = sm + slot(p - 1), *PP = sm + slot(p + 1);
        p0 = P0[id];
        const double xm = P0[id - 1], xp = P0[id + 1], ym = P0[id - X], yp = P0[id + X], zm = PM[id], zp = PP[id];
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

    double nmax = 0.0, acc = 0.0;
    const int tfirst = z0 - H, tlast = z1 + S - 2 + E;
    fetch(tfirst); stash(tfirst); fetch(tfirst + 1);
    __syncthreads();

    for (int t = tfirst; t <= tlast; ++t) {
        const int sel = (t + a.par0) & 1;                        // active column of row 0 (row 1: the other one)
        // ---- plane t+1 into the ring, plane t+2 on its way, next step's operator lines towards L2 ----
        stash(t + 1);
        fetch(t + 2);
#pragma unroll
        for (int s = 0; s < S; s += 2) {                         // plane t+1-s is new to stage s in the next step
            const int p = t + 1 - s, wz = zidx(p);
            if (wz != WAVE_NONE && anyld && (bx & 3) == 0)       // a block row covers 16 B: every fourth block asks for its 64 B
#pragma unroll
                for (int r = 0; r < 2; ++r) sweep_prefetch(a, gofs + a.s1 * r + a.s2 * (long)wz);
        }
        // ---- colour stages: stage s relaxes the active columns of plane t-s (operator data straight from global / L2) ----
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int p = t - s, RS = E + S - 1 - s;
            const int wz = (p >= z0 - RS && p <= z1 - 1 + RS) ? zidx(p) : WAVE_NONE;
            const bool zb = (p == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (p == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
            SweepCoef cf[2];
            bool run[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int q = sel ^ r;
                run[r] = wz != WAVE_NONE && depth[r][q] > s;
                if (run[r]) sweep_load(cf[r], a, gofs + a.s1 * r + q + a.s2 * (long)wz);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (run[r]) {
                    const int q = sel ^ r, id = sid + X * r + q;
                    double ax, dg, p0;
                    apply(p, id, gx0 + q, gy0 + r, cf[r], bnd || zb, ax, dg, p0);
                    if (dg != 0.0) sm[slot(p) + id] = p0 + (cf[r].rhs - ax) / dg;
                }
        }
        // ---- plane t-S+1 has passed every stage: write it out ----
        {
            const int r1 = t - S + 1;
            if (core && r1 >= z0 && r1 < z1) {
                const double *P0 = sm + slot(r1) + sid;
                double *dst = a.out + gofs + a.s2 * (long)r1;
#pragma unroll
                for (int r = 0; r < 2; ++r) *reinterpret_cast<sweep_d2 *>(dst + a.s1 * r) = *reinterpret_cast<const sweep_d2 *>(P0 + X * r);
            }
        }
        // ---- residual of plane t-S (its neighbours are final): red cells only, the black ones were just relaxed ----
        if (POST) {
            const int r0 = t - S;
            if (core && r0 >= z0 && r0 < z1) {
                const bool zb = (r0 == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (r0 == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
                double s2 = 0.0;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int q = sel ^ r, id = sid + X * r + q;
                    SweepCoef c; sweep_load(c, a, gofs + a.s1 * r + q + a.s2 * (long)r0);
                    double ax, dg, p0;
                    apply(r0, id, gx0 + q, gy0 + r, c, bnd || zb, ax, dg, p0);
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
        __syncthreads();
    }
    if (POST == 3) block_atomic_max(nmax, a.nrm);
}
