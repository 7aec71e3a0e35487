// vdn_mg_fused.cuh -- k_sweep3: the fused red-black multigrid smoother (included by vdn_mg.cu; also compiled as plain C++ by
// tests/emu/, where one OS thread plays each CUDA thread, so the kernel logic runs in the CPU-only test tier).
//
// One launch applies a full red-black Gauss-Seidel sweep (S = 2 colour stages) to a whole level while phi streams through
// shared memory ONCE (input array -> output array, ping-pong), with
//   PRE  = 1 : prolongation of the coarse correction added as a plane lands (k_prolong fused),
//   POST = 2 : residual of the finished plane averaged 2x2x2 into the coarse right-hand side, coarse phi zeroed
//              (k_residual + k_restrict fused; the fine residual is never stored),
//   POST = 3 : inf-norm of the residual reduced (the convergence test of the V-cycle, one atomic per CTA).
// A CTA owns a (TX x TY) tile of columns plus a halo of H cells and marches along z; in step t colour stage s relaxes plane
// t-s, so every stage sees exactly the neighbour values the plain sweep order gives it; halo cells (and, between ranks, the
// neighbour's cells held in the level's ghost layers, M_GHOST) are relaxed redundantly; periodic directions wrap by index.
// After the last (black) stage the residual of the black cells is zero to round-off -- a cell that was just relaxed satisfies
// its equation when its neighbours no longer change -- so the residual stage visits red cells only.
// HBM traffic per launch: phi in + out, rhs, 3 face-coefficient arrays = 48 B/cell (x tile halo) for two colour stages +
// residual + transfer operator, against 48 B/cell for EVERY colour stage (and again for the residual) of the plain kernels.
//
// Thread mapping (third generation; the captures that led here are profiles/r01_ncu_full_wave_v2 / _ksweep_v3 / _ksweep2_v4):
// a thread owns the two cells (x, 2j) and (x, 2j+1) of a column pair: one is red, one is black, so every thread has exactly
// ONE active cell per step and runs all colour stages on it (z-neighbours in program order, x/y-neighbours written one step
// earlier -> one barrier per plane).  Consecutive lanes are consecutive x: global accesses are coalesced 8-byte words,
// shared-memory rows are read conflict-free.  The operator data of stage 0 of step t+1 (new lines, HBM latency) is loaded
// into registers at the top of step t; stage 1 (and the residual stage) read lines that were streamed one (two) steps earlier
// and are L1/L2-resident; both are issued before the phi traffic of the step.
#pragma once

enum : int { M_GHOST = 0, M_NEU = 1, M_DIR = 2, M_WRAP = 3 };
constexpr int PUSH_DEPTH = 3;          // ghost layers a fused launch may read (H <= 3)

struct WaveArgs {
    int n[3]; long s1, s2, off;
    double h2[3]; int mode[3][2]; int par0;
    const double *rhs, *b0, *b1, *b2;
    const double *dinv;                  // 1 / diagonal of the operator per cell (0 where the diagonal is 0), boundary conditions included
    const double *in; double *out;
    const double *cphi; double *crhs, *czero; long cs1, cs2, coff;   // coarse level (PRE / POST == 2)
    double *nrm;
    int zchunk;
    // peer-memory mode (levels split across ranks, all level arrays in the symmetric heap of the peer-memory transport): the kernel PUSHES.
    // Every value it writes for a cell that lies within PUSH_DEPTH cells of a face shared with another rank -- the new phi, and under POST == 2
    // the coarse right-hand side and the zeroed coarse phi -- is also stored straight into the ghost layers of the SAME array on the rank(s)
    // that hold that cell as a ghost (face, edge and corner neighbours: up to 7 copies) -- by the CTA that produced it, after its march, from its own
    // L2-hot output.  Stores over NVLink are posted (nothing waits for them), so the exchange of a sweep costs no launch and no exposed latency;
    // the next launch then reads its ghost layers from local memory.
    // peer_delta[q]: byte distance from this rank's heap to the heap of the rank at process-grid offset (ox, oy, oz), q = (ox+1) + 3 (oy+1) +
    // 9 (oz+1) -- the heaps have one layout, so local pointer + delta = the same array there.  The kernel publishes "my stream has reached launch
    // `epoch`" (everything before it is complete) in its own flag word, and the CTAs that read ghost layers or push wait for the flags of the
    // neighbours: their earlier launches -- which wrote my ghost layers and read the ghost layers I am about to overwrite -- are complete.
    // Interior CTAs start at once, so the wait overlaps with their work.
    int p2p;
    unsigned peer_mask;                  // bit q: a rank exists at process-grid offset q
    long peer_delta[27];
    // flags: my_flag = word 0 of this rank's heap header (read remotely by the pull kernels of vdn_comm.cu); pub_flag[q] = the word in the heap
    // header of the rank at offset q where THIS rank announces its epoch (a posted remote store); wait_flag[q] = the word in this rank's own
    // header where that rank announces its epoch -- so the wait polls local memory and ends one NVLink latency after the neighbour's store
    unsigned long long *pub_flag[27]; const unsigned long long *wait_flag[27]; unsigned long long *my_flag; unsigned long long epoch;
    unsigned long long *wait_ns;         // measurement hook (may be null): [0] += ns an edge CTA waited for its first neighbour, [1] = max, [2] += 1
};


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

// relaxation operands of one cell: phi neighbours from the shared-memory ring, operator data in registers
// dinv: a relaxation is phi += (rhs - A phi) * dinv -- the reciprocal of the diagonal is computed once per solve and level (k_diag_inv in
// vdn_mg.cu) instead of a diagonal sum and an FP64 division (~35 of ~130 instructions) in every relaxation of every sweep
struct SweepCoef { double rhs, xl, xh, yl, yh, zl, zh, dinv; };

__device__ __forceinline__ void sweep_load(SweepCoef &c, const WaveArgs &a, long g)
{
    c.rhs = __ldg(a.rhs + g);
    c.xl = __ldg(a.b0 + g); c.xh = __ldg(a.b0 + g + 1);
    c.yl = __ldg(a.b1 + g); c.yh = __ldg(a.b1 + g + a.s1);
    c.zl = __ldg(a.b2 + g); c.zh = __ldg(a.b2 + g + a.s2);
    c.dinv = __ldg(a.dinv + g);
}

#ifndef VDN_EMU
__device__ __forceinline__ unsigned long long vdn_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
#ifdef VDN_EMU
inline unsigned long long vdn_globaltimer() { return 0; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
template <class T> inline T __ldcv(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
inline void __threadfence_system() { }
inline double emu_xor_buf[2048];
inline double __shfl_xor_sync(unsigned, double v, int m)
{
    emu_xor_buf[threadIdx.x] = v;
    __syncthreads();
    const double r = emu_xor_buf[threadIdx.x ^ m];
    __syncthreads();
    return r;
}
#endif

template <int PRE, int POST, int TX, int TY>
struct Sweep3Cfg {
    static constexpr int S = 2, E = POST ? 1 : 0, H = S + E, HE = (H + 1) & ~1;
    static constexpr int X = TX + 2 * HE, Y = TY + 2 * HE;      // shared-memory tile; row pairs / lane pairs aligned to even global indices
    static constexpr int BY = Y / 2, PLANE = X * Y;
    static constexpr int NPL = S + 2 + E;                        // phi ring: planes t-S-E .. t+1
    static constexpr size_t SMEM = sizeof(double) * PLANE * NPL;
    static constexpr int NT = ((X * BY + 31) / 32) * 32;
    static_assert(TX % 2 == 0 && TY % 2 == 0, "even tiles");
};

template <int PRE, int POST, int TX, int TY, bool P2P = false>
__global__ void __launch_bounds__(Sweep3Cfg<PRE, POST, TX, TY>::NT, 1) k_sweep3(const WaveArgs a)
{
    using C = Sweep3Cfg<PRE, POST, TX, TY>;
    constexpr int S = C::S, E = C::E, H = C::H, HE = C::HE, X = C::X, PLANE = C::PLANE, NPL = C::NPL;
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z0 = blockIdx.z * a.zchunk, z1 = min(z0 + a.zchunk, a.n[2]);
    const int n0 = a.n[0], n1 = a.n[1], n2 = a.n[2];
    const int mx0 = a.mode[0][0], mx1 = a.mode[0][1], my0 = a.mode[1][0], my1 = a.mode[1][1], mz0 = a.mode[2][0], mz1 = a.mode[2][1];
    const bool have = tid < X * C::BY;
    const int ty = have ? tid / X : 0, lx = have ? tid - ty * X : 0, ly0 = 2 * ty;
    const int gx = x0 - HE + lx, gy0 = y0 - HE + ly0;           // unwrapped global coordinates (gy0 even)
    const int sid = ly0 * X + lx;                               // shared-memory index of the pair's row-0 cell

    // per cell (row r): number of stages it may run (0: not loaded).  Stage s needs s+1 cells to the edge of the H-grown region.
    const int ux = lx - (HE - H);
    const int wx = wave_idx<H>(gx, n0, mx0, mx1);
    const bool okx = have && wx != WAVE_NONE && ux >= 0 && ux < TX + 2 * H;
    int depth[2], wy[2];
    bool ld[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int uy = ly0 + r - (HE - H);
        wy[r] = wave_idx<H>(gy0 + r, n1, my0, my1);
        ld[r] = okx && wy[r] != WAVE_NONE && uy >= 0 && uy < TY + 2 * H;
        const int d = min(min(ux, TX + 2 * H - 1 - ux), min(uy, TY + 2 * H - 1 - uy));
        depth[r] = ld[r] ? min(max(d, 0), H) : 0;
    }
    const int wyb = ld[0] ? wy[0] : wy[1] - 1;                  // rows of a pair exist together except on the masked outer ring
    const long gofs = a.off + (okx ? wx : 0) + a.s1 * (long)wyb;  // + s1 * row + s2 * plane
    const long cofs = PRE ? a.coff + ((okx ? wx : 0) >> 1) + a.cs1 * (long)(wyb >> 1) : 0;       // (arithmetic shifts: ghost indices are negative)
    const bool anyld = ld[0] || ld[1];
    const bool bndx = (gx == 0 && (mx0 == M_NEU || mx0 == M_DIR)) || (gx == n0 - 1 && (mx1 == M_NEU || mx1 == M_DIR));
    bool bndy[2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
        bndy[r] = (gy0 + r == 0 && (my0 == M_NEU || my0 == M_DIR)) || (gy0 + r == n1 - 1 && (my1 == M_NEU || my1 == M_DIR));
    const bool core = have && lx >= HE && lx < HE + TX && ly0 >= HE && ly0 < HE + TY && gx < n0 && gy0 < n1;

    auto zidx = [&](int p) { return (p >= z0 - H && p <= z1 - 1 + H) ? wave_idx<H>(p, n2, mz0, mz1) : WAVE_NONE; };
    auto ring = [](int p) { return ((p + 64 * NPL) % NPL) * PLANE; };

    // ---- phi prefetch registers: the pair's two cells of one plane (+ the coarse correction under them) ----
    double pf[2] = { 0.0, 0.0 }, pc = 0.0;
    auto fetch = [&](int wz) {
        pf[0] = 0.0; pf[1] = 0.0; pc = 0.0;
        if (wz != WAVE_NONE && anyld) {
            const double *src = a.in + gofs + a.s2 * (long)wz;      // ghost layers included (M_GHOST: filled by the neighbours' pushes / an exchange)
            if (ld[0]) pf[0] = src[0];
            if (ld[1]) pf[1] = src[a.s1];
            if (PRE) pc = __ldg(a.cphi + cofs + a.cs2 * (long)(wz >> 1));
        }
    };
    auto stash = [&](int o) {
        if (!have) return;
        sm[o + sid] = pf[0] + ((PRE && ld[0]) ? pc : 0.0);
        sm[o + sid + X] = pf[1] + ((PRE && ld[1]) ? pc : 0.0);
    };

    // A*phi and the diagonal at tile cell `id` of the plane at ring offset o0 (o0m / o0p: the planes below / above)
    auto apply = [&](int p, int o0, int o0m, int o0p, int id, int gy, bool general, const SweepCoef &c, double &ax, double &p0) {
        const double *P0 = sm + o0 + id;
        p0 = P0[0];
        const double xm = P0[-1], xp = P0[1], ym = P0[-X], yp = P0[X], zm = sm[o0m + id], zp = sm[o0p + id];
        if (!general) {
            ax = (c.xl * (p0 - xm) + c.xh * (p0 - xp)) * a.h2[0] + (c.yl * (p0 - ym) + c.yh * (p0 - yp)) * a.h2[1]
               + (c.zl * (p0 - zm) + c.zh * (p0 - zp)) * a.h2[2];
        } else {
            double dg = 0.0;
            ax = 0.0;
            wave_dir(c.xl, c.xh, a.h2[0], p0, xm, xp, gx == 0, gx == n0 - 1, mx0, mx1, ax, dg);
            wave_dir(c.yl, c.yh, a.h2[1], p0, ym, yp, gy == 0, gy == n1 - 1, my0, my1, ax, dg);
            wave_dir(c.zl, c.zh, a.h2[2], p0, zm, zp, p == 0, p == n2 - 1, mz0, mz1, ax, dg);
        }
    };
    // may stage s run on row r of plane p?
    auto runs = [&](int s, int p, int r, bool exists) {
        const int RS = E + S - 1 - s;
        return exists && p >= z0 - RS && p <= z1 - 1 + RS && (r ? depth[1] : depth[0]) > s;
    };

    double nmax = 0.0, acc = 0.0;
    const int tfirst = z0 - H, tlast = z1 + S - 2 + E;
    // ring offsets as rotating registers: at step t, oP[k] = offset of plane t+1-k; wzq[k] = array plane index of plane t+2-k
    int oP[NPL], wzq[S + 3 + E];
#pragma unroll
    for (int k = 0; k < NPL; ++k) oP[k] = ring(tfirst - k);             // state "step tfirst-1", rotated at the top of the loop
#pragma unroll
    for (int k = 0; k < S + 3 + E; ++k) wzq[k] = zidx(tfirst + 1 - k);
    if (P2P && a.my_flag) {
        // everything this rank's stream produced before this launch is complete: publish it; CTAs that read ghost layers or push wait until
        // the neighbours have published the same launch (their earlier launches are complete as well)
        constexpr int MG = 2 * PUSH_DEPTH;          // footprint of what the CTA reads (H) or pushes (fine cells over PUSH_DEPTH coarse cells)
        const bool edge = (x0 - MG < 0 && mx0 == M_GHOST) || (x0 + TX + MG > n0 && mx1 == M_GHOST) || (y0 - MG < 0 && my0 == M_GHOST) ||
                          (y0 + TY + MG > n1 && my1 == M_GHOST) || (z0 - MG < 0 && mz0 == M_GHOST) || (z1 + MG > n2 && mz1 == M_GHOST);
        // the CTAs that are dispatched first, and every CTA that is about to wait, announce the epoch to the neighbours
        const bool announce = edge || (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) < 32u;
        if (announce && tid == 0) { __threadfence_system(); *(volatile unsigned long long *)a.my_flag = a.epoch; }
        if (announce && tid < 27 && a.pub_flag[tid]) { __threadfence_system(); *(volatile unsigned long long *)a.pub_flag[tid] = a.epoch; }
        if (edge && tid < 27 && a.wait_flag[tid]) {
            const volatile unsigned long long *f = (const volatile unsigned long long *)a.wait_flag[tid];
            const unsigned long long t0 = a.wait_ns ? vdn_globaltimer() : 0ull;
            while (*f < a.epoch) { }
            __threadfence_system();
            if (a.wait_ns) {
                const unsigned long long dt = vdn_globaltimer() - t0;
                atomicAdd(a.wait_ns, dt); atomicMax(a.wait_ns + 1, dt); atomicAdd(a.wait_ns + 2, 1ull);
            }
        }
        __syncthreads();
    }
    fetch(wzq[1]); stash(oP[0]); fetch(wzq[0]);                        // planes tfirst (into the ring) and tfirst+1 (registers)
    // operator data of stage 0 of the first step
    SweepCoef c0n;
    bool run0n;
    {
        const int ra = (gx + tfirst + a.par0) & 1;
        run0n = runs(0, tfirst, ra, wzq[1] != WAVE_NONE);
        if (run0n) sweep_load(c0n, a, gofs + a.s1 * ra + a.s2 * (long)wzq[1]);
    }
    __syncthreads();

    for (int t = tfirst; t <= tlast; ++t) {
        const int ra = (gx + t + a.par0) & 1;                    // active row of this column in this step
        { const int x = oP[NPL - 1];
#pragma unroll
          for (int k = NPL - 1; k > 0; --k) oP[k] = oP[k - 1];
          oP[0] = x; }
#pragma unroll
        for (int k = S + 2 + E; k > 0; --k) wzq[k] = wzq[k - 1];
        wzq[0] = zidx(t + 2);
        // ---- operator data of stage 1: lines that stage 0 streamed one step ago (L1/L2-resident), issued before the phi traffic ----
        SweepCoef c1;
        const bool run1 = runs(1, t - 1, ra, wzq[3] != WAVE_NONE);
        if (run1) sweep_load(c1, a, gofs + a.s1 * ra + a.s2 * (long)wzq[3]);
        const int r0 = t - S;
        const bool run2 = POST && core && r0 >= z0 && r0 < z1;
        // ---- plane t+1 into the ring, plane t+2 on its way ----
        stash(oP[0]);
        fetch(wzq[0]);
        const int id = sid + (ra ? X : 0), gy = gy0 + ra;
        const bool bxy = bndx || (ra ? bndy[1] : bndy[0]);
        // ---- stage 0 on plane t: its operator data was requested in the middle of the previous step ----
        if (run0n) {
            const int p = t;
            const bool zb = (p == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (p == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
            double ax, p0;
            apply(p, oP[1], oP[2], oP[0], id, gy, bxy || zb, c0n, ax, p0);
            sm[oP[1] + id] = p0 + (c0n.rhs - ax) * c0n.dinv;
        }
        // ---- request the operator data of stage 0 of the NEXT step (new lines: HBM latency, half a step + the barrier to arrive) ----
        run0n = runs(0, t + 1, ra ^ 1, wzq[1] != WAVE_NONE);
        if (run0n) sweep_load(c0n, a, gofs + a.s1 * (ra ^ 1) + a.s2 * (long)wzq[1]);
        // ---- stage 1 on plane t-1 ----
        if (run1) {
            const int p = t - 1;
            const bool zb = (p == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (p == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
            double ax, p0;
            apply(p, oP[2], oP[3], oP[1], id, gy, bxy || zb, c1, ax, p0);
            sm[oP[2] + id] = p0 + (c1.rhs - ax) * c1.dinv;
        }
        // ---- plane t-S+1 has passed every stage: write it out ----
        {
            const int r1 = t - S + 1;
            if (core && r1 >= z0 && r1 < z1) {
                const long e = gofs + a.s2 * (long)r1;
                const double v0 = sm[oP[S] + sid], v1 = sm[oP[S] + sid + X];
                a.out[e] = v0;
                a.out[e + a.s1] = v1;
            }
        }
        // ---- residual of plane t-S: red cells only (the black ones were just relaxed) ----
        if (POST) {
            double res = 0.0;
            if (run2) {
                SweepCoef c2; sweep_load(c2, a, gofs + a.s1 * ra + a.s2 * (long)r0);        // streamed two steps ago
                const bool zb = (r0 == 0 && (mz0 == M_NEU || mz0 == M_DIR)) || (r0 == n2 - 1 && (mz1 == M_NEU || mz1 == M_DIR));
                double ax, p0;
                apply(r0, oP[S + 1], oP[(S + 2) % NPL], oP[S], id, gy, bxy || zb, c2, ax, p0);
                res = c2.rhs - ax;
                if (POST == 3) nmax = fmax(nmax, fabs(res));
            }
            if (POST == 2) {
                // the 2x2 block of a coarse cell: this column and its lane partner (x ^ 1) hold one red cell each
                const double s2 = res + __shfl_xor_sync(0xffffffffu, res, 1);
                if (run2) {
                    if ((r0 & 1) == 0) acc = s2;
                    else if ((gx & 1) == 0) {
                        const long cc = a.coff + (gx >> 1) + a.cs1 * (long)(gy0 >> 1) + a.cs2 * (long)(r0 >> 1);
                        const double cr = (acc + s2) * 0.125;
                        a.crhs[cc] = cr;
                        a.czero[cc] = 0.0;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (POST == 3) block_atomic_max(nmax, a.nrm);
    if (P2P) {
        // ---- peer-memory mode: what this CTA wrote within PUSH_DEPTH cells of a face shared with another rank goes into that rank's ghost layers
        // (face, edge and corner neighbours alike: the part of the CTA's box of cells that lies in the neighbour's ghost region).  Done after the
        // march, from the CTA's own output (L2-hot), so that the marching loop is the one of the single-GPU kernel. ----
        __syncthreads();
        const int cb[3][2] = { { x0, min(x0 + TX, n0) }, { y0, min(y0 + TY, n1) }, { z0, z1 } };
        for (int lev = 0; lev < (POST == 2 ? 2 : 1); ++lev) {         // 0: phi on this level; 1: coarse rhs and zeroed coarse phi
            const int nn[3] = { n0 >> lev, n1 >> lev, n2 >> lev };
            const long st1 = lev ? a.cs1 : a.s1, st2 = lev ? a.cs2 : a.s2, of = lev ? a.coff : a.off;
#pragma unroll 1
            for (int q = 0; q < 27; ++q) {
                if (q == 13 || !((a.peer_mask >> q) & 1)) continue;
                const int o[3] = { q % 3 - 1, (q / 3) % 3 - 1, q / 9 - 1 };
                int lo[3], ex[3];
                bool any = true;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    int l = cb[d][0] >> lev, h = cb[d][1] >> lev;
                    if (o[d] < 0) h = min(h, PUSH_DEPTH); else if (o[d] > 0) l = max(l, nn[d] - PUSH_DEPTH);
                    lo[d] = l; ex[d] = h - l;
                    if (ex[d] <= 0) any = false;
                }
                if (!any) continue;
                const long shift = -(long)o[0] * nn[0] - st1 * (long)(o[1] * nn[1]) - st2 * (long)(o[2] * nn[2]);
                const int cnt = ex[0] * ex[1] * ex[2];
                for (int t = tid; t < cnt; t += (int)blockDim.x) {
                    const int i = lo[0] + t % ex[0], j = lo[1] + (t / ex[0]) % ex[1], k = lo[2] + t / (ex[0] * ex[1]);
                    const long e = of + i + st1 * (long)j + st2 * (long)k;
                    if (lev == 0) {
                        ((double *)((char *)a.out + a.peer_delta[q]))[e + shift] = __ldcg(a.out + e);
                    } else {
                        ((double *)((char *)a.crhs + a.peer_delta[q]))[e + shift] = __ldcg(a.crhs + e);
                        ((double *)((char *)a.czero + a.peer_delta[q]))[e + shift] = 0.0;
                    }
                }
            }
        }
    }
}
