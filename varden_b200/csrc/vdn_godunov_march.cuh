// vdn_godunov_march.cuh -- the 3-D Godunov predictor as ONE plane-marching kernel per routine.
//
//   k_mkflux_march  = mkflux_3d  (mkflux.f90:1186-2567), slopes (slope.f90:148,291,437) included
//   k_velpred_march = velpred_3d (velpred.f90:1776-2765), slopes included
//
// The reference cycles two k-planes of scratch per box (velpred.f90:1986-2004, mkflux.f90:1407-1432); the device form
// of that idea: a CTA owns a tile of 32 x TYT columns of the S-layout (cells -1..n), of which the inner 30 x (TYT-2)
// produce output, and marches along z.  No intermediate ever reaches HBM: a field is read once, the edge states are
// written once.
//
// Everything is formulated CELL-CENTRED so that one thread = one column needs nothing from its neighbours except
// finished values:
//   * every left/right state of a face is a function of ONE cell (slopes, transverse terms, forces of that cell): the
//     thread of cell c computes the R state of its three lo faces and the L state of its three hi faces;
//   * a face value (upwind / Riemann pick) needs the L state of the lower neighbour: x -> warp shuffle (a tile row is one
//     warp), y -> shared memory (one exchange slot per quantity, conflict-free), z -> the thread's own previous iteration;
//   * a transverse term of cell c, e.g. (dt6/hy)(v_hi+v_lo)(q_hi-q_lo), is computed ONCE per cell and used by the four
//     states that contain it; it needs the face value of the upper neighbour (shuffle down / shared memory / next iteration).
//   * limited slopes: cen/lim/fromm (slope.f90:227-234) are evaluated once per cell and direction and exchanged the same way
//     (the z direction keeps them in registers across iterations).
//   * state that lives from one iteration to the next (the 1-D states of plane k-1, the z-face states of face k-1, in velpred also the z
//     window / slope parts) is kept in per-thread "keep" words of shared memory, not in registers: at 512 threads per CTA a thread has 128
//     registers, and the first build (all state in registers) spilled 140 B (mkflux) / 1 KB (velpred) per thread to local memory.
// Operation order inside every expression follows the reference (the file is compiled with -fmad=false), so the results
// are bit-identical to the staged kernels and to the CPU oracle; only the evaluation SITE of shared sub-expressions moves.
//
// Iteration k of the march (cell plane k, z-face k = lo face of that plane):
//   N(k)   slopes, 1-D extrapolation, simh_x/y (plane k), simh_z (face k)
//   T      s?xy, s?yx (plane k), s?zx, s?zy (face k), s?xz, s?yz (plane k-1: needs simh_z of faces k-1 and k)
//   F      sedge_z (face k), sedge_x, sedge_y (plane k-1)
// Boundary-condition overrides (table SURVEY A.4): at a lo face the pair becomes (f(r), f(r)), at a hi face (g(l), g(l)),
// so the owner of the R state (lo) resp. of the L state (hi) applies it before the exchange.
#pragma once

namespace march {

constexpr int TXT = 32;             // thread-tile width = one warp
constexpr int TXO = TXT - 2;        // output columns per tile row

#ifdef VDN_EMU
#define VDN_DYN_SMEM(name) double *name = (double *)emu_smem
#else
#define VDN_DYN_SMEM(name) extern __shared__ double name[]
#endif

// Synchronisation of the y exchange.  A row of the thread tile is one warp and only exchanges with the rows above and below it, so a
// CTA-wide barrier is more than the data flow needs: with MARCH_PAIRBAR every pair of adjacent rows owns one named barrier (ids 1..TYT-1);
// even rows meet their upper neighbour first, odd rows their lower one, so an exchange costs two 64-thread barriers per warp and the warps of
// a CTA may drift apart by a phase per row instead of all waiting for the slowest one.
#ifndef MARCH_PAIRBAR
#define MARCH_PAIRBAR 0
#endif
#ifndef MARCH_MINB
#define MARCH_MINB 1         // resident CTAs per SM the register allocation aims for
#endif
template <int TYT> __device__ __forceinline__ void row_sync(int ty)
{
#if MARCH_PAIRBAR && !defined(VDN_EMU)
    static_assert(TYT <= 16, "one hardware barrier per pair of rows");
    const int up = ty + 1, dn = ty;                       // barrier ids of the pairs (ty, ty+1) and (ty-1, ty)
    if ((ty & 1) == 0) {
        if (ty < TYT - 1) asm volatile("bar.sync %0, 64;" :: "r"(up) : "memory");
        if (ty > 0)       asm volatile("bar.sync %0, 64;" :: "r"(dn) : "memory");
    } else {
        asm volatile("bar.sync %0, 64;" :: "r"(dn) : "memory");
        if (ty < TYT - 1) asm volatile("bar.sync %0, 64;" :: "r"(up) : "memory");
    }
#else
    (void)ty; __syncthreads();
#endif
}

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ bool bc_overrides(int bc) { return bc == BC_INLET || bc == BC_OUTLET || bc == BC_SLIP_WALL || bc == BC_NO_SLIP_WALL; }

// min of two non-negative values (limiter): compare + select, no NaN handling needed
__device__ __forceinline__ double min_nn(double a, double b) { return a < b ? a : b; }

// slope.f90:227-234 for one cell from its three values
struct Parts { double cen, lim, flag, fromm; };
__device__ __forceinline__ Parts slope_parts3(double sm, double s0, double sp)
{
    Parts p;
    p.cen = HALF * (sp - sm);
    const double dmn = TWO * (s0 - sm), dpls = TWO * (sp - s0);
    const double l = min_nn(fabs(dmn), fabs(dpls));
    p.lim = (dpls * dmn > ZERO) ? l : ZERO;
    p.flag = copysign(ONE, p.cen);
    p.fromm = p.flag * min_nn(p.lim, fabs(p.cen));
    return p;
}
// one-sided boundary slopes from explicit values: v[-1], v[0], v[1], v[2] counted INTO the domain from the ghost cell
// (slope.f90:247-254 lo, :268-275 hi mirrored; 2nd order :193-200, :207-214)
__device__ __forceinline__ double slope_onesided(double g, double s0, double s1, double s2, bool hi, int order)
{
    // lo: g = s(is-1), s0 = s(is), s1 = s(is+1), s2 = s(is+2);  hi: g = s(ie+1), s0 = s(ie), s1 = s(ie-1), s2 = s(ie-2)
    double del, dmn, dpls;
    if (order == 2) {
        if (!hi) { del = (s1 + 3.0 * s0 - 4.0 * g) * (1.0 / 3.0); dpls = TWO * (s1 - s0); dmn = TWO * (s0 - g); }
        else     { del = -(s1 + 3.0 * s0 - 4.0 * g) * (1.0 / 3.0); dpls = TWO * (s0 - s1); dmn = TWO * (g - s0); }
    } else {
        const double two3rd = 2.0 / 3.0, tenth = 0.1;
        del = (-(16.0 / 15.0)) * g + HALF * s0 + two3rd * s1 - tenth * s2;
        if (hi) { del = -del; dmn = TWO * (s0 - s1); dpls = TWO * (g - s0); }
        else    { dmn = TWO * (s0 - g); dpls = TWO * (s1 - s0); }
    }
    double slim = fmin(fabs(dpls), fabs(dmn));
    slim = (dpls * dmn > ZERO) ? slim : ZERO;
    return copysign(ONE, del) * fmin(slim, fabs(del));
}
__device__ __forceinline__ double slope4_from(const Parts &p, double frm, double frp)     // slope.f90:238-240
{
    const double two3rd = 2.0 / 3.0, sixth = 1.0 / 6.0;
    const double ds = TWO * two3rd * p.cen - sixth * (frp + frm);
    return p.flag * min_nn(fabs(ds), p.lim);
}

// x / h with the reference's rounding: a true division unless h is a power of two (then the product with 1/h is exact)
__device__ __forceinline__ double div_h(double x, double h, double hinv, int hp2) { return hp2 ? x * hinv : x / h; }

// upwind pick with the face predicates precomputed (mkflux.f90:1520-1522)
struct FaceP { bool pos, big; };
__device__ __forceinline__ FaceP face_pred(double um, double eps) { FaceP p; p.pos = um > ZERO; p.big = fabs(um) > eps; return p; }
__device__ __forceinline__ double upw_p(double l, double r, FaceP p)
{
    const double v = p.pos ? l : r;
    const double savg = HALF * (l + r);
    return p.big ? v : savg;
}

// mkflux boundary override of the state that survives at a boundary face (mkflux.f90:1463-1515, :2356-2397):
// lo face -> both states become bc_lo(r), hi face -> both become bc_hi(l).  sg = s in the ghost cell.
__device__ __forceinline__ double mf_bc_lo(double v, double sg, int D, int bc, int is_vel, int comp)
{
    if (bc == BC_INLET) return sg;
    if (bc == BC_SLIP_WALL) return (is_vel && comp == D) ? ZERO : v;
    if (bc == BC_NO_SLIP_WALL) return is_vel ? ZERO : v;
    if (bc == BC_OUTLET) return (is_vel && comp == D) ? fmin(v, ZERO) : v;
    return v;
}
__device__ __forceinline__ double mf_bc_hi(double v, double sg, int D, int bc, int is_vel, int comp)
{
    if (bc == BC_INLET) return sg;
    if (bc == BC_SLIP_WALL) return (is_vel && comp == D) ? ZERO : v;
    if (bc == BC_NO_SLIP_WALL) return is_vel ? ZERO : v;
    if (bc == BC_OUTLET) return (is_vel && comp == D) ? fmax(v, ZERO) : v;
    return v;
}

constexpr int MARCH_MAXC = 3;

struct MfmArgs {
    Geo g;
    int nc;                                  // components handled by this launch (== template NC)
    const double *s[MARCH_MAXC]; int s_sy, s_sz;          // comp base pointers at local cell (0,0,0)
    const double *mac[3]; int m_sy[3], m_sz[3];
    const double *force[MARCH_MAXC]; int f_sy, f_sz;
    double *sedge[3][MARCH_MAXC]; double *flux[3][MARCH_MAXC]; int e_sy[3], e_sz[3];
    const double *eps;
    int is_vel, comp0, order, use_minion;
    int sbc[MARCH_MAXC][3][2];               // adv_bc[comp][d][side] (slopes)
    int zchunk;
    double dt2, c2[3], c3[3], c4[3], c6[3], hinv[3]; int hp2[3];
};

// ---- per-CTA shared memory: slot q holds one double per thread of the CTA ----
//  * exchange slots: a value written by row ty is read by row ty-1 / ty+1 after a barrier (conflict-free: consecutive lanes, consecutive words)
//  * keep slots: per-thread state that lives from one iteration to the next (the L/R states of plane k-1, the z-face states of face k-1 ...).
//    Holding it in registers costs more than 128 of them per thread (ncu, round 2 call 1: 23 local-memory loads + 11 stores per iteration,
//    long-scoreboard the top stall); a thread reads and writes only its own word, so no barrier is involved.
template <int TYT> struct YSlots {
    double *sm; int own, lo, hi;
    __device__ __forceinline__ YSlots(double *base, int tx, int ty) : sm(base), own(ty * TXT + tx),
        lo((ty > 0 ? ty - 1 : ty) * TXT + tx), hi((ty < TYT - 1 ? ty + 1 : ty) * TXT + tx) {}
    __device__ __forceinline__ void put(int q, double v) const { sm[q * (TXT * TYT) + own] = v; }
    __device__ __forceinline__ double get(int q) const { return sm[q * (TXT * TYT) + own]; }       // own word (keep slots)
    __device__ __forceinline__ double from_lo(int q) const { return sm[q * (TXT * TYT) + lo]; }    // value of row ty-1
    __device__ __forceinline__ double from_hi(int q) const { return sm[q * (TXT * TYT) + hi]; }    // value of row ty+1
};
// MK_LXP .. MK_RYP are double-buffered by iteration parity (set 4*(k&1)): plane k-1 is still read in F after plane k was stored in T
enum { MK_LXP = 0, MK_RXP, MK_LYP, MK_RYP, MK_LZH = 8, MK_LZXH, MK_LZYH, MK_LZF, MK_QZP, MK_XZXP, MK_XZYP, MK_FP, MK_NKEEP };
template <int NC> constexpr int mf_smem_slots() { return (4 + MK_NKEEP) * NC; }

// slope.f90:227-234 without the sign (recomputed from cen where it is needed: one LOP3)
struct PartsZ { double cen, lim, fromm; };

// ------------------------------------------------------------------------------------------
// mkflux_3d.  NC components per launch (they share the MAC velocities, eps and all index work); bit c of CONSMASK:
// component c is conservative (scalar_advance.f90:54-57).  GEN = false: slope_order 4, use_minion = F compiled in.
// HP2: every mesh spacing is a power of two, so x / h is the exact product x * (1/h) (no division sequence, no branch).
// grid = (tiles_x, tiles_y, z chunks), block = (32, TYT).
// ------------------------------------------------------------------------------------------
template <int NC, int CONSMASK, int TYT, bool GEN, bool HP2>
__global__ void __launch_bounds__(TXT * TYT, MARCH_MINB) k_mkflux_march(const MfmArgs a)
{
    VDN_DYN_SMEM(smem);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int n0 = a.g.n[0], n1 = a.g.n[1], n2 = a.g.n[2];
    const int it = (int)blockIdx.x * TXO - 1 + tx, jt = (int)blockIdx.y * (TYT - 2) - 1 + ty;     // this thread's column
    const int i = it < n0 ? it : n0, j = jt < n1 ? jt : n1;                                         // clamped for addressing
    // s itself is addressed one column further out: the slope of cell n needs fromm of cell n+1 (s has 3 ghost cells)
    const int is = it < n0 + 1 ? it : n0 + 1, js = jt < n1 + 1 ? jt : n1 + 1;
    const int ka = (int)blockIdx.z * a.zchunk;
    const int kb = (ka + a.zchunk < n2) ? ka + a.zchunk : n2;
    const bool lastchunk = kb == n2;
    const YSlots<TYT> Y(smem, tx, ty);
    constexpr int KB = 4 * NC;                            // first keep slot
    const int order = GEN ? a.order : 4;
    const bool minion = GEN ? (a.use_minion != 0) : false;

    const bool inner = tx >= 1 && tx <= TXT - 2 && ty >= 1 && ty <= TYT - 2 && it <= n0 && jt <= n1;
    const bool st_x = inner && jt < n1, st_y = inner && it < n0, st_z = inner && it < n0 && jt < n1;

    // physical-BC overrides on the region faces (periodic / rank-interior faces have none: those tiles never enter the BC code)
    const int bcx0 = a.g.pbc[0][0], bcx1 = a.g.pbc[0][1], bcy0 = a.g.pbc[1][0], bcy1 = a.g.pbc[1][1], bcz0 = a.g.pbc[2][0], bcz1 = a.g.pbc[2][1];
    const bool xlo = bc_overrides(bcx0) && i == 0, xhi_c = bc_overrides(bcx1) && i == n0 - 1, xhi_f = bc_overrides(bcx1) && i == n0;
    const bool ylo = bc_overrides(bcy0) && j == 0, yhi_c = bc_overrides(bcy1) && j == n1 - 1, yhi_f = bc_overrides(bcy1) && j == n1;
    const bool tile_x = (bc_overrides(bcx0) && blockIdx.x == 0) || (bc_overrides(bcx1) && (int)blockIdx.x * TXO + TXT >= n0);
    const bool tile_y = (bc_overrides(bcy0) && blockIdx.y == 0) || (bc_overrides(bcy1) && (int)blockIdx.y * (TYT - 2) + TYT >= n1);
    const bool tile_xy = tile_x || tile_y;

    // element offsets of (i, j, k = 0) in each layout
    const int so = is + a.s_sy * js, fo = i + a.f_sy * j;
    const int mo0 = i + a.m_sy[0] * j, mo1 = i + a.m_sy[1] * j, mo2 = i + a.m_sy[2] * j;
    const int eo0 = i + a.e_sy[0] * j, eo1 = i + a.e_sy[1] * j, eo2 = i + a.e_sy[2] * j;

    // eps: per reference box; the (i,j) part of the box index is fixed per thread
    const int nbt = a.g.nb[0] * a.g.nb[1] * a.g.nb[2];
    const int bij = nbt == 1 ? 0 : a.g.box1(0, i) + a.g.nb[0] * a.g.box1(1, j);
    auto divh = [&](double x, int d) { return HP2 ? x * a.hinv[d] : div_h(x, a.g.h[d], a.hinv[d], a.hp2[d]); };

    // ---- state carried in registers: the z window and the z slope parts; MAC data of plane k-1 ----
    double s0[NC], sp1[NC], sp2[NC];                   // s(k), s(k+1), s(k+2)
    PartsZ pz0[NC];                                    // slope parts of cell k (z)
    double frzm[NC];                                   // fromm_z(k-1)
    double umlP = 0, umhP = 0, vmlP = 0, vmhP = 0, wmlP = 0, wml = 0;
    FaceP PxP = { false, false }, PyP = { false, false };

    // prologue: z window around the first plane ka-1 and the slope parts that iteration needs
    {
        const int k = ka - 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double *sb = a.s[c] + so;
            const double sm3 = sb[(k - 2) * a.s_sz], sm2 = sb[(k - 1) * a.s_sz];
            s0[c] = sb[k * a.s_sz]; sp1[c] = sb[(k + 1) * a.s_sz]; sp2[c] = sb[(k + 2) * a.s_sz];
            const Parts p0 = slope_parts3(sm2, s0[c], sp1[c]);
            pz0[c].cen = p0.cen; pz0[c].lim = p0.lim; pz0[c].fromm = p0.fromm;
            frzm[c] = slope_parts3(sm3, sm2, s0[c]).fromm;
#pragma unroll
            for (int q = 0; q < MK_NKEEP; ++q) Y.put(KB + q * NC + c, ZERO);
        }
        wml = a.mac[2][mo2 + k * a.m_sz[2]];
    }

    for (int k = ka - 1; k <= kb; ++k) {
        const bool zlo = bc_overrides(bcz0) && k == 0, zhi_c = bc_overrides(bcz1) && k == n2 - 1, zhi_f = bc_overrides(bcz1) && k == n2;
        // uniform: can this CTA meet any BC work in this iteration (overrides, one-sided slopes next to a physical face)?
        const bool bnd = tile_xy || (bc_overrides(bcz0) && k <= 1) || (bc_overrides(bcz1) && k >= n2 - 2);
        const double eps = a.eps[nbt == 1 ? 0 : bij + a.g.nb[0] * a.g.nb[1] * a.g.box1(2, k)];
        const int pc = 4 * (k & 1), pp = 4 - pc;                      // keep-slot sets of plane k / plane k-1
        // ---- MAC velocities of cell plane k ----
        const double uml = a.mac[0][mo0 + k * a.m_sz[0]], umh = a.mac[0][mo0 + k * a.m_sz[0] + 1];
        const double vml = a.mac[1][mo1 + k * a.m_sz[1]], vmh = a.mac[1][mo1 + k * a.m_sz[1] + a.m_sy[1]];
        const double wmh = a.mac[2][mo2 + (k + 1) * a.m_sz[2]];
        const FaceP Px = face_pred(uml, eps), Py = face_pred(vml, eps), Pz = face_pred(wml, eps);
        // dt2*u/h at the lo and hi face of each direction
        const double txl = divh(a.dt2 * uml, 0), txh = divh(a.dt2 * umh, 0);
        const double tyl = divh(a.dt2 * vml, 1), tyh = divh(a.dt2 * vmh, 1);
        const double tzl = divh(a.dt2 * wml, 2), tzh = divh(a.dt2 * wmh, 2);
        // transverse factors of cell plane k (x, y) and k-1 (all)
        const double sux = umh + uml, suy = vmh + vml;
        const double suxP = umhP + umlP, suyP = vmhP + vmlP, suzP = wml + wmlP;

        double snext[NC], fk[NC];
        double lx[NC], rx[NC], ly[NC], ry[NC], lz[NC], rz[NC];      // 1-D states of cell plane k
        double qx[NC], qy[NC], qz[NC];                               // simh at the lo faces
        // ================= N(k): slopes and 1-D extrapolation =================
        Parts px[NC], py[NC], pz1[NC];
        double sxm[NC], sxp[NC], sym[NC], syp[NC];
        double frx_e[NC], fry_e[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double *sb = a.s[c] + so + k * a.s_sz;
            const int k3 = (k + 3 <= n2 + 2) ? 3 : 2;                 // s has 3 ghost planes
            snext[c] = sb[k3 * a.s_sz];
            sxm[c] = sb[-1]; sxp[c] = sb[1]; sym[c] = sb[-a.s_sy]; syp[c] = sb[a.s_sy];
            fk[c] = a.dt2 * a.force[c][fo + k * a.f_sz];
            frx_e[c] = fry_e[c] = ZERO;
            if (order != 0) {
                px[c] = slope_parts3(sxm[c], s0[c], sxp[c]);
                py[c] = slope_parts3(sym[c], s0[c], syp[c]);
                pz1[c] = slope_parts3(s0[c], sp1[c], sp2[c]);
                if (order == 4) {
                    // tile-edge columns: fromm of the cell just outside the thread tile
                    if (tx == 0 || tx == TXT - 1) {
                        const double s2 = sb[tx == 0 ? -2 : (it <= n0 ? 2 : 1)];
                        frx_e[c] = (tx == 0) ? slope_parts3(s2, sxm[c], s0[c]).fromm : slope_parts3(s0[c], sxp[c], s2).fromm;
                    }
                    if (ty == 0 || ty == TYT - 1) {
                        const double s2 = sb[ty == 0 ? -2 * a.s_sy : (jt <= n1 ? 2 : 1) * a.s_sy];
                        fry_e[c] = (ty == 0) ? slope_parts3(s2, sym[c], s0[c]).fromm : slope_parts3(s0[c], syp[c], s2).fromm;
                    }
                }
            }
        }
        // one-sided slopes at EXT_DIR / HOEXTRAP faces (slope.f90:243-283): the first interior cell takes the one-sided formula and
        // that value also replaces its fromm for the neighbour's 4th-order slope
        if (bnd && order != 0) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const double *sb = a.s[c] + so + k * a.s_sz;
                const bool bxl = a.sbc[c][0][0] == BC_EXT_DIR || a.sbc[c][0][0] == BC_HOEXTRAP, bxh = a.sbc[c][0][1] == BC_EXT_DIR || a.sbc[c][0][1] == BC_HOEXTRAP;
                const bool byl = a.sbc[c][1][0] == BC_EXT_DIR || a.sbc[c][1][0] == BC_HOEXTRAP, byh = a.sbc[c][1][1] == BC_EXT_DIR || a.sbc[c][1][1] == BC_HOEXTRAP;
                const bool bzl = a.sbc[c][2][0] == BC_EXT_DIR || a.sbc[c][2][0] == BC_HOEXTRAP, bzh = a.sbc[c][2][1] == BC_EXT_DIR || a.sbc[c][2][1] == BC_HOEXTRAP;
                if (bxl && i == 0)      px[c].fromm = slope_onesided(sxm[c], s0[c], sxp[c], sb[2], false, order);
                if (bxh && i == n0 - 1) px[c].fromm = slope_onesided(sxp[c], s0[c], sxm[c], sb[-2], true, order);
                if (bxh && tx == TXT - 1 && i + 1 == n0 - 1) frx_e[c] = slope_onesided(sb[2], sxp[c], s0[c], sxm[c], true, order);
                if (byl && j == 0)      py[c].fromm = slope_onesided(sym[c], s0[c], syp[c], sb[2 * a.s_sy], false, order);
                if (byh && j == n1 - 1) py[c].fromm = slope_onesided(syp[c], s0[c], sym[c], sb[-2 * a.s_sy], true, order);
                if (byh && ty == TYT - 1 && j + 1 == n1 - 1) fry_e[c] = slope_onesided(sb[2 * a.s_sy], syp[c], s0[c], sym[c], true, order);
                // z: pz1 belongs to cell k+1
                if (bzl && k + 1 == 0)      pz1[c].fromm = slope_onesided(s0[c], sp1[c], sp2[c], snext[c], false, order);
                if (bzh && k + 1 == n2 - 1) pz1[c].fromm = slope_onesided(sp2[c], sp1[c], s0[c], sb[-a.s_sz], true, order);
            }
        }
        // exchange fromm_y (y) -- x by shuffle, z in registers
        if (order == 4) {
#pragma unroll
            for (int c = 0; c < NC; ++c) Y.put(c, py[c].fromm);
        }
        row_sync<TYT>(ty);                                                                  // B1
        double slx[NC], sly[NC], slz[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (order == 4) {
                double fxm = shfl_up1(px[c].fromm), fxp = shfl_dn1(px[c].fromm);
                if (tx == 0) fxm = frx_e[c];
                if (tx == TXT - 1) fxp = frx_e[c];
                double fym = Y.from_lo(c), fyp = Y.from_hi(c);
                if (ty == 0) fym = fry_e[c];
                if (ty == TYT - 1) fyp = fry_e[c];
                slx[c] = slope4_from(px[c], fxm, fxp);
                sly[c] = slope4_from(py[c], fym, fyp);
                Parts pzc; pzc.cen = pz0[c].cen; pzc.lim = pz0[c].lim; pzc.flag = copysign(ONE, pz0[c].cen); pzc.fromm = pz0[c].fromm;
                slz[c] = slope4_from(pzc, frzm[c], pz1[c].fromm);
            } else if (order == 2) {
                slx[c] = px[c].fromm; sly[c] = py[c].fromm; slz[c] = pz0[c].fromm;
            } else { slx[c] = sly[c] = slz[c] = ZERO; }
        }
        if (bnd && order != 0) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const bool bxl = a.sbc[c][0][0] == BC_EXT_DIR || a.sbc[c][0][0] == BC_HOEXTRAP, bxh = a.sbc[c][0][1] == BC_EXT_DIR || a.sbc[c][0][1] == BC_HOEXTRAP;
                const bool byl = a.sbc[c][1][0] == BC_EXT_DIR || a.sbc[c][1][0] == BC_HOEXTRAP, byh = a.sbc[c][1][1] == BC_EXT_DIR || a.sbc[c][1][1] == BC_HOEXTRAP;
                const bool bzl = a.sbc[c][2][0] == BC_EXT_DIR || a.sbc[c][2][0] == BC_HOEXTRAP, bzh = a.sbc[c][2][1] == BC_EXT_DIR || a.sbc[c][2][1] == BC_HOEXTRAP;
                if ((bxl && i == 0) || (bxh && i == n0 - 1)) slx[c] = px[c].fromm;
                if ((bxl && i == -1) || (bxh && i == n0)) slx[c] = ZERO;
                if ((byl && j == 0) || (byh && j == n1 - 1)) sly[c] = py[c].fromm;
                if ((byl && j == -1) || (byh && j == n1)) sly[c] = ZERO;
                if ((bzl && k == 0) || (bzh && k == n2 - 1)) slz[c] = pz0[c].fromm;
                if ((bzl && k == -1) || (bzh && k == n2)) slz[c] = ZERO;
            }
        }
        // 1-D extrapolation to the faces of cell (i,j,k): mkflux.f90:1446-1447 (x), :1533-1534 (y), :1782-1783 (z)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int comp = a.comp0 + c;
            lx[c] = s0[c] + (HALF - txh) * slx[c]; rx[c] = s0[c] - (HALF + txl) * slx[c];
            ly[c] = s0[c] + (HALF - tyh) * sly[c]; ry[c] = s0[c] - (HALF + tyl) * sly[c];
            lz[c] = s0[c] + (HALF - tzh) * slz[c]; rz[c] = s0[c] - (HALF + tzl) * slz[c];
            if (minion) {
                lx[c] = lx[c] + fk[c]; rx[c] = rx[c] + fk[c]; ly[c] = ly[c] + fk[c]; ry[c] = ry[c] + fk[c]; lz[c] = lz[c] + fk[c]; rz[c] = rz[c] + fk[c];
            }
            if (bnd) {
                if (xlo) rx[c] = mf_bc_lo(rx[c], sxm[c], 0, bcx0, a.is_vel, comp);
                if (xhi_c) lx[c] = mf_bc_hi(lx[c], sxp[c], 0, bcx1, a.is_vel, comp);
                if (ylo) ry[c] = mf_bc_lo(ry[c], sym[c], 1, bcy0, a.is_vel, comp);
                if (yhi_c) ly[c] = mf_bc_hi(ly[c], syp[c], 1, bcy1, a.is_vel, comp);
                if (zlo) rz[c] = mf_bc_lo(rz[c], a.s[c][so + (k - 1) * a.s_sz], 2, bcz0, a.is_vel, comp);
                if (zhi_c) lz[c] = mf_bc_hi(lz[c], sp1[c], 2, bcz1, a.is_vel, comp);
            }
            Y.put(NC + c, ly[c]);
        }
        row_sync<TYT>(ty);                                                                  // B2
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double lxi = shfl_up1(lx[c]), lyi = Y.from_lo(NC + c), lzi = Y.get(KB + MK_LZH * NC + c);
            if (bnd) {
                if (xlo) lxi = rx[c];
                if (xhi_f) rx[c] = lxi;
                if (ylo) lyi = ry[c];
                if (yhi_f) ry[c] = lyi;
                if (zlo) lzi = rz[c];
                if (zhi_f) rz[c] = lzi;
            }
            qx[c] = upw_p(lxi, rx[c], Px); qy[c] = upw_p(lyi, ry[c], Py); qz[c] = upw_p(lzi, rz[c], Pz);
            Y.put(2 * NC + c, qy[c]);
        }
        row_sync<TYT>(ty);                                                                  // B3
        // ================= T: transverse terms and the six once-corrected states =================
        double Xxy[NC], Xyx[NC], Xzx[NC], Xzy[NC], Xxz[NC], Xyz[NC];
        double rxy[NC], ryx[NC], rzx[NC], rzy[NC], rxz[NC], ryz[NC], lxy[NC], lyx[NC], lxz[NC], lyz[NC];
        double lzxi[NC], lzyi[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int comp = a.comp0 + c;
            const bool cons = (CONSMASK >> c) & 1;
            const double qxh = shfl_dn1(qx[c]), qyh = Y.from_hi(2 * NC + c);
            const double qzP = Y.get(KB + MK_QZP * NC + c);
            const double lxP = Y.get(KB + (pp + MK_LXP) * NC + c), rxP = Y.get(KB + (pp + MK_RXP) * NC + c);
            const double lyP = Y.get(KB + (pp + MK_LYP) * NC + c), ryP = Y.get(KB + (pp + MK_RYP) * NC + c);
            lzxi[c] = Y.get(KB + MK_LZXH * NC + c); lzyi[c] = Y.get(KB + MK_LZYH * NC + c);
            double ttx, tty, ttz;
            if (cons) {          // mkflux.f90:1620-1622
                ttx = a.c3[0] * (qxh * umh - qx[c] * uml);
                tty = a.c3[1] * (qyh * vmh - qy[c] * vml);
                ttz = a.c3[2] * (qz[c] * wml - qzP * wmlP);
            } else {             // :1624-1626
                ttx = a.c6[0] * sux * (qxh - qx[c]);
                tty = a.c6[1] * suy * (qyh - qy[c]);
                ttz = a.c6[2] * suzP * (qz[c] - qzP);
            }
            rxy[c] = rx[c] - tty; lxy[c] = lx[c] - tty;
            ryx[c] = ry[c] - ttx; lyx[c] = ly[c] - ttx;
            rzx[c] = rz[c] - ttx; double lzxN = lz[c] - ttx;
            rzy[c] = rz[c] - tty; double lzyN = lz[c] - tty;
            rxz[c] = rxP - ttz; lxz[c] = lxP - ttz;
            ryz[c] = ryP - ttz; lyz[c] = lyP - ttz;
            if (bnd) {
                if (xlo || xhi_c) {
                    const double *sb = a.s[c] + so + k * a.s_sz;
                    const double sgP = sb[(xlo ? -1 : 1) - a.s_sz];
                    if (xlo) { rxy[c] = mf_bc_lo(rxy[c], sxm[c], 0, bcx0, a.is_vel, comp); rxz[c] = mf_bc_lo(rxz[c], sgP, 0, bcx0, a.is_vel, comp); }
                    else     { lxy[c] = mf_bc_hi(lxy[c], sxp[c], 0, bcx1, a.is_vel, comp); lxz[c] = mf_bc_hi(lxz[c], sgP, 0, bcx1, a.is_vel, comp); }
                }
                if (ylo || yhi_c) {
                    const double *sb = a.s[c] + so + k * a.s_sz;
                    const double sgP = sb[(ylo ? -a.s_sy : a.s_sy) - a.s_sz];
                    if (ylo) { ryx[c] = mf_bc_lo(ryx[c], sym[c], 1, bcy0, a.is_vel, comp); ryz[c] = mf_bc_lo(ryz[c], sgP, 1, bcy0, a.is_vel, comp); }
                    else     { lyx[c] = mf_bc_hi(lyx[c], syp[c], 1, bcy1, a.is_vel, comp); lyz[c] = mf_bc_hi(lyz[c], sgP, 1, bcy1, a.is_vel, comp); }
                }
                if (zlo) { const double sg = a.s[c][so + (k - 1) * a.s_sz];
                           rzx[c] = mf_bc_lo(rzx[c], sg, 2, bcz0, a.is_vel, comp); rzy[c] = mf_bc_lo(rzy[c], sg, 2, bcz0, a.is_vel, comp); }
                if (zhi_c) { lzxN = mf_bc_hi(lzxN, sp1[c], 2, bcz1, a.is_vel, comp); lzyN = mf_bc_hi(lzyN, sp1[c], 2, bcz1, a.is_vel, comp); }
            }
            // the L states this cell hands to z-face k+1, and the 1-D states the next iteration needs as "plane k-1"
            Y.put(KB + MK_LZXH * NC + c, lzxN); Y.put(KB + MK_LZYH * NC + c, lzyN);
            Y.put(KB + (pc + MK_LXP) * NC + c, lx[c]); Y.put(KB + (pc + MK_RXP) * NC + c, rx[c]);
            Y.put(KB + (pc + MK_LYP) * NC + c, ly[c]); Y.put(KB + (pc + MK_RYP) * NC + c, ry[c]);
            Y.put(KB + MK_LZH * NC + c, lz[c]); Y.put(KB + MK_QZP * NC + c, qz[c]);
            Y.put(3 * NC + c, lyx[c]); Y.put(c, lyz[c]);
        }
        row_sync<TYT>(ty);                                                                  // B4
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double lxyi = shfl_up1(lxy[c]), lxzi = shfl_up1(lxz[c]);
            double lyxi = Y.from_lo(3 * NC + c), lyzi = Y.from_lo(c);
            if (bnd) {
                if (xlo) { lxyi = rxy[c]; lxzi = rxz[c]; }
                if (xhi_f) { rxy[c] = lxyi; rxz[c] = lxzi; }
                if (ylo) { lyxi = ryx[c]; lyzi = ryz[c]; }
                if (yhi_f) { ryx[c] = lyxi; ryz[c] = lyzi; }
                if (zlo) { lzxi[c] = rzx[c]; lzyi[c] = rzy[c]; }
                if (zhi_f) { rzx[c] = lzxi[c]; rzy[c] = lzyi[c]; }
            }
            Xxy[c] = upw_p(lxyi, rxy[c], Px); Xyx[c] = upw_p(lyxi, ryx[c], Py);
            Xzx[c] = upw_p(lzxi[c], rzx[c], Pz); Xzy[c] = upw_p(lzyi[c], rzy[c], Pz);
            Xxz[c] = upw_p(lxzi, rxz[c], PxP); Xyz[c] = upw_p(lyzi, ryz[c], PyP);
            Y.put(NC + c, Xyx[c]); Y.put(2 * NC + c, Xyz[c]);
        }
        row_sync<TYT>(ty);                                                                  // B5
        // ================= F: final edge states =================
        double Ly[NC], Ry[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int comp = a.comp0 + c;
            const bool cons = (CONSMASK >> c) & 1;
            const double Xxyh = shfl_dn1(Xxy[c]), Xxzh = shfl_dn1(Xxz[c]);
            const double Xyxh = Y.from_hi(NC + c), Xyzh = Y.from_hi(2 * NC + c);
            const double LzH = Y.get(KB + MK_LZF * NC + c), fP = Y.get(KB + MK_FP * NC + c);
            const double XzxP = Y.get(KB + MK_XZXP * NC + c), XzyP = Y.get(KB + MK_XZYP * NC + c);
            // ---- sedge_z on face k: mkflux.f90:1870-1972 ----
            double Rz, LzN;
            if (cons) {
                const double A1 = a.c2[0] * (Xxyh * umh - Xxy[c] * uml), A2 = a.c2[1] * (Xyxh * vmh - Xyx[c] * vml);
                const double B1 = a.c2[0] * s0[c] * (umh - uml), B2 = a.c2[1] * s0[c] * (vmh - vml);
                Rz = rz[c] - A1 - A2 + B1 + B2; LzN = lz[c] - A1 - A2 + B1 + B2;
            } else {
                const double A1 = a.c4[0] * sux * (Xxyh - Xxy[c]), A2 = a.c4[1] * suy * (Xyxh - Xyx[c]);
                Rz = rz[c] - A1 - A2; LzN = lz[c] - A1 - A2;
            }
            if (!minion) { Rz = Rz + fk[c]; LzN = LzN + fk[c]; }
            if (k >= ka && (k < kb || lastchunk)) {
                double v = upw_p(LzH, Rz, Pz);
                if (zlo) v = mf_bc_lo(Rz, a.s[c][so + (k - 1) * a.s_sz], 2, bcz0, a.is_vel, comp);
                if (zhi_f) v = mf_bc_hi(LzH, s0[c], 2, bcz1, a.is_vel, comp);
                if (st_z) {
                    a.sedge[2][c][eo2 + k * a.e_sz[2]] = v;
                    if (cons) a.flux[2][c][eo2 + k * a.e_sz[2]] = v * wml;
                }
            }
            Y.put(KB + MK_LZF * NC + c, LzN); Y.put(KB + MK_FP * NC + c, fk[c]);
            Y.put(KB + MK_XZXP * NC + c, Xzx[c]); Y.put(KB + MK_XZYP * NC + c, Xzy[c]);
            // ---- sedge_x, sedge_y on plane k-1: mkflux.f90:2310-2408, :2414-2512 ----
            const double lxP = Y.get(KB + (pp + MK_LXP) * NC + c), rxP = Y.get(KB + (pp + MK_RXP) * NC + c);
            const double lyP = Y.get(KB + (pp + MK_LYP) * NC + c), ryP = Y.get(KB + (pp + MK_RYP) * NC + c);
            double Rx, Lx;
            if (cons) {
                const double sP = a.s[c][so + (k - 1) * a.s_sz];
                const double A1 = a.c2[1] * (Xyzh * vmhP - Xyz[c] * vmlP), A2 = a.c2[2] * (Xzy[c] * wml - XzyP * wmlP);
                const double B1 = a.c2[1] * sP * (vmhP - vmlP), B2 = a.c2[2] * sP * (wml - wmlP);
                Rx = rxP - A1 - A2 + B1 + B2; Lx = lxP - A1 - A2 + B1 + B2;
                const double C1 = a.c2[0] * (Xxzh * umhP - Xxz[c] * umlP), C2 = a.c2[2] * (Xzx[c] * wml - XzxP * wmlP);
                const double D1 = a.c2[0] * sP * (umhP - umlP), D2 = a.c2[2] * sP * (wml - wmlP);
                Ry[c] = ryP - C1 - C2 + D1 + D2; Ly[c] = lyP - C1 - C2 + D1 + D2;
            } else {
                const double A1 = a.c4[1] * suyP * (Xyzh - Xyz[c]), A2 = a.c4[2] * suzP * (Xzy[c] - XzyP);
                Rx = rxP - A1 - A2; Lx = lxP - A1 - A2;
                const double C1 = a.c4[0] * suxP * (Xxzh - Xxz[c]), C2 = a.c4[2] * suzP * (Xzx[c] - XzxP);
                Ry[c] = ryP - C1 - C2; Ly[c] = lyP - C1 - C2;
            }
            if (!minion) { Rx = Rx + fP; Lx = Lx + fP; Ry[c] = Ry[c] + fP; Ly[c] = Ly[c] + fP; }
            Y.put(3 * NC + c, Ly[c]);
            const double Lxi = shfl_up1(Lx);
            if (k > ka) {
                double v = upw_p(Lxi, Rx, PxP);
                if (bnd && (xlo || xhi_f)) {
                    const double *sb = a.s[c] + so + (k - 1) * a.s_sz;
                    if (xlo) v = mf_bc_lo(Rx, sb[-1], 0, bcx0, a.is_vel, comp);
                    else     v = mf_bc_hi(Lxi, sb[0], 0, bcx1, a.is_vel, comp);
                }
                if (st_x) {
                    a.sedge[0][c][eo0 + (k - 1) * a.e_sz[0]] = v;
                    if (cons) a.flux[0][c][eo0 + (k - 1) * a.e_sz[0]] = v * umlP;
                }
            }
        }
        row_sync<TYT>(ty);                                                                  // B6
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int comp = a.comp0 + c;
            const bool cons = (CONSMASK >> c) & 1;
            if (k > ka) {
                const double Lyi = Y.from_lo(3 * NC + c);
                double v = upw_p(Lyi, Ry[c], PyP);
                if (bnd && (ylo || yhi_f)) {
                    const double *sb = a.s[c] + so + (k - 1) * a.s_sz;
                    if (ylo) v = mf_bc_lo(Ry[c], sb[-a.s_sy], 1, bcy0, a.is_vel, comp);
                    else     v = mf_bc_hi(Lyi, sb[0], 1, bcy1, a.is_vel, comp);
                }
                if (st_y) {
                    a.sedge[1][c][eo1 + (k - 1) * a.e_sz[1]] = v;
                    if (cons) a.flux[1][c][eo1 + (k - 1) * a.e_sz[1]] = v * vmlP;
                }
            }
            // ---- rotate the register-carried state ----
            s0[c] = sp1[c]; sp1[c] = sp2[c]; sp2[c] = snext[c];
            frzm[c] = pz0[c].fromm; pz0[c].cen = pz1[c].cen; pz0[c].lim = pz1[c].lim; pz0[c].fromm = pz1[c].fromm;
        }
        umlP = uml; umhP = umh; vmlP = vml; vmhP = vmh; wmlP = wml; wml = wmh; PxP = Px; PyP = Py;
    }
}

// ------------------------------------------------------------------------------------------
// velpred_3d.  Same march; the advecting velocity of a face is the Riemann value uimh_D(D) of the normal extrapolation, so the
// "MAC sums" of the transverse terms come out of the N stage instead of being loaded.  All three components travel together.
// ------------------------------------------------------------------------------------------
struct VpmArgs {
    Geo g;
    const double *u[3]; int u_sy, u_sz;
    const double *force[3]; int f_sy, f_sz;
    double *out[3]; int o_sy[3], o_sz[3];
    const double *eps;
    int order, use_minion;
    int sbc[3][3][2];
    int zchunk;
    double dt2, c4[3], c6[3], hinv[3]; int hp2[3];
};
// keep slots (per-thread state across iterations, see YSlots): the 1-D states of plane k-1 (comps u, v; double-buffered by iteration parity)
// and the z-face states of face k-1
enum { VK_LX0 = 0, VK_LX1, VK_RX0, VK_RX1, VK_LY0, VK_LY1, VK_RY0, VK_RY1, VK_LZH0 = 16, VK_LZH1, VK_LZH2, VK_LZXH, VK_LZYH, VK_LZF,
       VK_QZ0, VK_QZ1, VK_QZ2, VK_XZXP, VK_XZYP, VK_FPX, VK_FPY,
       // the z window and z slope parts (7 per component): only the N stage touches them, so they are parked here during T and F
       VK_ZP0, VK_NKEEP = VK_ZP0 + 21 };
constexpr int VP_EX = 6;                                // exchange slots
constexpr int vp_smem_slots() { return VP_EX + VK_NKEEP; }

// velpred.f90:2044-2079 on the surviving state of a boundary face (normal extrapolation)
__device__ __forceinline__ double vp_bcn_lo(double v, double ug, bool isn, int bc)
{
    if (bc == BC_INLET) return ug;
    if (bc == BC_SLIP_WALL) return isn ? ZERO : v;
    if (bc == BC_NO_SLIP_WALL) return ZERO;
    if (bc == BC_OUTLET) return isn ? fmin(v, ZERO) : v;
    return v;
}
__device__ __forceinline__ double vp_bcn_hi(double v, double ug, bool isn, int bc, bool outlet_min)
{
    if (bc == BC_INLET) return ug;
    if (bc == BC_SLIP_WALL) return isn ? ZERO : v;
    if (bc == BC_NO_SLIP_WALL) return ZERO;
    if (bc == BC_OUTLET) return isn ? (outlet_min ? fmin(v, ZERO) : fmax(v, ZERO)) : v;      // velpred.f90:2075
    return v;
}
// velpred.f90:2202-2221 (transverse-once states): same rule on both sides
__device__ __forceinline__ double vp_bct(double v, double ug, int bc)
{
    if (bc == BC_INLET) return ug;
    if (bc == BC_NO_SLIP_WALL) return ZERO;
    return v;
}
__device__ __forceinline__ double riemann_p(double l, double r, double eps)      // velpred.f90:2084-2088
{
    const double uavg = HALF * (l + r);
    const bool test = ((l <= ZERO && r >= ZERO) || (fabs(l + r) < eps));
    const double v = (uavg > ZERO) ? l : r;
    return test ? ZERO : v;
}
struct TanP { bool pos, small; };
__device__ __forceinline__ TanP tan_pred(double un, double eps) { TanP p; p.pos = un > ZERO; p.small = fabs(un) < eps; return p; }
__device__ __forceinline__ double upt_p(double l, double r, TanP p)              // velpred.f90:2091-2093
{
    const double v = p.pos ? l : r;
    const double uavg = HALF * (l + r);
    return p.small ? uavg : v;
}

template <int TYT, bool GEN, bool HP2>
__global__ void __launch_bounds__(TXT * TYT, MARCH_MINB) k_velpred_march(const VpmArgs a)
{
    VDN_DYN_SMEM(smem);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int n0 = a.g.n[0], n1 = a.g.n[1], n2 = a.g.n[2];
    const int it = (int)blockIdx.x * TXO - 1 + tx, jt = (int)blockIdx.y * (TYT - 2) - 1 + ty;
    const int i = it < n0 ? it : n0, j = jt < n1 ? jt : n1;
    const int is = it < n0 + 1 ? it : n0 + 1, js = jt < n1 + 1 ? jt : n1 + 1;
    const int ka = (int)blockIdx.z * a.zchunk;
    const int kb = (ka + a.zchunk < n2) ? ka + a.zchunk : n2;
    const bool lastchunk = kb == n2;
    const YSlots<TYT> Y(smem, tx, ty);
    const int order = GEN ? a.order : 4;
    const bool minion = GEN ? (a.use_minion != 0) : false;

    const bool inner = tx >= 1 && tx <= TXT - 2 && ty >= 1 && ty <= TYT - 2 && it <= n0 && jt <= n1;
    const bool st_x = inner && jt < n1, st_y = inner && it < n0, st_z = inner && it < n0 && jt < n1;
    const int bcx0 = a.g.pbc[0][0], bcx1 = a.g.pbc[0][1], bcy0 = a.g.pbc[1][0], bcy1 = a.g.pbc[1][1], bcz0 = a.g.pbc[2][0], bcz1 = a.g.pbc[2][1];
    const bool xlo = bc_overrides(bcx0) && i == 0, xhi_c = bc_overrides(bcx1) && i == n0 - 1, xhi_f = bc_overrides(bcx1) && i == n0;
    const bool ylo = bc_overrides(bcy0) && j == 0, yhi_c = bc_overrides(bcy1) && j == n1 - 1, yhi_f = bc_overrides(bcy1) && j == n1;
    const bool tile_x = (bc_overrides(bcx0) && blockIdx.x == 0) || (bc_overrides(bcx1) && (int)blockIdx.x * TXO + TXT >= n0);
    const bool tile_y = (bc_overrides(bcy0) && blockIdx.y == 0) || (bc_overrides(bcy1) && (int)blockIdx.y * (TYT - 2) + TYT >= n1);
    const bool tile_xy = tile_x || tile_y;

    const int uo = is + a.u_sy * js, fo = i + a.f_sy * j;
    const int oo0 = i + a.o_sy[0] * j, oo1 = i + a.o_sy[1] * j, oo2 = i + a.o_sy[2] * j;
    const int nbt = a.g.nb[0] * a.g.nb[1] * a.g.nb[2];
    const int bij = nbt == 1 ? 0 : a.g.box1(0, i) + a.g.nb[0] * a.g.box1(1, j);

    auto divh = [&](double x, int d) { return HP2 ? x * a.hinv[d] : div_h(x, a.g.h[d], a.hinv[d], a.hp2[d]); };
    // ---- state carried in registers: the z window and the z slope parts; everything else of plane / face k-1 lives in the keep slots ----
    double nsxP = 0, nsyP = 0, epsP = 0;
    TanP tpxP = { false, false }, tpyP = { false, false };
    auto KS = [&](int q) { return VP_EX + q; };

    {
        const int k = ka - 1;
        double s0[3], sp1[3], sp2[3]; PartsZ pz0[3]; double frzm[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double *sb = a.u[c] + uo;
            const double sm3 = sb[(k - 2) * a.u_sz], sm2 = sb[(k - 1) * a.u_sz];
            s0[c] = sb[k * a.u_sz]; sp1[c] = sb[(k + 1) * a.u_sz]; sp2[c] = sb[(k + 2) * a.u_sz];
            const Parts p0 = slope_parts3(sm2, s0[c], sp1[c]);
            pz0[c].cen = p0.cen; pz0[c].lim = p0.lim; pz0[c].fromm = p0.fromm;
            frzm[c] = slope_parts3(sm3, sm2, s0[c]).fromm;
        }
#pragma unroll
        for (int q = 0; q < VK_ZP0; ++q) Y.put(KS(q), ZERO);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Y.put(KS(VK_ZP0 + 7 * c), s0[c]); Y.put(KS(VK_ZP0 + 7 * c + 1), sp1[c]); Y.put(KS(VK_ZP0 + 7 * c + 2), sp2[c]);
            Y.put(KS(VK_ZP0 + 7 * c + 3), pz0[c].cen); Y.put(KS(VK_ZP0 + 7 * c + 4), pz0[c].lim); Y.put(KS(VK_ZP0 + 7 * c + 5), pz0[c].fromm);
            Y.put(KS(VK_ZP0 + 7 * c + 6), frzm[c]);
        }
    }

    for (int k = ka - 1; k <= kb; ++k) {
        const bool zlo = bc_overrides(bcz0) && k == 0, zhi_c = bc_overrides(bcz1) && k == n2 - 1, zhi_f = bc_overrides(bcz1) && k == n2;
        const bool bnd = tile_xy || (bc_overrides(bcz0) && k <= 1) || (bc_overrides(bcz1) && k >= n2 - 2);
        const double eps = a.eps[nbt == 1 ? 0 : bij + a.g.nb[0] * a.g.nb[1] * a.g.box1(2, k)];
        const int pc = 8 * (k & 1), pp = 8 - pc;                      // keep-slot sets of plane k / plane k-1

        // ---- z direction first, one component at a time: un-park the z window / parts of cell k, finish the z slope, park the rotated
        // state for the next iteration -- nothing of it stays in registers except u(k) itself ----
        double s0[3], slz[3], fk[3];
        Parts px[3], py[3];
        double sxm[3], sxp[3], sym[3], syp[3], frx_e[3], fry_e[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double *sb = a.u[c] + uo + k * a.u_sz;
            const int k3 = (k + 3 <= n2 + 2) ? 3 : 2;
            const double snext = sb[k3 * a.u_sz];
            s0[c] = Y.get(KS(VK_ZP0 + 7 * c));
            const double sp1 = Y.get(KS(VK_ZP0 + 7 * c + 1)), sp2 = Y.get(KS(VK_ZP0 + 7 * c + 2));
            Parts pzc; pzc.cen = Y.get(KS(VK_ZP0 + 7 * c + 3)); pzc.lim = Y.get(KS(VK_ZP0 + 7 * c + 4)); pzc.fromm = Y.get(KS(VK_ZP0 + 7 * c + 5));
            pzc.flag = copysign(ONE, pzc.cen);
            const double frzm = Y.get(KS(VK_ZP0 + 7 * c + 6));
            Parts pz1 = slope_parts3(s0[c], sp1, sp2);
            const bool bzl = a.sbc[c][2][0] == BC_EXT_DIR || a.sbc[c][2][0] == BC_HOEXTRAP, bzh = a.sbc[c][2][1] == BC_EXT_DIR || a.sbc[c][2][1] == BC_HOEXTRAP;
            if (bnd && order != 0) {       // one-sided slopes next to an EXT_DIR / HOEXTRAP z face (pz1 belongs to cell k+1)
                if (bzl && k + 1 == 0)      pz1.fromm = slope_onesided(s0[c], sp1, sp2, snext, false, order);
                if (bzh && k + 1 == n2 - 1) pz1.fromm = slope_onesided(sp2, sp1, s0[c], sb[-a.u_sz], true, order);
            }
            slz[c] = order == 4 ? slope4_from(pzc, frzm, pz1.fromm) : (order == 2 ? pzc.fromm : ZERO);
            if (bnd && order != 0) {
                if ((bzl && k == 0) || (bzh && k == n2 - 1)) slz[c] = pzc.fromm;
                if ((bzl && k == -1) || (bzh && k == n2)) slz[c] = ZERO;
            }
            Y.put(KS(VK_ZP0 + 7 * c), sp1); Y.put(KS(VK_ZP0 + 7 * c + 1), sp2); Y.put(KS(VK_ZP0 + 7 * c + 2), snext);
            Y.put(KS(VK_ZP0 + 7 * c + 3), pz1.cen); Y.put(KS(VK_ZP0 + 7 * c + 4), pz1.lim); Y.put(KS(VK_ZP0 + 7 * c + 5), pz1.fromm);
            Y.put(KS(VK_ZP0 + 7 * c + 6), pzc.fromm);
            // ---- x, y: parts of this cell; fromm of the neighbours comes by shuffle / shared memory after the barrier ----
            sxm[c] = sb[-1]; sxp[c] = sb[1]; sym[c] = sb[-a.u_sy]; syp[c] = sb[a.u_sy];
            fk[c] = a.dt2 * a.force[c][fo + k * a.f_sz];
            frx_e[c] = fry_e[c] = ZERO;
            if (order != 0) {
                px[c] = slope_parts3(sxm[c], s0[c], sxp[c]);
                py[c] = slope_parts3(sym[c], s0[c], syp[c]);
                if (order == 4) {
                    if (tx == 0 || tx == TXT - 1) {
                        const double s2 = sb[tx == 0 ? -2 : (it <= n0 ? 2 : 1)];
                        frx_e[c] = (tx == 0) ? slope_parts3(s2, sxm[c], s0[c]).fromm : slope_parts3(s0[c], sxp[c], s2).fromm;
                    }
                    if (ty == 0 || ty == TYT - 1) {
                        const double s2 = sb[ty == 0 ? -2 * a.u_sy : (jt <= n1 ? 2 : 1) * a.u_sy];
                        fry_e[c] = (ty == 0) ? slope_parts3(s2, sym[c], s0[c]).fromm : slope_parts3(s0[c], syp[c], s2).fromm;
                    }
                }
            }
            if (bnd && order != 0) {
                const bool bxl = a.sbc[c][0][0] == BC_EXT_DIR || a.sbc[c][0][0] == BC_HOEXTRAP, bxh = a.sbc[c][0][1] == BC_EXT_DIR || a.sbc[c][0][1] == BC_HOEXTRAP;
                const bool byl = a.sbc[c][1][0] == BC_EXT_DIR || a.sbc[c][1][0] == BC_HOEXTRAP, byh = a.sbc[c][1][1] == BC_EXT_DIR || a.sbc[c][1][1] == BC_HOEXTRAP;
                if (bxl && i == 0)      px[c].fromm = slope_onesided(sxm[c], s0[c], sxp[c], sb[2], false, order);
                if (bxh && i == n0 - 1) px[c].fromm = slope_onesided(sxp[c], s0[c], sxm[c], sb[-2], true, order);
                if (bxh && tx == TXT - 1 && i + 1 == n0 - 1) frx_e[c] = slope_onesided(sb[2], sxp[c], s0[c], sxm[c], true, order);
                if (byl && j == 0)      py[c].fromm = slope_onesided(sym[c], s0[c], syp[c], sb[2 * a.u_sy], false, order);
                if (byh && j == n1 - 1) py[c].fromm = slope_onesided(syp[c], s0[c], sym[c], sb[-2 * a.u_sy], true, order);
                if (byh && ty == TYT - 1 && j + 1 == n1 - 1) fry_e[c] = slope_onesided(sb[2 * a.u_sy], syp[c], s0[c], sym[c], true, order);
            }
        }
        if (order == 4) {
#pragma unroll
            for (int c = 0; c < 3; ++c) Y.put(c, py[c].fromm);
        }
        row_sync<TYT>(ty);                                                                  // B1
        double slx[3], sly[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (order == 4) {
                double fxm = shfl_up1(px[c].fromm), fxp = shfl_dn1(px[c].fromm);
                if (tx == 0) fxm = frx_e[c];
                if (tx == TXT - 1) fxp = frx_e[c];
                double fym = Y.from_lo(c), fyp = Y.from_hi(c);
                if (ty == 0) fym = fry_e[c];
                if (ty == TYT - 1) fyp = fry_e[c];
                slx[c] = slope4_from(px[c], fxm, fxp);
                sly[c] = slope4_from(py[c], fym, fyp);
            } else if (order == 2) {
                slx[c] = px[c].fromm; sly[c] = py[c].fromm;
            } else { slx[c] = sly[c] = ZERO; }
        }
        if (bnd && order != 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const bool bxl = a.sbc[c][0][0] == BC_EXT_DIR || a.sbc[c][0][0] == BC_HOEXTRAP, bxh = a.sbc[c][0][1] == BC_EXT_DIR || a.sbc[c][0][1] == BC_HOEXTRAP;
                const bool byl = a.sbc[c][1][0] == BC_EXT_DIR || a.sbc[c][1][0] == BC_HOEXTRAP, byh = a.sbc[c][1][1] == BC_EXT_DIR || a.sbc[c][1][1] == BC_HOEXTRAP;
                if ((bxl && i == 0) || (bxh && i == n0 - 1)) slx[c] = px[c].fromm;
                if ((bxl && i == -1) || (bxh && i == n0)) slx[c] = ZERO;
                if ((byl && j == 0) || (byh && j == n1 - 1)) sly[c] = py[c].fromm;
                if ((byl && j == -1) || (byh && j == n1)) sly[c] = ZERO;
            }
        }
        // normal extrapolation of all comps to the six faces of the cell: velpred.f90:2022-2029 (x), :2108-2115 (y), :2286-2293 (z).
        // Operation-order quirks: x, z: dt2*max(0,u)/h ; y left state: dt2*max(0,u/h)
        const double clx = HALF - divh(a.dt2 * fmax(ZERO, s0[0]), 0);
        const double crx = HALF + divh(a.dt2 * fmin(ZERO, s0[0]), 0);
        const double cly = HALF - a.dt2 * fmax(ZERO, divh(s0[1], 1));
        const double cry = HALF + divh(a.dt2 * fmin(ZERO, s0[1]), 1);
        const double clz = HALF - divh(a.dt2 * fmax(ZERO, s0[2]), 2);
        const double crz = HALF + divh(a.dt2 * fmin(ZERO, s0[2]), 2);
        double lx[3], rx[3], ly[3], ry[3], lz[3], rz[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            lx[c] = s0[c] + clx * slx[c]; rx[c] = s0[c] - crx * slx[c];
            ly[c] = s0[c] + cly * sly[c]; ry[c] = s0[c] - cry * sly[c];
            lz[c] = s0[c] + clz * slz[c]; rz[c] = s0[c] - crz * slz[c];
            if (minion) {
                lx[c] = lx[c] + fk[c]; rx[c] = rx[c] + fk[c]; ly[c] = ly[c] + fk[c]; ry[c] = ry[c] + fk[c]; lz[c] = lz[c] + fk[c]; rz[c] = rz[c] + fk[c];
            }
            if (bnd) {
                if (xlo) rx[c] = vp_bcn_lo(rx[c], sxm[c], c == 0, bcx0);
                if (xhi_c) lx[c] = vp_bcn_hi(lx[c], sxp[c], c == 0, bcx1, true);
                if (ylo) ry[c] = vp_bcn_lo(ry[c], sym[c], c == 1, bcy0);
                if (yhi_c) ly[c] = vp_bcn_hi(ly[c], syp[c], c == 1, bcy1, false);
                if (zlo) rz[c] = vp_bcn_lo(rz[c], a.u[c][uo + (k - 1) * a.u_sz], c == 2, bcz0);
                if (zhi_c) lz[c] = vp_bcn_hi(lz[c], a.u[c][uo + (k + 1) * a.u_sz], c == 2, bcz1, false);
            }
            Y.put(3 + c, ly[c]);
        }
        row_sync<TYT>(ty);                                                                  // B2
        double qx[3], qy[3], qz[3];                      // uimh at the lo faces
        {
            double lxi[3], lyi[3], lzi[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lxi[c] = shfl_up1(lx[c]); lyi[c] = Y.from_lo(3 + c); lzi[c] = Y.get(KS(VK_LZH0 + c));
                if (bnd) {
                    if (xlo) lxi[c] = rx[c];
                    if (xhi_f) rx[c] = lxi[c];
                    if (ylo) lyi[c] = ry[c];
                    if (yhi_f) ry[c] = lyi[c];
                    if (zlo) lzi[c] = rz[c];
                    if (zhi_f) rz[c] = lzi[c];
                }
            }
            qx[0] = riemann_p(lxi[0], rx[0], eps); qy[1] = riemann_p(lyi[1], ry[1], eps); qz[2] = riemann_p(lzi[2], rz[2], eps);
            const TanP tpx = tan_pred(qx[0], eps), tpy = tan_pred(qy[1], eps), tpz = tan_pred(qz[2], eps);
            qx[1] = upt_p(lxi[1], rx[1], tpx); qx[2] = upt_p(lxi[2], rx[2], tpx);
            qy[0] = upt_p(lyi[0], ry[0], tpy); qy[2] = upt_p(lyi[2], ry[2], tpy);
            qz[0] = upt_p(lzi[0], rz[0], tpz); qz[1] = upt_p(lzi[1], rz[1], tpz);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) Y.put(c, qy[c]);
        row_sync<TYT>(ty);                                                                  // B3
        // ================= T =================
        double qxh[3], qyh[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { qxh[c] = shfl_dn1(qx[c]); qyh[c] = Y.from_hi(c); }
        double qzP[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) qzP[c] = Y.get(KS(VK_QZ0 + c));
        const double nsx = qxh[0] + qx[0], nsy = qyh[1] + qy[1], nszP = qz[2] + qzP[2];
        const double ttx_y = a.c6[0] * nsx * (qxh[1] - qx[1]), ttx_z = a.c6[0] * nsx * (qxh[2] - qx[2]);
        const double tty_x = a.c6[1] * nsy * (qyh[0] - qy[0]), tty_z = a.c6[1] * nsy * (qyh[2] - qy[2]);
        const double ttz_x = a.c6[2] * nszP * (qz[0] - qzP[0]), ttz_y = a.c6[2] * nszP * (qz[1] - qzP[1]);
        // wimhxy (x face, w, by y), wimhyx (y face, w, by x), vimhzx (z face, v, by x), uimhzy (z face, u, by y),
        // vimhxz (x face, v, by z; plane k-1), uimhyz (y face, u, by z; plane k-1)
        double rxy = rx[2] - tty_z, lxy = lx[2] - tty_z;
        double ryx = ry[2] - ttx_z, lyx = ly[2] - ttx_z;
        double rzx = rz[1] - ttx_y, lzxN = lz[1] - ttx_y;
        double rzy = rz[0] - tty_x, lzyN = lz[0] - tty_x;
        double rxz = Y.get(KS(pp + VK_RX1)) - ttz_y, lxz = Y.get(KS(pp + VK_LX1)) - ttz_y;
        double ryz = Y.get(KS(pp + VK_RY0)) - ttz_x, lyz = Y.get(KS(pp + VK_LY0)) - ttz_x;
        if (bnd) {
            if (xlo || xhi_c) {
                const int o = uo + k * a.u_sz + (xlo ? -1 : 1);
                if (xlo) { rxy = vp_bct(rxy, a.u[2][o], bcx0); rxz = vp_bct(rxz, a.u[1][o - a.u_sz], bcx0); }
                else     { lxy = vp_bct(lxy, a.u[2][o], bcx1); lxz = vp_bct(lxz, a.u[1][o - a.u_sz], bcx1); }
            }
            if (ylo || yhi_c) {
                const int o = uo + k * a.u_sz + (ylo ? -a.u_sy : a.u_sy);
                if (ylo) { ryx = vp_bct(ryx, a.u[2][o], bcy0); ryz = vp_bct(ryz, a.u[0][o - a.u_sz], bcy0); }
                else     { lyx = vp_bct(lyx, a.u[2][o], bcy1); lyz = vp_bct(lyz, a.u[0][o - a.u_sz], bcy1); }
            }
            if (zlo) { const int o = uo + (k - 1) * a.u_sz; rzx = vp_bct(rzx, a.u[1][o], bcz0); rzy = vp_bct(rzy, a.u[0][o], bcz0); }
            if (zhi_c) { const int o = uo + (k + 1) * a.u_sz; lzxN = vp_bct(lzxN, a.u[1][o], bcz1); lzyN = vp_bct(lzyN, a.u[0][o], bcz1); }
        }
        const double lzxH = Y.get(KS(VK_LZXH)), lzyH = Y.get(KS(VK_LZYH));
        // hand-over to the next iteration: the L states of z-face k+1 and the 1-D states of this plane (comps u, v)
        Y.put(KS(VK_LZXH), lzxN); Y.put(KS(VK_LZYH), lzyN);
        Y.put(KS(pc + VK_LX0), lx[0]); Y.put(KS(pc + VK_LX1), lx[1]); Y.put(KS(pc + VK_RX0), rx[0]); Y.put(KS(pc + VK_RX1), rx[1]);
        Y.put(KS(pc + VK_LY0), ly[0]); Y.put(KS(pc + VK_LY1), ly[1]); Y.put(KS(pc + VK_RY0), ry[0]); Y.put(KS(pc + VK_RY1), ry[1]);
#pragma unroll
        for (int c = 0; c < 3; ++c) { Y.put(KS(VK_LZH0 + c), lz[c]); Y.put(KS(VK_QZ0 + c), qz[c]); }
        Y.put(3, lyx); Y.put(4, lyz);
        row_sync<TYT>(ty);                                                                  // B4
        double Xxy, Xyx, Xzx, Xzy, Xxz, Xyz;
        {
            double lxyi = shfl_up1(lxy), lxzi = shfl_up1(lxz), lyxi = Y.from_lo(3), lyzi = Y.from_lo(4), lzxi = lzxH, lzyi = lzyH;
            if (bnd) {
                if (xlo) { lxyi = rxy; lxzi = rxz; }
                if (xhi_f) { rxy = lxyi; rxz = lxzi; }
                if (ylo) { lyxi = ryx; lyzi = ryz; }
                if (yhi_f) { ryx = lyxi; ryz = lyzi; }
                if (zlo) { lzxi = rzx; lzyi = rzy; }
                if (zhi_f) { rzx = lzxi; rzy = lzyi; }
            }
            const TanP tpx = tan_pred(qx[0], eps), tpy = tan_pred(qy[1], eps), tpz = tan_pred(qz[2], eps);
            Xxy = upt_p(lxyi, rxy, tpx); Xyx = upt_p(lyxi, ryx, tpy);
            Xzx = upt_p(lzxi, rzx, tpz); Xzy = upt_p(lzyi, rzy, tpz);
            Xxz = upt_p(lxzi, rxz, tpxP); Xyz = upt_p(lyzi, ryz, tpyP);
        }
        Y.put(0, Xyx); Y.put(1, Xyz);
        row_sync<TYT>(ty);                                                                  // B5
        // ================= F =================
        double Ly, Ry;
        {
            const double Xxyh = shfl_dn1(Xxy), Xxzh = shfl_dn1(Xxz), Xyxh = Y.from_hi(0), Xyzh = Y.from_hi(1);
            const double LzH = Y.get(KS(VK_LZF)), XzxP = Y.get(KS(VK_XZXP)), XzyP = Y.get(KS(VK_XZYP));
            const double fPx = Y.get(KS(VK_FPX)), fPy = Y.get(KS(VK_FPY));
            // wmac on face k: velpred.f90:2373-2419
            const double A1 = a.c4[0] * nsx * (Xxyh - Xxy), A2 = a.c4[1] * nsy * (Xyxh - Xyx);
            double Rz = rz[2] - A1 - A2, LzN = lz[2] - A1 - A2;
            if (!minion) { Rz = Rz + fk[2]; LzN = LzN + fk[2]; }
            if (k >= ka && (k < kb || lastchunk)) {
                double v = riemann_p(LzH, Rz, eps);
                if (zlo)   v = (bcz0 == BC_INLET) ? a.u[2][uo + (k - 1) * a.u_sz] : (bcz0 == BC_OUTLET ? fmin(Rz, ZERO) : ZERO);
                if (zhi_f) v = (bcz1 == BC_INLET) ? a.u[2][uo + k * a.u_sz] : (bcz1 == BC_OUTLET ? fmax(LzH, ZERO) : ZERO);
                if (st_z) a.out[2][oo2 + k * a.o_sz[2]] = v;
            }
            Y.put(KS(VK_LZF), LzN); Y.put(KS(VK_XZXP), Xzx); Y.put(KS(VK_XZYP), Xzy); Y.put(KS(VK_FPX), fk[0]); Y.put(KS(VK_FPY), fk[1]);
            // umac, vmac on plane k-1: velpred.f90:2617-2659, :2665-2707
            const double B1 = a.c4[1] * nsyP * (Xyzh - Xyz), B2 = a.c4[2] * nszP * (Xzy - XzyP);
            double Rx = Y.get(KS(pp + VK_RX0)) - B1 - B2, Lx = Y.get(KS(pp + VK_LX0)) - B1 - B2;
            const double C1 = a.c4[0] * nsxP * (Xxzh - Xxz), C2 = a.c4[2] * nszP * (Xzx - XzxP);
            Ry = Y.get(KS(pp + VK_RY1)) - C1 - C2; Ly = Y.get(KS(pp + VK_LY1)) - C1 - C2;
            if (!minion) { Rx = Rx + fPx; Lx = Lx + fPx; Ry = Ry + fPy; Ly = Ly + fPy; }
            Y.put(3, Ly);
            const double Lxi = shfl_up1(Lx);
            if (k > ka) {
                double v = riemann_p(Lxi, Rx, epsP);
                if (bnd) {
                    if (xlo)   v = (bcx0 == BC_INLET) ? a.u[0][uo + (k - 1) * a.u_sz - 1] : (bcx0 == BC_OUTLET ? fmin(Rx, ZERO) : ZERO);
                    if (xhi_f) v = (bcx1 == BC_INLET) ? a.u[0][uo + (k - 1) * a.u_sz] : (bcx1 == BC_OUTLET ? fmax(Lxi, ZERO) : ZERO);
                }
                if (st_x) a.out[0][oo0 + (k - 1) * a.o_sz[0]] = v;
            }
        }
        row_sync<TYT>(ty);                                                                  // B6
        if (k > ka) {
            const double Lyi = Y.from_lo(3);
            double v = riemann_p(Lyi, Ry, epsP);
            if (bnd) {
                if (ylo)   v = (bcy0 == BC_INLET) ? a.u[1][uo + (k - 1) * a.u_sz - a.u_sy] : (bcy0 == BC_OUTLET ? fmin(Ry, ZERO) : ZERO);
                if (yhi_f) v = (bcy1 == BC_INLET) ? a.u[1][uo + (k - 1) * a.u_sz] : (bcy1 == BC_OUTLET ? fmax(Lyi, ZERO) : ZERO);
            }
            if (st_y) a.out[1][oo1 + (k - 1) * a.o_sz[1]] = v;
        }
        // ---- rotate the register-carried state ----
        nsxP = nsx; nsyP = nsy; epsP = eps; tpxP = tan_pred(qx[0], eps); tpyP = tan_pred(qy[1], eps);

    }
}

} // namespace march

// ------------------------------------------------------------------------------------------
// host side: launch plan shared by vdn_godunov.cu (CUDA) and tests/emu (CPU execution of the same kernel source).
// L::run(kernel, grid, block, smem_bytes, args) launches; L::scope(name, alg_bytes, nlaunch) brackets for the profiler.
// ------------------------------------------------------------------------------------------
namespace march {

struct Plan { int ntx, nty, nzc, zchunk; };
// tiles over columns 0..n0 x 0..n1 (the hi faces live in column n); z chunks chosen so that the grid fills whole waves of
// `slots` resident CTAs with the least (chunk + 1 warm-up plane) work
inline Plan make_plan(const Geo &g, int tyt, int slots)
{
    Plan p;
    p.ntx = (g.n[0] + 1 + TXO - 1) / TXO; p.nty = (g.n[1] + 1 + (tyt - 2) - 1) / (tyt - 2);
    const long tiles = (long)p.ntx * p.nty;
    const int n2 = g.n[2];
    double best = 1e300; p.nzc = 1;
    for (int nzc = 1; nzc <= 64 && (n2 + nzc - 1) / nzc >= 8; ++nzc) {
        const int ch = (n2 + nzc - 1) / nzc;
        const long ctas = tiles * ((n2 + ch - 1) / ch);
        const long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (ch + 2);
        if (cost < best - 1e-9) { best = cost; p.nzc = (n2 + ch - 1) / ch; p.zchunk = ch; }
    }
    if (best == 1e300) { p.nzc = 1; p.zchunk = n2; }
    return p;
}
inline void fill_consts(const Geo &g, double dt, double &dt2, double *c2, double *c3, double *c4, double *c6, double *hinv, int *hp2)
{
    dt2 = HALF * dt;
    const double dt3 = dt / 3.0, dt4 = dt / 4.0, dt6 = dt / 6.0;
    for (int d = 0; d < 3; ++d) {
        c2[d] = dt2 / g.h[d]; c3[d] = dt3 / g.h[d]; c4[d] = dt4 / g.h[d]; c6[d] = dt6 / g.h[d];
        int e; hp2[d] = (std::frexp(g.h[d], &e) == 0.5) ? 1 : 0; hinv[d] = 1.0 / g.h[d];
    }
}

#ifndef MARCH_TYT
#define MARCH_TYT 16
#endif
#ifndef MARCH_TYT_VP
#define MARCH_TYT_VP MARCH_TYT
#endif
#ifndef MARCH_NCG
#define MARCH_NCG 1          // components per mkflux launch
#endif

template <int TYT, class L, class A>
void march_launch(L &launch, void (*k)(A), A &a, const Geo &g, size_t smem, int slots)
{
    const Plan p = make_plan(g, TYT, slots > 0 ? slots : launch.slots(k, TXT * TYT, smem));
    a.zchunk = p.zchunk;
    launch.run(k, dim3(p.ntx, p.nty, p.nzc), dim3(TXT, TYT, 1), smem, a);
}
template <int NC, class L>
void mkflux_march_group(L &launch, MfmArgs &a, int consmask, bool gen, int slots)
{
    constexpr int TYT = MARCH_TYT;
    const size_t smem = sizeof(double) * TXT * TYT * mf_smem_slots<NC>();
    a.nc = NC;
    const bool hp2 = a.hp2[0] && a.hp2[1] && a.hp2[2];
    // the common case (slope_order 4, no Minion forcing, power-of-two spacings) is compiled in; everything else takes the general instantiation
    if (consmask) { if (gen || !hp2) march_launch<TYT>(launch, k_mkflux_march<NC, 1, TYT, true, false>, a, a.g, smem, slots); else march_launch<TYT>(launch, k_mkflux_march<NC, 1, TYT, false, true>, a, a.g, smem, slots); }
    else          { if (gen || !hp2) march_launch<TYT>(launch, k_mkflux_march<NC, 0, TYT, true, false>, a, a.g, smem, slots); else march_launch<TYT>(launch, k_mkflux_march<NC, 0, TYT, false, true>, a, a.g, smem, slots); }
}

// mkflux.f90:16 for ncomp components of s.  adv_bc: [comp][3][2] of these components.  Conservative: comp 0 of the scalars.
template <class L>
void mkflux_march(L &launch, const Geo &g, const View &s, const View &force, const View *mac, const View *sedge, const View *flux,
                  const double *eps, double dt, int is_vel, int ncomp, int order, int use_minion, const int (*adv_bc)[3][2], int slots)
{
    MfmArgs a; memset(&a, 0, sizeof a);
    a.g = g; a.eps = eps; a.is_vel = is_vel; a.order = order; a.use_minion = use_minion;
    a.s_sy = s.sy; a.s_sz = s.sz; a.f_sy = force.sy; a.f_sz = force.sz;
    for (int d = 0; d < 3; ++d) { a.mac[d] = mac[d].p; a.m_sy[d] = mac[d].sy; a.m_sz[d] = mac[d].sz; a.e_sy[d] = sedge[d].sy; a.e_sz[d] = sedge[d].sz; }
    fill_consts(g, dt, a.dt2, a.c2, a.c3, a.c4, a.c6, a.hinv, a.hp2);
    const bool gen = !(order == 4 && !use_minion);
    const double cells = (double)g.n[0] * g.n[1] * g.n[2];
    for (int c0 = 0; c0 < ncomp; ) {
        const int nc = std::min(MARCH_NCG, ncomp - c0);
        const int consmask = (!is_vel && c0 == 0) ? 1 : 0;
        a.comp0 = c0;
        for (int c = 0; c < nc; ++c) {
            a.s[c] = s.p + (long)s.cs * (c0 + c); a.force[c] = force.p + (long)force.cs * (c0 + c);
            for (int d = 0; d < 3; ++d) {
                a.sedge[d][c] = sedge[d].p + (long)sedge[d].cs * (c0 + c);
                a.flux[d][c] = (consmask >> c) & 1 ? flux[d].p + (long)flux[d].cs * (c0 + c) : nullptr;
                for (int sd = 0; sd < 2; ++sd) a.sbc[c][d][sd] = adv_bc[c0 + c][d][sd];
            }
        }
        // the whole phase moves 136 B/cell (scalars: R s 2 + umac 3 + force 2 + mac_rhs 1, W sedge 6 + flux 3) resp. 152 B/cell
        // (velocity: R 3 + 3 + 3 + 1, W 9) of compulsory traffic; a launch is charged its share of the components (the MAC velocities are in fact
        // re-read by every launch, which the figure does not credit)
        const double bytes = cells * (is_vel ? 152.0 : 136.0) * (double)nc / (double)ncomp;
        auto ls = launch.scope(is_vel ? "mkflux_vel" : "mkflux_scal", bytes, 1);
        if (nc == 1) mkflux_march_group<1>(launch, a, consmask, gen, slots);
#if MARCH_NCG >= 2
        else if (nc == 2) mkflux_march_group<2>(launch, a, consmask, gen, slots);
#endif
#if MARCH_NCG >= 3
        else if (nc == 3) mkflux_march_group<3>(launch, a, consmask, gen, slots);
#endif
        c0 += nc;
    }
}

// velpred.f90:16 (3-D): u (ng 3, 3 comps), force (ng 1, 3 comps) -> umac_d (ng 1 face arrays), valid faces only
template <class L>
void velpred_march(L &launch, const Geo &g, const View &u, const View &force, const View *out, const double *eps, double dt,
                   int order, int use_minion, const int (*adv_bc)[3][2], int slots)
{
    constexpr int TYT = MARCH_TYT_VP;
    VpmArgs a; memset(&a, 0, sizeof a);
    a.g = g; a.eps = eps; a.order = order; a.use_minion = use_minion;
    a.u_sy = u.sy; a.u_sz = u.sz; a.f_sy = force.sy; a.f_sz = force.sz;
    double c2[3], c3[3];
    fill_consts(g, dt, a.dt2, c2, c3, a.c4, a.c6, a.hinv, a.hp2);
    for (int c = 0; c < 3; ++c) {
        a.u[c] = u.p + (long)u.cs * c; a.force[c] = force.p + (long)force.cs * c;
        a.out[c] = out[c].p; a.o_sy[c] = out[c].sy; a.o_sz[c] = out[c].sz;
        for (int d = 0; d < 3; ++d) for (int sd = 0; sd < 2; ++sd) a.sbc[c][d][sd] = adv_bc[c][d][sd];
    }
    const size_t smem = sizeof(double) * TXT * TYT * vp_smem_slots();
    // SURVEY 8(a) a2 bytes: R u 24 + force 24, W three face arrays 24
    auto ls = launch.scope("velpred", (double)g.n[0] * g.n[1] * g.n[2] * 72.0, 1);
    if (order == 4 && !use_minion && a.hp2[0] && a.hp2[1] && a.hp2[2]) march_launch<TYT>(launch, k_velpred_march<TYT, false, true>, a, g, smem, slots);
    else                                                                march_launch<TYT>(launch, k_velpred_march<TYT, true, false>, a, g, smem, slots);
}

} // namespace march
