// vdn_mg.cu -- MAC multigrid on the device: solves -div(beta grad phi) = rh on the rank's region.
//
// Replaces mac_multigrid (mac_multigrid.f90:19-66) -> ml_cc_solve (FBoxLib F_MG, third party, absent from the
// reference tree).  The algorithm is the one that call selects: V-cycles of red-black Gauss-Seidel (nu1 = nu2 = 2),
// cell-average restriction, piecewise-constant prolongation, BiCGStab bottom solve (bottom_solver_eps = 1e-3),
// stencil_order = 2 Dirichlet ghost, stop at |r|_inf <= eps*|rh|_inf.  F_MG's source is not available, so results
// are matched to the solver tolerance (parity bar: 10x eps), not bit-wise.
//
// Layout: every level uses one padded layout (n+2*MG_PAD per direction, cell (i,j,k) at off + i + sy*j + sz*k) shared by
// phi, rhs, res and the three face-coefficient arrays (b_d[idx] = beta on the LOW d-face of cell idx), so a
// stencil needs one index.  Level 0 aliases the context's PHI / RH / BETA_* fields (no copies).
// Physical BCs are synthesised in the stencil (Neumann: no flux; Dirichlet: 3*phi0 - phi1/3 one-sided), periodic
// directions owned by one rank wrap by index, so a single GPU needs no ghost-fill launches at all.
#include "vdn_ctx.h"
#include "vdn_comm.h"
#include "vdn_mg_fused.cuh"
#include <algorithm>


struct Lev {
    int n[3];
    long s[3];              // strides (1, sy, sz)
    long ntot, off;         // allocation size and offset of cell (0,0,0)
    double h2inv[3];
    int mode[3][2];
    int par0;               // parity of the global index of local cell (0,0,0)
    double *phi, *rhs, *res, *b[3];
    double *dinv;           // 1 / diagonal (fused smoother levels), recomputed with the coefficients at every solve
    double *alpha;          // Helmholtz solves (alpha - div beta grad) phi = rhs: the cell coefficient; null for the MAC projection
};

struct MG {
    int dim = 3, nlev = 0;
    std::vector<Lev> L;
    std::vector<double *> owned;     // device allocations to free
    bool singular = false;
    double *bot[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };   // BiCGStab vectors on the bottom level
    double *d_norm = nullptr;
    double h0[3];
    cudaGraphExec_t coarse_graph = nullptr;     // levels 1..bottom of one V-cycle (latency-bound launches), captured once
    int coarse_graph_launches = 0;
    int graph_level = 1;                        // first level executed by the captured graph
    bool distributed = false;                   // levels exchange halos with neighbour ranks
    bool push = false;                          // the fused levels fill their neighbours' ghost layers themselves (peer-memory transport)
    bool pushk = false;                         // ... by a push kernel after every sweep instead of stores inside the sweep (A/B hook)
    // agglomeration (multi-rank): local level agg_level is solved on `tail`, a whole-domain hierarchy every rank holds
    int agg_level = -1;
    MG *tail = nullptr;
    double *agg_send = nullptr, *agg_recv = nullptr;
    int *d_coords = nullptr;                    // [nranks][3] process-grid coordinates
    // fused smoother (k_sweep3): levels 0..nfused-1 of a 3-D hierarchy
    int nfused = 0;                             // number of leading levels that run the fused kernel
    int tile_force = -1;                        // test hook (vdn_mg_tune): force one tile shape
    int tail_from = -1;                         // first level of the single-CTA tail (k_tail); -1: none
    bool first_sweep_done = false;              // the first smoothing sweep of level 0 of the coming V-cycle has been launched already
    cudaEvent_t ev_norm = nullptr;              // the residual norm of the last V-cycle has reached the pinned host word
    int sm_count = 148;
    bool helm = false;                          // Helmholtz hierarchy (visc_solve / diff_scalar_solve): plain kernels with the alpha term
};

namespace {

const dim3 BLK(64, 4, 1);

// HELM: the operator is alpha - div(beta grad) (viscsolve.f90:98-100 -> mac_multigrid with alpha = rho or 1); the MAC projection
// instantiates HELM = false and its code is unchanged
template <int DIM, bool HELM = false>
__device__ __forceinline__ void cell_op(const Lev &L, const double *__restrict__ x, long c, const int (&ix)[3], double &Ax, double &dg)
{
    const double x0 = x[c];
    double a = 0.0, g = 0.0;
    if (HELM) { const double al = L.alpha[c]; a = al * x0; g = al; }
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const long st = L.s[d];
        const double h2 = L.h2inv[d];
        const double blo = L.b[d][c], bhi = L.b[d][c + st];
        const bool at_lo = ix[d] == 0, at_hi = ix[d] == L.n[d] - 1;
        const int mlo = L.mode[d][0], mhi = L.mode[d][1];
        if (at_lo && mlo == M_NEU) { }
        else if (at_lo && mlo == M_DIR) { a += blo * (3.0 * x0 - x[c + st] / 3.0) * h2; g += 3.0 * blo * h2; }
        else { const double xm = (at_lo && mlo == M_WRAP) ? x[c + (long)(L.n[d] - 1) * st] : x[c - st]; a += blo * (x0 - xm) * h2; g += blo * h2; }
        if (at_hi && mhi == M_NEU) { }
        else if (at_hi && mhi == M_DIR) { a += bhi * (3.0 * x0 - x[c - st] / 3.0) * h2; g += 3.0 * bhi * h2; }
        else { const double xp = (at_hi && mhi == M_WRAP) ? x[c - (long)(L.n[d] - 1) * st] : x[c + st]; a += bhi * (x0 - xp) * h2; g += bhi * h2; }
    }
    Ax = a; dg = g;
}

// one colour half-sweep; thread per colour cell
template <int DIM, bool HELM = false>
__global__ void k_gsrb(Lev L, int color)
{
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (j >= L.n[1]) return;
    const int i = 2 * i2 + ((j + k + color + L.par0) & 1);
    if (i >= L.n[0]) return;
    const long c = L.off + i + L.s[1] * j + L.s[2] * k;
    const int ix[3] = { i, j, k };
    double Ax, dg; cell_op<DIM, HELM>(L, L.phi, c, ix, Ax, dg);
    if (dg != 0.0) L.phi[c] += (L.rhs[c] - Ax) / dg;
}

// res = rhs - A phi; optional inf-norm via atomicMax on the bit pattern
template <int DIM, bool HELM = false>
__global__ void k_residual(Lev L, double *nrm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    double r = 0.0;
    if (i < L.n[0] && j < L.n[1]) {
        const long c = L.off + i + L.s[1] * j + L.s[2] * k;
        const int ix[3] = { i, j, k };
        double Ax, dg; cell_op<DIM, HELM>(L, L.phi, c, ix, Ax, dg);
        r = L.rhs[c] - Ax;
        L.res[c] = r;
        r = fabs(r);
    }
    if (nrm) block_atomic_max(r, nrm);
}

// coarse rhs = average of the fine residual; coarse phi = 0
template <int DIM>
__global__ void k_restrict(Lev F, Lev C)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= C.n[0] || j >= C.n[1]) return;
    const long cf = F.off + 2 * i + F.s[1] * (2 * j) + F.s[2] * (DIM == 3 ? 2 * k : 0);
    double s = F.res[cf] + F.res[cf + 1] + F.res[cf + F.s[1]] + F.res[cf + F.s[1] + 1];
    if (DIM == 3) s += F.res[cf + F.s[2]] + F.res[cf + F.s[2] + 1] + F.res[cf + F.s[2] + F.s[1]] + F.res[cf + F.s[2] + F.s[1] + 1];
    const long cc = C.off + i + C.s[1] * j + C.s[2] * k;
    C.rhs[cc] = s * (DIM == 3 ? 0.125 : 0.25);
    C.phi[cc] = 0.0;
}

// fine phi += piecewise-constant prolongation of coarse phi
template <int DIM>
__global__ void k_prolong(Lev F, Lev C)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= F.n[0] || j >= F.n[1]) return;
    const long cf = F.off + i + F.s[1] * j + F.s[2] * k;
    const long cc = C.off + (i >> 1) + C.s[1] * (j >> 1) + C.s[2] * (DIM == 3 ? (k >> 1) : 0);
    F.phi[cf] += C.phi[cc];
}

// dinv = 1 / diagonal of the operator (boundary conditions included, 0 where the diagonal vanishes) on the cells lo..hi of every direction:
// the level's own cells, plus the ghost layers that the fused smoother relaxes redundantly on levels split across ranks
template <int DIM>
__global__ void k_diag_inv(Lev L, int lo0, int lo1, int lo2, int hi0, int hi1, int hi2)
{
    const int i = lo0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = lo1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = lo2 + blockIdx.z;
    if (i > hi0 || j > hi1 || k > hi2) return;
    const long c = L.off + i + L.s[1] * j + L.s[2] * k;
    const int ix[3] = { i, j, k };
    double g = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double h2 = L.h2inv[d], blo = L.b[d][c], bhi = L.b[d][c + L.s[d]];
        const bool at_lo = ix[d] == 0, at_hi = ix[d] == L.n[d] - 1;
        const int mlo = L.mode[d][0], mhi = L.mode[d][1];
        if (at_lo && mlo == M_NEU) { } else if (at_lo && mlo == M_DIR) g += 3.0 * blo * h2; else g += blo * h2;
        if (at_hi && mhi == M_NEU) { } else if (at_hi && mhi == M_DIR) g += 3.0 * bhi * h2; else g += bhi * h2;
    }
    L.dinv[c] = g != 0.0 ? 1.0 / g : 0.0;
}

// coarse face coefficient = arithmetic mean of the fine faces it covers
template <int DIM>
__global__ void k_coarsen_beta(Lev F, Lev C, int d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= C.n[0] + (d == 0) || j >= C.n[1] + (d == 1)) return;
    const long cf = F.off + 2 * i + F.s[1] * (2 * j) + F.s[2] * (DIM == 3 ? 2 * k : 0);
    double s;
    if (DIM == 2) {
        const long t = F.s[1 - d];
        s = 0.5 * (F.b[d][cf] + F.b[d][cf + t]);
    } else {
        const long ta = F.s[(d + 1) % 3], tb = F.s[(d + 2) % 3];
        s = 0.25 * (F.b[d][cf] + F.b[d][cf + ta] + F.b[d][cf + tb] + F.b[d][cf + ta + tb]);
    }
    C.b[d][C.off + i + C.s[1] * j + C.s[2] * k] = s;
}

// ---- bottom solver: BiCGStab in ONE CTA (the coarsest level is a handful of cells); dot products use
// warp-shuffle + shared-memory block reductions ----
__device__ double block_sum(double v, double *sm)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    double t = (l < nw) ? sm[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;      // every thread of every warp holds the total
}
struct BotVec { double *r, *rh, *p, *v, *s, *t; };

template <int DIM, bool HELM = false>
__device__ void bottom_solve(const Lev &L, const BotVec &w, int maxit, double eps, int singular, double *sm)
{
    const long nc = (long)L.n[0] * L.n[1] * L.n[2];
    const int nt = blockDim.x, tid = threadIdx.x;
#define CELL(q, c, ix) const int ix##0 = (int)((q) % L.n[0]), ix##1 = (int)(((q) / L.n[0]) % L.n[1]), ix##2 = (int)((q) / ((long)L.n[0] * L.n[1])); \
                       const long c = L.off + ix##0 + L.s[1] * ix##1 + L.s[2] * ix##2; const int ix[3] = { ix##0, ix##1, ix##2 };
    if (singular) {         // project the null space out of the right-hand side
        double s = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; s += L.rhs[c]; }
        s = block_sum(s, sm) / (double)nc;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; L.rhs[c] -= s; }
    }
    double bn = 0.0;
    for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; const double b = L.rhs[c]; L.phi[c] = 0.0; w.r[c] = b; w.rh[c] = b; w.p[c] = b; bn += b * b; }
    bn = sqrt(block_sum(bn, sm));
    double rho = 1.0, alpha = 1.0, omega = 1.0;
    if (bn > 0.0)
    for (int it = 0; it < maxit; ++it) {
        double rho1 = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; rho1 += w.rh[c] * w.r[c]; }
        rho1 = block_sum(rho1, sm);
        if (rho1 == 0.0) break;
        if (it > 0) {
            const double beta = (rho1 / rho) * (alpha / omega);
            for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; w.p[c] = w.r[c] + beta * (w.p[c] - omega * w.v[c]); }
        }
        __syncthreads();
        double den = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) double Ax, dg; cell_op<DIM, HELM>(L, w.p, c, ix, Ax, dg); w.v[c] = Ax; den += w.rh[c] * Ax; }
        den = block_sum(den, sm);
        if (den == 0.0) break;
        alpha = rho1 / den;
        double sn = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; const double s = w.r[c] - alpha * w.v[c]; w.s[c] = s; sn += s * s; }
        sn = sqrt(block_sum(sn, sm));
        if (sn <= eps * bn) { for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; L.phi[c] += alpha * w.p[c]; } break; }
        __syncthreads();
        double ts = 0.0, tt = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) double Ax, dg; cell_op<DIM, HELM>(L, w.s, c, ix, Ax, dg); w.t[c] = Ax; ts += Ax * w.s[c]; tt += Ax * Ax; }
        ts = block_sum(ts, sm); tt = block_sum(tt, sm);
        if (tt == 0.0) { for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; L.phi[c] += alpha * w.p[c]; } break; }
        omega = ts / tt;
        double rn = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; L.phi[c] += alpha * w.p[c] + omega * w.s[c]; const double r = w.s[c] - omega * w.t[c]; w.r[c] = r; rn += r * r; }
        rn = sqrt(block_sum(rn, sm));
        rho = rho1;
        if (rn <= eps * bn || omega == 0.0) break;
        __syncthreads();
    }
    __syncthreads();
    if (singular) {
        double s = 0.0;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; s += L.phi[c]; }
        s = block_sum(s, sm) / (double)nc;
        for (long q = tid; q < nc; q += nt) { CELL(q, c, ix) (void)ix; L.phi[c] -= s; }
    }
#undef CELL
}
template <int DIM, bool HELM = false>
__global__ void __launch_bounds__(1024) k_bottom(Lev L, BotVec w, int maxit, double eps, int singular)
{
    __shared__ double sm[32];
    bottom_solve<DIM, HELM>(L, w, maxit, eps, singular, sm);
}

// ---- the tail of a V-cycle in ONE CTA: every level of at most TAIL_CELLS cells (8^3) -- smoothing, residual, restriction, the BiCGStab
// bottom solve, prolongation, smoothing -- with __syncthreads between the stages instead of a kernel launch per stage (44 launches of a few
// microseconds each per V-cycle at 256^3).  The level arrays stay in global memory (L1/L2-resident at these sizes). ----
constexpr long TAIL_CELLS = 512;       // 8^3: measured (r2 call 5) -- a 16^3 stage in one CTA costs as much as the launch it replaces
constexpr int TAIL_MAXLEV = 6;
struct TailArgs { int nl; Lev L[TAIL_MAXLEV]; BotVec w; int nu1, nu2, maxit, singular; double eps; };
template <int DIM>
__device__ __forceinline__ void tail_cell(const Lev &L, long q, long &c, int (&ix)[3])
{
    ix[0] = (int)(q % L.n[0]); ix[1] = (int)((q / L.n[0]) % L.n[1]); ix[2] = (int)(q / ((long)L.n[0] * L.n[1]));
    c = L.off + ix[0] + L.s[1] * ix[1] + L.s[2] * ix[2];
}
template <int DIM, bool HELM = false>
__device__ void tail_smooth(const Lev &L, int sweeps)
{
    const long nc = (long)L.n[0] * L.n[1] * L.n[2];
    for (int s = 0; s < sweeps; ++s)
        for (int color = 0; color < 2; ++color) {
            for (long q = threadIdx.x; q < nc; q += blockDim.x) {
                long c; int ix[3]; tail_cell<DIM>(L, q, c, ix);
                if (((ix[0] + ix[1] + ix[2] + color + L.par0) & 1) != 0) continue;
                double Ax, dg; cell_op<DIM, HELM>(L, L.phi, c, ix, Ax, dg);
                if (dg != 0.0) L.phi[c] += (L.rhs[c] - Ax) / dg;
            }
            __syncthreads();
        }
}
template <int DIM, bool HELM = false>
__global__ void __launch_bounds__(1024) k_tail(TailArgs a)
{
    __shared__ double sm[32];
    for (int l = 0; l + 1 < a.nl; ++l) {
        const Lev &F = a.L[l], &C = a.L[l + 1];
        tail_smooth<DIM, HELM>(F, a.nu1);
        const long nf = (long)F.n[0] * F.n[1] * F.n[2];
        for (long q = threadIdx.x; q < nf; q += blockDim.x) {
            long c; int ix[3]; tail_cell<DIM>(F, q, c, ix);
            double Ax, dg; cell_op<DIM, HELM>(F, F.phi, c, ix, Ax, dg);
            F.res[c] = F.rhs[c] - Ax;
        }
        __syncthreads();
        const long ncc = (long)C.n[0] * C.n[1] * C.n[2];
        for (long q = threadIdx.x; q < ncc; q += blockDim.x) {
            long cc; int ix[3]; tail_cell<DIM>(C, q, cc, ix);
            const long cf = F.off + 2 * ix[0] + F.s[1] * (2 * ix[1]) + F.s[2] * (DIM == 3 ? 2 * ix[2] : 0);
            double s = F.res[cf] + F.res[cf + 1] + F.res[cf + F.s[1]] + F.res[cf + F.s[1] + 1];
            if (DIM == 3) s += F.res[cf + F.s[2]] + F.res[cf + F.s[2] + 1] + F.res[cf + F.s[2] + F.s[1]] + F.res[cf + F.s[2] + F.s[1] + 1];
            C.rhs[cc] = s * (DIM == 3 ? 0.125 : 0.25);
            C.phi[cc] = 0.0;
        }
        __syncthreads();
    }
    bottom_solve<DIM, HELM>(a.L[a.nl - 1], a.w, a.maxit, a.eps, a.singular, sm);
    __syncthreads();
    for (int l = a.nl - 2; l >= 0; --l) {
        const Lev &F = a.L[l], &C = a.L[l + 1];
        const long nf = (long)F.n[0] * F.n[1] * F.n[2];
        for (long q = threadIdx.x; q < nf; q += blockDim.x) {
            long c; int ix[3]; tail_cell<DIM>(F, q, c, ix);
            const long cc = C.off + (ix[0] >> 1) + C.s[1] * (ix[1] >> 1) + C.s[2] * (DIM == 3 ? (ix[2] >> 1) : 0);
            F.phi[c] += C.phi[cc];
        }
        __syncthreads();
        tail_smooth<DIM, HELM>(F, a.nu2);
    }
}

template <class F> void for_dim(int dim, F f) { if (dim == 3) f(std::integral_constant<int, 3>()); else f(std::integral_constant<int, 2>()); }
// (dimension, Helmholtz) dispatch of the plain kernels
template <class F> void for_dim_h(int dim, bool helm, F f)
{
    if (helm) { if (dim == 3) f(std::integral_constant<int, 3>(), std::true_type()); else f(std::integral_constant<int, 2>(), std::true_type()); }
    else      { if (dim == 3) f(std::integral_constant<int, 3>(), std::false_type()); else f(std::integral_constant<int, 2>(), std::false_type()); }
}

dim3 cgrid(int nx, int ny, int nz) { return dim3(cdiv(nx, BLK.x), cdiv(ny, BLK.y), nz); }


// ---- agglomeration kernels: contiguous block <-> padded level array ----
struct BlkArgs { double *arr; long off, s1, s2; int ex[3]; double *buf; int nranks; const int *coords; int nloc[3]; int mine[3]; };
__global__ void k_blk_pack(BlkArgs a)           // buf[t] = arr(block of this rank)
{
    const long tot = (long)a.ex[0] * a.ex[1] * a.ex[2];
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t % a.ex[0]), j = (int)((t / a.ex[0]) % a.ex[1]), k = (int)(t / ((long)a.ex[0] * a.ex[1]));
        a.buf[t] = a.arr[a.off + i + a.s1 * j + a.s2 * k];
    }
}
__global__ void k_blk_unpack(BlkArgs a)         // arr(global) <- buf[r][t] for every rank r (blockIdx.y)
{
    const int r = blockIdx.y;
    const long tot = (long)a.ex[0] * a.ex[1] * a.ex[2];
    const int ox = a.coords[r * 3] * a.nloc[0], oy = a.coords[r * 3 + 1] * a.nloc[1], oz = a.coords[r * 3 + 2] * a.nloc[2];
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t % a.ex[0]), j = (int)((t / a.ex[0]) % a.ex[1]), k = (int)(t / ((long)a.ex[0] * a.ex[1]));
        a.arr[a.off + (ox + i) + a.s1 * (oy + j) + a.s2 * (oz + k)] = a.buf[(long)r * tot + t];
    }
}
struct ExtArgs { const double *src; long soff, ss1, ss2; double *dst; long doff, ds1, ds2; int ex[3]; int o[3]; };
__global__ void k_blk_extract(ExtArgs a)        // local phi <- my block of the global phi
{
    const long tot = (long)a.ex[0] * a.ex[1] * a.ex[2];
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t % a.ex[0]), j = (int)((t / a.ex[0]) % a.ex[1]), k = (int)(t / ((long)a.ex[0] * a.ex[1]));
        a.dst[a.doff + i + a.ds1 * j + a.ds2 * k] = a.src[a.soff + (a.o[0] + i) + a.ss1 * (a.o[1] + j) + a.ss2 * (a.o[2] + k)];
    }
}

// Build a hierarchy for a grid of n cells (spacing h, global index origin glo) with per-face modes.
// alias0: level 0 uses the context's PHI / RH / BETA_* storage.  max_levels < 0: coarsen as far as possible.
MG *mg_make(vdn_ctx *c, const int *n_in, const double *h_in, const int *glo_in, const int (*mode)[2], bool alias0, int max_levels, bool helm = false)
{
    MG *m = new MG();
    m->helm = helm;
    bool sym = false;                                    // any face shared with another rank
    for (int d = 0; d < c->dim; ++d) for (int s = 0; s < 2; ++s) if (mode[d][s] == M_GHOST) sym = true;
    m->dim = c->dim;
    int nn[3] = { n_in[0], n_in[1], n_in[2] };
    int nlev = 1;
    for (;;) {          // halve while every direction stays even and >= 2 afterwards (F_MG min_width = 2)
        bool ok = true;
        for (int d = 0; d < c->dim; ++d) if (nn[d] % 2 != 0 || nn[d] / 2 < 2) ok = false;
        if (!ok || (max_levels > 0 && nlev >= max_levels)) break;
        for (int d = 0; d < c->dim; ++d) nn[d] /= 2;
        ++nlev;
    }
    m->nlev = nlev; m->L.resize(nlev);
    m->singular = !helm;                                 // alpha > 0: never singular
    for (int d = 0; d < c->dim; ++d) for (int s = 0; s < 2; ++s) if (c->dom_bc[d][s] == BC_OUTLET) m->singular = false;
    int n[3] = { n_in[0], n_in[1], n_in[2] };
    double h[3] = { h_in[0], h_in[1], h_in[2] };
    int glo[3] = { glo_in[0], glo_in[1], glo_in[2] };
    for (int l = 0; l < nlev; ++l) {
        Lev &L = m->L[l];
        for (int d = 0; d < 3; ++d) { L.n[d] = n[d]; L.h2inv[d] = 1.0 / (h[d] * h[d]); }
        L.s[0] = 1; L.s[1] = n[0] + 2 * MG_PAD; L.s[2] = (long)(n[0] + 2 * MG_PAD) * (n[1] + 2 * MG_PAD);
        L.ntot = L.s[2] * (c->dim == 3 ? n[2] + 2 * MG_PAD : 1);
        L.off = MG_PAD * (1 + L.s[1] + (c->dim == 3 ? L.s[2] : 0));
        L.par0 = (glo[0] + glo[1] + (c->dim == 3 ? glo[2] : 0)) & 1;
        for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) { L.mode[d][s] = d < c->dim ? mode[d][s] : M_NEU; if (L.mode[d][s] == M_GHOST) m->distributed = true; }
        // rank-split hierarchies live in the symmetric heap of the peer-memory transport (every rank allocates the same sequence)
        auto dalloc = [&](long cnt) {
            bool owned = true; double *p;
            if (sym) p = comm_sym_alloc(c, sizeof(double) * cnt, &owned); else VDN_CUDA(cudaMalloc(&p, sizeof(double) * cnt));
            VDN_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * cnt, c->stream));
            if (owned) m->owned.push_back(p);
            return p; };
        if (l == 0 && alias0) {
            L.phi = c->f[VDN_PHI].base; L.rhs = c->f[VDN_RH].base;
            for (int d = 0; d < c->dim; ++d) L.b[d] = c->f[VDN_BETA_X + d].base;
            VDN_REQUIRE(c->f[VDN_RH].sy == L.s[1] && c->f[VDN_PHI].sy == L.s[1] && c->f[VDN_BETA_X].sy == L.s[1], "level-0 layout mismatch");
        } else {
            L.phi = dalloc(L.ntot); L.rhs = dalloc(L.ntot);
            for (int d = 0; d < c->dim; ++d) L.b[d] = dalloc(L.ntot);
        }
        for (int d = c->dim; d < 3; ++d) L.b[d] = nullptr;
        L.res = dalloc(L.ntot);
        L.dinv = (c->dim == 3 && !helm) ? dalloc(L.ntot) : nullptr;
        L.alpha = helm ? dalloc(L.ntot) : nullptr;
        for (int d = 0; d < c->dim; ++d) { n[d] /= 2; h[d] *= 2.0; glo[d] /= 2; }
    }
    if (!m->distributed)
        for (int l = 1; l < nlev; ++l) {
            const Lev &L = m->L[l];
            if ((long)L.n[0] * L.n[1] * L.n[2] <= TAIL_CELLS && nlev - l <= TAIL_MAXLEV) { m->tail_from = l; break; }
        }
    for (int q = 0; q < 6; ++q) { VDN_CUDA(cudaMalloc(&m->bot[q], sizeof(double) * m->L[nlev - 1].ntot)); VDN_CUDA(cudaMemsetAsync(m->bot[q], 0, sizeof(double) * m->L[nlev - 1].ntot, c->stream)); }
    VDN_CUDA(cudaMalloc(&m->d_norm, 64));
    return m;
}


// fused smoother on the leading (large) levels of a 3-D hierarchy.  Rank-local levels wrap by index; levels that are split
// across ranks relax up to 3 ghost layers redundantly after ONE deep halo exchange per launch.
void mg_pick_fused(vdn_ctx *c, MG *m)
{
    cudaDeviceProp pr; VDN_CUDA(cudaGetDeviceProperties(&pr, c->device)); m->sm_count = pr.multiProcessorCount;
    m->tile_force = c->mg_tile_force;
    if (c->dim != 3 || c->prm.mg_nu1 < 1 || c->prm.mg_nu2 < 1) return;
    const int last = m->tail ? m->agg_level : m->nlev - 1;      // the agglomerated / bottom level is never fused
    while (m->nfused < last) {
        const Lev &L = m->L[m->nfused];
        // rank-local hierarchies: the plain kernels inside the CUDA graph beat the fused one below 128^3 (round 1, b15_min64); levels split
        // across ranks: every plain colour half-sweep needs its own ghost exchange, so the fused kernel (which reads its neighbours' cells
        // itself) pays off down to 64^3
        const int fmin_ = (m->distributed && c->mg_fuse_min == 128) ? 64 : c->mg_fuse_min;
        if (std::min(L.n[0], std::min(L.n[1], L.n[2])) < std::max(fmin_, 16) || (L.n[0] | L.n[1] | L.n[2]) & 1) break;
        ++m->nfused;
    }
    m->push = m->distributed && m->nfused > 0 && comm_peer_mode(c) && comm_mg_xchg(c) == 0;
    m->pushk = m->distributed && m->nfused > 0 && comm_peer_mode(c) && comm_mg_xchg(c) == 3;
}

void mg_build(vdn_ctx *c)
{
    const Geo &g = c->geo;
    int mode[3][2];
    for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) {
        int e = d < c->dim ? c->ell_bc[d][s] : ELL_NEU;
        mode[d][s] = e == ELL_NEU ? M_NEU : e == ELL_DIR ? M_DIR : (e == ELL_PER && c->wrap[d]) ? M_WRAP : M_GHOST;
    }
    const int nr = comm_nranks(c);
    if (nr == 1) {
        MG *m = mg_make(c, g.n, g.h, c->rlo, mode, true, -1);
        c->mg = m;
        mg_pick_fused(c, m);
        return;
    }
    // multi-rank: distributed levels down to a local size of <= 32 cells per direction, then agglomerate
    int nloc[3] = { g.n[0], g.n[1], g.n[2] };
    int ndist = 1;
    for (;;) {
        int mx = 0; bool ok = true;
        for (int d = 0; d < c->dim; ++d) { mx = std::max(mx, nloc[d]); if (nloc[d] % 2 != 0 || nloc[d] / 2 < 2) ok = false; }
        if ((mx <= 32 && ndist >= 2) || !ok) break;     // keep at least one distributed level above the agglomerated one
        for (int d = 0; d < c->dim; ++d) nloc[d] /= 2;
        ++ndist;
    }
    VDN_REQUIRE(ndist >= 2, "multi-rank multigrid needs a local region that can be coarsened at least once (even sizes >= 4)");
    MG *m = mg_make(c, g.n, g.h, c->rlo, mode, true, ndist);
    c->mg = m;
    m->agg_level = m->nlev - 1;
    Lev &A = m->L[m->agg_level];
    const int *pg = comm_pgrid(c);
    int gn[3], gl[3] = { 0, 0, 0 }, gmode[3][2];
    double gh[3];
    for (int d = 0; d < 3; ++d) {
        gn[d] = d < c->dim ? A.n[d] * pg[d] : 1;
        gh[d] = d < c->dim ? g.h[d] * (double)(g.n[d] / A.n[d]) : 1.0;
        gl[d] = d < c->dim ? (c->dom_lo[d] / (g.n[d] / A.n[d])) : 0;
        for (int s = 0; s < 2; ++s) {
            int p = d < c->dim ? c->dom_bc[d][s] : BC_SLIP_WALL;
            gmode[d][s] = p == BC_PERIODIC ? M_WRAP : p == BC_OUTLET ? M_DIR : M_NEU;
        }
    }
    m->tail = mg_make(c, gn, gh, gl, gmode, false, -1);
    long mxblk = 1;
    for (int d = 0; d < c->dim; ++d) mxblk *= (A.n[d] + 1);
    VDN_CUDA(cudaMalloc(&m->agg_send, sizeof(double) * mxblk));
    VDN_CUDA(cudaMalloc(&m->agg_recv, sizeof(double) * mxblk * nr));
    std::vector<int> coords(3 * nr);
    for (int r = 0; r < nr; ++r) comm_coord_of(c, r, &coords[3 * r]);
    VDN_CUDA(cudaMalloc(&m->d_coords, sizeof(int) * 3 * nr));
    VDN_CUDA(cudaMemcpy(m->d_coords, coords.data(), sizeof(int) * 3 * nr, cudaMemcpyHostToDevice));
    mg_pick_fused(c, m);
}

// gather one array of the local agglomeration level (cells, or faces along fdir) into the tail's level 0 on every rank
void agg_gather(vdn_ctx *c, MG *m, double *local, double *global, int fdir)
{
    Lev &A = m->L[m->agg_level]; Lev &T = m->tail->L[0];
    const int nr = comm_nranks(c);
    BlkArgs a; a.arr = local; a.off = A.off; a.s1 = A.s[1]; a.s2 = A.s[2];
    for (int d = 0; d < 3; ++d) { a.ex[d] = A.n[d] + (d == fdir ? 1 : 0); a.nloc[d] = A.n[d]; a.mine[d] = 0; }
    a.buf = m->agg_send; a.nranks = nr; a.coords = m->d_coords;
    const long tot = (long)a.ex[0] * a.ex[1] * a.ex[2];
    const int nb = (int)std::min<long>(592, (tot + 255) / 256);
    LaunchScope ls(c, "mg_agglomerate", 0.0, 2);
    k_blk_pack<<<nb, 256, 0, c->stream>>>(a);
    comm_allgather(c, m->agg_send, m->agg_recv, (size_t)tot);
    BlkArgs u = a; u.arr = global; u.off = T.off; u.s1 = T.s[1]; u.s2 = T.s[2]; u.buf = m->agg_recv;
    k_blk_unpack<<<dim3(nb, nr), 256, 0, c->stream>>>(u);
    VDN_CUDA(cudaGetLastError());
}

void mg_halo(vdn_ctx *c, MG *m, Lev &L, double *x)
{
    if (!m->distributed) return;
    View v; v.p = x + L.off; v.sy = L.s[1]; v.sz = L.s[2]; v.cs = L.ntot;
    int dmask = 0;
    for (int d = 0; d < m->dim; ++d) if (L.mode[d][0] == M_GHOST || L.mode[d][1] == M_GHOST) dmask |= 1 << d;
    LaunchScope ls(c, "mg_halo_exchange", 0.0, 1);
    // the 7-point stencil of the plain kernels reads face neighbours only, but one plan serves every exchange (edges / corners: a few cells)
    comm_halo(c, v, L.n, m->dim, 1, 1, -1, dmask, true);
}

// ghost layers of depth ng of one level array, all split directions at once (x, then y over the x-ghosted range, then z: the
// tiles of the fused smoother also read edge and corner ghosts)
void mg_halo_deep(vdn_ctx *c, MG *m, Lev &L, double *x, int ng)
{
    if (!m->distributed) return;
    View v; v.p = x + L.off; v.sy = L.s[1]; v.sz = L.s[2]; v.cs = L.ntot;
    int dmask = 0;
    for (int d = 0; d < m->dim; ++d) if (L.mode[d][0] == M_GHOST || L.mode[d][1] == M_GHOST) dmask |= 1 << d;
    LaunchScope ls(c, "mg_halo_exchange", 0.0, 1);
    comm_halo(c, v, L.n, m->dim, ng, 1, -1, dmask, true);
}

void smooth(vdn_ctx *c, MG *m, int l, int sweeps)
{
    Lev &L = m->L[l];
    const double cells = (double)L.n[0] * L.n[1] * L.n[2];
    for (int s = 0; s < sweeps; ++s)
        for (int color = 0; color < 2; ++color) {
            mg_halo(c, m, L, L.phi);        // fill_boundary(phi) before each colour, as F_MG does
            // one colour half-sweep = R phi 8 + rhs 4 + beta 24 (3-D), W phi 4 = 40 B/cell
            LaunchScope ls(c, l == 0 ? "mg_gsrb_l0" : "mg_gsrb_coarse", cells * (m->dim == 3 ? 40.0 : 32.0));
            dim3 gr(cdiv((L.n[0] + 1) / 2, BLK.x), cdiv(L.n[1], BLK.y), L.n[2]);
            for_dim_h(m->dim, m->helm, [&](auto D, auto H) { k_gsrb<decltype(D)::value, decltype(H)::value><<<gr, BLK, 0, c->stream>>>(L, color); });
        }
}
void residual(vdn_ctx *c, MG *m, int l, double *nrm)
{
    Lev &L = m->L[l];
    const double cells = (double)L.n[0] * L.n[1] * L.n[2];
    mg_halo(c, m, L, L.phi);
    LaunchScope ls(c, l == 0 ? "mg_residual_l0" : "mg_residual_coarse", cells * (m->dim == 3 ? 48.0 : 40.0));
    if (nrm) VDN_CUDA(cudaMemsetAsync(nrm, 0, 8, c->stream));
    for_dim_h(m->dim, m->helm, [&](auto D, auto H) { k_residual<decltype(D)::value, decltype(H)::value><<<cgrid(L.n[0], L.n[1], L.n[2]), BLK, 0, c->stream>>>(L, nrm); });
}


// ---- fused smoother launcher ----
struct WaveVariant { const void *fn = nullptr; size_t smem = 0; int occ = 0; int H = 0, W = 0, HH = 0, TX = 0, TY = 0, NT = 0; };
constexpr int SWEEP_NCFG = 5;

// k_sweep3 variants (one column of cell pairs per thread, register-pipelined operator data): [tile cfg][pre][post index]
template <int PRE, int POST, int TX, int TY, bool P2P>
WaveVariant sweep3_variant()
{
    using C = Sweep3Cfg<PRE, POST, TX, TY>;
    WaveVariant v;
    v.fn = (const void *)k_sweep3<PRE, POST, TX, TY, P2P>;
    v.smem = C::SMEM;
    v.H = C::H; v.W = C::X; v.HH = C::Y; v.TX = TX; v.TY = TY; v.NT = C::NT;
    VDN_CUDA(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    VDN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v.occ, v.fn, C::NT, v.smem));
    VDN_REQUIRE(v.occ >= 1, "k_sweep3 variant does not fit on an SM");
    return v;
}
// p2p: the instantiation that reads neighbour ranks' cells from peer memory (levels split across ranks, peer-memory transport up)
WaveVariant &sweep3_get(int cfg, int pre, int post, bool p2p)
{
    static WaveVariant tab[2][SWEEP_NCFG][2][3];
    const int pi = post == 0 ? 0 : post == 2 ? 1 : 2;
    WaveVariant &v = tab[p2p ? 1 : 0][cfg][pre][pi];
    if (v.fn) return v;
#define SV3P(C, TX, TY, P) \
    if (cfg == C && p2p == P) { \
        if (pre == 0 && post == 0) v = sweep3_variant<0, 0, TX, TY, P>(); \
        if (pre == 0 && post == 2) v = sweep3_variant<0, 2, TX, TY, P>(); \
        if (pre == 0 && post == 3) v = sweep3_variant<0, 3, TX, TY, P>(); \
        if (pre == 1 && post == 0) v = sweep3_variant<1, 0, TX, TY, P>(); \
        if (pre == 1 && post == 2) v = sweep3_variant<1, 2, TX, TY, P>(); \
        if (pre == 1 && post == 3) v = sweep3_variant<1, 3, TX, TY, P>(); \
    }
#define SV3(C, TX, TY) SV3P(C, TX, TY, false) SV3P(C, TX, TY, true)
    SV3(0, 32, 32) SV3(1, 64, 16) SV3(2, 32, 16) SV3(3, 64, 14) SV3(4, 32, 24)     // 3, 4: 640-thread CTAs (20 warps: 96 registers, no spills)
#undef SV3
#undef SV3P
    VDN_REQUIRE(v.fn != nullptr, "no such k_sweep3 variant");
    return v;
}

// one fused launch on level l: one sweep reading L.phi, writing L.res; then the two buffers swap roles
void wave_launch(vdn_ctx *c, MG *m, int l, int pre, int post)
{
    Lev &L = m->L[l];
    const int nsw = 1;
    // pick tile shape and z-chunking: cost ~ waves * CTAs sharing an SM * iterations * plane cells
    int best_cfg = 0, best_ch = L.n[2]; double best = 1e300;
    // peer-memory mode: the kernel stores what it writes near a face shared with another rank into that rank's ghost layers as well (phi, and
    // under post == 2 the coarse right-hand side and the zeroed coarse phi), so the fused levels need no exchange launches inside a V-cycle
    WaveArgs a;
    a.p2p = 0; a.my_flag = nullptr; a.epoch = 0; a.peer_mask = 0;
    bool peer_mode = false;
    int dmask = 0;
    // the coarse correction under the fine ghost layers: pushed by the last sweep of the level below if that level is fused, exchanged otherwise
    // (before this launch takes its epoch: epochs are numbered in launch order)
    if ((m->push || m->pushk) && pre && l + 1 >= m->nfused) mg_halo_deep(c, m, m->L[l + 1], m->L[l + 1].phi, 2);
    if (m->distributed) for (int d = 0; d < m->dim; ++d) if (L.mode[d][0] == M_GHOST || L.mode[d][1] == M_GHOST) dmask |= 1 << d;
    a.wait_ns = nullptr;
    if (m->push) {
        const double *arrs[4] = { L.phi, L.res, post == 2 ? m->L[l + 1].rhs : nullptr, post == 2 ? m->L[l + 1].phi : nullptr };
        peer_mode = comm_peer_tables(c, arrs, 4, dmask, a.peer_delta, &a.peer_mask, a.pub_flag, a.wait_flag, &a.my_flag, &a.epoch);
        a.p2p = peer_mode ? 1 : 0;
    }
    VDN_REQUIRE(!m->distributed || peer_mode == m->push, "peer-memory mode of the fused smoother changed between launches");
    auto variant = [&](int cfg) -> const WaveVariant & { return sweep3_get(cfg, pre, post, peer_mode); };
    for (int cfg = 0; cfg < SWEEP_NCFG; ++cfg) {
        if (m->tile_force >= 0 && cfg != m->tile_force) continue;
        // measured defaults (profiles/r01_bench_256_v5_*): 32x24 (640 threads, 96 registers) for the plain / prolongating sweeps of the finest
        // level and for sweep + norm, 64x16 for the plain sweeps of smaller levels, 32x16 for sweep + residual + restriction (an 864-thread
        // 64x16 CTA is capped at 72 registers and spills)
        if (m->tile_force < 0 && cfg != (post == 2 ? 2 : (post == 3 || L.n[0] >= 256) ? 4 : 1)) continue;
        const WaveVariant &v = variant(cfg);
        const long ntiles = (long)cdiv(L.n[0], v.TX) * cdiv(L.n[1], v.TY);
        const long slots = (long)m->sm_count * v.occ;
        for (int nz = 1; nz <= std::max(1, L.n[2] / 8); ++nz) {
            const int ch = (cdiv(L.n[2], nz) + 1) & ~1;
            const long ctas = ntiles * cdiv(L.n[2], ch);
            const long waves = (ctas + slots - 1) / slots;
            const long per_sm = std::min<long>(v.occ, (ctas + m->sm_count - 1) / m->sm_count);
            const double cost = (double)waves * per_sm * (ch + 2 * v.H + 3) * v.W * v.HH;
            if (cost < best * (1.0 - 1e-9)) { best = cost; best_cfg = cfg; best_ch = ch; }
        }
    }
    const WaveVariant &v = variant(best_cfg);
    for (int d = 0; d < 3; ++d) { a.n[d] = L.n[d]; a.h2[d] = L.h2inv[d]; a.mode[d][0] = L.mode[d][0]; a.mode[d][1] = L.mode[d][1]; }
    a.s1 = L.s[1]; a.s2 = L.s[2]; a.off = L.off; a.par0 = L.par0;
    a.rhs = L.rhs; a.b0 = L.b[0]; a.b1 = L.b[1]; a.b2 = L.b[2]; a.dinv = L.dinv;
    a.in = L.phi; a.out = L.res;
    a.cphi = nullptr; a.crhs = nullptr; a.czero = nullptr; a.cs1 = a.cs2 = a.coff = 0;
    if (pre || post == 2) {
        Lev &C = m->L[l + 1];
        a.cphi = C.phi; a.crhs = C.rhs; a.czero = C.phi; a.cs1 = C.s[1]; a.cs2 = C.s[2]; a.coff = C.off;
    }
    a.nrm = m->d_norm; a.zchunk = best_ch;
    if (peer_mode) {
        c->comm_bytes += 8 * comm_halo_volume(c, L.n, m->dim, PUSH_DEPTH, dmask);
        if (post == 2) c->comm_bytes += 2 * 8 * comm_halo_volume(c, m->L[l + 1].n, m->dim, PUSH_DEPTH, dmask);
    } else if (!m->pushk) {
        if (pre) mg_halo_deep(c, m, m->L[l + 1], m->L[l + 1].phi, 2);   // the prolongation under 3 fine ghost layers reads 2 coarse ones
        mg_halo_deep(c, m, L, L.phi, v.H);      // neighbour-rank cells the tiles relax redundantly
    }
    if (post == 3) VDN_CUDA(cudaMemsetAsync(m->d_norm, 0, 8, c->stream));
    const double cells = (double)L.n[0] * L.n[1] * L.n[2];
    // SURVEY 8(a) a8 per stage: colour half-sweep 40, residual 48, restriction 9, prolongation 17 B/cell
    const double alg = cells * (nsw * 2 * 40.0 + (pre ? 17.0 : 0.0) + (post == 2 ? 48.0 + 9.0 : post == 3 ? 48.0 : 0.0));
    const char *name = l == 0 ? (post == 2 ? "mg_wave_down_l0" : post == 3 ? "mg_wave_up_l0" : pre ? "mg_wave_pro_l0" : "mg_wave_smooth_l0") : "mg_wave_coarse";
    if (peer_mode && c->d_dbg) a.wait_ns = c->d_dbg + 4 * (l == 0 ? (post == 2 ? 1 : post == 3 ? 3 : pre ? 2 : 0) : 4);
    LaunchScope ls(c, name, alg);
    dim3 grid(cdiv(L.n[0], v.TX), cdiv(L.n[1], v.TY), cdiv(L.n[2], best_ch));
    void *args[] = { (void *)&a };
    VDN_CUDA(cudaLaunchKernel(v.fn, grid, dim3(v.NT), args, v.smem, c->stream));
    std::swap(L.phi, L.res);
    if (m->pushk) {
        // A/B hook: the same ghost layers filled by a push kernel after the sweep
        LaunchScope lp(c, "mg_halo_exchange", 0.0, post == 2 ? 2 : 1);
        double *a1[1] = { L.phi };
        comm_push(c, a1, 1, L.off, (int)L.s[1], (int)L.s[2], L.n, m->dim, PUSH_DEPTH, dmask);
        if (post == 2) {
            Lev &C = m->L[l + 1];
            double *a2[2] = { C.rhs, C.phi };
            comm_push(c, a2, 2, C.off, (int)C.s[1], (int)C.s[2], C.n, m->dim, PUSH_DEPTH, dmask);
        }
    }
}

void mg_capture_coarse(vdn_ctx *c, MG *m, int from_level);

void vcycle(vdn_ctx *c, MG *m, int l)
{
    Lev &L = m->L[l];
    if (m->tail && l == m->agg_level) {
        // agglomerated coarse solve: every rank gathers the whole coarse right-hand side and finishes the V-cycle locally
        MG *t = m->tail;
        agg_gather(c, m, L.rhs, t->L[0].rhs, -1);
        VDN_CUDA(cudaMemsetAsync(t->L[0].phi, 0, sizeof(double) * t->L[0].ntot, c->stream));
        if (t->coarse_graph) {
            LaunchScope ls(c, "mg_coarse_levels_graph", 0.0, t->coarse_graph_launches);
            VDN_CUDA(cudaGraphLaunch(t->coarse_graph, c->stream));
        } else vcycle(c, t, 0);
        ExtArgs e; e.src = t->L[0].phi; e.soff = t->L[0].off; e.ss1 = t->L[0].s[1]; e.ss2 = t->L[0].s[2];
        e.dst = L.phi; e.doff = L.off; e.ds1 = L.s[1]; e.ds2 = L.s[2];
        const int *pc = comm_pcoord(c);
        for (int d = 0; d < 3; ++d) { e.ex[d] = L.n[d]; e.o[d] = pc[d] * L.n[d]; }
        LaunchScope ls(c, "mg_agglomerate", 0.0);
        const long tot = (long)L.n[0] * L.n[1] * L.n[2];
        k_blk_extract<<<(int)std::min<long>(592, (tot + 255) / 256), 256, 0, c->stream>>>(e);
        return;
    }
    if (l == m->tail_from) {
        LaunchScope ls(c, "mg_tail", 0.0);
        TailArgs a; a.nl = m->nlev - l;
        for (int q = 0; q < a.nl; ++q) a.L[q] = m->L[l + q];
        a.w = { m->bot[0], m->bot[1], m->bot[2], m->bot[3], m->bot[4], m->bot[5] };
        a.nu1 = c->prm.mg_nu1; a.nu2 = c->prm.mg_nu2; a.maxit = c->prm.mg_max_bottom_iter; a.eps = c->prm.mg_bottom_eps; a.singular = m->singular ? 1 : 0;
        for_dim_h(m->dim, m->helm, [&](auto D, auto H) { k_tail<decltype(D)::value, decltype(H)::value><<<1, 1024, 0, c->stream>>>(a); });
        return;
    }
    if (l == m->nlev - 1) {
        LaunchScope ls(c, "mg_bottom", 0.0);
        BotVec w = { m->bot[0], m->bot[1], m->bot[2], m->bot[3], m->bot[4], m->bot[5] };
        long nc = (long)L.n[0] * L.n[1] * L.n[2];
        int nt = nc >= 1024 ? 1024 : (int)std::max<long>(32, ((nc + 31) / 32) * 32);
        for_dim_h(m->dim, m->helm, [&](auto D, auto H) { k_bottom<decltype(D)::value, decltype(H)::value><<<1, nt, 0, c->stream>>>(L, w, c->prm.mg_max_bottom_iter, c->prm.mg_bottom_eps, m->singular ? 1 : 0); });
        return;
    }
    Lev &C = m->L[l + 1];
    if (l < m->nfused) {
        // fused path: [sweeps ... + residual + restriction] -> coarse -> [prolongation + sweeps ... (+ norm)]
        auto coarse = [&]() {
            if (m->coarse_graph && l + 1 == m->graph_level) {
                LaunchScope ls(c, "mg_coarse_levels_graph", 0.0, m->coarse_graph_launches);
                VDN_CUDA(cudaGraphLaunch(m->coarse_graph, c->stream));
            } else vcycle(c, m, l + 1);
        };
        if (l > 0 && !m->push && !m->pushk) mg_halo_deep(c, m, L, L.rhs, MG_PAD);        // restricted by the level above: valid cells only (peer-memory mode: pushed)
        int rem = c->prm.mg_nu1;
        if (l == 0 && m->first_sweep_done) { --rem; m->first_sweep_done = false; }      // launched ahead by the solve loop
        while (rem > 0) { --rem; wave_launch(c, m, l, 0, rem == 0 ? 2 : 0); }
        coarse();
        rem = c->prm.mg_nu2;
        bool first = true;
        while (rem > 0) { --rem; wave_launch(c, m, l, first ? 1 : 0, (rem == 0 && l == 0) ? 3 : 0); first = false; }
        return;
    }
    smooth(c, m, l, c->prm.mg_nu1);
    residual(c, m, l, nullptr);
    {
        LaunchScope ls(c, l == 0 ? "mg_restrict_l0" : "mg_restrict_coarse", (double)L.n[0] * L.n[1] * L.n[2] * 9.0);
        for_dim(m->dim, [&](auto D) { k_restrict<decltype(D)::value><<<cgrid(C.n[0], C.n[1], C.n[2]), BLK, 0, c->stream>>>(L, C); });
    }
    if (m->coarse_graph && l + 1 == m->graph_level) {
        LaunchScope ls(c, "mg_coarse_levels_graph", 0.0, m->coarse_graph_launches);
        VDN_CUDA(cudaGraphLaunch(m->coarse_graph, c->stream));
    } else
        vcycle(c, m, l + 1);
    {
        LaunchScope ls(c, l == 0 ? "mg_prolong_l0" : "mg_prolong_coarse", (double)L.n[0] * L.n[1] * L.n[2] * 17.0);
        for_dim(m->dim, [&](auto D) { k_prolong<decltype(D)::value><<<cgrid(L.n[0], L.n[1], L.n[2]), BLK, 0, c->stream>>>(L, C); });
    }
    smooth(c, m, l, c->prm.mg_nu2);
}

// capture levels from_level..bottom of the V-cycle into a CUDA graph: ~70 tiny launches become one
void mg_capture_coarse(vdn_ctx *c, MG *m, int from_level)
{
    if (m->nlev - from_level < 2 || m->coarse_graph || m->distributed) return;
    const bool prof = c->prof_on; c->prof_on = false;
    const long long l0 = c->launches;
    cudaGraph_t g = nullptr;
    VDN_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    try { vcycle(c, m, from_level); } catch (...) { cudaStreamEndCapture(c->stream, &g); if (g) cudaGraphDestroy(g); c->prof_on = prof; throw; }
    VDN_CUDA(cudaStreamEndCapture(c->stream, &g));
    m->coarse_graph_launches = (int)(c->launches - l0);
    m->graph_level = from_level;
    c->launches = l0;
    c->prof_on = prof;
    VDN_CUDA(cudaGraphInstantiate(&m->coarse_graph, g, 0));
    cudaGraphDestroy(g);
}

void coarsen_coefficients(vdn_ctx *c, MG *m)
{
    for (int l = 1; l < m->nlev; ++l) {
        Lev &F = m->L[l - 1], &C = m->L[l];
        LaunchScope ls(c, "mg_coarsen_beta", 0.0, m->dim);
        for (int d = 0; d < m->dim; ++d)
            for_dim(m->dim, [&](auto D) { k_coarsen_beta<decltype(D)::value><<<cgrid(C.n[0] + (d == 0), C.n[1] + (d == 1), C.n[2] + (d == 2 ? 1 : 0)), BLK, 0, c->stream>>>(F, C, d); });
    }
}

} // namespace

// ---- the Helmholtz solves of viscsolve.f90 (visc_solve :19, diff_scalar_solve :310), which reach the same
// mac_multigrid -> ml_cc_solve with alpha = rho (or 1) and beta = mu.  A separate hierarchy with its own level-0 arrays, the plain
// per-colour kernels with the alpha term (HELM instantiations), boundary modes per solved component, a non-zero initial guess. ----
namespace {
template <int DIM>
__global__ void k_coarsen_alpha(Lev F, Lev C)          // coarse alpha = average of the fine cells it covers (cell-average restriction)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= C.n[0] || j >= C.n[1]) return;
    const long cf = F.off + 2 * i + F.s[1] * (2 * j) + F.s[2] * (DIM == 3 ? 2 * k : 0);
    double s = F.alpha[cf] + F.alpha[cf + 1] + F.alpha[cf + F.s[1]] + F.alpha[cf + F.s[1] + 1];
    if (DIM == 3) s += F.alpha[cf + F.s[2]] + F.alpha[cf + F.s[2] + 1] + F.alpha[cf + F.s[2] + F.s[1]] + F.alpha[cf + F.s[2] + F.s[1] + 1];
    C.alpha[C.off + i + C.s[1] * j + C.s[2] * k] = s * (DIM == 3 ? 0.125 : 0.25);
}
__global__ void k_lev_absmax(Lev L, const double *x, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    double v = 0.0;
    if (i < L.n[0] && j < L.n[1]) v = fabs(x[L.off + i + L.s[1] * j + L.s[2] * k]);
    block_atomic_max(v, out);
}
} // namespace

static_assert((int)M_NEU == (int)VDN_MODE_NEU && (int)M_DIR == (int)VDN_MODE_DIR && (int)M_WRAP == (int)VDN_MODE_WRAP, "boundary mode codes");
// level-0 arrays of the Helmholtz hierarchy (built on first use): vdn_stream.cu fills them (st_helm_fill) and reads phi back
void mg_helm_level0(vdn_ctx *c, HelmLev0 *out)
{
    VDN_REQUIRE(!c->comm, "the Helmholtz solves (visc_solve / diff_scalar_solve) run on single-rank contexts only in this version");
    if (!c->mgh) {
        int mode[3][2];
        for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) mode[d][s] = M_NEU;
        c->mgh = mg_make(c, c->geo.n, c->geo.h, c->rlo, mode, false, -1, true);
    }
    const Lev &L = c->mgh->L[0];
    out->phi = L.phi; out->rhs = L.rhs; out->alpha = L.alpha;
    for (int d = 0; d < 3; ++d) out->b[d] = L.b[d];
    out->off = L.off; out->sy = L.s[1]; out->sz = L.s[2]; out->ntot = L.ntot;
}

// Solve with the level-0 arrays as filled by the caller; mode[d][side]: M_NEU / M_DIR / M_WRAP of the solved component.
int st_helm_solve(vdn_ctx *c, const int (*mode)[2], double rel_eps, double abs_eps, int *ncycles, double *resnorm)
{
    MG *m = c->mgh;
    VDN_REQUIRE(m != nullptr, "st_helm_solve before mg_helm_level0");
    for (int l = 0; l < m->nlev; ++l)
        for (int d = 0; d < 3; ++d) for (int s = 0; s < 2; ++s) m->L[l].mode[d][s] = d < m->dim ? mode[d][s] : M_NEU;
    coarsen_coefficients(c, m);
    for (int l = 1; l < m->nlev; ++l) {
        Lev &F = m->L[l - 1], &C = m->L[l];
        LaunchScope ls(c, "mg_coarsen_beta", 0.0);
        for_dim(m->dim, [&](auto D) { k_coarsen_alpha<decltype(D)::value><<<cgrid(C.n[0], C.n[1], C.n[2]), BLK, 0, c->stream>>>(F, C); });
    }
    VDN_CUDA(cudaGetLastError());
    Lev &L0 = m->L[0];
    auto pull_norm = [&]() {
        VDN_CUDA(cudaMemcpyAsync(c->h_pin, m->d_norm, 8, cudaMemcpyDeviceToHost, c->stream));
        VDN_CUDA(cudaStreamSynchronize(c->stream));
        return c->h_pin[0];
    };
    VDN_CUDA(cudaMemsetAsync(m->d_norm, 0, 8, c->stream));
    k_lev_absmax<<<cgrid(L0.n[0], L0.n[1], L0.n[2]), BLK, 0, c->stream>>>(L0, L0.rhs, m->d_norm);
    const double bnorm = pull_norm();
    auto res_norm = [&]() { residual(c, m, 0, m->d_norm); return pull_norm(); };
    double rn = res_norm();
    int cyc = 0;
    const bool talk = c->prm.mg_verbose != 0;
    if (talk) printf("vdn_mg (Helmholtz): levels %d  |rh| = %.6e  initial |r| = %.6e\n", m->nlev, bnorm, rn);
    auto converged = [&](double r) { return r <= rel_eps * bnorm || r <= abs_eps; };
    while (bnorm > 0.0 && !converged(rn) && cyc < c->prm.mg_max_cycles) {
        vcycle(c, m, 0);
        VDN_CUDA(cudaGetLastError());
        ++cyc;
        rn = res_norm();
        if (talk) printf("vdn_mg (Helmholtz): cycle %2d  |r|/|rh| = %.6e\n", cyc, rn / bnorm);
    }
    if (ncycles) *ncycles = cyc;
    if (resnorm) *resnorm = bnorm > 0.0 ? rn / bnorm : 0.0;
    return (bnorm > 0.0 && !converged(rn)) ? 1 : 0;
}

void mg_destroy(MG *m)
{
    if (!m) return;
    if (m->tail) mg_destroy(m->tail);
    if (m->coarse_graph) cudaGraphExecDestroy(m->coarse_graph);
    if (m->ev_norm) cudaEventDestroy(m->ev_norm);
    for (double *p : m->owned) cudaFree(p);
    for (int q = 0; q < 6; ++q) if (m->bot[q]) cudaFree(m->bot[q]);
    if (m->d_norm) cudaFree(m->d_norm);
    if (m->agg_send) cudaFree(m->agg_send);
    if (m->agg_recv) cudaFree(m->agg_recv);
    if (m->d_coords) cudaFree(m->d_coords);
    delete m;
}

// Solve with RH / BETA_* as right-hand side / coefficients and PHI as initial guess and result.
// phi_zero: the caller has just set PHI = 0, so the initial residual IS the right-hand side (no residual pass); bnorm_known >= 0: |rh|_inf as
// reduced by divumac in the pass that wrote rh (no separate norm pass).
int st_mac_solve(vdn_ctx *c, double rel_eps, double abs_eps, int *ncycles, double *resnorm, bool phi_zero, double bnorm_known)
{
    if (!c->mg) {
        mg_build(c);
        if (c->mg->tail) mg_capture_coarse(c, c->mg->tail, 0); else mg_capture_coarse(c, c->mg, std::max(1, c->mg->nfused));
    }
    MG *m = c->mg;
    // coefficient hierarchy
    coarsen_coefficients(c, m);
    if (m->tail) {
        Lev &A = m->L[m->agg_level];
        for (int d = 0; d < m->dim; ++d) agg_gather(c, m, A.b[d], m->tail->L[0].b[d], d);
        coarsen_coefficients(c, m->tail);
    }
    if (m->distributed)
        for (int l = 0; l < m->nfused; ++l) {
            for (int d = 0; d < m->dim; ++d) mg_halo_deep(c, m, m->L[l], m->L[l].b[d], MG_PAD);
            if (l == 0) mg_halo_deep(c, m, m->L[0], m->L[0].rhs, MG_PAD);
        }
    // inverse diagonal of the fused levels (after the coefficient ghost layers are in place: ghost cells are relaxed redundantly)
    for (int l = 0; l < m->nfused; ++l) {
        Lev &L = m->L[l];
        int lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            const int g = (d < m->dim && (L.mode[d][0] == M_GHOST || L.mode[d][1] == M_GHOST)) ? MG_PAD - 1 : 0;
            lo[d] = (d < m->dim && L.mode[d][0] == M_GHOST) ? -g : 0;
            hi[d] = L.n[d] - 1 + ((d < m->dim && L.mode[d][1] == M_GHOST) ? g : 0);
        }
        LaunchScope ls(c, "mg_diag_inv", (double)L.n[0] * L.n[1] * L.n[2] * 32.0);
        k_diag_inv<3><<<cgrid(hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1), BLK, 0, c->stream>>>(L, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    }
    VDN_CUDA(cudaGetLastError());
    const double bnorm = bnorm_known >= 0.0 ? bnorm_known : st_absmax_valid(c, VDN_RH);
    auto res_norm = [&]() {
        residual(c, m, 0, m->d_norm);
        VDN_CUDA(cudaMemcpyAsync(c->h_pin, m->d_norm, 8, cudaMemcpyDeviceToHost, c->stream));
        VDN_CUDA(cudaStreamSynchronize(c->stream));
        return comm_allreduce_max(c, c->h_pin[0]);
    };
    // peer-memory mode: the fused launches fill each other's ghost layers; the first one reads what the caller left there (phi = 0: zeros)
    if ((m->push || m->pushk) && !phi_zero) mg_halo_deep(c, m, m->L[0], m->L[0].phi, PUSH_DEPTH);
    double rn = phi_zero ? bnorm : res_norm();
    int cyc = 0;
    const bool talk = c->prm.mg_verbose && comm_rank(c) == 0;
    if (talk) printf("vdn_mg: levels %d%s  |rh| = %.6e  initial |r| = %.6e\n", m->nlev, m->tail ? " (+ agglomerated tail)" : "", bnorm, rn);
    auto converged = [&](double r) { return r <= rel_eps * bnorm || r <= abs_eps; };
    // Stagnation at the FP64 floor: the relative residual cannot fall below ~ eps_machine * max(beta) / h^2 * |phi| / |rh|, which grows with the
    // coefficient ratio and with n^2 (BASELINE config 5 at 512^3, density ratio 1000:1: contraction 0.47 per cycle down to 3e-10, then flat at
    // 1.4e-10, profiles/r02_config5_512_residuals.log).  Cycling on to mg_max_cycles cannot change the answer, so a solve whose residual has not
    // dropped by 20 % over three cycles AND sits within 10 x the tolerance (the parity bar of BASELINE.json:north_star) stops there; the caller
    // sees the residual it stopped at.  (F_MG would run into its iteration cap and abort.)
    double hist[3] = { 0.0, 0.0, 0.0 };
    auto stalled = [&](double r, int cyc) { return cyc >= 4 && hist[0] > 0.0 && r > 0.8 * hist[0] && (r <= 10.0 * rel_eps * bnorm || r <= 10.0 * abs_eps); };
    if (!m->ev_norm) VDN_CUDA(cudaEventCreateWithFlags(&m->ev_norm, cudaEventDisableTiming));
    bool stall = false;
    while (bnorm > 0.0 && !converged(rn) && !(stall = stalled(rn, cyc)) && cyc < c->prm.mg_max_cycles) {
        hist[0] = hist[1]; hist[1] = hist[2]; hist[2] = rn;
        vcycle(c, m, 0);
        VDN_CUDA(cudaGetLastError());
        ++cyc;
        if (m->nfused > 0) {
            // the up-leg kernel of level 0 already reduced |rhs - A phi|_inf: all-reduce it on the device and bring it to the host
            // asynchronously.  The convergence test costs a host round trip per V-cycle (and, between ranks, whatever their host threads
            // drift apart while they wait); it is hidden behind the FIRST smoothing sweep of the next cycle, which is launched before the
            // host looks at the norm.  If the solve turns out to be converged, that sweep has smoothed the solution once more -- harmless.
            comm_allreduce_max_dev(c, m->d_norm);
            VDN_CUDA(cudaMemcpyAsync(c->h_pin, m->d_norm, 8, cudaMemcpyDeviceToHost, c->stream));
            VDN_CUDA(cudaEventRecord(m->ev_norm, c->stream));
            if (c->prm.mg_nu1 >= 2 && cyc < c->prm.mg_max_cycles) { wave_launch(c, m, 0, 0, 0); m->first_sweep_done = true; }
            VDN_CUDA(cudaEventSynchronize(m->ev_norm));
            rn = c->h_pin[0];
        } else rn = res_norm();
        if (talk) printf("vdn_mg: cycle %2d  |r|/|rh| = %.6e\n", cyc, rn / bnorm);
    }
    m->first_sweep_done = false;
    if (m->nfused > 0 && m->L[0].phi != c->f[VDN_PHI].base) {       // odd number of ping-pong launches: result sits in the spare buffer
        VDN_CUDA(cudaMemcpyAsync(c->f[VDN_PHI].base, m->L[0].phi, sizeof(double) * m->L[0].ntot, cudaMemcpyDeviceToDevice, c->stream));
        std::swap(m->L[0].phi, m->L[0].res);
    }
    if (ncycles) *ncycles = cyc;
    if (resnorm) *resnorm = bnorm > 0.0 ? rn / bnorm : 0.0;
    if (talk && stall) printf("vdn_mg: stalled at the FP64 floor within 10 x the tolerance after %d cycles\n", cyc);
    return (bnorm > 0.0 && !converged(rn) && !stall) ? 1 : 0;
}
