// vdn_ctx.h -- host-side context of the B200 VARDEN hot path.
#pragma once
#include "vdn_common.cuh"
#include <array>


struct ProfEntry { std::string name; long long launches = 0; double ms = 0.0; double bytes = 0.0;
                   std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending; };

// ghost width of the multigrid level arrays (and of the PHI / RH / BETA_* fields that level 0 aliases): the fused wavefront
// smoother relaxes up to 3 layers of neighbour-rank cells redundantly instead of exchanging halos before every colour; 4 keeps
// every row of an even-sized level 32-byte aligned (16-byte pair loads, whole sectors)
constexpr int MG_PAD = 4;

struct MG;      // vdn_mg.cu
struct Comm;    // vdn_comm.cu

struct vdn_ctx {
    vdn_params prm;
    int dim = 3, device = 0;
    cudaStream_t stream = nullptr;
    Geo geo;
    int nboxes = 0;
    std::vector<std::array<int, 3>> box_lo, box_hi;    // global indices of the local reference boxes
    int rlo[3], rhi[3];                                 // region (global indices)
    int dom_lo[3], dom_hi[3], dom_bc[3][2];
    DField f[VDN_NFIELDS];
    int adv_bc[16][3][2];                               // [comp][d][side] on the region faces
    int ell_bc[3][2];                                   // pressure elliptic BC on the region faces
    bool wrap[3];                                       // periodic and owned by this rank alone in that direction
    // Godunov scratch arena of the staged 2-D kernels: NSCR arrays in the S-layout (cells -1..n, faces 0..n); unused in 3-D
    double *scratch = nullptr; int nscr = 0; long s_sy = 0, s_sz = 0, s_n = 0, s_off = 0;
    double *d_eps = nullptr;                            // per-box eps
    double *d_red = nullptr;                            // reduction scratch (device)
    unsigned long long *d_dbg = nullptr;                // measurement hook (vdn_debug_counters): 32 counters, allocated on first use
    double *h_pin = nullptr;                            // pinned host scalars
    double *stage = nullptr; size_t stage_bytes = 0;    // pinned staging buffer for pageable uploads
    double *xstage[4] = {}; size_t xstage_bytes[4] = {}; // device staging of host boxes (multi-box regions): [upload, upload on a copy stream, download, download on a copy stream]
    std::string err;
    long long launches = 0;
    long long comm_bytes = 0;                           // bytes this rank sent to other ranks (halo exchanges, all-gathers) since creation
    bool prof_on = false;
    std::vector<ProfEntry> prof;
    std::map<std::string, int> prof_idx;
    std::vector<cudaEvent_t> ev_pool;
    MG *mg = nullptr;
    MG *mgh = nullptr;                                  // Helmholtz hierarchy (visc_solve / diff_scalar_solve), built on first use
    Comm *comm = nullptr;
    long long umac_epoch = 0, eps_epoch = -1;          // UMAC_* write counter / counter value the per-box umac eps was computed at
    // vdn_advance_host: copy streams + per-field events (uploaded / final on the device)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_up[VDN_NFIELDS] = {}, ev_fin[VDN_NFIELDS] = {};
    const vdn_host_state *hio = nullptr;
    int mg_fuse_min = 128, mg_tile_force = -1;          // fused smoother: smallest level it runs on; test hook (vdn_mg_tune)
    int comm_mode = 0;                                  // test / measurement hook (vdn_comm_tune): 0 peer memory, fused levels push inside the sweep (default);
                                                        // 1 NCCL transport; 2 peer memory, pull kernel before every sweep; 3 peer memory, push kernel after every sweep
    bool lapu_set = false;                              // LAPU has been uploaded (required when visc_coef > 0)

    View S(int q) const { View v; v.sy = (int)s_sy; v.sz = (int)s_sz; v.cs = (int)s_n; v.p = scratch + (long)q * s_n + s_off; return v; }
    long ncells() const { return (long)geo.n[0] * geo.n[1] * geo.n[2]; }
};

// ---- profiling-aware launch bracket ----
struct LaunchScope {
    vdn_ctx *c; int idx = -1; cudaEvent_t e0 = nullptr, e1 = nullptr;
    LaunchScope(vdn_ctx *ctx, const char *name, double alg_bytes, int nlaunch = 1);
    ~LaunchScope();
};
void prof_collect(vdn_ctx *c);

// ---- stage implementations (each in its own .cu) ----
void ctx_require_comm(const vdn_ctx *c);                  // throws when the region is part of the domain and no communicator is set
void st_fill_boundary(vdn_ctx *c, int field);
void st_physbc(vdn_ctx *c, int field, int bccomp, bool same_boundary);
void st_mkvelforce(vdn_ctx *c, int rho_field, double visc_fac);
void st_mkscalforce(vdn_ctx *c, double diff_fac);
void st_velpred(vdn_ctx *c, double dt);
void st_mkflux(vdn_ctx *c, int is_vel, double dt);
void st_update(vdn_ctx *c, int is_vel, double dt);
void st_make_at_halftime(vdn_ctx *c);
double st_divumac(vdn_ctx *c, bool want_norm);
void st_mk_mac_coeffs(vdn_ctx *c);
void st_mkumac(vdn_ctx *c);
int  st_mac_solve(vdn_ctx *c, double rel_eps, double abs_eps, int *ncycles, double *resnorm, bool phi_zero = false, double bnorm_known = -1.0);
void st_setval(vdn_ctx *c, int field, double val);
// rows of a host box (staged flat on the device) <-> the region array
struct BoxCopyArgs { double *stage, *base; long cs; int hext[3], hofs[3], n[3], dofs[3], dext0, dext1, ncomp, upload; };
void st_box_copy(const BoxCopyArgs &a, cudaStream_t stream);
double st_estdt(vdn_ctx *c, double dtold, double cflfac, double max_dt_growth);   // estdt.f90:15-87 on the resident fields
void st_field_copy(vdn_ctx *c, int dst, int src);
double st_absmax_valid(vdn_ctx *c, int field);           // norm_inf over valid cells/faces, all comps
void mg_destroy(MG *mg);
// Helmholtz solves (vdn_mg.cu): level-0 arrays of the separate hierarchy, and the solve on them
enum : int { VDN_MODE_NEU = 1, VDN_MODE_DIR = 2, VDN_MODE_WRAP = 3 };      // = M_NEU / M_DIR / M_WRAP of vdn_mg_fused.cuh (checked in vdn_mg.cu)
struct HelmLev0 { double *phi, *rhs, *alpha, *b[3]; long off, sy, sz, ntot; };
void mg_helm_level0(vdn_ctx *c, HelmLev0 *out);
int  st_helm_solve(vdn_ctx *c, const int (*mode)[2], double rel_eps, double abs_eps, int *ncycles, double *resnorm);
int  st_visc_solve(vdn_ctx *c, double mu, int diffusion_type, int *ncycles, double *resnorm);          // viscsolve.f90:19
int  st_diff_scalar_solve(vdn_ctx *c, double mu, int icomp, int diffusion_type, int *ncycles, double *resnorm);   // viscsolve.f90:310
void comm_destroy(Comm *cm);
void comm_exchange_field(vdn_ctx *c, int field);          // ghost cells owned by neighbour ranks, every split direction at once (vdn_comm.cu)
const int *comm_pgrid_or_null(const vdn_ctx *c);
double comm_allreduce_max(vdn_ctx *c, double v);
double comm_allreduce_sum(vdn_ctx *c, double v);
