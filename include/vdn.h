/*
 * vdn.h -- C ABI of the B200-native VARDEN hot path (advection + MAC projection).
 *
 * This is the drop-in boundary: plain `extern "C"`, pointers and sizes only, `int`
 * status returns (0 = OK, non-zero = error, text via vdn_last_error).  Each entry point
 * names the reference interface it replaces (file:line under BoxLib-Codes/VARDEN).
 * The Fortran side keeps its module/procedure signatures and calls these through
 * ISO_C_BINDING (see fortran/vdn_iso_c.f90 and INTEGRATION.md).
 *
 * Host array convention (what `dataptr(mf,i)` points at in the reference):
 *   a(lo1-ng:hi1+ng, lo2-ng:hi2+ng, lo3-ng:hi3+ng, ncomp)   Fortran order, FP64,
 *   face ("edge") multifabs in direction d have hi_d+1; 2-D has a unit third extent.
 *
 * Process model: one context per GPU / MPI rank.  A context owns the device mirror of
 * the rank-local boxes, stored as ONE merged array per field over the rank's region
 * (the bounding box of its boxes, which must tile it).  Not re-entrant per context.
 */
#ifndef VDN_H
#define VDN_H

#ifdef __cplusplus
extern "C" {
#endif

/* physical BC codes = FBoxLib bc_module (define_bc_tower.f90:158-340, inputs bcx_lo ...) */
#define VDN_BC_PERIODIC     (-1)
#define VDN_BC_INTERIOR     0
#define VDN_BC_INLET        11
#define VDN_BC_OUTLET       12
#define VDN_BC_SYMMETRY     13
#define VDN_BC_SLIP_WALL    14
#define VDN_BC_NO_SLIP_WALL 15

/* Device-resident fields.  (ng, ncomp, centring) are fixed per field:
 *   UOLD/UNEW ng3 dm comps; SOLD/SNEW ng3 nscal comps; GP, EXT_VEL_FORCE, VEL_FORCE ng1 dm comps;
 *   EXT_SCAL_FORCE, SCAL_FORCE ng1 nscal comps; LAPU ng0 dm comps; MAC_RHS, RHOHALF, PHI ng1 1 comp;
 *   UMAC_X/Y/Z face, ng1, 1 comp; RH ng0; BETA_X/Y/Z face ng0;
 *   SEDGE_X/Y/Z face ng0 nscal comps; SFLUX_X/Y/Z face ng0 1 comp (density, the only conservative comp:
 *   scalar_advance.f90:54-57; the reference's other flux comps stay 0); UEDGE_X/Y/Z face ng0 dm comps.
 *   SEDGE_* and UEDGE_* share storage (the scalar and velocity phases never overlap). */
enum vdn_field {
    VDN_UOLD = 0, VDN_SOLD, VDN_UNEW, VDN_SNEW, VDN_GP,
    VDN_EXT_VEL_FORCE, VDN_EXT_SCAL_FORCE, VDN_LAPU,
    VDN_UMAC_X, VDN_UMAC_Y, VDN_UMAC_Z,
    VDN_MAC_RHS, VDN_RHOHALF, VDN_VEL_FORCE, VDN_SCAL_FORCE,
    VDN_RH, VDN_PHI, VDN_BETA_X, VDN_BETA_Y, VDN_BETA_Z,
    VDN_SEDGE_X, VDN_SEDGE_Y, VDN_SEDGE_Z,
    VDN_SFLUX_X, VDN_SFLUX_Y, VDN_SFLUX_Z,
    VDN_UEDGE_X, VDN_UEDGE_Y, VDN_UEDGE_Z,
    VDN_NFIELDS
};

/* The probin values the path reads (src/_parameters; probin_module uses in slope.f90:15,
 * velpred.f90:130, mkforce.f90:147, multifab_physbc.f90 u_bc..trac_bc). */
typedef struct vdn_params {
    int    nscal;              /* _parameters: nscal (2) */
    int    slope_order;        /* 0, 2 or 4 (4) */
    int    use_minion;         /* 0 */
    int    boussinesq;         /* 0 */
    int    stencil_order;      /* 2 (only 2 is implemented) */
    int    mg_verbose;         /* >0: per-cycle residual log on stdout */
    int    mg_nu1, mg_nu2;     /* pre/post smoothing sweeps (F_MG default 2,2) */
    int    mg_max_cycles;      /* V-cycle cap (100) */
    int    mg_max_bottom_iter; /* BiCGStab cap (100) */
    double mg_bottom_eps;      /* mac_multigrid.f90:56 -> 1.d-3 */
    double visc_coef, diff_coef;
    double bc_val[5][3][2];    /* u_bc, v_bc, w_bc, rho_bc, trac_bc (dir, side): EXT_DIR constants */
} vdn_params;

typedef struct vdn_ctx vdn_ctx;

/* Fill a vdn_params with the reference defaults (src/_parameters). */
void vdn_params_default(vdn_params *p);

/*
 * Create the device mirror of one level's rank-local layout.
 *   replaces: multifab_build / layout on the device side (advance_timestep.f90:66-80, macproject.f90:47-58)
 *             and bc_tower_level_build (define_bc_tower.f90:62-105; BC tables are derived here from phys_bc).
 * box_lo/box_hi: [nboxes][3] inclusive cell indices of the local boxes (lwb/upb of get_box);
 * dom_lo/dom_hi: level problem domain; phys_bc[3][2]: bcx_lo.. codes; dx[3]; device: CUDA ordinal.
 */
int vdn_ctx_create(const vdn_params *prm, int dim, int nboxes, const int *box_lo, const int *box_hi,
                   const int *dom_lo, const int *dom_hi, const int *phys_bc, const double *dx,
                   int device, vdn_ctx **out);
void vdn_ctx_destroy(vdn_ctx *ctx);
const char *vdn_last_error(const vdn_ctx *ctx);   /* ctx may be NULL: last create error */

/*
 * Multi-GPU (one context per rank).  region_lo/hi: [nranks][3] cell bounds of every rank's region
 * (a tensor-product decomposition of the domain); nccl_unique_id: 128 bytes from ncclGetUniqueId,
 * broadcast by the caller (torch.distributed / MPI).   replaces: layout_build_ba + parallel (MPI) in FBoxLib.
 */
int vdn_ctx_set_comm(vdn_ctx *ctx, int rank, int nranks, const int *region_lo, const int *region_hi,
                     const void *nccl_unique_id);
/* CUDA devices visible to this process (the Fortran shim maps MPI rank -> device = rank mod count); 0 when there is none */
int vdn_device_count(void);
/* rank 0 creates the 128-byte id (ncclGetUniqueId) that the caller broadcasts to every rank */
int vdn_nccl_unique_id(void *out128);
/* host-only: neighbour ranks nbr[3][2] (-1 = physical boundary, own rank = periodic self-wrap), process grid and this
 * rank's coordinates for a tensor-product decomposition.  Returns 2 if the regions are not such a decomposition. */
int vdn_comm_plan(int dim, int rank, int nranks, const int *region_lo, const int *region_hi,
                  const int *dom_lo, const int *dom_hi, const int *phys_bc, int *nbr, int *pgrid, int *pcoord);

/* host-only: the general exchange plan (cell- or face-centred arrays, multigrid level arrays with carry_n = 1); recv_shift: peer index = my
 * index + shift for every received box (what the peer-memory transport reads); see vdn_comm.cu */
int vdn_halo_plan_ex(int dim, const int *pgrid, const int *pcoord, const int *periodic, const int *coord2rank, const int *n, int ng,
                     int dmask, int nodal, int carry_n,
                     int *nsend, int *send_peer, int *send_lo, int *send_n, int *nrecv, int *recv_peer, int *recv_lo, int *recv_n, int *recv_shift);
/* host-only: message plan of the single-phase ghost-layer exchange of the multigrid level arrays (faces, edges and corners to the
 * up-to-26 neighbour ranks in one NCCL group).  Arrays hold up to 26 entries; *_lo / *_n are [entry][3] local index boxes.  Entries are
 * in issue order: NCCL matches the messages of a pair of ranks first-in first-out.  Returns 1 if more than 26 messages would be needed. */
int vdn_halo_plan(int dim, const int *pgrid, const int *pcoord, const int *periodic, const int *coord2rank, const int *n, int ng, int dmask,
                  int *nsend, int *send_peer, int *send_lo, int *send_n, int *nrecv, int *recv_peer, int *recv_lo, int *recv_n);

/* Path-boundary copies (SURVEY 8(b) "Copies"): host box array <-> region array, ghosts included
 * where they lie outside the region's valid area.  `host` has `ng` ghosts and `ncomp` comps and
 * must match the field's fixed (ng, ncomp). */
int vdn_field_upload(vdn_ctx *ctx, int field, int ibox, const double *host, int ng, int ncomp);
int vdn_field_download(vdn_ctx *ctx, int field, int ibox, double *host, int ng, int ncomp);
int vdn_field_setval(vdn_ctx *ctx, int field, double val);       /* multifab setval(all=.true.) */
int vdn_sync(vdn_ctx *ctx);
/* the cudaStream_t every kernel of this context is launched on (for CUDA-event timing by the caller) */
void *vdn_get_stream(vdn_ctx *ctx);

/* ---- stage calls: same names / argument meaning as the reference module procedures ---- */

/* multifab_fill_boundary (FBoxLib; call sites velpred.f90:110, macproject.f90:117,492) */
int vdn_fill_boundary(vdn_ctx *ctx, int field);
/* ml_restrict_and_fill for nlevs==1 = fill_boundary + multifab_physbc (multifab_physbc.f90:17).
 * bccomp: 0-based index into the adv_bc table (0..dm-1 vel, dm.. scalars, dm+nscal press, dm+nscal+1 extrap). */
int vdn_fill_and_physbc(vdn_ctx *ctx, int field, int bccomp, int same_boundary);

/* mkvelforce (mkforce.f90:18): VEL_FORCE = ext [*s2] + (visc_coef*visc_fac*lapu - gp)/rho; rho_field = VDN_SOLD or VDN_RHOHALF */
int vdn_mkvelforce(vdn_ctx *ctx, int rho_field, double visc_fac);
/* mkscalforce (mkforce.f90:238): SCAL_FORCE, laps = 0 (diff_coef == 0 path) */
int vdn_mkscalforce(vdn_ctx *ctx, double diff_fac);
/* velpred (velpred.f90:16): UOLD, VEL_FORCE -> UMAC_* (+ fill_boundary) */
int vdn_velpred(vdn_ctx *ctx, double dt);
/* macproject (macproject.f90:20): UMAC_*, rho = SOLD comp 1, MAC_RHS -> projected UMAC_*, PHI.
 * rel_eps <= 0 selects the reference's 1.d-10 (macproject.f90:92).  Returns 2 if not converged. */
int vdn_macproject(vdn_ctx *ctx, double rel_eps, double abs_eps, int *ncycles, double *resnorm);
/* mkflux (mkflux.f90:16): is_vel=0: SOLD,SCAL_FORCE -> SEDGE_*,SFLUX_* (density conservative);
 *                         is_vel=1: UOLD,VEL_FORCE,MAC_RHS -> UEDGE_* */
int vdn_mkflux(vdn_ctx *ctx, int is_vel, double dt);
/* update (update.f90:16) + ghost fill: is_vel=0 -> SNEW, is_vel=1 -> UNEW */
int vdn_update(vdn_ctx *ctx, int is_vel, double dt);
/* make_at_halftime (make_at_halftime.f90:18): RHOHALF = 0.5*(SOLD_1 + SNEW_1) + ghost fill */
int vdn_make_at_halftime(vdn_ctx *ctx);

/* ---- the driver's per-step glue around the path, on the resident fields ----
 * estdt (estdt.f90:15-87, per-box kernels :89-181): dt for the coming step from UOLD, SOLD (density), GP and EXT_VEL_FORCE -- the advective
 * limit dx/max|u|, the forcing limit sqrt(2 dx / max|gp/rho - f|) (each only above the reference's single-precision eps 1.0e-8), min(dx) when
 * nothing limits, times cflfac, capped by max_dt_growth * dtold when dtold > 0.  Reduced over all ranks (the reference's parallel_reduce). */
int vdn_estdt(vdn_ctx *ctx, double dtold, double cflfac, double max_dt_growth, double *dt);
/* multifab_copy_c of a whole field on the device, ghost cells included (varden.f90:321-324: uold <- unew, sold <- snew); both fields must
 * have the same layout.  With vdn_fill_and_physbc (varden.f90:291-300) this keeps a run resident between steps wherever the Fortran driver
 * does not need the host copy. */
int vdn_field_copy(vdn_ctx *ctx, int dst_field, int src_field);

/* ---- the implicit viscous / diffusive solves that follow the path (single-rank contexts in this version) ----
 * visc_solve (viscsolve.f90:19): for every velocity component d, (RHOHALF - mu div grad) u_d = RHOHALF u_d [+ mu LAPU_d] + (1/3) visc_mu_dt
 * d(MAC_RHS)/dx_d on UNEW, boundary types of component d (define_bc_tower.f90:254-340), Dirichlet data from UNEW's ghost cells, initial
 * guess UNEW, rel. tolerance 1.d-12 (viscsolve.f90:88); then UNEW's ghost cells are refilled (:105).  mu = (1/2) dt visc_coef with
 * diffusion_type 1 (Crank-Nicolson; needs LAPU) or dt visc_coef with diffusion_type 2 (velocity_advance.f90:105-111).  ncycles: V-cycles
 * summed over the components; resnorm: the largest final relative residual.  Returns 2 when the V-cycle cap is hit. */
int vdn_visc_solve(vdn_ctx *ctx, double mu, int diffusion_type, int *ncycles, double *resnorm);
/* diff_scalar_solve (viscsolve.f90:310): (1 - mu div grad) s = s on component icomp (0-based) of SNEW, boundary types of that scalar;
 * diffusion_type 2 only in this version (the Crank-Nicolson form needs the explicit term laps, for which the context has no field yet). */
int vdn_diff_scalar_solve(vdn_ctx *ctx, double mu, int icomp, int diffusion_type, int *ncycles, double *resnorm);

/* The whole device-resident path advance_timestep.f90:95-124:
 * advance_premac -> macproject -> scalar_advance -> make_at_halftime -> velocity_advance. */
int vdn_advance(vdn_ctx *ctx, double dt, double mac_rel_eps, int *mac_cycles, double *mac_resnorm);

/* The same pass driven from HOST multifabs, i.e. what the Fortran advance_timestep hands over and takes back
 * (advance_timestep.f90:26-27 arguments; SURVEY 8(b) "Copies"): every pointer array has one entry per local box
 * (dataptr(mf,i)), ghost widths / components as in the field table above.  Uploads run on a copy stream in the order
 * the stages first read them (ext_vel_force, gp, sold | uold | ext_scal_force) and each stage waits only for its own
 * inputs; snew and rhohalf travel back while velocity_advance is still running, unew last.  Returns when every output
 * array is complete on the host.  Host arrays should be page-locked (cudaHostRegister once per layout) for the copies to
 * overlap; pageable memory works but serialises. */
typedef struct vdn_host_state {
    const double *const *uold, *const *sold, *const *gp, *const *ext_vel_force, *const *ext_scal_force;   /* in  */
    double *const *unew, *const *snew, *const *rhohalf;                                                    /* out */
    /* optional inputs, NULL when absent: lapu (ng 0, dm comps) = the explicit viscous term the reference computes before the path
     * (get_explicit_diffusive_term, advance_timestep.f90:84-88; REQUIRED when visc_coef > 0), mac_rhs (ng 1, 1 comp; advance_timestep.f90:66,
     * zero in VARDEN unless a divergence constraint is set) */
    const double *const *lapu, *const *mac_rhs;
} vdn_host_state;
int vdn_advance_host(vdn_ctx *ctx, double dt, double mac_rel_eps, const vdn_host_state *hs, int *mac_cycles, double *mac_resnorm);

/* ---- building blocks exposed for parity tests (macproject.f90 internal procedures) ---- */
int vdn_divumac(vdn_ctx *ctx, double *rhmax);     /* RH = MAC_RHS - div(UMAC), macproject.f90:137 */
int vdn_mk_mac_coeffs(vdn_ctx *ctx);              /* BETA_* from SOLD comp 1, macproject.f90:280 */
int vdn_mac_solve(vdn_ctx *ctx, double rel_eps, double abs_eps, int *ncycles, double *resnorm); /* mac_multigrid.f90:19 */
int vdn_mkumac(vdn_ctx *ctx);                     /* macproject.f90:403 */

/* test hook: smallest level size (cells per direction) the fused smoother runs on (default 128; the parity tests lower it so that small
 * grids exercise the production kernel) and a forced tile shape of it (-1 = the measured default per launch kind).  Drops the cached
 * multigrid hierarchy; takes effect at the next solve. */
int vdn_mg_tune(vdn_ctx *ctx, int fuse_min, int tile);

/* test / measurement hook, before vdn_ctx_set_comm: transport of the ghost exchanges between ranks.
 *   0 (default) peer memory (CUDA-IPC mapped symmetric heap): one pull kernel per field exchange; the fused multigrid sweeps store their
 *               boundary results straight into the neighbours' ghost layers (no exchange launches inside a V-cycle)
 *   1           NCCL (pack + grouped send/recv + unpack)
 *   2           peer memory, one pull kernel before every fused sweep
 *   3           peer memory, one push kernel after every fused sweep */
int vdn_comm_tune(vdn_ctx *ctx, int force_nccl);

/* measurement hook: 32 device counters -- per fused-smoother kernel family f (0 smooth, 1 down, 2 pro, 3 up of level 0, 4 coarser levels):
 * out32[4f] = ns its boundary CTAs spent waiting for a neighbour rank's flag, [4f+1] = longest wait, [4f+2] = number of waits.  The first call
 * switches the accounting on; every call returns the counters and resets them. */
int vdn_debug_counters(vdn_ctx *ctx, unsigned long long *out32);

/* ---- measurement: per-kernel-family CUDA-event timing on the launching stream ---- */
int vdn_prof_enable(vdn_ctx *ctx, int on);        /* resets counters */
int vdn_prof_count(vdn_ctx *ctx);                 /* number of kernel families seen */
/* name (<=63 chars), launches, total device ms, algorithmic bytes (SURVEY 8(a) figure x cells) */
int vdn_prof_get(vdn_ctx *ctx, int idx, char *name, long long *launches, double *ms, double *alg_bytes);
long long vdn_launch_count(vdn_ctx *ctx);         /* kernels launched since creation */
long long vdn_comm_bytes(vdn_ctx *ctx);           /* bytes this rank sent to other ranks over NVLink since creation (halo exchanges, all-gathers) */

#ifdef __cplusplus
}
#endif
#endif
