! fortran/vdn_iso_c.f90 -- ISO_C_BINDING interface to libvdn.so (include/vdn.h) and the drop-in replacement
! of the device-resident segment of advance_timestep (src/advance_timestep.f90:95-124).
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build container has no Fortran compiler and no FBoxLib/AMReX
! (SURVEY.md, BASELINE.md section 2).  The C ABI it binds is exercised through ctypes by tests/ instead;
! the array layout contract is stated in include/vdn.h ("Host array convention").
!
! Two levels of integration (see INTEGRATION.md); both keep every caller's source unchanged up to the point named:
!   (1) literal drop-in: fortran/vdn_modules.f90 supplies velpred_module, mkflux_module, update_module, macproject_module and
!       mac_multigrid_module with the reference's own procedure names and argument lists (each call copies its multifabs in and out);
!   (2) fused: in advance_timestep replace the five calls advance_premac / macproject / scalar_advance / make_at_halftime /
!       velocity_advance (advance_timestep.f90:95-124) by
!         call vdn_advance_path(mla,sold,uold,snew,unew,gp,ext_vel_force,ext_scal_force,rhohalf,umac,lapu,mac_rhs,the_bc_tower,dt,dx)
!       which keeps the fields device-resident across them (copies only at the path boundary).
! Everything before (lapu, print_old) and after (hgproject, print_new, timers) stays the reference Fortran.

module vdn_iso_c

  use iso_c_binding
  implicit none

  ! enum vdn_field (include/vdn.h)
  integer(c_int), parameter :: VDN_UOLD = 0, VDN_SOLD = 1, VDN_UNEW = 2, VDN_SNEW = 3, VDN_GP = 4, &
       VDN_EXT_VEL_FORCE = 5, VDN_EXT_SCAL_FORCE = 6, VDN_LAPU = 7, VDN_UMAC_X = 8, VDN_UMAC_Y = 9, VDN_UMAC_Z = 10, &
       VDN_MAC_RHS = 11, VDN_RHOHALF = 12, VDN_VEL_FORCE = 13, VDN_SCAL_FORCE = 14, VDN_RH = 15, VDN_PHI = 16, &
       VDN_BETA_X = 17, VDN_BETA_Y = 18, VDN_BETA_Z = 19, VDN_SEDGE_X = 20, VDN_SEDGE_Y = 21, VDN_SEDGE_Z = 22, &
       VDN_SFLUX_X = 23, VDN_SFLUX_Y = 24, VDN_SFLUX_Z = 25, VDN_UEDGE_X = 26, VDN_UEDGE_Y = 27, VDN_UEDGE_Z = 28

  type, bind(c) :: vdn_params
     integer(c_int) :: nscal, slope_order, use_minion, boussinesq, stencil_order, mg_verbose
     integer(c_int) :: mg_nu1, mg_nu2, mg_max_cycles, mg_max_bottom_iter
     real(c_double) :: mg_bottom_eps, visc_coef, diff_coef
     real(c_double) :: bc_val(2,3,5)     ! C bc_val[5][3][2]: (side, dir, {u,v,w,rho,trac})
  end type vdn_params

  ! struct vdn_host_state (include/vdn.h): arrays of nboxes pointers, one per local fab (c_loc(dataptr(mf,i)))
  type, bind(c) :: vdn_host_state
     type(c_ptr) :: uold, sold, gp, ext_vel_force, ext_scal_force      ! in
     type(c_ptr) :: unew, snew, rhohalf                                ! out
     type(c_ptr) :: lapu, mac_rhs                                      ! optional in (c_null_ptr when absent)
  end type vdn_host_state

  interface
     subroutine vdn_params_default(p) bind(c, name='vdn_params_default')
       import :: vdn_params
       type(vdn_params), intent(out) :: p
     end subroutine vdn_params_default

     integer(c_int) function vdn_ctx_create(prm, dim, nboxes, box_lo, box_hi, dom_lo, dom_hi, phys_bc, dx, device, ctx) &
          bind(c, name='vdn_ctx_create')
       import :: vdn_params, c_int, c_double, c_ptr
       type(vdn_params), intent(in) :: prm
       integer(c_int), value :: dim, nboxes, device
       integer(c_int), intent(in) :: box_lo(3,*), box_hi(3,*), dom_lo(3), dom_hi(3), phys_bc(2,3)
       real(c_double), intent(in) :: dx(3)
       type(c_ptr), intent(out) :: ctx
     end function vdn_ctx_create

     subroutine vdn_ctx_destroy(ctx) bind(c, name='vdn_ctx_destroy')
       import :: c_ptr
       type(c_ptr), value :: ctx
     end subroutine vdn_ctx_destroy

     type(c_ptr) function vdn_last_error(ctx) bind(c, name='vdn_last_error')
       import :: c_ptr
       type(c_ptr), value :: ctx
     end function vdn_last_error

     integer(c_int) function vdn_field_upload(ctx, field, ibox, host, ng, ncomp) bind(c, name='vdn_field_upload')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx, host
       integer(c_int), value :: field, ibox, ng, ncomp
     end function vdn_field_upload

     integer(c_int) function vdn_field_download(ctx, field, ibox, host, ng, ncomp) bind(c, name='vdn_field_download')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx, host
       integer(c_int), value :: field, ibox, ng, ncomp
     end function vdn_field_download

     integer(c_int) function vdn_advance(ctx, dt, mac_rel_eps, mac_cycles, mac_resnorm) bind(c, name='vdn_advance')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: dt, mac_rel_eps
       integer(c_int), intent(out) :: mac_cycles
       real(c_double), intent(out) :: mac_resnorm
     end function vdn_advance

     ! the same pass from / to HOST multifabs, copies overlapped with the stages (include/vdn.h: vdn_advance_host)
     integer(c_int) function vdn_advance_host(ctx, dt, mac_rel_eps, hs, mac_cycles, mac_resnorm) bind(c, name='vdn_advance_host')
       import :: c_ptr, c_int, c_double, vdn_host_state
       type(c_ptr), value :: ctx
       real(c_double), value :: dt, mac_rel_eps
       type(vdn_host_state), intent(in) :: hs
       integer(c_int), intent(out) :: mac_cycles
       real(c_double), intent(out) :: mac_resnorm
     end function vdn_advance_host

     ! stage-wise entry points (same names as the reference procedures)
     integer(c_int) function vdn_mkvelforce(ctx, rho_field, visc_fac) bind(c, name='vdn_mkvelforce')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: rho_field
       real(c_double), value :: visc_fac
     end function vdn_mkvelforce
     integer(c_int) function vdn_velpred(ctx, dt) bind(c, name='vdn_velpred')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: dt
     end function vdn_velpred
     integer(c_int) function vdn_macproject(ctx, rel_eps, abs_eps, ncycles, resnorm) bind(c, name='vdn_macproject')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: rel_eps, abs_eps
       integer(c_int), intent(out) :: ncycles
       real(c_double), intent(out) :: resnorm
     end function vdn_macproject
     integer(c_int) function vdn_mkflux(ctx, is_vel, dt) bind(c, name='vdn_mkflux')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: is_vel
       real(c_double), value :: dt
     end function vdn_mkflux
     integer(c_int) function vdn_update(ctx, is_vel, dt) bind(c, name='vdn_update')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: is_vel
       real(c_double), value :: dt
     end function vdn_update
     integer(c_int) function vdn_mac_solve(ctx, rel_eps, abs_eps, ncycles, resnorm) bind(c, name='vdn_mac_solve')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: rel_eps, abs_eps
       integer(c_int), intent(out) :: ncycles
       real(c_double), intent(out) :: resnorm
     end function vdn_mac_solve
     integer(c_int) function vdn_fill_boundary(ctx, field) bind(c, name='vdn_fill_boundary')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
       integer(c_int), value :: field
     end function vdn_fill_boundary
     integer(c_int) function vdn_mkscalforce(ctx, diff_fac) bind(c, name='vdn_mkscalforce')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: diff_fac
     end function vdn_mkscalforce
     integer(c_int) function vdn_make_at_halftime(ctx) bind(c, name='vdn_make_at_halftime')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
     end function vdn_make_at_halftime

     ! multi-rank: one context per MPI rank, NCCL communicator from a unique id created on rank 0 and broadcast by the caller
     ! SURVEY 8(f) row 3: the driver's per-step glue on the resident fields (include/vdn.h: vdn_estdt, vdn_field_copy)
     integer(c_int) function vdn_estdt(ctx, dtold, cflfac, max_dt_growth, dt) bind(c, name='vdn_estdt')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: dtold, cflfac, max_dt_growth
       real(c_double), intent(out) :: dt
     end function vdn_estdt
     integer(c_int) function vdn_field_copy(ctx, dst_field, src_field) bind(c, name='vdn_field_copy')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: dst_field, src_field
     end function vdn_field_copy

     ! SURVEY 8(f) row 2: the implicit viscous / diffusive solves (include/vdn.h: vdn_visc_solve, vdn_diff_scalar_solve)
     integer(c_int) function vdn_visc_solve(ctx, mu, diffusion_type, ncycles, resnorm) bind(c, name='vdn_visc_solve')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: mu
       integer(c_int), value :: diffusion_type
       integer(c_int), intent(out) :: ncycles
       real(c_double), intent(out) :: resnorm
     end function vdn_visc_solve
     integer(c_int) function vdn_diff_scalar_solve(ctx, mu, icomp, diffusion_type, ncycles, resnorm) bind(c, name='vdn_diff_scalar_solve')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), value :: mu
       integer(c_int), value :: icomp, diffusion_type
       integer(c_int), intent(out) :: ncycles
       real(c_double), intent(out) :: resnorm
     end function vdn_diff_scalar_solve

     integer(c_int) function vdn_device_count() bind(c, name='vdn_device_count')
       import :: c_int
     end function vdn_device_count
     integer(c_int) function vdn_nccl_unique_id(out128) bind(c, name='vdn_nccl_unique_id')
       import :: c_int
       integer(c_int), intent(out) :: out128(32)          ! 128 bytes
     end function vdn_nccl_unique_id
     integer(c_int) function vdn_ctx_set_comm(ctx, rank, nranks, region_lo, region_hi, nccl_unique_id) bind(c, name='vdn_ctx_set_comm')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
       integer(c_int), value :: rank, nranks
       integer(c_int), intent(in) :: region_lo(3,*), region_hi(3,*), nccl_unique_id(32)
     end function vdn_ctx_set_comm
  end interface

end module vdn_iso_c


module vdn_path_module

  use iso_c_binding
  use vdn_iso_c
  use bl_types
  use multifab_module
  use ml_layout_module
  use define_bc_module
  use bl_error_module
  use parallel

  implicit none
  private
  public :: vdn_advance_path, vdn_path_finalize, vdn_ctx_for, vdn_ctx_current, vdn_put, vdn_get, vdn_check

  type(c_ptr), save :: ctx = c_null_ptr      ! one context per MPI rank; rebuilt after regrid (call vdn_path_finalize)
  type(c_ptr), allocatable, target, save :: fab_tab(:,:)   ! (local fab, multifab slot): host pointers handed to vdn_advance_host

contains

  ! error convention of the reference: bl_error aborts the run (velocity_advance.f90:113, multifab_physbc.f90:125);
  ! the library's message (vdn_last_error) becomes the bl_error text
  subroutine vdn_check(c, rc)
    type(c_ptr), intent(in) :: c
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: text
    type(c_ptr) :: p
    integer :: k
    if (rc == 0) return
    text = ''
    p = vdn_last_error(c)
    if (c_associated(p)) then
       call c_f_pointer(p, msg, [512])
       do k = 1, 512
          if (msg(k) == c_null_char) exit
          text(k:k) = msg(k)
       end do
    end if
    call bl_error('libvdn: ' // trim(text))
  end subroutine vdn_check

  ! the rank's context for this layout (created on first use).  Device = local rank modulo the GPUs of the node; with more
  ! than one MPI rank every rank's region is gathered and the NCCL unique id of rank 0 is broadcast (FBoxLib parallel module)
  function vdn_ctx_for(mla, mf, dx) result(c)
    type(ml_layout), intent(in) :: mla
    type(multifab) , intent(in) :: mf
    real(dp_t)     , intent(in) :: dx(:,:)
    type(c_ptr) :: c
    if (.not. c_associated(ctx)) call vdn_path_init(mla, mf, dx)
    c = ctx
  end function vdn_ctx_for

  ! the context that an earlier path call created (estdt has no ml_layout argument; the driver first calls it at istep = 2,
  ! varden.f90:302, after the first advance_timestep has built the context)
  function vdn_ctx_current() result(c)
    type(c_ptr) :: c
    if (.not. c_associated(ctx)) call bl_error('libvdn: no context yet (estdt before the first advance_timestep)')
    c = ctx
  end function vdn_ctx_current

  subroutine vdn_path_init(mla, mf, dx)
    use probin_module, only: nscal, slope_order, use_minion, boussinesq, stencil_order, mg_verbose, visc_coef, diff_coef, &
                             u_bc, v_bc, w_bc, rho_bc, trac_bc, bcx_lo, bcx_hi, bcy_lo, bcy_hi, bcz_lo, bcz_hi
    type(ml_layout), intent(in) :: mla
    type(multifab) , intent(in) :: mf
    real(dp_t)     , intent(in) :: dx(:,:)
    type(vdn_params) :: prm
    integer(c_int), allocatable :: blo(:,:), bhi(:,:), rlo(:,:), rhi(:,:)
    integer(c_int) :: dlo(3), dhi(3), pbc(2,3), myreg(6), uid(32), ndev
    integer, allocatable :: allreg(:)
    real(c_double) :: cdx(3)
    type(box) :: pd
    integer :: i, d, dm, nb, np, me

    dm = mla%dim
    nb = nfabs(mf)
    allocate(blo(3,nb), bhi(3,nb)); blo = 0; bhi = 0
    do i = 1, nb
       blo(1:dm,i) = lwb(get_box(mf,i)); bhi(1:dm,i) = upb(get_box(mf,i))
    end do
    pd = ml_layout_get_pd(mla,1)
    dlo = 0; dhi = 0; dlo(1:dm) = lwb(pd); dhi(1:dm) = upb(pd)
    cdx = 0.d0; cdx(1:dm) = dx(1,1:dm)
    ! domain BCs = the inputs bcx_lo .. bcz_hi (what define_bc_tower.f90:129-154 puts into phys_bc_level_array(0,:,:))
    pbc = 0
    pbc(1,1) = bcx_lo; pbc(2,1) = bcx_hi; pbc(1,2) = bcy_lo; pbc(2,2) = bcy_hi
    if (dm == 3) then
       pbc(1,3) = bcz_lo; pbc(2,3) = bcz_hi
    end if
    call vdn_params_default(prm)
    prm%nscal = nscal; prm%slope_order = slope_order; prm%use_minion = merge(1,0,use_minion)
    prm%boussinesq = boussinesq; prm%stencil_order = stencil_order; prm%mg_verbose = mg_verbose
    prm%visc_coef = visc_coef; prm%diff_coef = diff_coef
    do d = 1, dm
       prm%bc_val(:,d,1) = u_bc(d,:); prm%bc_val(:,d,2) = v_bc(d,:); prm%bc_val(:,d,3) = w_bc(d,:)
       prm%bc_val(:,d,4) = rho_bc(d,:); prm%bc_val(:,d,5) = trac_bc(d,:)
    end do
    np = parallel_nprocs(); me = parallel_myproc()
    ndev = vdn_device_count()
    if (ndev < 1) call bl_error('libvdn: no CUDA device (the hot path has no CPU fallback)')
    call vdn_check(c_null_ptr, vdn_ctx_create(prm, int(dm,c_int), int(nb,c_int), blo, bhi, dlo, dhi, pbc, cdx, &
                                              int(mod(me, ndev),c_int), ctx))
    if (np > 1) then
       ! every rank's region = bounding box of its boxes (they must tile it: a block distribution of boxarray_maxsize boxes)
       myreg(1:3) = minval(blo, dim=2); myreg(4:6) = maxval(bhi, dim=2)
       allocate(allreg(6*np), rlo(3,np), rhi(3,np))
       call parallel_allgather(int(myreg), allreg, 6)
       do i = 1, np
          rlo(:,i) = allreg(6*(i-1)+1:6*(i-1)+3); rhi(:,i) = allreg(6*(i-1)+4:6*(i-1)+6)
       end do
       uid = 0
       if (me == parallel_IOProcessorNode()) call vdn_check(ctx, vdn_nccl_unique_id(uid))
       call parallel_bcast(uid, parallel_IOProcessorNode())
       call vdn_check(ctx, vdn_ctx_set_comm(ctx, int(me,c_int), int(np,c_int), rlo, rhi, uid))
    end if
  end subroutine vdn_path_init

  subroutine vdn_path_finalize()
    if (c_associated(ctx)) call vdn_ctx_destroy(ctx)
    ctx = c_null_ptr
    if (allocated(fab_tab)) deallocate(fab_tab)
  end subroutine vdn_path_finalize

  ! host multifab -> device field (all local fabs); nc: number of leading components to copy (default: all)
  subroutine vdn_put(c, field, mf, nc)
    type(c_ptr), intent(in) :: c
    integer(c_int), intent(in) :: field
    type(multifab), intent(in) :: mf
    integer, intent(in), optional :: nc
    real(dp_t), pointer :: p(:,:,:,:)
    integer :: i, n
    n = ncomp(mf); if (present(nc)) n = nc
    do i = 1, nfabs(mf)
       p => dataptr(mf, i)
       call vdn_check(c, vdn_field_upload(c, field, int(i-1,c_int), c_loc(p(lbound(p,1),lbound(p,2),lbound(p,3),1)), &
                                          int(nghost(mf),c_int), int(n,c_int)))
    end do
  end subroutine vdn_put

  subroutine vdn_get(c, field, mf, nc)
    type(c_ptr), intent(in) :: c
    integer(c_int), intent(in) :: field
    type(multifab), intent(inout) :: mf
    integer, intent(in), optional :: nc
    real(dp_t), pointer :: p(:,:,:,:)
    integer :: i, n
    n = ncomp(mf); if (present(nc)) n = nc
    do i = 1, nfabs(mf)
       p => dataptr(mf, i)
       call vdn_check(c, vdn_field_download(c, field, int(i-1,c_int), c_loc(p(lbound(p,1),lbound(p,2),lbound(p,3),1)), &
                                            int(nghost(mf),c_int), int(n,c_int)))
    end do
  end subroutine vdn_get

  ! c_ptr array of the local fabs of mf (slot = which of the ten multifabs of a step; storage lives in the module)
  function fabptrs(mf, slot) result(p)
    type(multifab), intent(in) :: mf
    integer, intent(in) :: slot
    type(c_ptr) :: p
    real(dp_t), pointer :: q(:,:,:,:)
    integer :: i
    if (.not. allocated(fab_tab)) allocate(fab_tab(nfabs(mf), 10))
    do i = 1, nfabs(mf)
       q => dataptr(mf, i)
       fab_tab(i, slot) = c_loc(q(lbound(q,1),lbound(q,2),lbound(q,3),1))
    end do
    p = c_loc(fab_tab(1, slot))
  end function fabptrs

  ! Replaces advance_timestep.f90:95-124 for nlevs == 1 (diff_coef == 0; visc_coef > 0 needs lapu, which the reference computes
  ! at advance_timestep.f90:84-88 and hands over here; visc_solve after the path stays the reference's).
  subroutine vdn_advance_path(mla,sold,uold,snew,unew,gp,ext_vel_force,ext_scal_force,rhohalf,umac,lapu,mac_rhs,the_bc_tower,dt,dx)
    use probin_module, only: visc_coef
    type(ml_layout), intent(in   ) :: mla
    type(multifab) , intent(in   ) :: sold(:), uold(:), gp(:), ext_vel_force(:), ext_scal_force(:), lapu(:), mac_rhs(:)
    type(multifab) , intent(inout) :: snew(:), unew(:), rhohalf(:), umac(:,:)
    type(bc_tower) , intent(in   ) :: the_bc_tower
    real(dp_t)     , intent(in   ) :: dt, dx(:,:)
    integer(c_int) :: ncyc
    real(c_double) :: res
    type(vdn_host_state) :: hs
    type(c_ptr) :: c
    integer :: d

    if (mla%nlevel /= 1) call bl_error('vdn_advance_path: single-level only')
    c = vdn_ctx_for(mla, sold(1), dx)
    ! copies only at the path boundary (BASELINE.json north_star): one call that overlaps the copies with the stages.
    ! fabptrs(mf) returns a c_ptr array with c_loc(dataptr(mf,i)) for the local fabs (module storage, valid until the call returns).
    hs%uold = fabptrs(uold(1), 1); hs%sold = fabptrs(sold(1), 2); hs%gp = fabptrs(gp(1), 3)
    hs%ext_vel_force = fabptrs(ext_vel_force(1), 4); hs%ext_scal_force = fabptrs(ext_scal_force(1), 5)
    hs%unew = fabptrs(unew(1), 6); hs%snew = fabptrs(snew(1), 7); hs%rhohalf = fabptrs(rhohalf(1), 8)
    hs%lapu = c_null_ptr; hs%mac_rhs = c_null_ptr
    if (visc_coef > 0.d0) hs%lapu = fabptrs(lapu(1), 9)
    if (norm_inf(mac_rhs(1)) > 0.d0) hs%mac_rhs = fabptrs(mac_rhs(1), 10)
    call vdn_check(c, vdn_advance_host(c, real(dt,c_double), -1.0_c_double, hs, ncyc, res))
    do d = 1, mla%dim
       call vdn_get(c, VDN_UMAC_X + int(d-1,c_int), umac(1,d))     ! diagnostics only
    end do
  end subroutine vdn_advance_path

end module vdn_path_module
