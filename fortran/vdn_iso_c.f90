! fortran/vdn_iso_c.f90 -- ISO_C_BINDING interface to libvdn.so (include/vdn.h) and the drop-in replacement
! of the device-resident segment of advance_timestep (src/advance_timestep.f90:95-124).
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build container has no Fortran compiler and no FBoxLib/AMReX
! (SURVEY.md, BASELINE.md section 2).  The C ABI it binds is exercised through ctypes by tests/ instead;
! the array layout contract is stated in include/vdn.h ("Host array convention").
!
! Usage inside VARDEN (see INTEGRATION.md): add this file to src/GPackage.mak, link with -lvdn -lcudart, and
! in advance_timestep replace the calls
!     advance_premac / macproject / scalar_advance / make_at_halftime / velocity_advance
! by   call vdn_advance_path(mla,sold,uold,snew,unew,gp,ext_vel_force,ext_scal_force,rhohalf,umac,the_bc_tower,dt,dx)
! Everything before (lapu, print_old) and after (hgproject, print_new, timers) stays the reference Fortran.

module vdn_iso_c

  use iso_c_binding
  implicit none

  ! enum vdn_field (include/vdn.h)
  integer(c_int), parameter :: VDN_UOLD = 0, VDN_SOLD = 1, VDN_UNEW = 2, VDN_SNEW = 3, VDN_GP = 4, &
       VDN_EXT_VEL_FORCE = 5, VDN_EXT_SCAL_FORCE = 6, VDN_LAPU = 7, VDN_UMAC_X = 8, VDN_UMAC_Y = 9, VDN_UMAC_Z = 10, &
       VDN_MAC_RHS = 11, VDN_RHOHALF = 12, VDN_VEL_FORCE = 13, VDN_SCAL_FORCE = 14, VDN_RH = 15, VDN_PHI = 16

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

  implicit none
  private
  public :: vdn_advance_path, vdn_path_finalize

  type(c_ptr), save :: ctx = c_null_ptr      ! one context per MPI rank; rebuilt after regrid (call vdn_path_finalize)
  type(c_ptr), allocatable, target, save :: fab_tab(:,:)   ! (local fab, multifab slot): host pointers handed to vdn_advance_host

contains

  subroutine vdn_check(rc)
    integer(c_int), intent(in) :: rc
    ! error convention of the reference: bl_error aborts the run (velocity_advance.f90:113, multifab_physbc.f90:125)
    if (rc /= 0) call bl_error('libvdn: device hot path failed (see vdn_last_error)')
  end subroutine vdn_check

  subroutine vdn_path_init(mla, sold, the_bc_tower, dx)
    use probin_module, only: nscal, slope_order, use_minion, boussinesq, stencil_order, mg_verbose, visc_coef, diff_coef, &
                             u_bc, v_bc, w_bc, rho_bc, trac_bc
    type(ml_layout), intent(in) :: mla
    type(multifab) , intent(in) :: sold(:)
    type(bc_tower) , intent(in) :: the_bc_tower
    real(dp_t)     , intent(in) :: dx(:,:)
    type(vdn_params) :: prm
    integer(c_int), allocatable :: blo(:,:), bhi(:,:)
    integer(c_int) :: dlo(3), dhi(3), pbc(2,3)
    real(c_double) :: cdx(3)
    type(box) :: pd
    integer :: i, d, dm, nb

    dm = mla%dim
    nb = nfabs(sold(1))
    allocate(blo(3,nb), bhi(3,nb)); blo = 0; bhi = 0
    do i = 1, nb
       blo(1:dm,i) = lwb(get_box(sold(1),i)); bhi(1:dm,i) = upb(get_box(sold(1),i))
    end do
    pd = ml_layout_get_pd(mla,1)
    dlo = 0; dhi = 0; dlo(1:dm) = lwb(pd); dhi(1:dm) = upb(pd)
    pbc = 0; cdx = 0.d0; cdx(1:dm) = dx(1,1:dm)
    do d = 1, dm      ! domain BCs: phys_bc_level_array(0,:,:) (define_bc_tower.f90:142-146)
       pbc(1,d) = the_bc_tower%bc_tower_array(1)%phys_bc_level_array(0,d,1)
       pbc(2,d) = the_bc_tower%bc_tower_array(1)%phys_bc_level_array(0,d,2)
    end do
    call vdn_params_default(prm)
    prm%nscal = nscal; prm%slope_order = slope_order; prm%use_minion = merge(1,0,use_minion)
    prm%boussinesq = boussinesq; prm%stencil_order = stencil_order; prm%mg_verbose = mg_verbose
    prm%visc_coef = visc_coef; prm%diff_coef = diff_coef
    do d = 1, dm
       prm%bc_val(:,d,1) = u_bc(d,:); prm%bc_val(:,d,2) = v_bc(d,:); prm%bc_val(:,d,3) = w_bc(d,:)
       prm%bc_val(:,d,4) = rho_bc(d,:); prm%bc_val(:,d,5) = trac_bc(d,:)
    end do
    ! device ordinal = local MPI rank modulo GPUs per node; 0 in the single-rank case
    call vdn_check(vdn_ctx_create(prm, int(dm,c_int), int(nb,c_int), blo, bhi, dlo, dhi, pbc, cdx, 0_c_int, ctx))
  end subroutine vdn_path_init

  subroutine vdn_path_finalize()
    if (c_associated(ctx)) call vdn_ctx_destroy(ctx)
    ctx = c_null_ptr
    if (allocated(fab_tab)) deallocate(fab_tab)
  end subroutine vdn_path_finalize

  subroutine put(field, mf)
    integer(c_int), intent(in) :: field
    type(multifab), intent(in) :: mf
    real(dp_t), pointer :: p(:,:,:,:)
    integer :: i
    do i = 1, nfabs(mf)
       p => dataptr(mf, i)
       call vdn_check(vdn_field_upload(ctx, field, int(i-1,c_int), c_loc(p(lbound(p,1),lbound(p,2),lbound(p,3),1)), &
                                       int(nghost(mf),c_int), int(ncomp(mf),c_int)))
    end do
  end subroutine put

  subroutine get(field, mf, nc)
    integer(c_int), intent(in) :: field
    type(multifab), intent(inout) :: mf
    integer, intent(in) :: nc
    real(dp_t), pointer :: p(:,:,:,:)
    integer :: i
    do i = 1, nfabs(mf)
       p => dataptr(mf, i)
       call vdn_check(vdn_field_download(ctx, field, int(i-1,c_int), c_loc(p(lbound(p,1),lbound(p,2),lbound(p,3),1)), &
                                         int(nghost(mf),c_int), int(nc,c_int)))
    end do
  end subroutine get

  ! c_ptr array of the local fabs of mf (slot = which of the eight multifabs of a step; storage lives in the module)
  function fabptrs(mf, slot) result(p)
    type(multifab), intent(in) :: mf
    integer, intent(in) :: slot
    type(c_ptr) :: p
    real(dp_t), pointer :: q(:,:,:,:)
    integer :: i
    if (.not. allocated(fab_tab)) allocate(fab_tab(nfabs(mf), 8))
    do i = 1, nfabs(mf)
       q => dataptr(mf, i)
       fab_tab(i, slot) = c_loc(q(lbound(q,1),lbound(q,2),lbound(q,3),1))
    end do
    p = c_loc(fab_tab(1, slot))
  end function fabptrs

  ! Replaces advance_timestep.f90:95-124 for nlevs == 1 and visc_coef == diff_coef == 0.
  subroutine vdn_advance_path(mla,sold,uold,snew,unew,gp,ext_vel_force,ext_scal_force,rhohalf,umac,the_bc_tower,dt,dx)
    type(ml_layout), intent(in   ) :: mla
    type(multifab) , intent(in   ) :: sold(:), uold(:), gp(:), ext_vel_force(:), ext_scal_force(:)
    type(multifab) , intent(inout) :: snew(:), unew(:), rhohalf(:), umac(:,:)
    type(bc_tower) , intent(in   ) :: the_bc_tower
    real(dp_t)     , intent(in   ) :: dt, dx(:,:)
    integer(c_int) :: ncyc
    real(c_double) :: res
    type(vdn_host_state) :: hs
    integer :: d

    if (mla%nlevel /= 1) call bl_error('vdn_advance_path: single-level only')
    if (.not. c_associated(ctx)) call vdn_path_init(mla, sold, the_bc_tower, dx)
    ! copies only at the path boundary (BASELINE.json north_star).  Sequential form:
    !   put(uold, sold, gp, ext_vel_force, ext_scal_force); vdn_advance; get(unew, snew, rhohalf)
    ! Pipelined form (default): one call that overlaps the copies with the stages.  fabptrs(mf) returns a
    ! c_ptr array with c_loc(dataptr(mf,i)) for the local fabs (kept in module storage until the call returns).
    hs%uold = fabptrs(uold(1), 1); hs%sold = fabptrs(sold(1), 2); hs%gp = fabptrs(gp(1), 3)
    hs%ext_vel_force = fabptrs(ext_vel_force(1), 4); hs%ext_scal_force = fabptrs(ext_scal_force(1), 5)
    hs%unew = fabptrs(unew(1), 6); hs%snew = fabptrs(snew(1), 7); hs%rhohalf = fabptrs(rhohalf(1), 8)
    call vdn_check(vdn_advance_host(ctx, real(dt,c_double), -1.0_c_double, hs, ncyc, res))
    do d = 1, mla%dim
       call get(VDN_UMAC_X + int(d-1,c_int), umac(1,d), 1)     ! diagnostics only
    end do
  end subroutine vdn_advance_path

end module vdn_path_module
