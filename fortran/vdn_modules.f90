! fortran/vdn_modules.f90 -- SAME-NAMED drop-in replacements of the reference's hot-path modules (SURVEY 8(b)):
!
!     velpred_module        :: velpred        (replaces src/velpred.f90:16        argument list unchanged)
!     mkflux_module         :: mkflux         (replaces src/mkflux.f90:16-17)
!     update_module         :: update         (replaces src/update.f90:16-17)
!     macproject_module     :: macproject     (replaces src/macproject.f90:20)
!     mac_multigrid_module  :: mac_multigrid  (replaces src/mac_multigrid.f90:19-20)
!     estdt_module          :: estdt          (replaces src/estdt.f90:15; SURVEY 8(f) row 3)
!     viscous_module        :: visc_solve, diff_scalar_solve   (replaces src/viscsolve.f90:19,310; SURVEY 8(f) row 2; single MPI rank)
!
! A maintainer removes velpred.f90, mkflux.f90, update.f90, macproject.f90, mac_multigrid.f90 from src/GPackage.mak and adds
! vdn_iso_c.f90 + this file; every caller (advance_premac.f90:51, scalar_advance.f90:102,118, velocity_advance.f90:76,92,
! advance_timestep.f90:100) compiles unchanged.  Each procedure gathers c_loc(dataptr(mf,i)) per local fab, copies its INPUT
! multifabs to the device mirror, runs the stage through the C ABI (include/vdn.h) and copies its OUTPUT multifabs back, so the
! host multifabs stay authoritative exactly as in the reference.  This is the literal drop-in (more PCIe traffic: every stage pays
! its own copies); the fused vdn_advance_path of vdn_iso_c.f90 keeps the fields resident across the five calls and is the
! performance path.  Single level only (nlevs == 1): the multi-level hooks (create_umac_grown, ml_edge_restriction) stay the
! reference's and are out of scope (SURVEY 8(f) row 4).
!
! NOT COMPILED IN THIS REPOSITORY'S CI (no Fortran compiler, no FBoxLib in the container); the identical call sequence is driven
! through the C ABI by tests/mock_driver.c with Fortran-layout host arrays.

module velpred_module

  use bl_types
  use multifab_module
  use define_bc_module
  use ml_layout_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_for, vdn_put, vdn_get, vdn_check

  implicit none
  private
  public :: velpred

contains

  subroutine velpred(nlevs,u,umac,force,dx,dt,the_bc_level,mla)

    integer        , intent(in   ) :: nlevs
    type(multifab) , intent(in   ) :: u(:)
    type(multifab) , intent(inout) :: umac(:,:)
    type(multifab) , intent(in   ) :: force(:)
    real(kind=dp_t), intent(in   ) :: dx(:,:),dt
    type(bc_level) , intent(in   ) :: the_bc_level(:)
    type(ml_layout), intent(in   ) :: mla

    type(c_ptr) :: ctx
    integer :: d
    type(bl_prof_timer), save :: bpt

    call build(bpt,"velpred")
    if (nlevs /= 1) call bl_error('velpred (libvdn): single-level only')
    ctx = vdn_ctx_for(mla, u(1), dx)
    call vdn_put(ctx, VDN_UOLD, u(1))
    call vdn_put(ctx, VDN_VEL_FORCE, force(1))
    call vdn_check(ctx, vdn_velpred(ctx, real(dt,c_double)))          ! includes multifab_fill_boundary(umac), velpred.f90:107-112
    do d = 1, mla%dim
       call vdn_get(ctx, VDN_UMAC_X + int(d-1,c_int), umac(1,d))
    end do
    call destroy(bpt)

  end subroutine velpred

end module velpred_module


module mkflux_module

  use bl_types
  use multifab_module
  use ml_layout_module
  use define_bc_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_for, vdn_put, vdn_get, vdn_check

  implicit none
  private
  public :: mkflux

contains

  subroutine mkflux(mla,sold,sedge,flux,umac,force,mac_rhs,dx,dt,the_bc_level, &
                    is_vel,is_conservative)

    type(ml_layout), intent(in   ) :: mla
    type(multifab) , intent(in   ) :: sold(:)
    type(multifab) , intent(inout) :: sedge(:,:)
    type(multifab) , intent(inout) :: flux(:,:)
    type(multifab) , intent(in   ) :: umac(:,:)
    type(multifab) , intent(in   ) :: force(:)
    type(multifab) , intent(in   ) :: mac_rhs(:)
    real(kind=dp_t), intent(in   ) :: dx(:,:),dt
    type(bc_level) , intent(in   ) :: the_bc_level(:)
    logical        , intent(in   ) :: is_vel,is_conservative(:)

    type(c_ptr) :: ctx
    integer :: d, c
    type(bl_prof_timer), save :: bpt

    call build(bpt,"mkflux")
    if (mla%nlevel /= 1) call bl_error('mkflux (libvdn): single-level only')
    ! the device path implements the two call sites of the reference: velocity (no conservative component,
    ! velocity_advance.f90:51) and scalars (density conservative, tracers not, scalar_advance.f90:54-57)
    if (is_vel) then
       if (any(is_conservative)) call bl_error('mkflux (libvdn): conservative velocity components are not supported')
    else
       if (.not. is_conservative(1)) call bl_error('mkflux (libvdn): the density must be conservative')
       do c = 2, size(is_conservative)
          if (is_conservative(c)) call bl_error('mkflux (libvdn): only the density may be conservative')
       end do
    end if
    ctx = vdn_ctx_for(mla, sold(1), dx)
    if (is_vel) then
       call vdn_put(ctx, VDN_UOLD, sold(1)); call vdn_put(ctx, VDN_VEL_FORCE, force(1))
    else
       call vdn_put(ctx, VDN_SOLD, sold(1)); call vdn_put(ctx, VDN_SCAL_FORCE, force(1))
    end if
    call vdn_put(ctx, VDN_MAC_RHS, mac_rhs(1))
    do d = 1, mla%dim
       call vdn_put(ctx, VDN_UMAC_X + int(d-1,c_int), umac(1,d))
    end do
    call vdn_check(ctx, vdn_mkflux(ctx, merge(1_c_int, 0_c_int, is_vel), real(dt,c_double)))
    do d = 1, mla%dim
       if (is_vel) then
          call vdn_get(ctx, VDN_UEDGE_X + int(d-1,c_int), sedge(1,d))
       else
          call vdn_get(ctx, VDN_SEDGE_X + int(d-1,c_int), sedge(1,d))
          call vdn_get(ctx, VDN_SFLUX_X + int(d-1,c_int), flux(1,d), 1)   ! component 1 only: the others stay 0 (SURVEY Q14)
       end if
    end do
    call destroy(bpt)

  end subroutine mkflux

end module mkflux_module


module update_module

  use bl_types
  use multifab_module
  use define_bc_module
  use ml_layout_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_for, vdn_put, vdn_get, vdn_check

  implicit none
  private
  public :: update

contains

  subroutine update(mla,sold,umac,sedge,flux,force,snew,dx,dt,is_vel,is_cons, &
                    the_bc_level)

    type(ml_layout)   , intent(in   ) :: mla
    type(multifab)    , intent(in   ) :: sold(:)
    type(multifab)    , intent(in   ) :: umac(:,:)
    type(multifab)    , intent(in   ) :: sedge(:,:)
    type(multifab)    , intent(in   ) :: flux(:,:)
    type(multifab)    , intent(in   ) :: force(:)
    type(multifab)    , intent(inout) :: snew(:)
    real(kind = dp_t) , intent(in   ) :: dx(:,:),dt
    logical           , intent(in   ) :: is_vel,is_cons(:)
    type(bc_level)    , intent(in   ) :: the_bc_level(:)

    type(c_ptr) :: ctx
    integer :: d
    type(bl_prof_timer), save :: bpt

    call build(bpt,"update")
    if (mla%nlevel /= 1) call bl_error('update (libvdn): single-level only')
    ctx = vdn_ctx_for(mla, sold(1), dx)
    if (is_vel) then
       call vdn_put(ctx, VDN_UOLD, sold(1)); call vdn_put(ctx, VDN_VEL_FORCE, force(1))
    else
       call vdn_put(ctx, VDN_SOLD, sold(1)); call vdn_put(ctx, VDN_SCAL_FORCE, force(1))
    end if
    do d = 1, mla%dim
       call vdn_put(ctx, VDN_UMAC_X + int(d-1,c_int), umac(1,d))
       if (is_vel) then
          call vdn_put(ctx, VDN_UEDGE_X + int(d-1,c_int), sedge(1,d))
       else
          call vdn_put(ctx, VDN_SEDGE_X + int(d-1,c_int), sedge(1,d))
          call vdn_put(ctx, VDN_SFLUX_X + int(d-1,c_int), flux(1,d), 1)
       end if
    end do
    ! update_2d/3d + ml_restrict_and_fill (update.f90:103-107): ghost cells of snew included
    call vdn_check(ctx, vdn_update(ctx, merge(1_c_int, 0_c_int, is_vel), real(dt,c_double)))
    if (is_vel) then
       call vdn_get(ctx, VDN_UNEW, snew(1))
    else
       call vdn_get(ctx, VDN_SNEW, snew(1))
    end if
    call destroy(bpt)

  end subroutine update

end module update_module


module mac_multigrid_module

  use bl_types
  use ml_layout_module
  use define_bc_module
  use multifab_module
  use bndry_reg_module
  use bl_constants_module
  use bl_prof_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_for, vdn_put, vdn_get, vdn_check

  implicit none
  private
  public :: mac_multigrid

contains

  subroutine mac_multigrid(mla,rh,phi,fine_flx,alpha,beta,dx,the_bc_tower,bc_comp,&
                           stencil_order,rel_solver_eps,abs_solver_eps)

    type(ml_layout), intent(in   ) :: mla
    type(multifab) , intent(inout) :: rh(:),phi(:)
    type(bndry_reg), intent(inout) :: fine_flx(:)
    type(multifab) , intent(in   ) :: alpha(:), beta(:,:)
    real(dp_t)     , intent(in   ) :: dx(:,:)
    type(bc_tower) , intent(in   ) :: the_bc_tower
    integer        , intent(in   ) :: bc_comp
    integer        , intent(in   ) :: stencil_order
    real(dp_t)     , intent(in   ) :: rel_solver_eps
    real(dp_t)     , intent(in   ) :: abs_solver_eps

    type(c_ptr) :: ctx
    integer(c_int) :: ncyc, rc
    real(c_double) :: res
    integer :: d, i, dm
    type(bl_prof_timer), save :: bpt

    call build(bpt, "mac_multigrid")
    if (mla%nlevel /= 1) call bl_error('mac_multigrid (libvdn): single-level only')
    if (stencil_order /= 2) call bl_error('mac_multigrid (libvdn): stencil_order = 2 only')
    if (norm_inf(alpha(1)) /= ZERO) call bl_error('mac_multigrid (libvdn): alpha must be zero (the MAC projection; visc_solve is SURVEY 8(f) row 2)')
    dm = mla%dim
    ctx = vdn_ctx_for(mla, phi(1), dx)
    call vdn_put(ctx, VDN_RH, rh(1))
    call vdn_put(ctx, VDN_PHI, phi(1))                                 ! initial guess (macproject.f90:57 sets it to zero)
    do d = 1, dm
       call vdn_put(ctx, VDN_BETA_X + int(d-1,c_int), beta(1,d))
    end do
    rc = vdn_mac_solve(ctx, real(rel_solver_eps,c_double), real(abs_solver_eps,c_double), ncyc, res)
    if (rc /= 0 .and. rc /= 2) call vdn_check(ctx, rc)                  ! 2 = V-cycle cap reached: F_MG warns and carries on
    call vdn_check(ctx, vdn_fill_boundary(ctx, VDN_PHI))               ! ghost cells of phi for the box-boundary fluxes below
    call vdn_get(ctx, VDN_PHI, phi(1))
    ! fine_flx: what ml_cc_solve returns on every box's outer faces and mkumac consumes (macproject.f90:608-609,621-622,634-635):
    !   umac(lo) -= lo_flx*dx ,  umac(hi+1) += hi_flx*dx     <=>   lo_flx = beta*grad(phi)/dx ,  hi_flx = -beta*grad(phi)/dx
    ! with grad(phi) the stencil gradient of the face: neighbour difference, 0 on a Neumann face, the stencil_order = 2
    ! one-sided formula on a Dirichlet face (SURVEY Q7).
    do i = 1, nfabs(phi(1))
       call fill_fine_flx(i)
    end do
    call destroy(bpt)

  contains

    subroutine fill_fine_flx(i)
      integer, intent(in) :: i
      real(dp_t), pointer :: pp(:,:,:,:), bp(:,:,:,:), fp(:,:,:,:)
      integer :: lo(3), hi(3), d, side, ii, jj, kk, i1, j1, k1, e(3), bc
      real(dp_t) :: g, h
      lo = 1; hi = 1
      lo(1:dm) = lwb(get_box(phi(1),i)); hi(1:dm) = upb(get_box(phi(1),i))
      pp => dataptr(phi(1), i)
      do d = 1, dm
         bp => dataptr(beta(1,d), i)
         h = dx(1,d)
         e = 0; e(d) = 1
         do side = 0, 1
            fp => dataptr(fine_flx(1)%bmf(d,side), i)
            bc = the_bc_tower%bc_tower_array(1)%ell_bc_level_array(i,d,side+1,bc_comp)
            do kk = lo(3), merge(lo(3), hi(3), d == 3)
               do jj = lo(2), merge(lo(2), hi(2), d == 2)
                  do ii = lo(1), merge(lo(1), hi(1), d == 1)
                     ! face cell: first cell inside the box next to the face
                     i1 = merge(merge(lo(1), hi(1), side == 0), ii, d == 1)
                     j1 = merge(merge(lo(2), hi(2), side == 0), jj, d == 2)
                     k1 = merge(merge(lo(3), hi(3), side == 0), kk, d == 3)
                     if (bc == BC_NEU) then
                        g = ZERO
                     else if (bc == BC_DIR .and. side == 0) then
                        g =  (3.d0*pp(i1,j1,k1,1) - pp(i1+e(1),j1+e(2),k1+e(3),1)/3.d0) / h
                     else if (bc == BC_DIR) then
                        g = -(3.d0*pp(i1,j1,k1,1) - pp(i1-e(1),j1-e(2),k1-e(3),1)/3.d0) / h
                     else if (side == 0) then
                        g = (pp(i1,j1,k1,1) - pp(i1-e(1),j1-e(2),k1-e(3),1)) / h
                     else
                        g = (pp(i1+e(1),j1+e(2),k1+e(3),1) - pp(i1,j1,k1,1)) / h
                     end if
                     if (side == 0) then
                        fp(lbound(fp,1)+merge(0,ii-lo(1),d==1), lbound(fp,2)+merge(0,jj-lo(2),d==2), lbound(fp,3)+merge(0,kk-lo(3),d==3), 1) = &
                             bp(i1,j1,k1,1) * g / h
                     else
                        fp(lbound(fp,1)+merge(0,ii-lo(1),d==1), lbound(fp,2)+merge(0,jj-lo(2),d==2), lbound(fp,3)+merge(0,kk-lo(3),d==3), 1) = &
                             -bp(i1+e(1),j1+e(2),k1+e(3),1) * g / h
                     end if
                  end do
               end do
            end do
         end do
      end do
    end subroutine fill_fine_flx

  end subroutine mac_multigrid

end module mac_multigrid_module


module macproject_module

  use bl_types
  use bl_constants_module
  use multifab_module
  use define_bc_module
  use ml_layout_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_for, vdn_put, vdn_get, vdn_check

  implicit none
  private
  public :: macproject

contains

  subroutine macproject(mla,umac,rho,mac_rhs,dx,the_bc_tower,bc_comp)

    use probin_module, only : stencil_order, use_hypre

    type(ml_layout), intent(in   ) :: mla
    type(multifab ), intent(inout) :: umac(:,:)
    type(multifab ), intent(inout) :: rho(:)
    type(multifab ), intent(inout) :: mac_rhs(:)
    real(dp_t)     , intent(in   ) :: dx(:,:)
    type(bc_tower ), intent(in   ) :: the_bc_tower
    integer        , intent(in   ) :: bc_comp

    type(c_ptr) :: ctx
    integer(c_int) :: ncyc, rc
    real(c_double) :: res
    integer :: d

    if (mla%nlevel /= 1) call bl_error('macproject (libvdn): single-level only')
    if (use_hypre == 1) call bl_error('macproject (libvdn): use_hypre = 1 selects the HYPRE path, which stays the reference''s')
    ctx = vdn_ctx_for(mla, rho(1), dx)
    ! rho is sold at both call sites (advance_timestep.f90:100): component 1 = density, ghost cells filled (varden.f90:293,298)
    call vdn_put(ctx, VDN_SOLD, rho(1))
    call vdn_put(ctx, VDN_MAC_RHS, mac_rhs(1))
    do d = 1, mla%dim
       call vdn_put(ctx, VDN_UMAC_X + int(d-1,c_int), umac(1,d))
    end do
    ! divumac -> mk_mac_coeffs -> mac_multigrid (rel 1.d-10, abs -1: macproject.f90:91-93) -> mkumac -> fill_boundary(umac)
    rc = vdn_macproject(ctx, 1.0e-10_c_double, -1.0_c_double, ncyc, res)
    if (rc /= 0 .and. rc /= 2) call vdn_check(ctx, rc)
    do d = 1, mla%dim
       call vdn_get(ctx, VDN_UMAC_X + int(d-1,c_int), umac(1,d))
    end do

  end subroutine macproject

end module macproject_module


module estdt_module

  use bl_types
  use multifab_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_current, vdn_put, vdn_check

  implicit none
  private
  public :: estdt

contains

  ! estdt.f90:15 -- argument list unchanged.  The six maxima are reduced on the device over the rank's boxes and all-reduced over the ranks
  ! inside the library (the reference's parallel_reduce(dt, dt_proc, MPI_MIN), estdt.f90:68).  u, s, gp and ext_vel_force are what the driver
  ! hands to advance_timestep a few lines later (varden.f90:302-316), so the copies made here are the ones the path needs anyway: a driver
  ! that uses the fused vdn_advance_path can skip its own upload of these four fields for this step.
  subroutine estdt (lev, u, s, gp, ext_vel_force, dx, dtold, dt)

    use probin_module, only: max_dt_growth, cflfac, verbose

    type(multifab) , intent( in) :: u,s,gp,ext_vel_force
    real(kind=dp_t), intent( in) :: dx(:)
    real(kind=dp_t), intent( in) :: dtold
    real(kind=dp_t), intent(out) :: dt
    integer        , intent( in) :: lev

    type(c_ptr) :: ctx
    real(c_double) :: dt_c
    type(bl_prof_timer), save :: bpt

    call build(bpt,"estdt")
    if (lev /= 1) call bl_error('estdt (libvdn): single-level only')
    ctx = vdn_ctx_current()
    call vdn_put(ctx, VDN_UOLD, u)
    call vdn_put(ctx, VDN_SOLD, s)
    call vdn_put(ctx, VDN_GP, gp)
    call vdn_put(ctx, VDN_EXT_VEL_FORCE, ext_vel_force)
    call vdn_check(ctx, vdn_estdt(ctx, real(dtold,c_double), real(cflfac,c_double), real(max_dt_growth,c_double), dt_c))
    dt = dt_c
    if (parallel_IOProcessor() .and. verbose .ge. 1) write(6,1000) lev,dt
1000 format("Computing dt at level ",i2," to be ... ",e15.8)
    call destroy(bpt)

  end subroutine estdt

end module estdt_module


module viscous_module

  use bl_types
  use multifab_module
  use ml_layout_module
  use define_bc_module
  use vdn_iso_c
  use vdn_path_module, only : vdn_ctx_for, vdn_put, vdn_get, vdn_check

  implicit none
  private
  public :: visc_solve, diff_scalar_solve

contains

  ! viscsolve.f90:19 -- argument list unchanged.  One Helmholtz solve per velocity component on the device (alpha = rho, beta = mu,
  ! boundary types of the component, Dirichlet data from unew's ghost cells, rel. tolerance 1.d-12), then unew's ghost cells refilled.
  subroutine visc_solve(mla,unew,lapu,rho,mac_rhs,dx,mu,the_bc_tower)

    use probin_module, only : diffusion_type

    type(ml_layout), intent(in   ) :: mla
    type(multifab ), intent(inout) :: unew(:)
    type(multifab ), intent(in   ) :: lapu(:)
    type(multifab ), intent(in   ) :: rho(:)
    real(dp_t)     , intent(in   ) :: dx(:,:),mu
    type(bc_tower ), intent(in   ) :: the_bc_tower
    type(multifab ), intent(in   ) :: mac_rhs(:)

    type(c_ptr) :: ctx
    integer(c_int) :: ncyc, rc
    real(c_double) :: res
    type(bl_prof_timer), save :: bpt

    call build(bpt,"visc_solve")
    if (mla%nlevel /= 1) call bl_error('visc_solve (libvdn): single-level only')
    if (parallel_nprocs() /= 1) call bl_error('visc_solve (libvdn): single MPI rank only in this version')
    ctx = vdn_ctx_for(mla, unew(1), dx)
    call vdn_put(ctx, VDN_UNEW, unew(1))
    call vdn_put(ctx, VDN_RHOHALF, rho(1), 1)             ! rhohalf: component 1 (velocity_advance.f90:117)
    call vdn_put(ctx, VDN_MAC_RHS, mac_rhs(1))
    if (diffusion_type == 1) call vdn_put(ctx, VDN_LAPU, lapu(1))
    rc = vdn_visc_solve(ctx, real(mu,c_double), int(diffusion_type,c_int), ncyc, res)
    call vdn_check(ctx, rc)
    call vdn_get(ctx, VDN_UNEW, unew(1))
    call destroy(bpt)

  end subroutine visc_solve

  ! viscsolve.f90:310 -- argument list unchanged; backward Euler (diffusion_type = 2) only in this version
  subroutine diff_scalar_solve(mla,snew,laps,dx,mu,the_bc_tower,icomp,bc_comp)

    use probin_module, only : diffusion_type

    type(ml_layout), intent(in   ) :: mla
    type(multifab ), intent(inout) :: snew(:)
    type(multifab ), intent(in   ) :: laps(:)
    real(dp_t)     , intent(in   ) :: dx(:,:)
    real(dp_t)     , intent(in   ) :: mu
    type(bc_tower ), intent(in   ) :: the_bc_tower
    integer        , intent(in   ) :: icomp,bc_comp

    type(c_ptr) :: ctx
    integer(c_int) :: ncyc, rc
    real(c_double) :: res
    type(bl_prof_timer), save :: bpt

    call build(bpt,"diff_scalar_solve")
    if (mla%nlevel /= 1) call bl_error('diff_scalar_solve (libvdn): single-level only')
    if (parallel_nprocs() /= 1) call bl_error('diff_scalar_solve (libvdn): single MPI rank only in this version')
    if (diffusion_type /= 2) call bl_error('diff_scalar_solve (libvdn): diffusion_type = 2 only in this version')
    ctx = vdn_ctx_for(mla, snew(1), dx)
    call vdn_put(ctx, VDN_SNEW, snew(1))
    rc = vdn_diff_scalar_solve(ctx, real(mu,c_double), int(icomp-1,c_int), int(diffusion_type,c_int), ncyc, res)
    call vdn_check(ctx, rc)
    call vdn_get(ctx, VDN_SNEW, snew(1))
    call destroy(bpt)

  end subroutine diff_scalar_solve

end module viscous_module
