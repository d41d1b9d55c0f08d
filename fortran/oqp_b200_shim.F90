!> ISO_C_BINDING shim over libopenqp_b200.so (include/oqp_b200.h).
!> Shipped as SOURCE ONLY: this image has no Fortran compiler, so the file is untested here; the C ABI it binds is
!> exercised through tests/test_gpu_parity.py.  It is written to be dropped next to
!> source/modules/routec_bridge.F90 and used from scf_addons.F90::fock_jk and the TDHF / MRSF drivers
!> (INTEGRATION.md shows the call-site edits).  Integers crossing the ABI are c_int (32 bit) although OpenQP is built
!> with -fdefault-integer-8, exactly like the existing seam (routec_bridge.F90:33-40).
module oqp_b200_shim
  use, intrinsic :: iso_c_binding
  use precision, only: dp
  use basis_tools, only: basis_set
  implicit none
  private
  public :: oqpb_int2_t

  interface
    integer(c_int) function oqpb_ctx_create(ctx, device) bind(C, name="oqpb_ctx_create")
      import; type(c_ptr), intent(out) :: ctx; integer(c_int), value :: device
    end function
    subroutine oqpb_ctx_destroy(ctx) bind(C, name="oqpb_ctx_destroy")
      import; type(c_ptr), value :: ctx
    end subroutine
    integer(c_int) function oqpb_set_basis(ctx, nshell, nprim, am, harmonic, ncontr, g_offset, ao_offset, naos, &
                                           ex, cc, centers, harmonic_active) bind(C, name="oqpb_set_basis")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nshell, nprim, harmonic_active
      integer(c_int), intent(in) :: am(*), harmonic(*), ncontr(*), g_offset(*), ao_offset(*), naos(*)
      real(c_double), intent(in) :: ex(*), cc(*), centers(*)
    end function
    integer(c_int) function oqpb_set_cutoff(ctx, cutoff) bind(C, name="oqpb_set_cutoff")
      import; type(c_ptr), value :: ctx; real(c_double), value :: cutoff
    end function
    integer(c_int) function oqpb_set_screening(ctx, schwarz_in) bind(C, name="oqpb_set_screening")
      import; type(c_ptr), value :: ctx; type(c_ptr), value :: schwarz_in
    end function
    integer(c_int) function oqpb_set_partition(ctx, rank, nranks) bind(C, name="oqpb_set_partition")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: rank, nranks
    end function
    integer(c_int) function oqpb_fock(ctx, urohf, d, f, nfocks, se, sc, post, nskipped) bind(C, name="oqpb_fock")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: urohf, nfocks, post
      real(c_double), intent(in) :: d(*); real(c_double), intent(out) :: f(*)
      real(c_double), value :: se, sc; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_fock_cam(ctx, urohf, d, f, nfocks, alpha, beta, mu, alpha_coulomb, beta_coulomb, post, &
                                          nskipped) bind(C, name="oqpb_fock_cam")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: urohf, nfocks, post
      real(c_double), intent(in) :: d(*); real(c_double), intent(out) :: f(*)
      real(c_double), value :: alpha, beta, mu, alpha_coulomb, beta_coulomb; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_td(ctx, d2, nvec, flags, se, sc, apb, amb, nskipped) bind(C, name="oqpb_jk_td")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nvec, flags
      real(c_double), intent(in) :: d2(*); real(c_double), intent(out) :: apb(*), amb(*)
      real(c_double), value :: se, sc; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_mrsf(ctx, d3, nvec, ncomp, se, sc, f3, nskipped) bind(C, name="oqpb_jk_mrsf")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nvec, ncomp
      real(c_double), intent(in) :: d3(*); real(c_double), intent(out) :: f3(*)
      real(c_double), value :: se, sc; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_td_cam(ctx, d2, nvec, flags, alpha, beta, mu, alpha_coulomb, beta_coulomb, apb, amb, &
                                           nskipped) bind(C, name="oqpb_jk_td_cam")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nvec, flags
      real(c_double), intent(in) :: d2(*); real(c_double), intent(out) :: apb(*), amb(*)
      real(c_double), value :: alpha, beta, mu, alpha_coulomb, beta_coulomb; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_mrsf_cam(ctx, d3, nvec, ncomp, alpha, beta, mu, alpha_coulomb, f3, nskipped) &
        bind(C, name="oqpb_jk_mrsf_cam")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nvec, ncomp
      real(c_double), intent(in) :: d3(*); real(c_double), intent(out) :: f3(*)
      real(c_double), value :: alpha, beta, mu, alpha_coulomb; integer(c_long_long), intent(out) :: nskipped
    end function
    !> device-resident variant (d3_dev, f3_dev = CUDA device pointers, e.g. from a routec_sig session)
    integer(c_int) function oqpb_jk_mrsf_dev(ctx, d3_dev, nvec, ncomp, se, sc, f3_dev) bind(C, name="oqpb_jk_mrsf_dev")
      import; type(c_ptr), value :: ctx, d3_dev, f3_dev; integer(c_int), value :: nvec, ncomp
      real(c_double), value :: se, sc
    end function
  end interface

  !> Same public surface as int2_compute_t (int2.F90:137-185): init / set_screening / set_cutoff / clean, `skipped`;
  !> run() is split per consumer because the consumers' digestion runs on the device.
  type :: oqpb_int2_t
    type(c_ptr) :: ctx = c_null_ptr
    integer :: skipped = 0
    logical :: ok = .false.
  contains
    procedure :: init => shim_init
    procedure :: set_screening => shim_set_screening
    procedure :: set_cutoff => shim_set_cutoff
    procedure :: run_fock => shim_run_fock      !< int2_rhf_data_t / int2_urohf_data_t
    procedure :: run_fock_cam => shim_run_fock_cam  !< same consumers through int2_run_cam (int2.F90:538-584)
    procedure :: run_td => shim_run_td          !< int2_td_data_t
    procedure :: run_mrsf => shim_run_mrsf      !< int2_mrsf_data_t
    procedure :: run_mrsf_cam => shim_run_mrsf_cam  !< int2_mrsf_data_t through int2_run_cam (pass 2: component 7 exchange)
    procedure :: clean => shim_clean
  end type

contains

  !> int2_compute_t%init (int2.F90:245-289).  info /= 0 => caller keeps the native driver.
  subroutine shim_init(this, basis, cutoff, harmonic_active, rank, nranks, info)
    class(oqpb_int2_t), intent(inout) :: this
    type(basis_set), intent(in) :: basis
    real(dp), intent(in) :: cutoff
    logical, intent(in) :: harmonic_active
    integer, intent(in) :: rank, nranks
    integer, intent(out) :: info
    integer(c_int), allocatable :: am(:), hm(:), nc(:), g0(:), ao(:), na(:)
    real(c_double), allocatable :: cen(:)
    integer :: n, np, i
    n = basis%nshell
    np = basis%g_offset(n) + basis%ncontr(n) - 1
    info = oqpb_ctx_create(this%ctx, 0_c_int)
    if (info /= 0) return
    allocate(am(n), hm(n), nc(n), g0(n), ao(n), na(n), cen(3*n))
    am = int(basis%am(1:n), c_int); hm = int(basis%harmonic(1:n), c_int); nc = int(basis%ncontr(1:n), c_int)
    g0 = int(basis%g_offset(1:n) - 1, c_int)      ! 0-based offsets
    ao = int(basis%ao_offset(1:n) - 1, c_int)
    na = int(basis%naos(1:n), c_int)
    do i = 1, n
      cen(3*i-2:3*i) = basis%shell_centers(i, 1:3)
    end do
    info = oqpb_set_basis(this%ctx, int(n, c_int), int(np, c_int), am, hm, nc, g0, ao, na, basis%ex, basis%cc, cen, &
                          merge(1_c_int, 0_c_int, harmonic_active))
    if (info == 0) info = oqpb_set_cutoff(this%ctx, cutoff)
    if (info == 0) info = oqpb_set_partition(this%ctx, int(rank, c_int), int(nranks, c_int))
    this%ok = info == 0
  end subroutine

  subroutine shim_set_screening(this, info)
    class(oqpb_int2_t), intent(inout) :: this
    integer, intent(out) :: info
    info = oqpb_set_screening(this%ctx, c_null_ptr)
  end subroutine

  subroutine shim_set_cutoff(this, cutoff, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), intent(in) :: cutoff
    integer, intent(out) :: info
    info = oqpb_set_cutoff(this%ctx, cutoff)
  end subroutine

  !> d, f: packed (ntri, nfocks) as in fock_jk; raw accumulators unless post (scf_addons.F90:1177-1185)
  subroutine shim_run_fock(this, urohf, d, f, scale_exchange, scale_coulomb, post, info)
    class(oqpb_int2_t), intent(inout) :: this
    logical, intent(in) :: urohf, post
    real(dp), contiguous, intent(in) :: d(:,:)
    real(dp), contiguous, intent(out) :: f(:,:)
    real(dp), intent(in) :: scale_exchange, scale_coulomb
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    info = oqpb_fock(this%ctx, merge(1_c_int, 0_c_int, urohf), d, f, int(size(d, 2), c_int), scale_exchange, &
                     scale_coulomb, merge(1_c_int, 0_c_int, post), ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  !> run(consumer, cam=.true., alpha, beta, mu): regular pass + Erf-attenuated pass into the same f
  subroutine shim_run_fock_cam(this, urohf, d, f, alpha, beta, mu, post, info, alpha_coulomb, beta_coulomb)
    class(oqpb_int2_t), intent(inout) :: this
    logical, intent(in) :: urohf, post
    real(dp), contiguous, intent(in) :: d(:,:)
    real(dp), contiguous, intent(out) :: f(:,:)
    real(dp), intent(in) :: alpha, beta, mu
    integer, intent(out) :: info
    real(dp), intent(in), optional :: alpha_coulomb, beta_coulomb
    integer(c_long_long) :: ns
    real(dp) :: ac, bc
    ac = 1.0_dp; if (present(alpha_coulomb)) ac = alpha_coulomb
    bc = 0.0_dp; if (present(beta_coulomb)) bc = beta_coulomb
    info = oqpb_fock_cam(this%ctx, merge(1_c_int, 0_c_int, urohf), d, f, int(size(d, 2), c_int), alpha, beta, mu, ac, bc, &
                         merge(1_c_int, 0_c_int, post), ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  subroutine shim_run_td(this, d2, int_apb, int_amb, tamm_dancoff, tamm_dancoff_coulomb, se, sc, apb, amb, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), contiguous, intent(in) :: d2(:,:,:)
    logical, intent(in) :: int_apb, int_amb, tamm_dancoff, tamm_dancoff_coulomb
    real(dp), intent(in) :: se, sc
    real(dp), contiguous, intent(out) :: apb(:,:,:), amb(:,:,:)
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    integer(c_int) :: flags
    flags = merge(1, 0, int_apb) + merge(2, 0, int_amb) + merge(4, 0, tamm_dancoff) + merge(8, 0, tamm_dancoff_coulomb)
    info = oqpb_jk_td(this%ctx, d2, int(size(d2, 3), c_int), flags, se, sc, apb, amb, ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  subroutine shim_run_mrsf(this, d3, se, sc, f3, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), contiguous, intent(in) :: d3(:,:,:,:)
    real(dp), intent(in) :: se, sc
    real(dp), contiguous, intent(out) :: f3(:,:,:,:)
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    info = oqpb_jk_mrsf(this%ctx, d3, int(size(d3, 1), c_int), int(size(d3, 2), c_int), se, sc, f3, ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  !> run(int2_data, cam=.true., alpha, beta, mu[, alpha_coulomb]) with the MRSF consumer (tdhf_mrsf_lib.F90:279-326)
  subroutine shim_run_mrsf_cam(this, d3, alpha, beta, mu, alpha_coulomb, f3, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), contiguous, intent(in) :: d3(:,:,:,:)
    real(dp), intent(in) :: alpha, beta, mu, alpha_coulomb
    real(dp), contiguous, intent(out) :: f3(:,:,:,:)
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    info = oqpb_jk_mrsf_cam(this%ctx, d3, int(size(d3, 1), c_int), int(size(d3, 2), c_int), alpha, beta, mu, alpha_coulomb, f3, ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  subroutine shim_clean(this)
    class(oqpb_int2_t), intent(inout) :: this
    if (c_associated(this%ctx)) call oqpb_ctx_destroy(this%ctx)
    this%ctx = c_null_ptr
    this%ok = .false.
  end subroutine

end module oqp_b200_shim
