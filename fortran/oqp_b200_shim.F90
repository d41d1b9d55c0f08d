!> ISO_C_BINDING shim over libopenqp_b200.so (include/oqp_b200.h).
!> Shipped as SOURCE ONLY: this image has no Fortran compiler, so the file is untested here; the C ABI it binds is
!> exercised from C (tests/c/routec_driver.c) and through ctypes (tests/test_gpu_parity.py).  It is written to be dropped next to
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
  ! plain C entries a call site may need directly (INTEGRATION.md): the legacy-seam registration, the CAM screening setup,
  ! the device-pointer entries and the generic J/K engine
  public :: oqpb_set_default_ctx, oqpb_set_default_scftype, oqpb_set_screening_cam, oqpb_fock_dev, oqpb_fock_post_dev, &
            oqpb_jk_mrsf_dev, oqpb_jk, oqpb_ctx_ndevices

  interface
    integer(c_int) function oqpb_ctx_create(ctx, device) bind(C, name="oqpb_ctx_create")
      import; type(c_ptr), intent(out) :: ctx; integer(c_int), value :: device
    end function
    !> one context over ndev GPUs of the node (devices = c_null_ptr: 0 .. ndev-1), NCCL all-reduce inside the library
    integer(c_int) function oqpb_ctx_create_multi(ctx, ndev, devices) bind(C, name="oqpb_ctx_create_multi")
      import; type(c_ptr), intent(out) :: ctx; integer(c_int), value :: ndev; type(c_ptr), value :: devices
    end function
    integer(c_int) function oqpb_ctx_ndevices(ctx) bind(C, name="oqpb_ctx_ndevices")
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function oqpb_set_default_ctx(ctx) bind(C, name="oqpb_set_default_ctx")
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function oqpb_set_default_scftype(urohf) bind(C, name="oqpb_set_default_scftype")
      import; integer(c_int), value :: urohf
    end function
    integer(c_int) function oqpb_set_screening_cam(ctx, mu, schwarz_att_in) bind(C, name="oqpb_set_screening_cam")
      import; type(c_ptr), value :: ctx; real(c_double), value :: mu; type(c_ptr), value :: schwarz_att_in
    end function
    integer(c_int) function oqpb_fock_dev(ctx, urohf, d_dev, f_dev, nfocks, se, sc) bind(C, name="oqpb_fock_dev")
      import; type(c_ptr), value :: ctx, d_dev, f_dev; integer(c_int), value :: urohf, nfocks; real(c_double), value :: se, sc
    end function
    integer(c_int) function oqpb_fock_post_dev(ctx, f_dev, nfocks) bind(C, name="oqpb_fock_post_dev")
      import; type(c_ptr), value :: ctx, f_dev; integer(c_int), value :: nfocks
    end function
    integer(c_int) function oqpb_jk(ctx, n, p, want_j, want_k, j, k, nskipped) bind(C, name="oqpb_jk")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: n
      real(c_double), intent(in) :: p(*); integer(c_int), intent(in) :: want_j(*), want_k(*)
      real(c_double), intent(inout) :: j(*), k(*); integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_tdgrd(ctx, d2, flags, se, sc, apb, amb, nskipped) bind(C, name="oqpb_jk_tdgrd")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: flags
      real(c_double), intent(in) :: d2(*); real(c_double), intent(out) :: apb(*), amb(*)
      real(c_double), value :: se, sc; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_rpagrd(ctx, nspin, np, nm, nt, xpy, xmy, t, se, sc, hpp, hpt, hmm, nskipped) &
        bind(C, name="oqpb_jk_rpagrd")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nspin, np, nm, nt
      type(c_ptr), value :: xpy, xmy, t, hpp, hpt, hmm   ! c_loc of the arrays, c_null_ptr when the count is 0
      real(c_double), value :: se, sc; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_umrsf(ctx, d3, nvec, se, sc, f3, nskipped) bind(C, name="oqpb_jk_umrsf")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nvec
      real(c_double), intent(in) :: d3(*); real(c_double), intent(out) :: f3(*)
      real(c_double), value :: se, sc; integer(c_long_long), intent(out) :: nskipped
    end function
    integer(c_int) function oqpb_jk_umrsf_cam(ctx, d3, nvec, alpha, beta, mu, alpha_coulomb, f3, nskipped) &
        bind(C, name="oqpb_jk_umrsf_cam")
      import; type(c_ptr), value :: ctx; integer(c_int), value :: nvec
      real(c_double), intent(in) :: d3(*); real(c_double), intent(out) :: f3(*)
      real(c_double), value :: alpha, beta, mu, alpha_coulomb; integer(c_long_long), intent(out) :: nskipped
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

  !> Same public surface as int2_compute_t (int2.F90:137-185): init / set_screening / set_cutoff / run / clean, `skipped`.
  !> run(consumer, cam, alpha, beta, mu, ...) takes the reference's own consumer objects (int2_rhf_data_t, int2_urohf_data_t,
  !> int2_td_data_t, int2_tdgrd_data_t, int2_mrsf_data_t, int2_umrsf_data_t) and fills the arrays their callers read
  !> (f(:,:,1), apb/amb(:,:,:,1), f3(:,:,:,:,1)), so a call site only swaps the driver object; the run_* procedures are the
  !> same entries with plain arrays.  The context is kept alive across fock_jk calls: init() on an unchanged basis and
  !> geometry is a no-op, so the pair table and the Schwarz matrix are built once per geometry, not once per SCF iteration.
  type :: oqpb_int2_t
    type(c_ptr) :: ctx = c_null_ptr
    integer :: skipped = 0
    logical :: ok = .false.
    integer :: nshell_cached = -1
    real(dp) :: cutoff_cached = -1.0_dp
    real(dp), allocatable :: centers_cached(:)
  contains
    procedure :: init => shim_init
    procedure :: set_screening => shim_set_screening
    procedure :: set_cutoff => shim_set_cutoff
    procedure :: run => shim_run                !< run(consumer [, cam, alpha, beta, mu, alpha_coulomb, beta_coulomb], info)
    procedure :: run_td_cam => shim_run_td_cam  !< int2_td_data_t through int2_run_cam
    procedure :: run_tdgrd => shim_run_tdgrd    !< int2_tdgrd_data_t
    procedure :: run_umrsf => shim_run_umrsf    !< int2_umrsf_data_t
    procedure :: run_fock => shim_run_fock      !< int2_rhf_data_t / int2_urohf_data_t
    procedure :: run_fock_cam => shim_run_fock_cam  !< same consumers through int2_run_cam (int2.F90:538-584)
    procedure :: run_td => shim_run_td          !< int2_td_data_t
    procedure :: run_mrsf => shim_run_mrsf      !< int2_mrsf_data_t
    procedure :: run_mrsf_cam => shim_run_mrsf_cam  !< int2_mrsf_data_t through int2_run_cam (pass 2: component 7 exchange)
    procedure :: clean => shim_clean
  end type

contains

  !> int2_compute_t%init (int2.F90:245-289).  info /= 0 => caller keeps the native driver.
  !> ngpus (optional, default 1): > 1 creates ONE context over that many GPUs of the node (oqpb_ctx_create_multi, NCCL
  !> all-reduce inside the library) -- the route for a non-MPI OpenQP; with MPI, one rank per GPU: the device is
  !> local_rank = mod(rank, devices per node) unless `device` is given.
  !> Called again with an unchanged basis, geometry and cutoff (every SCF iteration through fock_jk) it returns at once:
  !> the context, its pair table and its Schwarz matrix are reused.
  subroutine shim_init(this, basis, cutoff, harmonic_active, rank, nranks, info, device, ngpus, ranks_per_node)
    class(oqpb_int2_t), intent(inout) :: this
    type(basis_set), intent(in) :: basis
    real(dp), intent(in) :: cutoff
    logical, intent(in) :: harmonic_active
    integer, intent(in) :: rank, nranks
    integer, intent(out) :: info
    integer, intent(in), optional :: device, ngpus, ranks_per_node
    integer(c_int), allocatable :: am(:), hm(:), nc(:), g0(:), ao(:), na(:)
    real(c_double), allocatable :: cen(:)
    integer :: n, np, i, dev, ng, rpn
    n = basis%nshell
    np = basis%g_offset(n) + basis%ncontr(n) - 1
    allocate(cen(3*n))
    do i = 1, n
      cen(3*i-2:3*i) = basis%shell_centers(i, 1:3)
    end do
    ! unchanged basis / geometry / cutoff: keep the context (and the screening data set_screening built on it)
    if (this%ok .and. c_associated(this%ctx) .and. n == this%nshell_cached .and. cutoff == this%cutoff_cached) then
      if (allocated(this%centers_cached)) then
        if (size(this%centers_cached) == 3*n) then
          if (all(this%centers_cached == cen)) then
            info = 0
            return
          end if
        end if
      end if
    end if
    call this%clean()
    ng = 1; if (present(ngpus)) ng = ngpus
    rpn = max(1, nranks); if (present(ranks_per_node)) rpn = max(1, ranks_per_node)
    dev = mod(rank, rpn); if (present(device)) dev = device
    if (ng > 1) then
      info = oqpb_ctx_create_multi(this%ctx, int(ng, c_int), c_null_ptr)
    else
      info = oqpb_ctx_create(this%ctx, int(dev, c_int))
    end if
    if (info /= 0) return
    allocate(am(n), hm(n), nc(n), g0(n), ao(n), na(n))
    am = int(basis%am(1:n), c_int); hm = int(basis%harmonic(1:n), c_int); nc = int(basis%ncontr(1:n), c_int)
    g0 = int(basis%g_offset(1:n) - 1, c_int)      ! 0-based offsets
    ao = int(basis%ao_offset(1:n) - 1, c_int)
    na = int(basis%naos(1:n), c_int)
    info = oqpb_set_basis(this%ctx, int(n, c_int), int(np, c_int), am, hm, nc, g0, ao, na, basis%ex, basis%cc, cen, &
                          merge(1_c_int, 0_c_int, harmonic_active))
    if (info == 0) info = oqpb_set_cutoff(this%ctx, cutoff)
    if (info == 0) info = oqpb_set_partition(this%ctx, int(rank, c_int), int(nranks, c_int))
    this%ok = info == 0
    if (this%ok) then
      this%nshell_cached = n
      this%cutoff_cached = cutoff
      this%centers_cached = cen
    end if
  end subroutine

  !> int2_compute_t%run(int2_data, cam, alpha, beta, mu, alpha_coulomb, beta_coulomb) (int2.F90:500-536): dispatch on the
  !> reference's own consumer type, results where the callers read them (thread slot 1).  info /= 0: not handled, the
  !> caller runs its native driver.  (Needs `use int2_compute`, `use tdhf_lib`, `use tdhf_mrsf_lib` of the host code.)
  subroutine shim_run(this, consumer, info, cam, alpha, beta, mu, alpha_coulomb, beta_coulomb)
    use int2_compute, only: int2_compute_data_t, int2_rhf_data_t, int2_urohf_data_t
    use tdhf_lib, only: int2_td_data_t, int2_tdgrd_data_t
    use tdhf_mrsf_lib, only: int2_mrsf_data_t, int2_umrsf_data_t
    class(oqpb_int2_t), intent(inout) :: this
    class(int2_compute_data_t), intent(inout) :: consumer
    integer, intent(out) :: info
    logical, intent(in), optional :: cam
    real(dp), intent(in), optional :: alpha, beta, mu, alpha_coulomb, beta_coulomb
    logical :: is_cam
    real(dp) :: al, be, m, ac, bc
    integer :: nbf, nv
    is_cam = .false.; if (present(cam)) is_cam = cam
    al = 1.0_dp; if (present(alpha)) al = alpha
    be = 0.0_dp; if (present(beta)) be = beta
    m = 0.0_dp; if (present(mu)) m = mu
    ac = 1.0_dp; if (present(alpha_coulomb)) ac = alpha_coulomb
    bc = 0.0_dp; if (present(beta_coulomb)) bc = beta_coulomb
    info = 1
    select type (consumer)
    class is (int2_urohf_data_t)
      if (allocated(consumer%f)) deallocate(consumer%f)
      allocate(consumer%f(size(consumer%d, 1), size(consumer%d, 2), 1))
      if (is_cam) then
        call this%run_fock_cam(.true., consumer%d, consumer%f(:,:,1), al, be, m, .false., info, ac, bc)
      else
        call this%run_fock(.true., consumer%d, consumer%f(:,:,1), consumer%scale_exchange, consumer%scale_coulomb, .false., info)
      end if
    class is (int2_rhf_data_t)
      if (allocated(consumer%f)) deallocate(consumer%f)
      allocate(consumer%f(size(consumer%d, 1), size(consumer%d, 2), 1))
      if (is_cam) then
        call this%run_fock_cam(.false., consumer%d, consumer%f(:,:,1), al, be, m, .false., info, ac, bc)
      else
        call this%run_fock(.false., consumer%d, consumer%f(:,:,1), consumer%scale_exchange, consumer%scale_coulomb, .false., info)
      end if
    class is (int2_tdgrd_data_t)   ! before its parent int2_td_data_t
      if (is_cam) return
      nbf = size(consumer%d2, 1)
      if (allocated(consumer%apb)) deallocate(consumer%apb, consumer%amb)
      allocate(consumer%apb(nbf, nbf, 2, 1), consumer%amb(nbf, nbf, 2, 1))
      call this%run_tdgrd(consumer%d2, consumer%int_apb, consumer%int_amb, consumer%scale_exchange, consumer%scale_coulomb, &
                          consumer%apb(:,:,:,1), consumer%amb(:,:,:,1), info)
    class is (int2_td_data_t)
      nbf = size(consumer%d2, 1); nv = size(consumer%d2, 3)
      if (allocated(consumer%apb)) deallocate(consumer%apb, consumer%amb)
      allocate(consumer%apb(nbf, nbf, nv, 1), consumer%amb(nbf, nbf, nv, 1))
      if (is_cam) then
        call this%run_td_cam(consumer%d2, consumer%int_apb, consumer%int_amb, consumer%tamm_dancoff, &
                             consumer%tamm_dancoff_coulomb, al, be, m, ac, bc, consumer%apb(:,:,:,1), consumer%amb(:,:,:,1), info)
      else
        call this%run_td(consumer%d2, consumer%int_apb, consumer%int_amb, consumer%tamm_dancoff, consumer%tamm_dancoff_coulomb, &
                         consumer%scale_exchange, consumer%scale_coulomb, consumer%apb(:,:,:,1), consumer%amb(:,:,:,1), info)
      end if
    class is (int2_umrsf_data_t)   ! before its parent int2_mrsf_data_t
      if (allocated(consumer%f3)) deallocate(consumer%f3)
      allocate(consumer%f3(size(consumer%d3, 1), size(consumer%d3, 2), size(consumer%d3, 3), size(consumer%d3, 4), 1))
      call this%run_umrsf(consumer%d3, is_cam, merge(al, consumer%scale_exchange, is_cam), be, m, &
                          merge(ac, consumer%scale_coulomb, is_cam), consumer%f3(:,:,:,:,1), info)
    class is (int2_mrsf_data_t)
      if (allocated(consumer%f3)) deallocate(consumer%f3)
      allocate(consumer%f3(size(consumer%d3, 1), size(consumer%d3, 2), size(consumer%d3, 3), size(consumer%d3, 4), 1))
      if (is_cam) then
        call this%run_mrsf_cam(consumer%d3, al, be, m, ac, consumer%f3(:,:,:,:,1), info)
      else
        call this%run_mrsf(consumer%d3, consumer%scale_exchange, consumer%scale_coulomb, consumer%f3(:,:,:,:,1), info)
      end if
    end select
  end subroutine

  !> int2_td_data_t through int2_run_cam: the same update in both passes (tdhf_lib.F90:140-224)
  subroutine shim_run_td_cam(this, d2, int_apb, int_amb, tamm_dancoff, tamm_dancoff_coulomb, alpha, beta, mu, alpha_coulomb, &
                             beta_coulomb, apb, amb, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), contiguous, intent(in) :: d2(:,:,:)
    logical, intent(in) :: int_apb, int_amb, tamm_dancoff, tamm_dancoff_coulomb
    real(dp), intent(in) :: alpha, beta, mu, alpha_coulomb, beta_coulomb
    real(dp), contiguous, intent(out) :: apb(:,:,:), amb(:,:,:)
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    integer(c_int) :: flags
    flags = merge(1, 0, int_apb) + merge(2, 0, int_amb) + merge(4, 0, tamm_dancoff) + merge(8, 0, tamm_dancoff_coulomb)
    info = oqpb_jk_td_cam(this%ctx, d2, int(size(d2, 3), c_int), flags, alpha, beta, mu, alpha_coulomb, beta_coulomb, apb, amb, ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  !> int2_tdgrd_data_t (tdhf_lib.F90:228-295): d2, apb, amb (nbf, nbf, 2)
  subroutine shim_run_tdgrd(this, d2, int_apb, int_amb, se, sc, apb, amb, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), contiguous, intent(in) :: d2(:,:,:)
    logical, intent(in) :: int_apb, int_amb
    real(dp), intent(in) :: se, sc
    real(dp), contiguous, intent(out) :: apb(:,:,:), amb(:,:,:)
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    info = oqpb_jk_tdgrd(this%ctx, d2, int(merge(1, 0, int_apb) + merge(2, 0, int_amb), c_int), se, sc, apb, amb, ns)
    if (info == 0) this%skipped = int(ns)
  end subroutine

  !> int2_umrsf_data_t (tdhf_mrsf_lib.F90:337-426): d3, f3 (nvec, 11, nbf, nbf); cam: pass 2 = component 11 with beta
  subroutine shim_run_umrsf(this, d3, cam, se, beta, mu, sc, f3, info)
    class(oqpb_int2_t), intent(inout) :: this
    real(dp), contiguous, intent(in) :: d3(:,:,:,:)
    logical, intent(in) :: cam
    real(dp), intent(in) :: se, beta, mu, sc
    real(dp), contiguous, intent(out) :: f3(:,:,:,:)
    integer, intent(out) :: info
    integer(c_long_long) :: ns
    if (cam) then
      info = oqpb_jk_umrsf_cam(this%ctx, d3, int(size(d3, 1), c_int), se, beta, mu, sc, f3, ns)
    else
      info = oqpb_jk_umrsf(this%ctx, d3, int(size(d3, 1), c_int), se, sc, f3, ns)
    end if
    if (info == 0) this%skipped = int(ns)
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
    this%nshell_cached = -1
    if (allocated(this%centers_cached)) deallocate(this%centers_cached)
  end subroutine

end module oqp_b200_shim
