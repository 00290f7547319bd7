!===============================================================================
! od_b200.f90 -- ISO_C_BINDING layer between OptaDOS (Fortran) and liboptados_b200.so
!
! Drop this module into optados/src/, add od_b200.o to OBJS in src/Makefile before dos_utils.o, link with
! -loptados_b200, and replace the BODIES of the private work-horses by the calls shown in INTEGRATION.md:
!   calculate_dos                  dos_utils.f90:1140    -> b200_calculate_dos
!   calculate_dos_at_e             dos_utils.f90:1651    -> b200_calculate_dos_at_e
!   dos_utils_merge / _merge_at_e  dos_utils.f90:1013    -> (merged inside the call; or b200_allreduce)
!   calculate_jdos                 jdos_utils.f90:280    -> b200_calculate_jdos
!   jdos_utils_merge               jdos_utils.f90:407    -> (merged inside the call)
!   make_weights                   optics.f90:157        -> b200_make_weights
!   core_prepare_matrix_elements   core.f90:84           -> b200_core_weights
!   projection_merge               projection_utils.f90:64 -> b200_projection_merge
! Every public entry point (dos_utils_calculate, jdos_utils_calculate, ...) and the file contract stay as
! they are.  This file has NOT been compiled in the build image (it has no Fortran compiler); it is written
! against include/optados_b200.h, whose symbols tests/test_abi.py checks in the built library.
!===============================================================================
module od_b200
  use, intrinsic :: iso_c_binding
  use od_constants, only: dp
  implicit none
  private

  type(c_ptr), save, public :: b200_ctx = c_null_ptr

  integer(c_int), parameter, public :: OD_MEM_HOST = 0, OD_MEM_DEVICE = 1, OD_MEM_HOST_DEFERRED = 2
  integer(c_int), parameter, public :: OD_GEOM_POLYCRYS = 0, OD_GEOM_POLAR = 1, OD_GEOM_UNPOLAR = 2, OD_GEOM_TENSOR = 3
  integer(c_int), parameter, public :: OD_CORE_ABSORPTION = 0, OD_CORE_EMISSION = 1, OD_CORE_ALL = 2
  integer(c_int), parameter, public :: OD_EFERMI_OPTADOS = 0, OD_EFERMI_FILE = 1, OD_EFERMI_USER = 2, OD_EFERMI_INSULATOR = 3

  type, bind(c), public :: od_broadening
    real(c_double) :: adaptive_smearing
    real(c_double) :: fixed_smearing
    real(c_double) :: linear_smearing
    integer(c_int) :: finite_bin_correction
    integer(c_int) :: numerical_intdos
    integer(c_int) :: hybrid_linear
    real(c_double) :: hybrid_linear_grad_tol
  end type od_broadening

  ! ---- the parameter / result structures of the task drivers (include/optados_b200.h, same member order)
  type, bind(c), public :: od_dos_params
    integer(c_int) :: fixed, adaptive, linear
    type(od_broadening) :: br
    real(c_double) :: dos_spacing
    integer(c_int) :: dos_nbins
    integer(c_int) :: have_dos_min_energy, have_dos_max_energy
    real(c_double) :: dos_min_energy, dos_max_energy
    integer(c_int) :: efermi_choice
    real(c_double) :: efermi_user, efermi_castep
    integer(c_int) :: nspins_electrons
    real(c_double) :: num_electrons(2)
    integer(c_int) :: dos_per_volume
    real(c_double) :: cell_volume
  end type od_dos_params
  type, bind(c), public :: od_dos_result
    integer(c_int) :: nbins
    real(c_double) :: emin, delta_bins
    real(c_double) :: efermi_fixed, efermi_adaptive, efermi_linear
    integer(c_int) :: have_fixed, have_adaptive, have_linear, have_weighted
    integer(c_int) :: weighted_norb
  end type od_dos_result
  type, bind(c), public :: od_jdos_params
    integer(c_int) :: fixed, adaptive, linear
    type(od_broadening) :: br
    real(c_double) :: jdos_spacing, jdos_max_energy, scissor_op
    integer(c_int) :: num_exclude_bands
    type(c_ptr) :: exclude_bands          ! c_loc of an integer(c_int) array, or c_null_ptr
    real(c_double) :: efermi
    integer(c_int) :: dos_per_volume
    real(c_double) :: cell_volume
  end type od_jdos_params
  type, bind(c), public :: od_jdos_result
    integer(c_int) :: nbins
    real(c_double) :: delta_bins
    integer(c_int) :: have_fixed, have_adaptive, have_linear, have_weighted, n_geom
  end type od_jdos_result
  type, bind(c), public :: od_optics_params
    type(od_jdos_params) :: jdos
    integer(c_int) :: geom
    real(c_double) :: qdir(3)
    real(c_double) :: lossfn_gaussian
    integer(c_int) :: intraband
    real(c_double) :: drude_broadening
    real(c_double) :: dos_delta_bins
  end type od_optics_params
  type, bind(c), public :: od_optics_result
    integer(c_int) :: nbins, n_geom
    real(c_double) :: delta_bins
    integer(c_int) :: have_derived, have_loss_fn_broadened
    real(c_double) :: n_eff, n_eff2, n_eff3
    integer(c_int) :: n_parts, n_loss_fn
  end type od_optics_result
  type, bind(c), public :: od_core_params
    type(od_dos_params) :: dos
    integer(c_int) :: core_type, polar
    real(c_double) :: qdir(3)
    integer(c_int) :: lai_broadening
    real(c_double) :: lai_gaussian_width, lai_lorentzian_width, lai_lorentzian_scale, lai_lorentzian_offset
    integer(c_int) :: set_efermi_zero
    real(c_double) :: core_chemical_shift
  end type od_core_params
  type, bind(c), public :: od_core_result
    integer(c_int) :: nbins, norb
    real(c_double) :: emin, delta_bins, efermi
    integer(c_int) :: have_broadened
  end type od_core_result
  type, bind(c), public :: od_bandgap_result
    real(c_double) :: vbm_energy, cbm_energy, thermal_bandgap
    integer(c_long_long) :: thermal_vbm_k, thermal_cbm_k
    integer(c_int) :: vbm_multiplicity(2), cbm_multiplicity(2), optical_multiplicity(2)
    real(c_double) :: optical_bandgap(2), average_bandgap(2)
    real(c_double) :: weighted_average
    integer(c_int) :: thermal_multiplicity
  end type od_bandgap_result

  interface
    ! ---- the task drivers: one call = one of the reference's public L3 entry points
    subroutine od_dos_params_defaults(p) bind(c, name='od_dos_params_defaults')
      import :: od_dos_params
      type(od_dos_params), intent(out) :: p
    end subroutine
    integer(c_int) function od_dos_utils_calculate(ctx, p, mw, mw_norb, mw_nb, mw_nk, mw_mem, res) bind(c, name='od_dos_utils_calculate')
      import :: c_int, c_ptr, od_dos_params, od_dos_result
      type(c_ptr), value :: ctx
      type(od_dos_params), intent(inout) :: p
      type(c_ptr), value :: mw             ! c_loc(matrix_weights) or c_null_ptr
      integer(c_int), value :: mw_norb, mw_nb, mw_nk, mw_mem
      type(od_dos_result), intent(out) :: res
    end function
    integer(c_int) function od_dos_get(ctx, dos_type, dos, intdos) bind(c, name='od_dos_get')
      import :: c_int, c_ptr, c_double, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char), value :: dos_type
      real(c_double), intent(out) :: dos(*), intdos(*)
    end function
    integer(c_int) function od_dos_get_weighted(ctx, weighted_dos) bind(c, name='od_dos_get_weighted')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: weighted_dos(*)
    end function
    integer(c_int) function od_dos_utils_set_efermi(ctx, p, efermi) bind(c, name='od_dos_utils_set_efermi')
      import :: c_int, c_ptr, c_double, od_dos_params
      type(c_ptr), value :: ctx
      type(od_dos_params), intent(inout) :: p
      real(c_double), intent(out) :: efermi
    end function
    integer(c_int) function od_dos_utils_calculate_at_e(ctx, p, energy, delta_bins, mw, mw_norb, mw_nb, mw_nk, mw_mem, dos_at_e, &
                                                        weighted_dos_at_e) bind(c, name='od_dos_utils_calculate_at_e')
      import :: c_int, c_ptr, c_double, od_dos_params
      type(c_ptr), value :: ctx
      type(od_dos_params), intent(in) :: p
      real(c_double), value :: energy, delta_bins
      type(c_ptr), value :: mw, weighted_dos_at_e
      integer(c_int), value :: mw_norb, mw_nb, mw_nk, mw_mem
      real(c_double), intent(out) :: dos_at_e(3, *)
    end function
    integer(c_int) function od_dos_utils_compute_bandgap(ctx, efermi, num_electrons, nk_global, k_first, res) &
      bind(c, name='od_dos_utils_compute_bandgap')
      import :: c_int, c_ptr, c_double, c_long_long, od_bandgap_result
      type(c_ptr), value :: ctx
      real(c_double), value :: efermi
      real(c_double), intent(in) :: num_electrons(*)
      integer(c_long_long), value :: nk_global, k_first
      type(od_bandgap_result), intent(out) :: res
    end function
    integer(c_int) function od_dos_utils_compute_band_energies(ctx, efermi, band_energy_dos, band_energy_castep) &
      bind(c, name='od_dos_utils_compute_band_energies')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: efermi
      real(c_double), intent(out) :: band_energy_dos(3), band_energy_castep
    end function
    subroutine od_jdos_params_defaults(p) bind(c, name='od_jdos_params_defaults')
      import :: od_jdos_params
      type(od_jdos_params), intent(out) :: p
    end subroutine
    integer(c_int) function od_jdos_utils_calculate(ctx, p, mw, n_geom, mw_mem, optical_mat, ome_mem, geom, qdir, symops, num_symm, res) &
      bind(c, name='od_jdos_utils_calculate')
      import :: c_int, c_ptr, c_double, od_jdos_params, od_jdos_result
      type(c_ptr), value :: ctx
      type(od_jdos_params), intent(inout) :: p
      type(c_ptr), value :: mw, optical_mat, symops   ! exactly one of mw / optical_mat may be non-null
      integer(c_int), value :: n_geom, mw_mem, ome_mem, geom, num_symm
      real(c_double), intent(in) :: qdir(3)
      type(od_jdos_result), intent(out) :: res
    end function
    integer(c_int) function od_jdos_get(ctx, jdos_type, jdos) bind(c, name='od_jdos_get')
      import :: c_int, c_ptr, c_double, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char), value :: jdos_type
      real(c_double), intent(out) :: jdos(*)
    end function
    integer(c_int) function od_jdos_get_weighted(ctx, weighted_jdos) bind(c, name='od_jdos_get_weighted')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: weighted_jdos(*)
    end function
    integer(c_int) function od_calculate_jdos_optics(ctx, jdos_type, delta_bins, nbins, br, efermi, scissor_op, exclude_bands, &
                                                     num_exclude_bands, optical_mat, ome_mem, geom, qdir, symops, num_symm, merge, &
                                                     jdos, weighted_jdos) bind(c, name='od_calculate_jdos_optics')
      import :: c_int, c_ptr, c_double, c_char, od_broadening
      type(c_ptr), value :: ctx
      character(kind=c_char), value :: jdos_type
      real(c_double), value :: delta_bins, efermi, scissor_op
      integer(c_int), value :: nbins, num_exclude_bands, ome_mem, geom, num_symm, merge
      type(od_broadening), intent(in) :: br
      integer(c_int), intent(in) :: exclude_bands(*)
      complex(c_double_complex), intent(in) :: optical_mat(*)
      real(c_double), intent(in) :: qdir(3)
      type(c_ptr), value :: symops
      real(c_double), intent(out) :: jdos(*), weighted_jdos(*)
    end function
    subroutine od_optics_params_defaults(p) bind(c, name='od_optics_params_defaults')
      import :: od_optics_params
      type(od_optics_params), intent(out) :: p
    end subroutine
    integer(c_int) function od_optics_calculate(ctx, p, optical_mat, ome_mem, symops, num_symm, res) bind(c, name='od_optics_calculate')
      import :: c_int, c_ptr, c_double, od_optics_params, od_optics_result
      type(c_ptr), value :: ctx
      type(od_optics_params), intent(inout) :: p
      complex(c_double_complex), intent(in) :: optical_mat(*)
      integer(c_int), value :: ome_mem, num_symm
      type(c_ptr), value :: symops
      type(od_optics_result), intent(out) :: res
    end function
    integer(c_int) function od_optics_get(ctx, E, epsilon, conduct, refract, loss_fn, absorp, reflect) bind(c, name='od_optics_get')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: E, epsilon, conduct, refract, loss_fn, absorp, reflect   ! c_loc(array) or c_null_ptr
    end function
    subroutine od_core_params_defaults(p) bind(c, name='od_core_params_defaults')
      import :: od_core_params
      type(od_core_params), intent(out) :: p
    end subroutine
    integer(c_int) function od_core_calculate(ctx, p, elnes_mat, elnes_mem, norb, symops, num_sym, res) bind(c, name='od_core_calculate')
      import :: c_int, c_ptr, c_double, od_core_params, od_core_result
      type(c_ptr), value :: ctx
      type(od_core_params), intent(inout) :: p
      complex(c_double_complex), intent(in) :: elnes_mat(*)
      integer(c_int), value :: elnes_mem, norb, num_sym
      type(c_ptr), value :: symops
      type(od_core_result), intent(out) :: res
    end function
    integer(c_int) function od_core_get(ctx, E, weighted_dos, weighted_dos_broadened) bind(c, name='od_core_get')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: E, weighted_dos, weighted_dos_broadened
    end function
    integer(c_int) function od_host_register(hptr, nbytes) bind(c, name='od_host_register')
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: hptr
      integer(c_size_t), value :: nbytes
    end function
    integer(c_int) function od_host_unregister(hptr) bind(c, name='od_host_unregister')
      import :: c_int, c_ptr
      type(c_ptr), value :: hptr
    end function
    integer(c_int) function od_abi_sizeof(struct_name) bind(c, name='od_abi_sizeof')
      import :: c_int, c_char
      character(kind=c_char), intent(in) :: struct_name(*)            ! null-terminated
    end function
    integer(c_int) function od_run_seed(ctx, dir, seedname) bind(c, name='od_run_seed')
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char), intent(in) :: dir(*), seedname(*)       ! null-terminated
    end function
    integer(c_int) function od_ctx_create(device, ctx) bind(c, name='od_ctx_create')
      import :: c_int, c_ptr
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function od_comm_get_unique_id(id128) bind(c, name='od_comm_get_unique_id')
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id128(128)
    end function
    integer(c_int) function od_ctx_create_rank(device, rank, nranks, id128, ctx) bind(c, name='od_ctx_create_rank')
      import :: c_int, c_ptr, c_char
      integer(c_int), value :: device, rank, nranks
      character(kind=c_char), intent(in) :: id128(128)
      type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function od_ctx_destroy(ctx) bind(c, name='od_ctx_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    type(c_ptr) function od_last_error(ctx) bind(c, name='od_last_error')
      import :: c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function od_allreduce(ctx, buf, n, op) bind(c, name='od_allreduce')
      import :: c_int, c_ptr, c_double, c_size_t, c_char
      type(c_ptr), value :: ctx
      real(c_double), intent(inout) :: buf(*)
      integer(c_size_t), value :: n
      character(kind=c_char), value :: op
    end function
    integer(c_int) function od_set_system(ctx, nk, ns, nb, recip, kgrid, kweight, eps) bind(c, name='od_set_system')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nk, ns, nb
      real(c_double), intent(in) :: recip(3, 3), kweight(*)
      integer(c_int), intent(in) :: kgrid(3)
      real(c_double), value :: eps
    end function
    integer(c_int) function od_set_bands(ctx, band_energy, mem) bind(c, name='od_set_bands')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: band_energy(*)
      integer(c_int), value :: mem
    end function
    integer(c_int) function od_set_gradients(ctx, band_gradient, mem) bind(c, name='od_set_gradients')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: band_gradient(*)
      integer(c_int), value :: mem
    end function
    integer(c_int) function od_calculate_dos(ctx, dos_type, emin, delta_bins, nbins, br, mw, norb, mw_nb, mw_nk, &
                                             mw_mem, merge, dos, intdos, wdos) bind(c, name='od_calculate_dos')
      import :: c_int, c_ptr, c_double, c_char, od_broadening
      type(c_ptr), value :: ctx
      character(kind=c_char), value :: dos_type
      real(c_double), value :: emin, delta_bins
      integer(c_int), value :: nbins, norb, mw_nb, mw_nk, mw_mem, merge
      type(od_broadening), intent(in) :: br
      type(c_ptr), value :: mw, wdos          ! c_null_ptr when there are no weights
      real(c_double), intent(out) :: dos(*), intdos(*)
    end function
    integer(c_int) function od_calculate_dos_at_e(ctx, dos_type, energy, delta_bins, br, mw, norb, mw_nb, mw_nk, &
                                                  mw_mem, merge, dos_at_e, wdos_at_e) bind(c, name='od_calculate_dos_at_e')
      import :: c_int, c_ptr, c_double, c_char, od_broadening
      type(c_ptr), value :: ctx
      character(kind=c_char), value :: dos_type
      real(c_double), value :: energy, delta_bins
      type(od_broadening), intent(in) :: br
      type(c_ptr), value :: mw, wdos_at_e
      integer(c_int), value :: norb, mw_nb, mw_nk, mw_mem, merge
      real(c_double), intent(out) :: dos_at_e(*)
    end function
    integer(c_int) function od_calculate_jdos(ctx, jdos_type, delta_bins, nbins, br, efermi, scissor_op, exclude_bands, &
                                              num_exclude_bands, mw, n_geom, mw_mem, merge, jdos, wjdos) &
      bind(c, name='od_calculate_jdos')
      import :: c_int, c_ptr, c_double, c_char, od_broadening
      type(c_ptr), value :: ctx
      character(kind=c_char), value :: jdos_type
      real(c_double), value :: delta_bins, efermi, scissor_op
      integer(c_int), value :: nbins, num_exclude_bands, n_geom, mw_mem, merge
      type(od_broadening), intent(in) :: br
      integer(c_int), intent(in) :: exclude_bands(*)
      type(c_ptr), value :: mw, wjdos
      real(c_double), intent(out) :: jdos(*)
    end function
    integer(c_int) function od_make_weights(ctx, optical_mat, ome_mem, geom, qdir, symops, num_symm, efermi, scissor_op, &
                                            mw, out_mem) bind(c, name='od_make_weights')
      import :: c_int, c_ptr, c_double, c_double_complex
      type(c_ptr), value :: ctx
      complex(c_double_complex), intent(in) :: optical_mat(*)
      integer(c_int), value :: ome_mem, geom, num_symm, out_mem
      real(c_double), intent(in) :: qdir(3), symops(*)
      real(c_double), value :: efermi, scissor_op
      real(c_double), intent(out) :: mw(*)
    end function
    integer(c_int) function od_core_prepare_matrix_elements(ctx, elnes_mat, elnes_mem, norb, polar, qdir, symops, num_sym, &
                                                            core_type, efermi, mw, out_mem) &
      bind(c, name='od_core_prepare_matrix_elements')
      import :: c_int, c_ptr, c_double, c_double_complex
      type(c_ptr), value :: ctx
      complex(c_double_complex), intent(in) :: elnes_mat(*)
      integer(c_int), value :: elnes_mem, norb, polar, num_sym, core_type, out_mem
      real(c_double), intent(in) :: qdir(3), symops(*)
      real(c_double), value :: efermi
      real(c_double), intent(out) :: mw(*)
    end function
    integer(c_int) function od_projection_merge(ctx, pdos_weights, pw_mem, mask, norb_in, pw_nb, nproj, mw, out_mem) &
      bind(c, name='od_projection_merge')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: pdos_weights(*)
      integer(c_int), value :: pw_mem, norb_in, pw_nb, nproj, out_mem
      integer(c_int), intent(in) :: mask(norb_in, nproj)
      real(c_double), intent(out) :: mw(*)
    end function
    ! ---- root-only O(B^2) post-processing (optics.f90:569-621, :676-697; core.f90:655-732)
    integer(c_int) function od_calc_epsilon_1(ctx, E, nbins, n_geom, epsilon_2, epsilon_1) bind(c, name='od_calc_epsilon_1')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbins, n_geom
      real(c_double), intent(in) :: E(nbins), epsilon_2(nbins, n_geom)
      real(c_double), intent(out) :: epsilon_1(nbins, n_geom)
    end function
    integer(c_int) function od_calc_loss_fn(ctx, E, nbins, epsilon_1, epsilon_2, lossfn_gaussian, loss_fn, loss_fn_broadened) &
      bind(c, name='od_calc_loss_fn')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbins
      real(c_double), value :: lossfn_gaussian
      real(c_double), intent(in) :: E(nbins), epsilon_1(nbins), epsilon_2(nbins)
      real(c_double), intent(out) :: loss_fn(nbins), loss_fn_broadened(nbins)
    end function
    integer(c_int) function od_core_lai_broadening(ctx, E, nbins, ncols, efermi, gaussian_width, lorentzian_width, &
                                                   lorentzian_scale, lorentzian_offset, weighted_dos, weighted_dos_broadened) &
      bind(c, name='od_core_lai_broadening')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbins, ncols
      real(c_double), value :: efermi, gaussian_width, lorentzian_width, lorentzian_scale, lorentzian_offset
      real(c_double), intent(in) :: E(nbins), weighted_dos(nbins, ncols)
      real(c_double), intent(out) :: weighted_dos_broadened(nbins, ncols)
    end function
  end interface

  public :: b200_setup, b200_end, b200_upload_bands, b200_upload_gradients, b200_check
  public :: b200_calculate_dos, b200_calculate_dos_at_e, b200_calculate_jdos
  public :: b200_make_weights, b200_core_weights, b200_projection_merge, b200_broadening
  public :: od_allreduce, od_comm_get_unique_id
  public :: od_calc_epsilon_1, od_calc_loss_fn, od_core_lai_broadening
  public :: od_dos_params_defaults, od_dos_utils_calculate, od_dos_get, od_dos_get_weighted, od_dos_utils_set_efermi
  public :: od_dos_utils_calculate_at_e, od_dos_utils_compute_bandgap, od_dos_utils_compute_band_energies
  public :: od_jdos_params_defaults, od_jdos_utils_calculate, od_jdos_get, od_jdos_get_weighted, od_calculate_jdos_optics
  public :: od_optics_params_defaults, od_optics_calculate, od_optics_get
  public :: od_core_params_defaults, od_core_calculate, od_core_get, od_run_seed

contains

  !> forwards a non-zero status to io_error with the library's message (io.f90:91)
  subroutine b200_check(ierr)
    use od_io, only: io_error
    integer(c_int), intent(in) :: ierr
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: text
    integer :: i
    if (ierr == 0) return
    call c_f_pointer(od_last_error(b200_ctx), msg, [512])
    text = ' '
    do i = 1, 512
      if (msg(i) == c_null_char) exit
      text(i:i) = msg(i)
    end do
    call io_error('optados_b200: '//trim(text))
  end subroutine b200_check

  !> after comms_setup (comms.F90:85) and elec_read_band_energy: one context per MPI rank, one GPU per rank.
  !> id128 is created on root with od_comm_get_unique_id and broadcast with comms_bcast (as 128 characters).
  subroutine b200_setup(local_device, id128)
    use od_comms, only: my_node_id, num_nodes
    use od_cell, only: recip_lattice, kpoint_grid_dim, kpoint_weight, num_kpoints_on_node
    use od_electronic, only: nspins, nbands, electrons_per_state
    use od_io, only: io_error
    integer, intent(in) :: local_device
    character(kind=c_char), intent(in) :: id128(128)
    type(od_dos_params) :: chk_dos
    type(od_jdos_params) :: chk_jdos
    type(od_optics_params) :: chk_optics
    type(od_core_params) :: chk_core
    type(od_bandgap_result) :: chk_gap
    ! the derived types above must have the layout the library was compiled with
    if (c_sizeof(chk_dos) /= od_abi_sizeof('od_dos_params'//c_null_char) .or. &
        c_sizeof(chk_jdos) /= od_abi_sizeof('od_jdos_params'//c_null_char) .or. &
        c_sizeof(chk_optics) /= od_abi_sizeof('od_optics_params'//c_null_char) .or. &
        c_sizeof(chk_core) /= od_abi_sizeof('od_core_params'//c_null_char) .or. &
        c_sizeof(chk_gap) /= od_abi_sizeof('od_bandgap_result'//c_null_char)) &
      call io_error('optados_b200: od_b200.f90 and liboptados_b200.so disagree on a structure layout (rebuild both)')
    if (num_nodes > 1) then
      call b200_check(od_ctx_create_rank(int(local_device, c_int), int(my_node_id, c_int), int(num_nodes, c_int), id128, b200_ctx))
    else
      call b200_check(od_ctx_create(int(local_device, c_int), b200_ctx))
    end if
    call b200_check(od_set_system(b200_ctx, int(num_kpoints_on_node(my_node_id), c_int), int(nspins, c_int), &
                                  int(nbands, c_int), recip_lattice, int(kpoint_grid_dim, c_int), kpoint_weight, &
                                  real(electrons_per_state, c_double)))
  end subroutine b200_setup

  subroutine b200_end()
    integer(c_int) :: ierr
    ierr = od_ctx_destroy(b200_ctx)
    b200_ctx = c_null_ptr
  end subroutine b200_end

  subroutine b200_upload_bands()
    use od_electronic, only: band_energy
    call b200_check(od_set_bands(b200_ctx, band_energy, OD_MEM_HOST))
  end subroutine b200_upload_bands

  !> call once after elec_read_band_gradient (electronic.f90:170)
  !> band_gradient is the big input (12 GB at BASELINE config 2).  Page-lock the Fortran array once (od_host_register: a
  !> pageable source is staged through a bounce buffer at a fraction of the PCIe rate) and hand the library the HOST
  !> pointer: with OD_MEM_HOST_DEFERRED it is uploaded slab by slab inside calculate_dos, under the kernels.
  subroutine b200_upload_gradients()
    use od_electronic, only: band_gradient
    integer(c_int) :: ierr
    ierr = od_host_register(c_loc(band_gradient), int(size(band_gradient), c_size_t)*8_c_size_t)   ! (failure = stay pageable)
    call b200_check(od_set_gradients(b200_ctx, band_gradient, OD_MEM_HOST_DEFERRED))
  end subroutine b200_upload_gradients

  function b200_broadening() result(br)
    use od_parameters, only: adaptive_smearing, fixed_smearing, linear_smearing, finite_bin_correction, &
      numerical_intdos, hybrid_linear, hybrid_linear_grad_tol
    type(od_broadening) :: br
    br%adaptive_smearing = adaptive_smearing
    br%fixed_smearing = fixed_smearing
    br%linear_smearing = linear_smearing
    br%finite_bin_correction = merge(1_c_int, 0_c_int, finite_bin_correction)
    br%numerical_intdos = merge(1_c_int, 0_c_int, numerical_intdos)
    br%hybrid_linear = merge(1_c_int, 0_c_int, hybrid_linear)
    br%hybrid_linear_grad_tol = hybrid_linear_grad_tol
  end function b200_broadening

  !> body of calculate_dos (dos_utils.f90:1140) followed by dos_utils_merge (:1013).
  !> E(1) and delta_bins are the module variables set by setup_energy_scale (:994-997).
  subroutine b200_calculate_dos(dos_type, E1, delta_bins, dos, intdos, matrix_weights, weighted_dos)
    use od_parameters, only: dos_nbins
    use od_electronic, only: nspins
    character(len=1), intent(in) :: dos_type
    real(dp), intent(in) :: E1, delta_bins
    real(dp), intent(out), allocatable :: dos(:, :), intdos(:, :)
    real(dp), intent(in), optional, target :: matrix_weights(:, :, :, :)
    real(dp), intent(out), allocatable, optional, target :: weighted_dos(:, :, :)
    type(c_ptr) :: pmw, pw
    integer(c_int) :: norb, mwnb, mwnk
    allocate (dos(dos_nbins, nspins), intdos(dos_nbins, nspins))
    pmw = c_null_ptr; pw = c_null_ptr; norb = 0; mwnb = 0; mwnk = 0
    if (present(matrix_weights)) then
      norb = size(matrix_weights, 1); mwnb = size(matrix_weights, 2); mwnk = size(matrix_weights, 3)
      allocate (weighted_dos(dos_nbins, nspins, norb))
      pmw = c_loc(matrix_weights); pw = c_loc(weighted_dos)
    end if
    call b200_check(od_calculate_dos(b200_ctx, dos_type, E1, delta_bins, int(dos_nbins, c_int), b200_broadening(), &
                                     pmw, norb, mwnb, mwnk, OD_MEM_HOST, 1_c_int, dos, intdos, pw))
  end subroutine b200_calculate_dos

  !> body of calculate_dos_at_e (dos_utils.f90:1651) + dos_utils_merge_at_e (:1803)
  subroutine b200_calculate_dos_at_e(dos_type, energy, delta_bins, dos_at_e, matrix_weights, weighted_dos_at_e)
    character(len=1), intent(in) :: dos_type
    real(dp), intent(in) :: energy, delta_bins
    real(dp), intent(out) :: dos_at_e(:)
    real(dp), intent(in), optional, target :: matrix_weights(:, :, :, :)
    real(dp), intent(out), optional, target :: weighted_dos_at_e(:, :)
    type(c_ptr) :: pmw, pw
    integer(c_int) :: norb, mwnb, mwnk
    pmw = c_null_ptr; pw = c_null_ptr; norb = 0; mwnb = 0; mwnk = 0
    if (present(matrix_weights) .and. present(weighted_dos_at_e)) then
      norb = size(matrix_weights, 1); mwnb = size(matrix_weights, 2); mwnk = size(matrix_weights, 3)
      pmw = c_loc(matrix_weights); pw = c_loc(weighted_dos_at_e)
    end if
    call b200_check(od_calculate_dos_at_e(b200_ctx, dos_type, energy, delta_bins, b200_broadening(), pmw, norb, mwnb, &
                                          mwnk, OD_MEM_HOST, 1_c_int, dos_at_e, pw))
  end subroutine b200_calculate_dos_at_e

  !> body of calculate_jdos (jdos_utils.f90:280) + jdos_utils_merge (:407)
  subroutine b200_calculate_jdos(jdos_type, delta_bins, jdos_nbins, jdos, matrix_weights, weighted_jdos)
    use od_parameters, only: scissor_op, exclude_bands, num_exclude_bands
    use od_electronic, only: nspins, efermi
    character(len=1), intent(in) :: jdos_type
    real(dp), intent(in) :: delta_bins
    integer, intent(in) :: jdos_nbins
    real(dp), intent(out), allocatable :: jdos(:, :)
    real(dp), intent(in), optional, target :: matrix_weights(:, :, :, :, :)
    real(dp), intent(inout), allocatable, optional, target :: weighted_jdos(:, :, :)
    type(c_ptr) :: pmw, pw
    integer(c_int) :: ngeom
    integer(c_int) :: dummy(1)
    allocate (jdos(jdos_nbins, nspins))
    pmw = c_null_ptr; pw = c_null_ptr; ngeom = 0; dummy = 0
    if (present(matrix_weights)) then
      ngeom = size(matrix_weights, 5)
      allocate (weighted_jdos(jdos_nbins, nspins, ngeom))
      pmw = c_loc(matrix_weights); pw = c_loc(weighted_jdos)
    end if
    if (num_exclude_bands > 0) then
      call b200_check(od_calculate_jdos(b200_ctx, jdos_type, delta_bins, int(jdos_nbins, c_int), b200_broadening(), efermi, &
                                        scissor_op, int(exclude_bands, c_int), int(num_exclude_bands, c_int), pmw, ngeom, &
                                        OD_MEM_HOST, 1_c_int, jdos, pw))
    else
      call b200_check(od_calculate_jdos(b200_ctx, jdos_type, delta_bins, int(jdos_nbins, c_int), b200_broadening(), efermi, &
                                        scissor_op, dummy, 0_c_int, pmw, ngeom, OD_MEM_HOST, 1_c_int, jdos, pw))
    end if
  end subroutine b200_calculate_jdos

  !> body of make_weights (optics.f90:157): geom 0 polycrys, 1 polar, 2 unpolar, 3 tensor
  subroutine b200_make_weights(geom, matrix_weights)
    use od_electronic, only: optical_mat, efermi
    use od_cell, only: num_crystal_symmetry_operations, crystal_symmetry_operations
    use od_parameters, only: optics_qdir, scissor_op
    integer, intent(in) :: geom
    real(dp), intent(out) :: matrix_weights(:, :, :, :, :)
    real(c_double) :: nosym(9)
    nosym = 0.0_dp
    if (num_crystal_symmetry_operations > 0) then
      call b200_check(od_make_weights(b200_ctx, optical_mat, OD_MEM_HOST, int(geom, c_int), optics_qdir, &
                                      crystal_symmetry_operations, int(num_crystal_symmetry_operations, c_int), efermi, &
                                      scissor_op, matrix_weights, OD_MEM_HOST))
    else
      call b200_check(od_make_weights(b200_ctx, optical_mat, OD_MEM_HOST, int(geom, c_int), optics_qdir, nosym, 0_c_int, &
                                      efermi, scissor_op, matrix_weights, OD_MEM_HOST))
    end if
  end subroutine b200_make_weights

  !> body of core_prepare_matrix_elements (core.f90:84): core_type 0 absorption, 1 emission, 2 all
  subroutine b200_core_weights(polar, core_type, matrix_weights)
    use od_electronic, only: elnes_mat, elnes_mwab, efermi
    use od_cell, only: num_crystal_symmetry_operations, crystal_symmetry_operations
    use od_parameters, only: core_qdir
    logical, intent(in) :: polar
    integer, intent(in) :: core_type
    real(dp), intent(out) :: matrix_weights(:, :, :, :)
    call b200_check(od_core_prepare_matrix_elements(b200_ctx, elnes_mat, OD_MEM_HOST, int(elnes_mwab%norbitals, c_int), &
                                                    merge(1_c_int, 0_c_int, polar), core_qdir, crystal_symmetry_operations, &
                                                    int(num_crystal_symmetry_operations, c_int), int(core_type, c_int), &
                                                    efermi, matrix_weights, OD_MEM_HOST))
  end subroutine b200_core_weights

  !> body of projection_merge (projection_utils.f90:64): the 4-index lookup flattened to mask(orb, proj)
  subroutine b200_projection_merge(projection_array, num_proj, matrix_weights)
    use od_electronic, only: pdos_orbital, pdos_weights, pdos_mwab
    integer, intent(in) :: projection_array(:, :, :, :), num_proj
    real(dp), intent(out) :: matrix_weights(:, :, :, :)
    integer(c_int), allocatable :: mask(:, :)
    integer :: orb, p
    allocate (mask(pdos_mwab%norbitals, num_proj))
    do p = 1, num_proj
      do orb = 1, pdos_mwab%norbitals
        mask(orb, p) = projection_array(pdos_orbital%species_no(orb), pdos_orbital%rank_in_species(orb), &
                                        pdos_orbital%am_channel(orb) + 1, p)
      end do
    end do
    call b200_check(od_projection_merge(b200_ctx, pdos_weights, OD_MEM_HOST, mask, int(pdos_mwab%norbitals, c_int), &
                                        int(pdos_mwab%nbands, c_int), int(num_proj, c_int), matrix_weights, OD_MEM_HOST))
  end subroutine b200_projection_merge

end module od_b200
