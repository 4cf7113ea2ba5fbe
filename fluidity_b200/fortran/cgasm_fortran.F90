!  cgasm_fortran.F90 -- ISO_C_BINDING interface module for libcgasm.so (include/cgasm.h).
!
!  This is the reference-side half of the drop-in boundary: a maintainer adds this file to
!  femtools/ (it follows the conventions of femtools/Node_Owner_Finder_Fortran.F90:61-98 and
!  femtools/Integer_set.F90:15-60: bind(c) interfaces, integer handle, flat arrays, integer
!  status) and replaces the two colour/element loops by the calls shown in INTEGRATION.md.
!
!  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler is available here); the C
!  side it binds to is exercised through the identical ctypes binding fluidity_b200/cgasm.py.
!  Every dummy argument below maps 1:1 to the C prototype of the same name.
#include "fdebug.h"
module cgasm_interface
  use iso_c_binding
  implicit none

  private
  public :: cgasm_momentum_opts, cgasm_advdiff_opts
  public :: cgasm_create, cgasm_destroy, cgasm_set_coordinates, cgasm_set_sparsity, &
       & cgasm_build_sparsity, cgasm_get_sparsity, cgasm_set_colouring, cgasm_build_colouring, &
       & cgasm_get_colouring, cgasm_set_scatter, cgasm_set_field, cgasm_get_field, &
       & cgasm_momentum, cgasm_advdiff, cgasm_momentum_dev, cgasm_advdiff_dev, cgasm_momentum_advdiff_dev, &
       & cgasm_momentum_fetch, cgasm_advdiff_fetch, cgasm_momentum_identical_blocks, &
       & cgasm_momentum_mass_fetch, cgasm_momentum_mass_dev, &
       & cgasm_momentum_fetch_blocks, cgasm_momentum_element, &
       & cgasm_advdiff_element, cgasm_synchronize, cgasm_set_async, cgasm_halo_create, cgasm_halo_update, cgasm_halo_set_overlap, &
       & cgasm_coo_pattern_dev, cgasm_coo_values_dev, cgasm_coo_fetch, &
       & cgasm_nccl_unique_id, cgasm_last_error, cgasm_last_error_string, cgasm_set_surface, cgasm_advdiff_surface_dev, &
       & cgasm_advdiff_dirichlet_dev, cgasm_momentum_surface_dev, cgasm_momentum_dirichlet_dev, &
       & cgasm_correct_masslumped_velocity, cgasm_cmc_build_sparsity, &
       & cgasm_cmc_get_sparsity, cgasm_cmc_set_sparsity, cgasm_cmc_dev, cgasm_cmc_fetch, &
       & cgasm_kmk_dev, cgasm_kmk_fetch
  public :: CGASM_OK, CGASM_EUNSUPPORTED
  public :: CGASM_F_NU, CGASM_F_OLDU, CGASM_F_DENSITY, CGASM_F_VISCOSITY, CGASM_F_BUOYANCY, &
       & CGASM_F_HB_DENSITY, CGASM_F_GRAVITY, CGASM_F_ABSORPTION, CGASM_F_SOURCE, CGASM_F_T, &
       & CGASM_F_T_DIFFUSIVITY, CGASM_F_T_SOURCE, CGASM_F_T_ABSORPTION
  public :: CGASM_SCATTER_ATOMIC, CGASM_SCATTER_COLOURED, CGASM_SCATTER_WARPAGG, &
       & CGASM_SCATTER_TILED, CGASM_SCATTER_GATHER, CGASM_SCATTER_STRIP

  integer(c_int), parameter :: CGASM_OK = 0, CGASM_EUNSUPPORTED = 3
  integer(c_int), parameter :: CGASM_F_NU = 0, CGASM_F_OLDU = 1, CGASM_F_DENSITY = 2, &
       & CGASM_F_VISCOSITY = 3, CGASM_F_BUOYANCY = 4, CGASM_F_HB_DENSITY = 5, &
       & CGASM_F_GRAVITY = 6, CGASM_F_ABSORPTION = 7, CGASM_F_SOURCE = 8, CGASM_F_T = 9, &
       & CGASM_F_T_DIFFUSIVITY = 10, CGASM_F_T_SOURCE = 11, CGASM_F_T_ABSORPTION = 12
  integer(c_int), parameter :: CGASM_SCATTER_ATOMIC = 0, CGASM_SCATTER_COLOURED = 1, &
       & CGASM_SCATTER_WARPAGG = 2, CGASM_SCATTER_TILED = 3, CGASM_SCATTER_GATHER = 4, &
       & CGASM_SCATTER_STRIP = 5

  !! Mirrors struct cgasm_momentum_opts: the module-level switches of
  !! assemble/Momentum_CG.F90:83-178. Logicals travel as integer(c_int) (0/1).
  type, bind(c) :: cgasm_momentum_opts
     real(c_double) :: dt, theta, beta, gravity_magnitude, nu_bar_scale, fs_sf
     integer(c_int) :: lump_mass, exclude_mass, exclude_advection, integrate_advection_by_parts
     integer(c_int) :: have_source, lump_source, have_gravity, subtract_out_reference_profile
     integer(c_int) :: have_absorption, lump_absorption, pressure_corrected_absorption
     integer(c_int) :: have_viscosity, viscosity_shape, assemble_inverse_masslump
     integer(c_int) :: assemble_ct_matrix_here, stabilisation_scheme, nu_bar_scheme
     integer(c_int) :: have_les, multiphase, on_sphere, move_mesh, have_coriolis
     integer(c_int) :: have_geostrophic_pressure, have_surfacetension, have_vertical_stabilization
     integer(c_int) :: have_swe_bottom_drag, have_wd_abs, have_temperature_dependent_viscosity
     integer(c_int) :: stress_form, partial_stress_form, radial_gravity, vel_lump_on_submesh
     integer(c_int) :: cmc_lump_on_submesh, abs_lump_on_submesh, assemble_mass_matrix
     integer(c_int) :: integrate_continuity_by_parts, have_surface_fs_stabilisation
  end type cgasm_momentum_opts

  !! Mirrors struct cgasm_advdiff_opts: assemble/Advection_Diffusion_CG.F90:77-123.
  type, bind(c) :: cgasm_advdiff_opts
     real(c_double) :: dt, theta, beta, nu_bar_scale
     integer(c_int) :: have_mass, lump_mass, have_advection, integrate_advection_by_parts
     integer(c_int) :: have_source, have_absorption, have_diffusivity, diffusivity_shape
     integer(c_int) :: stabilisation_scheme, nu_bar_scheme
     integer(c_int) :: move_mesh, multiphase, equation_type_not_advdiff
  end type cgasm_advdiff_opts

  interface
     function cgasm_create(id, device, dim, loc, ngi, n_nodes, n_elements, ndglno, n, dn, weight) &
          & bind(c, name="cgasm_create") result(stat)
       use iso_c_binding
       integer(c_int), intent(out) :: id
       integer(c_int), value :: device, dim, loc, ngi, n_nodes, n_elements
       integer(c_int), dimension(*), intent(in) :: ndglno
       real(c_double), dimension(*), intent(in) :: n, dn, weight
       integer(c_int) :: stat
     end function cgasm_create

     function cgasm_destroy(id) bind(c, name="cgasm_destroy") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int) :: stat
     end function cgasm_destroy

     function cgasm_last_error() bind(c, name="cgasm_last_error") result(msg)
       use iso_c_binding
       type(c_ptr) :: msg
     end function cgasm_last_error

     function cgasm_set_coordinates(id, x) bind(c, name="cgasm_set_coordinates") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       real(c_double), dimension(*), intent(in) :: x
       integer(c_int) :: stat
     end function cgasm_set_coordinates

     function cgasm_build_sparsity(id, nnz) bind(c, name="cgasm_build_sparsity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), intent(out) :: nnz
       integer(c_int) :: stat
     end function cgasm_build_sparsity

     function cgasm_get_sparsity(id, findrm, colm, centrm) bind(c, name="cgasm_get_sparsity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), dimension(*), intent(out) :: findrm, colm, centrm
       integer(c_int) :: stat
     end function cgasm_get_sparsity

     function cgasm_set_sparsity(id, rows, nnz, findrm, colm) bind(c, name="cgasm_set_sparsity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, rows, nnz
       integer(c_int), dimension(*), intent(in) :: findrm, colm
       integer(c_int) :: stat
     end function cgasm_set_sparsity

     function cgasm_build_colouring(id, ncolours) bind(c, name="cgasm_build_colouring") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), intent(out) :: ncolours
       integer(c_int) :: stat
     end function cgasm_build_colouring

     function cgasm_get_colouring(id, colour_ptr, colour_elements) bind(c, name="cgasm_get_colouring") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), dimension(*), intent(out) :: colour_ptr, colour_elements
       integer(c_int) :: stat
     end function cgasm_get_colouring

     function cgasm_set_colouring(id, ncolours, colour_ptr, colour_elements) &
          & bind(c, name="cgasm_set_colouring") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, ncolours
       integer(c_int), dimension(*), intent(in) :: colour_ptr, colour_elements
       integer(c_int) :: stat
     end function cgasm_set_colouring

     function cgasm_set_scatter(id, variant) bind(c, name="cgasm_set_scatter") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, variant
       integer(c_int) :: stat
     end function cgasm_set_scatter

     function cgasm_set_field(id, slot, rank, field_type, val, n_val_nodes) &
          & bind(c, name="cgasm_set_field") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, slot, rank, field_type, n_val_nodes
       real(c_double), dimension(*), intent(in) :: val
       integer(c_int) :: stat
     end function cgasm_set_field

     function cgasm_get_field(id, slot, val, n_val_nodes) bind(c, name="cgasm_get_field") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, slot, n_val_nodes
       real(c_double), dimension(*), intent(out) :: val
       integer(c_int) :: stat
     end function cgasm_get_field

     function cgasm_momentum(id, opts, big_m, rhs, masslump, ct_m) bind(c, name="cgasm_momentum") result(stat)
       use iso_c_binding
       import :: cgasm_momentum_opts
       integer(c_int), value :: id
       type(cgasm_momentum_opts), intent(in) :: opts
       real(c_double), dimension(*), intent(out) :: big_m, rhs
       type(c_ptr), value :: masslump, ct_m   ! c_loc(array) or c_null_ptr
       integer(c_int) :: stat
     end function cgasm_momentum

     function cgasm_advdiff(id, opts, matrix_val, rhs) bind(c, name="cgasm_advdiff") result(stat)
       use iso_c_binding
       import :: cgasm_advdiff_opts
       integer(c_int), value :: id
       type(cgasm_advdiff_opts), intent(in) :: opts
       real(c_double), dimension(*), intent(out) :: matrix_val, rhs
       integer(c_int) :: stat
     end function cgasm_advdiff

     function cgasm_momentum_dev(id, opts) bind(c, name="cgasm_momentum_dev") result(stat)
       use iso_c_binding
       import :: cgasm_momentum_opts
       integer(c_int), value :: id
       type(cgasm_momentum_opts), intent(in) :: opts
       integer(c_int) :: stat
     end function cgasm_momentum_dev

     function cgasm_advdiff_dev(id, opts) bind(c, name="cgasm_advdiff_dev") result(stat)
       use iso_c_binding
       import :: cgasm_advdiff_opts
       integer(c_int), value :: id
       type(cgasm_advdiff_opts), intent(in) :: opts
       integer(c_int) :: stat
     end function cgasm_advdiff_dev

     function cgasm_momentum_advdiff_dev(id, mopts, aopts) bind(c, name="cgasm_momentum_advdiff_dev") result(stat)
       use iso_c_binding
       import :: cgasm_momentum_opts, cgasm_advdiff_opts
       integer(c_int), value :: id
       type(cgasm_momentum_opts), intent(in) :: mopts
       type(cgasm_advdiff_opts), intent(in) :: aopts
       integer(c_int) :: stat
     end function cgasm_momentum_advdiff_dev

     ! the `mass` matrix of construct_momentum_cg (assemble_mass_matrix): dim diagonal blocks of nnz values
     function cgasm_momentum_mass_fetch(id, mass) bind(c, name="cgasm_momentum_mass_fetch") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       real(c_double), intent(out) :: mass(*)
       integer(c_int) :: stat
     end function cgasm_momentum_mass_fetch

     function cgasm_momentum_mass_dev(id, mass_dev) bind(c, name="cgasm_momentum_mass_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       type(c_ptr), intent(out) :: mass_dev
       integer(c_int) :: stat
     end function cgasm_momentum_mass_dev

     function cgasm_momentum_fetch(id, big_m, rhs, masslump, ct_m) bind(c, name="cgasm_momentum_fetch") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       type(c_ptr), value :: big_m, rhs, masslump, ct_m
       integer(c_int) :: stat
     end function cgasm_momentum_fetch

     function cgasm_momentum_identical_blocks(id, identical) &
          & bind(c, name="cgasm_momentum_identical_blocks") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), intent(out) :: identical
       integer(c_int) :: stat
     end function cgasm_momentum_identical_blocks

     function cgasm_momentum_fetch_blocks(id, first_block, nblocks, big_m) &
          & bind(c, name="cgasm_momentum_fetch_blocks") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, first_block, nblocks
       real(c_double), dimension(*), intent(out) :: big_m
       integer(c_int) :: stat
     end function cgasm_momentum_fetch_blocks

     function cgasm_advdiff_fetch(id, matrix_val, rhs) bind(c, name="cgasm_advdiff_fetch") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       type(c_ptr), value :: matrix_val, rhs
       integer(c_int) :: stat
     end function cgasm_advdiff_fetch

     function cgasm_momentum_element(id, opts, ele, big_m_tensor_addto, rhs_addto, mass_lump, grad_p_u_mat) &
          & bind(c, name="cgasm_momentum_element") result(stat)
       use iso_c_binding
       import :: cgasm_momentum_opts
       integer(c_int), value :: id, ele
       type(cgasm_momentum_opts), intent(in) :: opts
       real(c_double), dimension(*), intent(out) :: big_m_tensor_addto, rhs_addto, mass_lump, grad_p_u_mat
       integer(c_int) :: stat
     end function cgasm_momentum_element

     function cgasm_advdiff_element(id, opts, ele, matrix_addto, rhs_addto) &
          & bind(c, name="cgasm_advdiff_element") result(stat)
       use iso_c_binding
       import :: cgasm_advdiff_opts
       integer(c_int), value :: id, ele
       type(cgasm_advdiff_opts), intent(in) :: opts
       real(c_double), dimension(*), intent(out) :: matrix_addto, rhs_addto
       integer(c_int) :: stat
     end function cgasm_advdiff_element

     !! Boundary faces of the mesh: sndgln = face_global_nodes of every surface element, face_ele, and the
     !! tables of mesh%faces%shape (n, dn, quadrature%weight)
     function cgasm_set_surface(id, n_faces, sloc, sngi, sndgln, face_ele, n_f, dn_f, weight_f) &
          & bind(c, name="cgasm_set_surface") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, n_faces, sloc, sngi
       integer(c_int), dimension(*), intent(in) :: sndgln, face_ele
       real(c_double), dimension(*), intent(in) :: n_f, dn_f, weight_f
       integer(c_int) :: stat
     end function cgasm_set_surface

     !! Face loop of assemble_advection_diffusion_cg (Advection_Diffusion_CG.F90:609-643), added to the
     !! device-resident result of cgasm_advdiff_dev. t_bc, t_bc_2: (sloc, n_faces) = ele_val(t_bc, face)
     function cgasm_advdiff_surface_dev(id, opts, bc_type, t_bc, t_bc_2) &
          & bind(c, name="cgasm_advdiff_surface_dev") result(stat)
       use iso_c_binding
       import :: cgasm_advdiff_opts
       integer(c_int), value :: id
       type(cgasm_advdiff_opts), intent(in) :: opts
       integer(c_int), dimension(*), intent(in) :: bc_type
       type(c_ptr), value :: t_bc, t_bc_2   ! c_loc of the (sloc, n_faces) values, or c_null_ptr if no face reads them
       integer(c_int) :: stat
     end function cgasm_advdiff_surface_dev

     !! apply_dirichlet_conditions (Boundary_Conditions.F90:1982-2024) for one boundary condition
     function cgasm_advdiff_dirichlet_dev(id, n, nodes, values, have_dt, dt) &
          & bind(c, name="cgasm_advdiff_dirichlet_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, n, have_dt
       integer(c_int), dimension(*), intent(in) :: nodes
       real(c_double), dimension(*), intent(in) :: values
       real(c_double), value :: dt
       integer(c_int) :: stat
     end function cgasm_advdiff_dirichlet_dev

     !! surface_element_loop of construct_momentum_cg (Momentum_CG.F90:795-812): velocity_bc_type(dim, n_faces),
     !! velocity_bc(dim, sloc, n_faces), pressure_bc_type(n_faces)
     function cgasm_momentum_surface_dev(id, opts, velocity_bc_type, velocity_bc, pressure_bc_type) &
          & bind(c, name="cgasm_momentum_surface_dev") result(stat)
       use iso_c_binding
       import :: cgasm_momentum_opts
       integer(c_int), value :: id
       type(cgasm_momentum_opts), intent(in) :: opts
       integer(c_int), dimension(*), intent(in) :: velocity_bc_type
       type(c_ptr), value :: velocity_bc, pressure_bc_type   ! c_loc(...) or c_null_ptr
       integer(c_int) :: stat
     end function cgasm_momentum_surface_dev

     !! Lumped-mass pressure matrix (assemble_masslumped_cmc, Assemble_CMC.F90:119-135) on the second-order
     !! sparsity of get_csr_sparsity_secondorder: adopt cmc_m%sparsity (findrm, colm) or build the same pattern
     function cgasm_cmc_set_sparsity(id, rows, nnz2, findrm2, colm2) bind(c, name="cgasm_cmc_set_sparsity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, rows, nnz2
       integer(c_int), dimension(*), intent(in) :: findrm2, colm2
       integer(c_int) :: stat
     end function cgasm_cmc_set_sparsity

     function cgasm_cmc_build_sparsity(id, nnz2) bind(c, name="cgasm_cmc_build_sparsity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_long_long), intent(out) :: nnz2
       integer(c_int) :: stat
     end function cgasm_cmc_build_sparsity

     function cgasm_cmc_get_sparsity(id, findrm2, colm2) bind(c, name="cgasm_cmc_get_sparsity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), dimension(*), intent(out) :: findrm2, colm2
       integer(c_int) :: stat
     end function cgasm_cmc_get_sparsity

     !! ct_m: the dim blocks of ct_m%val(1,d)%ptr back to back; inverse_masslump%val(dim, nodes). Pass
     !! c_null_ptr-associated arrays (or use the *_resident wrapper of INTEGRATION.md) to use the device copies
     function cgasm_cmc_dev(id, ct_m, inverse_masslump) bind(c, name="cgasm_cmc_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       type(c_ptr), value :: ct_m, inverse_masslump
       integer(c_int) :: stat
     end function cgasm_cmc_dev

     function cgasm_cmc_fetch(id, cmc_val) bind(c, name="cgasm_cmc_fetch") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       real(c_double), dimension(*), intent(out) :: cmc_val
       integer(c_int) :: stat
     end function cgasm_cmc_fetch

     !! P1-P1 stabilisation (assemble_kmk_matrix, Momentum_CG.F90:2707-2766) on the same second-order sparsity;
     !! kt / p_masslump: c_null_ptr to skip
     function cgasm_kmk_dev(id, theta_pg) bind(c, name="cgasm_kmk_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       real(c_double), value :: theta_pg
       integer(c_int) :: stat
     end function cgasm_kmk_dev

     function cgasm_kmk_fetch(id, kmk, kt, p_masslump) bind(c, name="cgasm_kmk_fetch") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       real(c_double), dimension(*), intent(out) :: kmk
       type(c_ptr), value :: kt, p_masslump
       integer(c_int) :: stat
     end function cgasm_kmk_fetch

     !! on /= 0: uploads and result downloads are queued (two streams); cgasm_synchronize waits
     function cgasm_set_async(id, on) bind(c, name="cgasm_set_async") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, on
       integer(c_int) :: stat
     end function cgasm_set_async

     function cgasm_synchronize(id) bind(c, name="cgasm_synchronize") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int) :: stat
     end function cgasm_synchronize

     function cgasm_nccl_unique_id(out128) bind(c, name="cgasm_nccl_unique_id") result(stat)
       use iso_c_binding
       character(kind=c_char), dimension(128), intent(out) :: out128
       integer(c_int) :: stat
     end function cgasm_nccl_unique_id

     function cgasm_halo_create(id, nprocs, rank, nsend, sends, nrecv, recvs, nccl_unique_id) &
          & bind(c, name="cgasm_halo_create") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, nprocs, rank
       integer(c_int), dimension(*), intent(in) :: nsend, sends, nrecv, recvs
       character(kind=c_char), dimension(128), intent(in) :: nccl_unique_id
       integer(c_int) :: stat
     end function cgasm_halo_create

     function cgasm_halo_update(id, slot_mask) bind(c, name="cgasm_halo_update") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), value :: slot_mask
       integer(c_int) :: stat
     end function cgasm_halo_update

     function cgasm_halo_set_overlap(id, on) bind(c, name="cgasm_halo_set_overlap") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       integer(c_int), value :: on
       integer(c_int) :: stat
     end function cgasm_halo_set_overlap

     ! strong Dirichlet conditions of the velocity on the resident big_m / rhs (lift_boundary_conditions)
     function cgasm_momentum_dirichlet_dev(id, n, nodes, comps, values) bind(c, name="cgasm_momentum_dirichlet_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, n
       integer(c_int), intent(in) :: nodes(*), comps(*)
       real(c_double), intent(in) :: values(*)
       integer(c_int) :: stat
     end function cgasm_momentum_dirichlet_dev

     ! correct_masslumped_velocity: u += inverse_masslump * ct_m^T delta_p (ct_m: c_null_ptr = the resident one)
     function cgasm_correct_masslumped_velocity(id, ct_m, inverse_masslump, delta_p, u) &
          & bind(c, name="cgasm_correct_masslumped_velocity") result(stat)
       use iso_c_binding
       integer(c_int), value :: id
       type(c_ptr), value :: ct_m
       real(c_double), intent(in) :: inverse_masslump(*), delta_p(*)
       real(c_double), intent(inout) :: u(*)
       integer(c_int) :: stat
     end function cgasm_correct_masslumped_velocity

     ! device hand-off to PETSc: the (i, j) pattern once per sparsity for MatSetPreallocationCOO, the values per assembly
     ! for MatSetValuesCOO. row_gnn2unn / col_gnn2unn = petsc_numbering%gnn2unn of the matrix (col: c_null_ptr = rows).
     function cgasm_coo_pattern_dev(id, which, row_gnn2unn, col_gnn2unn, compact, ncoo, coo_i_dev, coo_j_dev) &
          & bind(c, name="cgasm_coo_pattern_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, which
       integer(c_int), intent(in) :: row_gnn2unn(*)
       type(c_ptr), value :: col_gnn2unn
       integer(c_int), value :: compact
       integer(c_long_long), intent(out) :: ncoo
       type(c_ptr), intent(out) :: coo_i_dev, coo_j_dev
       integer(c_int) :: stat
     end function cgasm_coo_pattern_dev

     function cgasm_coo_values_dev(id, which, coo_v_dev) bind(c, name="cgasm_coo_values_dev") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, which
       type(c_ptr), intent(out) :: coo_v_dev
       integer(c_int) :: stat
     end function cgasm_coo_values_dev

     function cgasm_coo_fetch(id, which, ncoo, coo_i, coo_j, coo_v) bind(c, name="cgasm_coo_fetch") result(stat)
       use iso_c_binding
       integer(c_int), value :: id, which
       integer(c_long_long), value :: ncoo
       type(c_ptr), value :: coo_i, coo_j, coo_v
       integer(c_int) :: stat
     end function cgasm_coo_fetch
  end interface

contains

  !! cgasm_last_error as a Fortran string (for FLAbort in the shim, INTEGRATION.md section 3)
  function cgasm_last_error_string() result(msg)
    use iso_c_binding
    character(len=512) :: msg
    type(c_ptr) :: p
    character(kind=c_char), pointer :: c(:)
    integer :: i
    msg = ""
    p = cgasm_last_error()
    if (.not. c_associated(p)) return
    call c_f_pointer(p, c, (/ 512 /))
    do i = 1, 512
       if (c(i) == c_null_char) exit
       msg(i:i) = c(i)
    end do
  end function cgasm_last_error_string

end module cgasm_interface
