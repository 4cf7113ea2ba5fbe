/*
 * cgasm.h -- C ABI of the B200-native continuous-Galerkin element assembly.
 *
 * One call into this library replaces one element loop of the reference:
 *
 *   cgasm_momentum  <->  assemble/Momentum_CG.F90:726-752  (colour loop around
 *                        construct_momentum_element_cg, :1193-1490)
 *   cgasm_advdiff   <->  assemble/Advection_Diffusion_CG.F90:574-598 (colour loop
 *                        around assemble_advection_diffusion_element_cg, :702-865)
 *
 * Conventions follow the reference's own Fortran<->C seams (integer handle + flat
 * arrays + integer status, femtools/Node_Owner_Finder_Fortran.F90:61-98):
 *   - every pointer argument is HOST memory owned by the caller unless the name ends
 *     in _dev;
 *   - integers are 32-bit and indices are 1-BASED exactly as Fortran holds them
 *     (mesh%ndglno, findrm, colm, halo send/receive lists);
 *   - reals are FP64; vector fields are val(dim, nodes), tensor fields
 *     val(dim, dim, nodes), column-major (femtools/Fields_Data_Types.F90:154-233);
 *   - every function returns 0 on success or a CGASM_E* code; the library never aborts
 *     (the Fortran side decides to FLAbort) and never falls back to a CPU path.
 *
 * All work of one handle is issued on one CUDA stream of one device.
 */
#ifndef CGASM_H
#define CGASM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------ */
enum {
  CGASM_OK = 0,
  CGASM_EHANDLE = 1,     /* unknown / destroyed handle                              */
  CGASM_EARG = 2,        /* invalid argument (bad dim, null pointer, bad index ...) */
  CGASM_EUNSUPPORTED = 3,/* option combination outside the device path: caller must */
                         /* keep the Fortran loop for this assembly                 */
  CGASM_ESTATE = 4,      /* call order violated (e.g. assemble before coordinates)  */
  CGASM_ECUDA = 5,       /* CUDA runtime error (cgasm_last_error has the text)      */
  CGASM_ENCCL = 6,       /* NCCL error                                              */
  CGASM_ENODEVICE = 7    /* no usable sm_100 device                                 */
};

/* ---- field slots (cgasm_set_field) ---------------------------------------------
 * Names are the dummy arguments of construct_momentum_element_cg
 * (Momentum_CG.F90:1193-1203) and assemble_advection_diffusion_element_cg
 * (Advection_Diffusion_CG.F90:702-706). */
enum {
  CGASM_F_NU = 0,          /* vector: nonlinear (advecting) velocity; the tracer's      */
                           /*         `velocity` is the same NonlinearVelocity field    */
  CGASM_F_OLDU = 1,        /* vector: old velocity                                      */
  CGASM_F_DENSITY = 2,     /* scalar                                                    */
  CGASM_F_VISCOSITY = 3,   /* tensor                                                    */
  CGASM_F_BUOYANCY = 4,    /* scalar                                                    */
  CGASM_F_HB_DENSITY = 5,  /* scalar (subtract_out_reference_profile)                   */
  CGASM_F_GRAVITY = 6,     /* vector: gravity direction                                 */
  CGASM_F_ABSORPTION = 7,  /* vector                                                    */
  CGASM_F_SOURCE = 8,      /* vector                                                    */
  CGASM_F_T = 9,           /* scalar: tracer                                            */
  CGASM_F_T_DIFFUSIVITY = 10, /* tensor                                                 */
  CGASM_F_T_SOURCE = 11,   /* scalar                                                    */
  CGASM_F_T_ABSORPTION = 12,  /* scalar                                                 */
  CGASM_F_NSLOTS = 13
};

/* field_type, femtools/Fields_Data_Types.F90 (FIELD_TYPE_NORMAL / _CONSTANT) */
enum { CGASM_FIELD_NORMAL = 0, CGASM_FIELD_CONSTANT = 1 };

/* stabilisation_scheme, assemble/Upwind_Stabilisation.F90 + Momentum_CG.F90:118 */
enum { CGASM_STAB_NONE = 0, CGASM_STAB_STREAMLINE_UPWIND = 1, CGASM_STAB_SUPG = 2 };
/* nu_bar_scheme, assemble/Upwind_Stabilisation.F90:47-48 */
enum { CGASM_NU_BAR_OPTIMAL = 1, CGASM_NU_BAR_DOUBLY_ASYMPTOTIC = 2,
       CGASM_NU_BAR_CRITICAL_RULE = 3, CGASM_NU_BAR_UNITY = 4 };
/* viscosity / diffusivity tensor shape, femtools/Field_Options.F90:1062-1088 */
enum { CGASM_TENSOR_ISOTROPIC = 0, CGASM_TENSOR_DIAGONAL = 1, CGASM_TENSOR_FULL = 2 };

/* how local contributions reach the CSR values (all give the same sums up to FP64
 * summation order; north-star asks for the variants to be compared) */
enum {
  CGASM_SCATTER_ATOMIC = 0,   /* one thread per element, red.global.add.f64             */
  CGASM_SCATTER_COLOURED = 1, /* femtools/Colouring.F90 colours, plain stores           */
  CGASM_SCATTER_WARPAGG = 2,  /* warp-aggregated atomics (match.any on the slot)        */
  CGASM_SCATTER_TILED = 3,    /* node-tile owner-computes, shared-memory accumulate,    */
                              /* every CSR value written exactly once                   */
  CGASM_SCATTER_GATHER = 4,   /* two passes: element kernel streams local rows to a     */
                              /* staging buffer, row kernel gathers them (no atomics,   */
                              /* no colouring, no redundant element math)               */
  CGASM_SCATTER_STRIP = 5     /* row owner, elements streamed as a strip around the     */
                              /* node (one node load per entry, register accumulators,  */
                              /* rhs as one sparse dot per row); option sets outside    */
                              /* the common one run the GATHER kernels                  */
};

/* ---- option structs: the module-level switches read once per assembly ----------
 * Momentum_CG.F90:83-178 (declarations), :335-669 (population). int = Fortran logical. */
typedef struct cgasm_momentum_opts {
  double dt;
  double theta;
  double beta;                 /* conservative_advection                             */
  double gravity_magnitude;
  double nu_bar_scale;
  double fs_sf;                /* get_surface_stab_scale_factor(u), Momentum_CG.F90:782 (surface loop) */
  int lump_mass;
  int exclude_mass;
  int exclude_advection;
  int integrate_advection_by_parts;
  int have_source;
  int lump_source;
  int have_gravity;
  int subtract_out_reference_profile;
  int have_absorption;
  int lump_absorption;
  int pressure_corrected_absorption;
  int have_viscosity;
  int viscosity_shape;         /* CGASM_TENSOR_*                                     */
  int assemble_inverse_masslump;
  int assemble_ct_matrix_here;
  int stabilisation_scheme;    /* CGASM_STAB_*                                       */
  int nu_bar_scheme;           /* CGASM_NU_BAR_*                                     */
  /* switches the device path does NOT implement; any non-zero => CGASM_EUNSUPPORTED
   * (Momentum_CG.F90 names): */
  int have_les, multiphase, on_sphere, move_mesh, have_coriolis,
      have_geostrophic_pressure, have_surfacetension, have_vertical_stabilization,
      have_swe_bottom_drag, have_wd_abs, have_temperature_dependent_viscosity,
      stress_form, partial_stress_form, radial_gravity, vel_lump_on_submesh,
      cmc_lump_on_submesh, abs_lump_on_submesh;
  /* implemented since round 2 (they keep their place in the struct): the `mass` matrix (cgasm_momentum_mass_fetch) and
   * continuity by parts (volume form in the element loop, boundary blocks in cgasm_momentum_surface_dev) */
  int assemble_mass_matrix, integrate_continuity_by_parts;
  /* have_fs_stab(u): free-surface stabilisation of the surface loop (Momentum_CG.F90:1108-1178, scale fs_sf above) on
   * faces of type FREE_SURFACE: cgasm_momentum_surface_dev adds it to big_m, rhs and (pressure-corrected absorption
   * with lumped mass) masslump; needs the gravity direction field and have_gravity. on_sphere stays unsupported. */
  int have_surface_fs_stabilisation;
} cgasm_momentum_opts;

/* Advection_Diffusion_CG.F90:77-123 (declarations), :384-556 (population). */
typedef struct cgasm_advdiff_opts {
  double dt;
  double theta;
  double beta;
  double nu_bar_scale;
  int have_mass;
  int lump_mass;
  int have_advection;
  int integrate_advection_by_parts;
  int have_source;             /* and not add_src_directly_to_rhs                     */
  int have_absorption;
  int have_diffusivity;
  int diffusivity_shape;       /* CGASM_TENSOR_ISOTROPIC or CGASM_TENSOR_FULL         */
  int stabilisation_scheme;
  int nu_bar_scheme;
  /* unsupported on the device path => CGASM_EUNSUPPORTED */
  int move_mesh, multiphase, equation_type_not_advdiff;
} cgasm_advdiff_opts;

/* ---- lifecycle ------------------------------------------------------------------- */

/* Copies the P1 simplex mesh and the reference-element tables to the device.
 * ndglno: mesh%ndglno, loc*n_elements, 1-based (femtools/Fields_Data_Types.F90:59-100).
 * n(loc,ngi), dn(loc,ngi,dim), weight(ngi): element_type tables, column-major
 * (femtools/Elements.F90:38-65). Only loc == dim+1 (P1) with dim in {2,3} is accepted.
 * device < 0 selects the current CUDA device. */
int cgasm_create(int* id, int device, int dim, int loc, int ngi, int n_nodes, int n_elements,
                 const int* ndglno, const double* n, const double* dn, const double* weight);

/* Frees every device and host resource of the handle. */
int cgasm_destroy(int id);

/* Text of the last error raised on this host thread (never NULL). */
const char* cgasm_last_error(void);

/* (Re)uploads Coordinate%val(dim, n_nodes); call again when X%refcount%id or
 * EVENT_MESH_MOVEMENT changes (femtools/Transform_elements.F90:131-168). */
int cgasm_set_coordinates(int id, const double* X);

/* ---- sparsity (femtools/Sparsity_Patterns.F90:47-85,299-428) ----------------------- */

/* Builds the first-order node-node sparsity of the mesh on its own. *nnz out. */
int cgasm_build_sparsity(int id, int* nnz);
/* Copies it out for the bit-exact comparison with the reference: findrm(n_nodes+1),
 * colm(nnz), centrm(n_nodes), all 1-based (centrm as lists2csr_sparsity, :412-426). */
int cgasm_get_sparsity(int id, int* findrm, int* colm, int* centrm);
/* Alternatively adopt the reference-built pattern (must have sorted rows). */
int cgasm_set_sparsity(int id, int rows, int nnz, const int* findrm, const int* colm);

/* ---- colouring (femtools/Colouring.F90:85-199,250-262) ---------------------------- */

/* Greedy CG1 element colouring built internally (used by CGASM_SCATTER_COLOURED). */
int cgasm_build_colouring(int id, int* ncolours);
/* colour_ptr(ncolours+1) (1-based offsets into colour_elements), colour_elements(n_elements)
 * (1-based element ids ascending inside a colour == fetch(colours(clr), nnid)). */
int cgasm_get_colouring(int id, int* colour_ptr, int* colour_elements);
int cgasm_set_colouring(int id, int ncolours, const int* colour_ptr, const int* colour_elements);

/* Selects CGASM_SCATTER_*; builds whatever plan that variant needs. */
int cgasm_set_scatter(int id, int variant);

/* ---- fields ------------------------------------------------------------------------- */

/* rank 0/1/2 = scalar/vector/tensor; field_type CGASM_FIELD_*; n_val_nodes = n_nodes for a
 * NORMAL field, 1 for a CONSTANT one (preprocessor/Populate_State.F90:1678-1724). */
int cgasm_set_field(int id, int slot, int rank, int field_type, const double* val,
                    int n_val_nodes);
/* Reads a resident field back (used after cgasm_halo_update). */
int cgasm_get_field(int id, int slot, double* val, int n_val_nodes);

/* ---- the two element loops --------------------------------------------------------------
 * Outputs are OVERWRITTEN with the assembled sums (the reference zeroes them just before
 * the loop: Momentum_Equation.F90:593-606, Advection_Diffusion_CG.F90:560-561).
 *
 * big_m: dim diagonal blocks, block d (0-based) at big_m[d*nnz .. (d+1)*nnz), entries in the
 *        order of colm -- the val(d,d)%ptr arrays of a block_csr_matrix
 *        (femtools/Sparse_Tools.F90:117-155); the Fortran shim inserts them into the
 *        petsc_csr_matrix row-wise.
 * rhs(dim, n_nodes); masslump(dim, n_nodes) or NULL; ct_m: dim blocks (1,d) of nnz, or NULL.
 */
int cgasm_momentum(int id, const cgasm_momentum_opts* opts, double* big_m, double* rhs,
                   double* masslump, double* ct_m);
int cgasm_advdiff(int id, const cgasm_advdiff_opts* opts, double* matrix_val, double* rhs);

/* Device-resident flavour: assemble into the handle's own device buffers and return without
 * copying (this is what stays on the GPU between the loop and a device-side consumer).
 * The *_fetch calls copy the last result to host buffers. */
int cgasm_momentum_dev(int id, const cgasm_momentum_opts* opts);
int cgasm_advdiff_dev(int id, const cgasm_advdiff_opts* opts);
/* Both element loops of a time step in one call (a tracer advected by the same nonlinear velocity the momentum
 * equation uses). When both option sets are the common ones of the STRIP variant -- lumped or excluded momentum mass,
 * plain advection, constant isotropic viscosity / diffusivity, constant gravity direction, no absorption or sources,
 * no ct_m -- ONE kernel assembles both systems (shared strip, staged node records and element geometry); any other
 * combination runs cgasm_momentum_dev then cgasm_advdiff_dev. The results are those of the two calls; fetch them
 * with cgasm_momentum_fetch / cgasm_advdiff_fetch. cgasm_last_kernel_ms afterwards covers the whole call. */
int cgasm_momentum_advdiff_dev(int id, const cgasm_momentum_opts* mopts, const cgasm_advdiff_opts* aopts);
int cgasm_momentum_fetch(int id, double* big_m, double* rhs, double* masslump, double* ct_m);
int cgasm_advdiff_fetch(int id, double* matrix_val, double* rhs);
/* 1 if the dim diagonal blocks of the last momentum result are identical (no absorption term:
 * mass, advection and tensor-form viscosity add the same loc x loc matrix to every (d,d) block,
 * Momentum_CG.F90:1550,1711,2312-2317), so one block can be fetched and inserted dim times. */
/* The `mass` matrix of construct_momentum_cg when opts.assemble_mass_matrix = 1 (Momentum_CG.F90:1567-1571: the
 * density-weighted consistent mass matrix on every diagonal block, plus dt*theta*absorption_mat with
 * pressure_corrected_absorption, :2073-2078), as [dim][nnz] diagonal blocks in colm order like big_m. */
int cgasm_momentum_mass_fetch(int id, double* mass);
int cgasm_momentum_mass_dev(int id, double** mass_dev);
int cgasm_momentum_identical_blocks(int id, int* identical);
/* Copies blocks first_block .. first_block+nblocks-1 (0-based) of big_m to the host. */
int cgasm_momentum_fetch_blocks(int id, int first_block, int nblocks, double* big_m);
/* Raw device pointers of the last result (NULL if that output was not assembled). */
int cgasm_momentum_result_dev(int id, double** big_m_dev, double** rhs_dev,
                              double** masslump_dev, double** ct_m_dev);
int cgasm_advdiff_result_dev(int id, double** matrix_dev, double** rhs_dev);

/* Single element, for element-matrix parity checks: big_m_tensor_addto(dim,dim,loc,loc) with
 * the lumped diagonal already folded in (Momentum_CG.F90:1462), rhs_addto(dim,loc),
 * mass_lump(loc), grad_p_u_mat(dim,loc,loc); tracer matrix_addto(loc,loc), rhs_addto(loc).
 * ele is 1-based. Computed on the device by the same code as the assembly kernels. */
int cgasm_momentum_element(int id, const cgasm_momentum_opts* opts, int ele,
                           double* big_m_tensor_addto, double* rhs_addto, double* mass_lump,
                           double* grad_p_u_mat);
int cgasm_advdiff_element(int id, const cgasm_advdiff_opts* opts, int ele,
                          double* matrix_addto, double* rhs_addto);

/* ---- surface-element loops and strong Dirichlet conditions: the step right after the element loops ----
 * (SURVEY.md 8(f) #1). They ADD to the device-resident result of the preceding cgasm_advdiff_dev /
 * cgasm_momentum_dev call, so a shim replaces
 *     element loop; face loop; apply_dirichlet_conditions          (Advection_Diffusion_CG.F90:569-645)
 * by  cgasm_advdiff_dev; cgasm_advdiff_surface_dev; cgasm_advdiff_dirichlet_dev; cgasm_advdiff_fetch.
 *
 * cgasm_set_surface: the boundary faces of the mesh (mesh%faces: surface_element_count, face_global_nodes,
 * face_ele, faces%shape; femtools/Fields_Base.F90:597-608,1317-1325,1917-1932). sndgln(sloc*n_faces) 1-based global nodes,
 * face_ele(n_faces) 1-based owning element; n_f(sloc,sngi), dn_f(sloc,sngi,dim-1), weight_f(sngi): the face
 * element's tables, column-major as in cgasm_create. P1 simplex faces only (sloc = dim, sngi <= 4). */
int cgasm_set_surface(int id, int n_faces, int sloc, int sngi, const int* sndgln, const int* face_ele,
                      const double* n_f, const double* dn_f, const double* weight_f);

/* Tracer boundary-condition types of assemble/Advection_Diffusion_CG.F90:74-75 */
enum {
  CGASM_TBC_NONE = 0,
  CGASM_TBC_NEUMANN = 1,
  CGASM_TBC_WEAKDIRICHLET = 2,
  CGASM_TBC_INTERNAL = 3,
  CGASM_TBC_ROBIN = 4
};
/* Face loop of assemble_advection_diffusion_cg (Advection_Diffusion_CG.F90:609-643 calling
 * assemble_advection_diffusion_face_cg :1228-1379): by-parts advection boundary term (weak Dirichlet or
 * free), Neumann and Robin conditions of the diffusive term. bc_type(n_faces); t_bc, t_bc_2(sloc, n_faces) =
 * ele_val of the "entire boundary condition" surface fields (either may be NULL if no face reads it). Does
 * nothing unless (integrate_advection_by_parts and have_advection) or have_diffusivity, like the reference.
 * CGASM_EUNSUPPORTED: a weak Dirichlet face with have_diffusivity (the reference FLExits, :1375). */
int cgasm_advdiff_surface_dev(int id, const cgasm_advdiff_opts* opts, const int* bc_type, const double* t_bc,
                              const double* t_bc_2);
/* apply_dirichlet_conditions for a scalar field (femtools/Boundary_Conditions.F90:1982-2024), one call per
 * boundary condition: rhs(nodes(j)) = (values(j) - T(nodes(j))) / dt if have_dt, else values(j). nodes are
 * 1-based. The matrix itself is untouched: the reference only flags the rows inactive (set_inactive), which
 * stays with the caller's csr_matrix. */
int cgasm_advdiff_dirichlet_dev(int id, int n, const int* nodes, const double* values, int have_dt, double dt);

/* Strong Dirichlet conditions of the velocity on big_m / rhs, device-resident (apply_dirichlet_conditions for a
 * petsc_csr_matrix: femtools/Boundary_Conditions.F90:2198-2218 = collect_vector_dirichlet_conditions :2125-2178 +
 * lift_boundary_conditions, femtools/Sparse_Tools_Petsc.F90:1139-1254 = MatZeroRowsColumns with pivot 1, then
 * fix_scaling). ONE call carries every strong condition: nodes(n) 1-based, comps(n) 1-based component, values(n) = what
 * collect_vector_dirichlet_conditions writes into rhs ((bc - u)/dt in acceleration form, else bc); a pair listed twice
 * takes the later value. For every listed (r, d): rhs(d, i) -= big_m_d(i, r) * value for the rows i not listed
 * themselves, row r and column r of block d are zeroed except the diagonal, which keeps its value, and
 * rhs(d, r) = diagonal * value. */
int cgasm_momentum_dirichlet_dev(int id, int n, const int* nodes, const int* comps, const double* values);

/* correct_masslumped_velocity (assemble/Momentum_CG.F90:2544-2575): u(d, :) += inverse_masslump(d, :) * (ct_m(1,d)^T
 * delta_p), every component. ct_m: host [dim][nnz] blocks or NULL = the ct_m left on the device by the last
 * cgasm_momentum_dev with assemble_ct_matrix_here; inverse_masslump host (dim, n_nodes) (the caller inverts masslump as
 * Momentum_Equation does); delta_p host (n_nodes); u host (dim, n_nodes), corrected in place. Same summation order as the
 * reference's mult_T. The halo_update(u) that follows stays with the caller. */
int cgasm_correct_masslumped_velocity(int id, const double* ct_m, const double* inverse_masslump, const double* delta_p,
                                      double* u);

/* Velocity boundary-condition types of assemble/Momentum_CG.F90:138-140 */
enum {
  CGASM_VBC_NONE = 0,
  CGASM_VBC_WEAKDIRICHLET = 1,
  CGASM_VBC_NO_NORMAL_FLOW = 2,
  CGASM_VBC_INTERNAL = 3,
  CGASM_VBC_FREE_SURFACE = 4,
  CGASM_VBC_FLUX = 5
};
/* surface_element_loop of construct_momentum_cg (Momentum_CG.F90:795-812 calling
 * construct_momentum_surface_element_cg :959-1191), the branches inside the device path's guard: by-parts
 * advection boundary term (:1029-1071) and flux conditions (:1180-1187). velocity_bc_type(dim, n_faces),
 * velocity_bc(dim, sloc, n_faces) = ele_val of the boundary-condition surface field (NULL if no face is weak
 * Dirichlet or flux), pressure_bc_type(n_faces) or NULL (only enters the skip rule :799-803). The caller keeps
 * the Fortran loop when have_fs_stab(u) (free-surface stabilisation) or integrate_continuity_by_parts. */
int cgasm_momentum_surface_dev(int id, const cgasm_momentum_opts* opts, const int* velocity_bc_type,
                               const double* velocity_bc, const int* pressure_bc_type);

/* ---- lumped-mass pressure matrix next to the momentum loop -------------------------
 * cmc_m = C_P^T M_L^-1 C: assemble_masslumped_cmc (assemble/Assemble_CMC.F90:119-135) ->
 * mult_div_vector_div_T (femtools/Sparse_Matrices_Fields.F90:590-671), for P1-P1 with ctp_m = ct_m (single
 * phase, not compressible). Values live on the SECOND-ORDER sparsity get_csr_sparsity_secondorder builds with
 * make_sparsity_mult (femtools/Sparsity_Patterns.F90:150-210): adopt the reference's pattern with
 * cgasm_cmc_set_sparsity(rows, nnz2, findrm, colm) (1-based, sorted rows) or let the library build the same
 * pattern (cgasm_cmc_build_sparsity; cgasm_cmc_get_sparsity returns it 1-based for the bit-exact check). */
int cgasm_cmc_build_sparsity(int id, long long* nnz2);
int cgasm_cmc_get_sparsity(int id, int* findrm2, int* colm2);
int cgasm_cmc_set_sparsity(int id, int rows, int nnz2, const int* findrm2, const int* colm2);
/* ct_m: host [dim][nnz] blocks on the first-order sparsity, or NULL = the ct_m left on the device by the last
 * cgasm_momentum_dev with assemble_ct_matrix_here. inverse_masslump: host (dim, n_nodes) -- what the caller
 * holds after invert() and apply_dirichlet_conditions_inverse_mass (Momentum_CG.F90:873-876) -- or NULL =
 * 1 / (lumped mass of the last cgasm_momentum_dev), i.e. no strong Dirichlet rows. */
int cgasm_cmc_dev(int id, const double* ct_m, const double* inverse_masslump);
int cgasm_cmc_fetch(int id, double* cmc_val);            /* nnz2 values in the order of the second-order colm */
int cgasm_cmc_result_dev(int id, double** cmc_val_dev);  /* raw device pointer of the last result */
/* P1-P1 pressure stabilisation on the same second-order sparsity: assemble_kmk_matrix
 * (assemble/Momentum_CG.F90:2707-2766). kt = sum_e 0.5 dshape_tensor_dshape(dp, h_bar, dp, detwei) with h_bar the
 * element's edge-length tensor (get_edge_lengths -> simplex_tensor, femtools/Metric_tools.F90:852-941), then
 * kmk = kt diag(1 / (theta_pg p_masslump)) kt^T (mult_div_invscalar_div_T, femtools/Sparse_Matrices_Fields.F90:673-748)
 * with p_masslump = get_lumped_mass(pressure mesh). Needs coordinates, the first-order sparsity (the pressure mesh is
 * the P1 mesh of the handle) and the second-order one. cgasm_kmk_fetch: kmk (nnz2 values, colm2 order), and if not
 * NULL kt (nnz) and p_masslump (n_nodes). add_kmk_matrix / add_kmk_rhs (:2768-2790) stay with the caller. */
int cgasm_kmk_dev(int id, double theta_pg);
int cgasm_kmk_fetch(int id, double* kmk, double* kt, double* p_masslump);
int cgasm_kmk_result_dev(int id, double** kmk_dev);
/* Diagnostics (host only, no GPU): the second-order pattern of a first-order one (both 1-based). findrm2
 * (n_nodes+1) is always written; colm2 only if *needed <= capacity. */
int cgasm_cmc_sparsity_host(int n_nodes, const int* findrm, const int* colm, int* findrm2, int* colm2,
                            long long capacity, long long* needed);
/* Diagnostics (host only, no GPU): the plan of the expansion kernel for 1-based first- and second-order patterns.
 * tpos(nnz): 0-based position of the transposed entry of every first-order entry; pptr(n_nodes+1): offsets into
 * slots; slots: for row i, for every column k of row i in order, for every entry (k, j) of row k in order, the 0-based
 * position of j in second-order row i. *needed = number of slots, or -1 if the patterns admit no plan (the merge
 * kernel runs). slots are written only if *needed <= capacity. */
int cgasm_cmc_expand_plan_host(int n_nodes, const int* findrm, const int* colm, const int* findrm2, const int* colm2,
                               int* tpos, long long* pptr, unsigned short* slots, long long capacity, long long* needed,
                               int* n2max);

/* Asynchronous host flavour. With on != 0: cgasm_set_field returns once the upload is queued (val must
 * stay valid, ideally pinned, until cgasm_synchronize) and the *_fetch calls queue their device -> host
 * copies on a second stream behind the result and return at once, so the next element loop and the
 * uploads of its fields overlap with the download of the previous result (PCIe is full duplex). The
 * host buffers are valid after cgasm_synchronize. A new cgasm_momentum_dev / cgasm_advdiff_dev waits on
 * the device for a pending download of the buffers it overwrites. Default: off (every call blocks). */
int cgasm_set_async(int id, int on);

/* Blocks until everything queued on the handle's streams has finished. */
int cgasm_synchronize(int id);
/* The handle's cudaStream_t (as void*), so callers can record events on it. */
int cgasm_stream(int id, void** stream);
/* Number of CUDA kernels this library has launched on the handle so far. */
int cgasm_launch_count(int id, long long* launches);
/* Which kernel family the most recent cgasm_momentum_dev / cgasm_advdiff_dev ran (diagnostics: the option set
 * decides whether a scatter variant's own kernels apply or a slower general path of the same variant). */
enum cgasm_path {
  CGASM_PATH_NONE = 0,
  CGASM_PATH_ELEMENT = 1,       /* thread per element: ATOMIC / COLOURED / WARPAGG */
  CGASM_PATH_TILED = 2,
  CGASM_PATH_GATHER_STAGED = 3, /* two passes: element kernel -> row records -> row kernel */
  CGASM_PATH_GATHER_ROWS = 4,   /* single-pass row kernels (direct / walk) */
  CGASM_PATH_STRIP = 5,         /* strip kernels, node records fetched per entry */
  CGASM_PATH_STRIP_STAGED = 6   /* strip kernels, node records staged in shared memory (+ additive passes) */
};
int cgasm_last_path(int id, int* momentum_path, int* advdiff_path);
/* Diagnostics: shape of the row-block plan behind CGASM_SCATTER_GATHER / STRIP (after cgasm_set_scatter).
 * stats(8) = row blocks, rows per block, longest CSR row, strip entries per (row, element) pair, staged STRIP plan
 * usable (0/1), largest number of distinct nodes a block touches, staged node capacity per block (NL), bytes of
 * shared memory per block of the staged momentum kernel. */
int cgasm_plan_stats(int id, double* stats);
/* Device time in ms of the most recent cgasm_*_dev call (CUDA events on the handle stream). */
int cgasm_last_kernel_ms(int id, float* ms);

/* Diagnostics (host only, no GPU): the strip ordering CGASM_SCATTER_STRIP uses for every row of
 * a P1 simplex mesh. ndglno as in cgasm_create. row_ptr(n_nodes+1): 0-based offsets; entries:
 * pairs {node (1-based), meta} with meta bits 0-7 = 0-based CSR slot of the node inside the row,
 * bit 8 = "the last loc-1 pushed nodes plus the row node form an element: compute it now".
 * At most `capacity` pairs are written; *needed returns the total. */
int cgasm_strip_plan_host(int loc, int n_nodes, int n_elements, const int* ndglno,
                          long long* row_ptr, int* entries, long long capacity, long long* needed);

/* Diagnostics (host only, no GPU): the row blocks the GATHER / STRIP variants work on. Nodes are quantised
 * to a lattice as fine as the mesh (lattice_scale(dim) = cells per unit length, out, may be NULL), ordered
 * along the Morton curve and cut into blocks of at most block_rows rows (bricks of the lattice where the
 * mesh is structured). rows(nblocks*block_rows): 1-based node of every row slot, 0 = padding; written only
 * if *nblocks <= capacity_blocks. X(dim, n_nodes) and ndglno as in cgasm_create / cgasm_set_coordinates. */
int cgasm_row_blocks_host(int dim, int n_nodes, int n_elements, const int* ndglno, const double* X,
                          int block_rows, int* rows, int capacity_blocks, int* nblocks,
                          double* lattice_scale);

/* ---- device hand-off to PETSc (femtools/Sparse_Tools_Petsc.F90:848-879, femtools/Petsc_Tools.F90:141-306) --------
 * The assembled matrices as COO triplets in PETSc's universal numbering, on the device, in the layout
 * MatSetPreallocationCOO(A, ncoo, i, j) / MatSetValuesCOO(A, v, INSERT_VALUES) take (device pointers with PETSc's
 * CUDA matrix types): no host round trip of the values.
 *   which        CGASM_COO_MOMENTUM: the dim diagonal blocks of big_m, entries ordered [block d][CSR entry];
 *                CGASM_COO_TRACER: the tracer matrix.
 *   row_gnn2unn  petsc_numbering%gnn2unn(n_nodes, nfields) of the rows as Fluidity built it (column-major, 0-based
 *                universal numbers; nfields = dim for big_m, 1 for the tracer); -1 = masked (the rows of nodes this
 *                process does not own, Sparse_Tools_Petsc.F90:220-227, and ghost nodes, Petsc_Tools.F90:255-297).
 *   col_gnn2unn  the column numbering, NULL = the row numbering.
 *   compact      0: ncoo = nblocks * nnz, masked entries keep index -1 (PETSc ignores negative indices) and the values
 *                   are the assembly's own result buffer (cgasm_coo_values_dev returns it: zero copy);
 *                1: masked entries are removed; cgasm_coo_values_dev gathers the values with one kernel per call.
 * The pattern belongs to the current sparsity: cgasm_build_sparsity / cgasm_set_sparsity drop it. */
enum cgasm_coo_matrix { CGASM_COO_MOMENTUM = 0, CGASM_COO_TRACER = 1 };
int cgasm_coo_pattern_dev(int id, int which, const int* row_gnn2unn, const int* col_gnn2unn, int compact,
                          long long* ncoo, int** coo_i_dev, int** coo_j_dev);
/* Values of the most recent assembly in the order of the pattern (device pointer, valid until the next assembly). */
int cgasm_coo_values_dev(int id, int which, double** coo_v_dev);
/* The same triplets copied to the host (any of the three may be NULL): for callers without a CUDA-aware PETSc. */
int cgasm_coo_fetch(int id, int which, long long ncoo, int* coo_i, int* coo_j, double* coo_v);

/* Diagnostics (host only, no GPU): wall-clock seconds of the host phases of a handle's set-up on this mesh --
 * times(6) = connectivity conversion, node->element adjacency, sparsity, Morton order, row blocks, strip plans
 * (the staged STRIP plan); entries_per_pair (may be NULL) = strip entries per (row, element) pair. */
int cgasm_plan_host_timing(int dim, int n_nodes, int n_elements, const int* ndglno, const double* X,
                           double* times, double* entries_per_pair);

/* Diagnostics (host only, no GPU): shape of the staged STRIP plan of a mesh -- stats(5) = strip entries per (row, element)
 * pair, largest number of distinct nodes a row block touches, shared-memory wavefronts per quarter-warp read of the
 * staged node records (1 = bank-conflict free), entries a warp walks / entries its rows need, element computations a
 * warp executes / computations its rows need (SIMT efficiency of the row-owner loop on unstructured meshes). */
int cgasm_plan_host_stats(int dim, int n_nodes, int n_elements, const int* ndglno, const double* X, double* stats);

/* ---- halo update (femtools/Halos_Communications.F90:320-412,497-567) -------------------
 * nprocs neighbours; sends/recvs are the concatenated 1-based node lists of
 * halo%sends(p) / halo%receives(p), nsend/nrecv their lengths per process p = 0..nprocs-1
 * (entry for p == rank is 0). nccl_unique_id: the 128-byte ncclUniqueId made by rank 0. */
int cgasm_halo_create(int id, int nprocs, int rank, const int* nsend, const int* sends,
                      const int* nrecv, const int* recvs, const void* nccl_unique_id);
/* Fills 128 bytes with a fresh ncclUniqueId (call on one rank, broadcast by the host). */
int cgasm_nccl_unique_id(void* out128);
/* halo_update of the resident fields whose bit (1 << slot) is set, in place: one pack kernel, one NCCL group
 * (one message per neighbour carrying every field), one unpack kernel that also refreshes the packed node
 * records of the received nodes. */
int cgasm_halo_update(int id, unsigned slot_mask);
/* Overlap (Halos_Communications.F90 has none: halo_update blocks in MPI_Waitall; SURVEY.md section 5 asks for
 * it). on = 1: cgasm_halo_update queues the exchange on a stream of its own and returns; the next
 * cgasm_momentum_dev / cgasm_advdiff_dev with the STRIP variant first launches the row blocks that read no
 * received node, then makes the compute stream wait for the exchange and launches the rest. Every other call
 * that reads or writes resident fields waits for a pending exchange first. Results are identical. Default 0. */
int cgasm_halo_set_overlap(int id, int on);

#ifdef __cplusplus
}
#endif
#endif /* CGASM_H */
