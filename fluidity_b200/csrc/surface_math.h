// surface_math.h -- per-face arithmetic of the surface-element loops, written once for
// the device kernels in surface.cu. Plain C++ with no CUDA dependency beyond the CG_HD qualifier, so the same
// functions also compile with g++ into the CPU harness of tests/test_surface_math.py, which checks them
// against the oracle without a GPU (the harness is test code; the product only ever runs them on the device).
//
// P1 simplex faces: sloc = dim nodes, the face quadrature has sngi <= 4 points. All face integrals are
// contractions sum_g a_ig b_jg c_g over the face's own tables (mesh%faces%shape), as in femtools/FETools.F90.
#pragma once
#include <math.h>

#include "../../include/cgasm.h"

#if defined(__CUDACC__)
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD inline
#endif

namespace cgasm {

constexpr int kMaxSloc = 3, kMaxSngi = 4;

struct SurfTables {  // faces%shape: n(sloc,sngi), dn(sloc,sngi,dim-1), quadrature%weight(sngi), column-major
  int sloc, sngi;
  double n[kMaxSloc * kMaxSngi];
  double dn[kMaxSloc * kMaxSngi * 2];
  double w[kMaxSngi];
};

// transform_facet_to_physical_full for a linear simplex facet (femtools/Transform_elements.F90:1353-1525):
// J = X_f . dn(:,1,:) at the first quadrature point only, detJ = facet measure / reference measure,
// normal = facet_normal(J, facet centroid - element centroid) (:1445, :1527-1553). Xf[i][a]: position of face
// node i; Xc: centroid of the owning element.
template <int DIM>
CG_HD void facet_geometry(const SurfTables& t, const double (&Xf)[DIM][DIM], const double (&Xc)[DIM], double& detJ,
                          double (&nrm)[DIM]) {
  double J[DIM][DIM - 1];
  for (int k = 0; k < DIM - 1; k++)
    for (int a = 0; a < DIM; a++) {
      double s = 0.0;
      for (int i = 0; i < DIM; i++) s += Xf[i][a] * t.dn[i + DIM * (0 + t.sngi * k)];
      J[a][k] = s;
    }
  if constexpr (DIM == 2) {
    detJ = sqrt(J[0][0] * J[0][0] + J[1][0] * J[1][0]);
    nrm[0] = -J[1][0];
    nrm[1] = J[0][0];
  } else {
    // cross_product(J(:,1), J(:,2)); detJ is its length (:1484-1489)
    nrm[0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
    nrm[1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
    nrm[2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
    detJ = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
  }
  double dotp = 0.0;
  for (int a = 0; a < DIM; a++) {
    double cf = 0.0;
    for (int i = 0; i < DIM; i++) cf += Xf[i][a];
    dotp += nrm[a] * (cf / DIM - Xc[a]);
  }
  double nn = 0.0;
  for (int a = 0; a < DIM; a++) {
    nrm[a] *= dotp;
    nn += nrm[a] * nrm[a];
  }
  nn = sqrt(nn);
  for (int a = 0; a < DIM; a++) nrm[a] /= nn;
}

// shape_shape(face_shape, face_shape, c) (FETools.F90:206-226): M_ij = sum_g N_ig N_jg c_g
template <int DIM>
CG_HD void face_shape_shape(const SurfTables& t, const double* c, double (&M)[DIM][DIM]) {
  for (int i = 0; i < DIM; i++)
    for (int j = 0; j < DIM; j++) {
      double s = 0.0;
      for (int g = 0; g < t.sngi; g++) s += (t.n[i + DIM * g] * t.n[j + DIM * g]) * c[g];
      M[i][j] = s;
    }
}
// face_val_at_quad (Fields_Base.F90:2374-2400): q_g = sum_i v_i N_ig
template <int DIM>
CG_HD double face_at_quad(const SurfTables& t, const double (&v)[DIM], int g) {
  double s = 0.0;
  for (int i = 0; i < DIM; i++) s += v[i] * t.n[i + DIM * g];
  return s;
}

// assemble_advection_diffusion_face_cg (assemble/Advection_Diffusion_CG.F90:1228-1379), default equation
// type, static mesh. Tf: T at the face nodes; Uf[i][a]: velocity; bc, bc2: ele_val(t_bc), ele_val(t_bc_2).
// A[i][j], r[i] are overwritten. The caller has excluded INTERNAL faces and weak Dirichlet + diffusivity.
template <int DIM>
CG_HD void advdiff_face(const SurfTables& t, const cgasm_advdiff_opts& o, int bc_type, const double (&Xf)[DIM][DIM],
                        const double (&Xc)[DIM], const double (&Tf)[DIM], const double (&Uf)[DIM][DIM],
                        const double (&bc)[DIM], const double (&bc2)[DIM], double (&A)[DIM][DIM], double (&r)[DIM]) {
  const double dt_theta = o.dt * o.theta;
  const bool implicit = fabs(dt_theta) > 2.220446049250313e-16;  // epsilon(0.0), :1329,1365
  for (int i = 0; i < DIM; i++) {
    r[i] = 0.0;
    for (int j = 0; j < DIM; j++) A[i][j] = 0.0;
  }
  const bool by_parts = o.have_advection && o.integrate_advection_by_parts;
  const bool flux_bc = o.have_diffusivity && (bc_type == CGASM_TBC_NEUMANN || bc_type == CGASM_TBC_ROBIN);
  if (!by_parts && !flux_bc) return;
  double detJ, nrm[DIM], c[kMaxSngi], M[DIM][DIM];
  facet_geometry<DIM>(t, Xf, Xc, detJ, nrm);
  if (by_parts) {  // add_advection_face_cg :1285-1340
    double un[DIM];
    for (int i = 0; i < DIM; i++) {
      un[i] = 0.0;
      for (int a = 0; a < DIM; a++) un[i] += Uf[i][a] * nrm[a];
    }
    for (int g = 0; g < t.sngi; g++) c[g] = detJ * t.w[g] * face_at_quad<DIM>(t, un, g);
    face_shape_shape<DIM>(t, c, M);
    for (int i = 0; i < DIM; i++) {
      double mt = 0.0, mb = 0.0;
      for (int j = 0; j < DIM; j++) {
        mt += M[i][j] * Tf[j];
        mb += M[i][j] * (bc[j] - Tf[j]);
      }
      if (implicit) {
        if (bc_type == CGASM_TBC_WEAKDIRICHLET) r[i] -= o.theta * mb;
        else
          for (int j = 0; j < DIM; j++) A[i][j] += dt_theta * M[i][j];
      }
      r[i] -= mt;
    }
  }
  if (flux_bc) {  // add_diffusivity_face_cg :1342-1379
    for (int i = 0; i < DIM; i++) {
      double s = 0.0;
      for (int g = 0; g < t.sngi; g++) s += t.n[i + DIM * g] * (detJ * t.w[g] * face_at_quad<DIM>(t, bc, g));
      r[i] += s;
    }
    if (bc_type == CGASM_TBC_ROBIN) {
      for (int g = 0; g < t.sngi; g++) c[g] = detJ * t.w[g] * face_at_quad<DIM>(t, bc2, g);
      face_shape_shape<DIM>(t, c, M);
      for (int i = 0; i < DIM; i++) {
        double mt = 0.0;
        for (int j = 0; j < DIM; j++) {
          mt += M[i][j] * Tf[j];
          if (implicit) A[i][j] += dt_theta * M[i][j];
        }
        r[i] -= mt;
      }
    }
  }
}

// Skip rule of the momentum surface loop (assemble/Momentum_CG.F90:799-803)
template <int DIM>
CG_HD bool momentum_face_skipped(const int (&bt)[DIM], int pressure_bc_type) {
  int sum = 0;
  bool internal = false;
  for (int d = 0; d < DIM; d++) {
    sum += bt[d];
    internal = internal || bt[d] == CGASM_VBC_INTERNAL;
  }
  return ((bt[0] == CGASM_VBC_NO_NORMAL_FLOW && sum == CGASM_VBC_NO_NORMAL_FLOW) || internal) && pressure_bc_type == 0;
}

// construct_momentum_surface_element_cg (assemble/Momentum_CG.F90:959-1191): by-parts advection boundary
// term :1029-1071 and flux conditions :1180-1187. Uf: nu, Of: oldu, rho: density at the face nodes;
// bc[i][d] = velocity_bc(d, i). B[d][i][j] (added to block (d,d)), r[d][i] overwritten.
// Free-surface stabilisation (:1108-1176, not on the sphere) on faces of type FREE_SURFACE: Gf = gravity direction at
// the face nodes, ml[d][i] = what goes to masslump (pressure-corrected absorption with lumped mass).
template <int DIM>
CG_HD void momentum_face(const SurfTables& t, const cgasm_momentum_opts& o, const int (&bt)[DIM],
                         const double (&Xf)[DIM][DIM], const double (&Xc)[DIM], const double (&Uf)[DIM][DIM],
                         const double (&Of)[DIM][DIM], const double (&rho)[DIM], const double (&bc)[DIM][DIM],
                         const double (&Gf)[DIM][DIM], double (&B)[DIM][DIM][DIM], double (&r)[DIM][DIM],
                         double (&ml)[DIM][DIM]) {
  for (int d = 0; d < DIM; d++)
    for (int i = 0; i < DIM; i++) {
      r[d][i] = ml[d][i] = 0.0;
      for (int j = 0; j < DIM; j++) B[d][i][j] = 0.0;
    }
  double detJ, nrm[DIM], c[kMaxSngi], M[DIM][DIM];
  facet_geometry<DIM>(t, Xf, Xc, detJ, nrm);
  if (bt[0] == CGASM_VBC_FREE_SURFACE && o.have_surface_fs_stabilisation) {
    // fs_surfacestab(d,i,j) = sum_g N_i N_j detwei rho_g dt g_mag fs_sf (n . up_g) up_g(d), up = -gravity direction
    const double dtt = o.dt * o.theta;
    double fs[DIM][DIM][DIM];
    for (int d = 0; d < DIM; d++)
      for (int i = 0; i < DIM; i++)
        for (int j = 0; j < DIM; j++) fs[d][i][j] = 0.0;
    for (int g = 0; g < t.sngi; g++) {
      double up[DIM], nk = 0.0;
      for (int a = 0; a < DIM; a++) {
        double gcol[DIM];
        for (int i = 0; i < DIM; i++) gcol[i] = Gf[i][a];
        up[a] = -face_at_quad<DIM>(t, gcol, g);
        nk += nrm[a] * up[a];
      }
      const double cg = detJ * t.w[g] * face_at_quad<DIM>(t, rho, g);
      for (int d = 0; d < DIM; d++) {
        const double vec = o.dt * o.gravity_magnitude * (o.fs_sf * nk * up[d]);
        for (int i = 0; i < DIM; i++)
          for (int j = 0; j < DIM; j++) fs[d][i][j] += t.n[i + DIM * g] * t.n[j + DIM * g] * cg * vec;
      }
    }
    for (int d = 0; d < DIM; d++)
      for (int i = 0; i < DIM; i++) {
        if (o.lump_mass) {
          double l = 0.0;
          for (int j = 0; j < DIM; j++) l += fs[d][i][j];
          B[d][i][i] += dtt * l;
          r[d][i] -= l * Of[i][d];
          if (o.pressure_corrected_absorption) ml[d][i] += dtt * l;
        } else {
          double v = 0.0;
          for (int j = 0; j < DIM; j++) {
            B[d][i][j] += dtt * fs[d][i][j];
            v += fs[d][i][j] * Of[j][d];
          }
          r[d][i] -= v;
        }
      }
  }
  if (bt[0] != CGASM_VBC_NO_NORMAL_FLOW && o.integrate_advection_by_parts && !o.exclude_advection) {
    double un[DIM];
    for (int i = 0; i < DIM; i++) {
      un[i] = 0.0;
      for (int a = 0; a < DIM; a++) un[i] += Uf[i][a] * nrm[a];
    }
    for (int g = 0; g < t.sngi; g++) c[g] = detJ * t.w[g] * face_at_quad<DIM>(t, un, g) * face_at_quad<DIM>(t, rho, g);
    face_shape_shape<DIM>(t, c, M);
    const double dtt = o.dt * o.theta;
    for (int d = 0; d < DIM; d++)
      for (int i = 0; i < DIM; i++) {
        double s = 0.0;
        if (bt[d] == CGASM_VBC_WEAKDIRICHLET) {
          for (int j = 0; j < DIM; j++) s += M[i][j] * bc[j][d];
        } else {
          for (int j = 0; j < DIM; j++) {
            s += M[i][j] * Of[j][d];
            B[d][i][j] += dtt * M[i][j];
          }
        }
        r[d][i] -= s;
      }
  }
  for (int d = 0; d < DIM; d++)
    if (bt[d] == CGASM_VBC_FLUX) {
      double bd[DIM];
      for (int i = 0; i < DIM; i++) bd[i] = bc[i][d];
      for (int i = 0; i < DIM; i++) {
        double s = 0.0;
        for (int g = 0; g < t.sngi; g++) s += t.n[i + DIM * g] * (face_at_quad<DIM>(t, bd, g) * (detJ * t.w[g]));
        r[d][i] += s;
      }
    }
}

// The continuity half of construct_momentum_surface_element_cg with integrate_continuity_by_parts
// (assemble/Momentum_CG.F90:1073-1088): ct_mat_bdy = shape_shape_vector(p_shape, u_shape, detwei_bdy, normal_bdy),
// CB[d][i][j] = normal_d sum_g N_i N_j detwei (P1 faces: the normal is constant), added to block d of ct_m at
// (pressure node i, velocity node j) on faces that are neither no-normal-flow nor free-surface.
template <int DIM>
CG_HD void momentum_face_ct(const SurfTables& t, const double (&Xf)[DIM][DIM], const double (&Xc)[DIM],
                            double (&CB)[DIM][DIM][DIM]) {
  double detJ, nrm[DIM], c[kMaxSngi], M[DIM][DIM];
  facet_geometry<DIM>(t, Xf, Xc, detJ, nrm);
  for (int g = 0; g < t.sngi; g++) c[g] = detJ * t.w[g];
  face_shape_shape<DIM>(t, c, M);
  for (int d = 0; d < DIM; d++)
    for (int i = 0; i < DIM; i++)
      for (int j = 0; j < DIM; j++) CB[d][i][j] = nrm[d] * M[i][j];
}

// csr_sparsity_pos on a sorted row (femtools/Sparse_Tools.F90:2438-2497), 0-based: position of column j in
// row i, or -1
CG_HD int csr_pos0(const int* findrm, const int* colm, int i, int j) {
  int lo = findrm[i], hi = findrm[i + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int c = colm[mid];
    if (c == j) return mid;
    if (c < j) lo = mid + 1;
    else hi = mid - 1;
  }
  return -1;
}

}  // namespace cgasm
