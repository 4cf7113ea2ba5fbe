/*
 * cg_oracle.c -- CPU restatement of Fluidity's CG element-assembly hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (fluidity_b200/, include/) may call,
 * link or import this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * The reference's element loops (Fortran + PETSc) cannot be compiled in this image (no Fortran
 * compiler, no PETSc; only its C++ halo IO and fldecomp writer build, see oracle/Makefile `ref`),
 * so this is a line-for-line restatement in plain C of the reference algorithm,
 * keeping its loop nests and summation order (gi innermost in the FETools contractions,
 * elements in ascending order, (iloc,jloc) row-major scatter). Every function cites the
 * reference file:line it follows (paths relative to the Fluidity source tree).
 *
 * PARITY PINNING: (1) the reference's own unit tests pin the tables, the transform, the
 * colouring and the block addto (tests/test_oracle_golden.py re-runs those known answers
 * against this file). (2) Outputs of the reference itself: python/fluidity/state_types.py,
 * the reference's Python implementation of transform_to_physical / shape_shape /
 * shape_dshape / ele_val_at_quad / addto, imported unmodified by
 * tests/golden/make_pyref_golden.py, pins detwei, the physical gradients, the momentum mass
 * and lumped mass, the tracer mass matrix and grad_p_u_mat / ct_m, per element and
 * assembled (tests/golden/pyref_*.npz, tests/test_pyref_golden.py), and supplies every
 * ingredient of the common-option-set matrices (the `c_*` arrays). (3) The contractions of
 * the advection / viscosity / absorption / source / buoyancy terms exist only in the
 * Fortran, which cannot run here: for those the VALUES are "parity unpinned" against a real
 * Fortran run, beyond (2)'s ingredients, the closed forms and the independent numpy
 * evaluation in tests/test_oracle_closed_forms.py.
 *
 * Array conventions are the Fortran ones: column-major, 1-based node/element numbers in
 * the integer arrays (ndglno, findrm, colm), FP64 reals.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../include/cgasm.h" /* option structs + enums only */

#define MAXDIM 3
#define MAXLOC 4
#define MAXNGI 5

/* column-major helpers */
#define N_(i, g) n[(i) + loc * (g)]                      /* n(loc,ngi)        */
#define DN_(i, g, k) dn[(i) + loc * ((g) + ngi * (k))]     /* dn(loc,ngi,dim)   */
#define DS_(i, g, k) dshape[(i) + loc * ((g) + ngi * (k))] /* dshape(loc,ngi,dim) */

typedef struct orc_field {
  const double* val; /* scalar val(N), vector val(dim,N), tensor val(dim,dim,N) */
  int field_type;    /* CGASM_FIELD_NORMAL / CGASM_FIELD_CONSTANT               */
} orc_field;

typedef struct orc_mesh {
  int dim, loc, ngi, n_nodes, n_elements;
  const int* ndglno;    /* loc*n_elements, 1-based */
  const double* n;      /* n(loc,ngi) */
  const double* dn;     /* dn(loc,ngi,dim) */
  const double* weight; /* weight(ngi) */
  const double* X;      /* Coordinate%val(dim,n_nodes) */
} orc_mesh;

/* ------------------------------------------------------------------------------------
 * Tables
 * ------------------------------------------------------------------------------------ */

/* Degree-3 simplex quadrature. femtools/Quadrature.F90:690-708 (tet, 5 points),
 * :951-970 (triangle, 4 points); permutations :1723-1733 (tet_permutations 7,8) and
 * :1815-1830 (tri_permutations 4,5); expansion expand_quadrature_template :572-605:
 *   l(k+dk, j) = coords(permutation(j,k)),  weight(dk+1:dk+nperm) = generator weight.
 * l is (ngi, loc) column-major: l[g + ngi*j]. Returns ngi. */
int orc_quadrature_degree3(int dim, double* l, double* weight) {
  if (dim == 3) {
    const int ngi = 5;
    /* generator 1: permutation (1,1,1,1), coords (0.25), weight -2/15 */
    for (int j = 0; j < 4; j++) l[0 + ngi * j] = 0.25;
    weight[0] = -0.133333333333333333333333333333333;
    /* generator 2: coords(1)=1/6, coords(2)=1-3*coords(1); columns of tet_permutations(8):
       (1,1,1,2),(1,1,2,1),(1,2,1,1),(2,1,1,1) */
    double coords[2];
    coords[0] = 0.166666666666666666666666666666666;
    coords[1] = 1.0 - 3.0 * coords[0];
    static const int p8[4][4] = {{1, 1, 1, 2}, {1, 1, 2, 1}, {1, 2, 1, 1}, {2, 1, 1, 1}};
    for (int k = 0; k < 4; k++) {
      for (int j = 0; j < 4; j++) l[(k + 1) + ngi * j] = coords[p8[k][j] - 1];
      weight[k + 1] = 0.075;
    }
    return ngi;
  } else if (dim == 2) {
    const int ngi = 4;
    for (int j = 0; j < 3; j++) l[0 + ngi * j] = 0.333333333333333333333333333333333;
    weight[0] = -0.28125;
    double coords[2];
    coords[0] = 0.2;
    coords[1] = 1.0 - 2.0 * coords[0];
    static const int p5[3][3] = {{1, 1, 2}, {1, 2, 1}, {2, 1, 1}};
    for (int k = 0; k < 3; k++) {
      for (int j = 0; j < 3; j++) l[(k + 1) + ngi * j] = coords[p5[k][j] - 1];
      weight[k + 1] = 0.260416666666666666666666666666666;
    }
    return ngi;
  }
  return 0;
}

/* P1 Lagrange simplex shape tables. femtools/Elements.F90:511-513 (n = l for P1),
 * eval_dshape_simplex :615-663 with diffl4 = -1 :693-717 and the raw polynomials of
 * femtools/Shape_Functions.F90:230-235: local node i <-> barycentric coordinate i
 * (femtools/Element_Numbering.F90:389-450), so dn(i,g,k) = delta_ik for i <= dim and
 * dn(loc,g,k) = -1. */
void orc_shape_p1(int dim, int ngi, const double* l, double* n, double* dn) {
  const int loc = dim + 1;
  for (int g = 0; g < ngi; g++) {
    for (int i = 0; i < loc; i++) {
      N_(i, g) = l[g + ngi * i];
      for (int k = 0; k < dim; k++) {
        double d;
        if (i < dim) d = (i == k) ? 1.0 : 0.0; /* dP_i/dL_k * P_loc(=1) */
        else d = -1.0;                         /* dl4dl(k) * dP_loc/dL_loc */
        DN_(i, g, k) = d;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * Gathers. femtools/Fields_Base.F90:1822-1832 (ele_nodes), :2047-2147 (ele_val, constant
 * fields broadcast node 1), :2256-2310 (ele_val_at_quad = matmul(ele_val, shape%n)),
 * :2557-2575 (ele_div_at_quad).
 * ------------------------------------------------------------------------------------ */
static inline const int* ele_nodes(const orc_mesh* m, int ele /*1-based*/) {
  return m->ndglno + (size_t)m->loc * (size_t)(ele - 1);
}

static void ele_val_scalar(const orc_mesh* m, const orc_field* f, int ele, double* out) {
  const int* nd = ele_nodes(m, ele);
  for (int i = 0; i < m->loc; i++)
    out[i] = (f->field_type == CGASM_FIELD_CONSTANT) ? f->val[0] : f->val[nd[i] - 1];
}

/* out(dim,loc) */
static void ele_val_vector(const orc_mesh* m, const orc_field* f, int ele, double* out) {
  const int dim = m->dim;
  const int* nd = ele_nodes(m, ele);
  for (int i = 0; i < m->loc; i++)
    for (int d = 0; d < dim; d++)
      out[d + dim * i] = (f->field_type == CGASM_FIELD_CONSTANT)
                             ? f->val[d]
                             : f->val[d + (size_t)dim * (size_t)(nd[i] - 1)];
}

/* out(dim,dim,loc) */
static void ele_val_tensor(const orc_mesh* m, const orc_field* f, int ele, double* out) {
  const int dd = m->dim * m->dim;
  const int* nd = ele_nodes(m, ele);
  for (int i = 0; i < m->loc; i++)
    for (int a = 0; a < dd; a++)
      out[a + dd * i] = (f->field_type == CGASM_FIELD_CONSTANT)
                            ? f->val[a]
                            : f->val[a + (size_t)dd * (size_t)(nd[i] - 1)];
}

/* quad(g) = sum_i val(i) n(i,g) */
static void at_quad_scalar(const orc_mesh* m, const double* ev, double* q) {
  const int loc = m->loc, ngi = m->ngi;
  const double* n = m->n;
  for (int g = 0; g < ngi; g++) {
    double s = 0.0;
    for (int i = 0; i < loc; i++) s += ev[i] * N_(i, g);
    q[g] = s;
  }
}
/* quad(c,g) = sum_i val(c,i) n(i,g), ncomp components */
static void at_quad_multi(const orc_mesh* m, int ncomp, const double* ev, double* q) {
  const int loc = m->loc, ngi = m->ngi;
  const double* n = m->n;
  for (int g = 0; g < ngi; g++)
    for (int c = 0; c < ncomp; c++) {
      double s = 0.0;
      for (int i = 0; i < loc; i++) s += ev[c + ncomp * i] * N_(i, g);
      q[c + ncomp * g] = s;
    }
}
/* Fields_Base.F90:2557-2575: quad_div = sum_d matmul(ele_val(field,d), dn(:,:,d)) */
static void div_at_quad(const orc_mesh* m, const double* ev /*(dim,loc)*/, const double* dshape,
                        double* q) {
  const int dim = m->dim, loc = m->loc, ngi = m->ngi;
  for (int g = 0; g < ngi; g++) q[g] = 0.0;
  for (int d = 0; d < dim; d++)
    for (int g = 0; g < ngi; g++) {
      double s = 0.0;
      for (int i = 0; i < loc; i++) s += ev[d + dim * i] * DS_(i, g, d);
      q[g] = q[g] + s;
    }
}

/* ------------------------------------------------------------------------------------
 * transform_to_physical_full, uncached branch: femtools/Transform_elements.F90:807-887
 * (linear simplex => Jacobian formed at gi==1 only, dshape copied to every gi).
 * X_val(dim,loc); dshape(loc,ngi,dim); detwei(ngi); J(dim,dim,ngi) optional.
 * ------------------------------------------------------------------------------------ */
static int cyc3(int i) { return ((i - 1) % 3) + 1; } /* femtools: cyc3(4)=1, cyc3(5)=2 */

void orc_transform_to_physical(int dim, int ngi, const double* X_val, const double* dn,
                               const double* weight, double* dshape, double* detwei,
                               double* J) {
  const int loc = dim + 1;
  double JT[MAXDIM * MAXDIM], invJ[MAXDIM * MAXDIM], detJ;
#define JT_(a, k) JT[(a) + dim * (k)]
#define IJ_(a, k) invJ[(a) + dim * (k)]
  /* J_local_T = matmul(X_val, dn(:,1,:))   (:828) */
  for (int a = 0; a < dim; a++)
    for (int k = 0; k < dim; k++) {
      double s = 0.0;
      for (int i = 0; i < loc; i++) s += X_val[a + dim * i] * DN_(i, 0, k);
      JT_(a, k) = s;
    }
  if (dim == 2) {
    /* reshape((/J22,-J12,-J21,J11/),(/2,2/))  (:839-840), column-major fill */
    IJ_(0, 0) = JT_(1, 1);
    IJ_(1, 0) = -JT_(0, 1);
    IJ_(0, 1) = -JT_(1, 0);
    IJ_(1, 1) = JT_(0, 0);
  } else {
    /* (:843-847) */
    for (int i = 1; i <= 3; i++)
      for (int k = 1; k <= 3; k++)
        IJ_(i - 1, k - 1) = JT_(cyc3(i + 1) - 1, cyc3(k + 1) - 1) * JT_(cyc3(i + 2) - 1, cyc3(k + 2) - 1) -
                            JT_(cyc3(i + 2) - 1, cyc3(k + 1) - 1) * JT_(cyc3(i + 1) - 1, cyc3(k + 2) - 1);
  }
  /* detJ = dot_product(J_local_T(:,1), invJ_local(:,1))  (:853) */
  detJ = 0.0;
  for (int a = 0; a < dim; a++) detJ += JT_(a, 0) * IJ_(a, 0);
  /* invJ = invJ/detJ (:856) */
  for (int a = 0; a < dim * dim; a++) invJ[a] = invJ[a] / detJ;
  /* dshape(i,1,:) = matmul(invJ, dn(i,1,:)) (:864-866); other gi copy gi=1 (:868) */
  for (int i = 0; i < loc; i++)
    for (int a = 0; a < dim; a++) {
      double s = 0.0;
      for (int k = 0; k < dim; k++) s += IJ_(a, k) * DN_(i, 0, k);
      DS_(i, 0, a) = s;
    }
  for (int g = 1; g < ngi; g++)
    for (int i = 0; i < loc; i++)
      for (int a = 0; a < dim; a++) DS_(i, g, a) = DS_(i, 0, a);
  /* detwei(gi) = abs(detJ)*weight(gi) (:873) */
  if (detwei)
    for (int g = 0; g < ngi; g++) detwei[g] = fabs(detJ) * weight[g];
  /* J(:,:,gi) = transpose(J_local_T) (:878-882) */
  if (J)
    for (int g = 0; g < ngi; g++)
      for (int a = 0; a < dim; a++)
        for (int k = 0; k < dim; k++) J[a + dim * (k + dim * g)] = JT_(k, a);
#undef JT_
#undef IJ_
}

/* ------------------------------------------------------------------------------------
 * FETools local integrals, femtools/FETools.F90 (line ranges at each function).
 * All R are (loc,loc) column-major R[i + loc*j] unless stated.
 * ------------------------------------------------------------------------------------ */
#define R_(i, j) R[(i) + loc * (j)]

/* shape_dshape :332-362 (non-INLINE_MATMUL branch):
 * R(1:dim,i,j) = matmul(detwei, spread(n(i,:),2,dim)*dshape(j,:,:)) */
static void shape_dshape(int dim, int loc, int ngi, const double* n, const double* dshape,
                         const double* detwei, double* R /*(dim,loc,loc)*/) {
  for (int j = 0; j < loc; j++)
    for (int i = 0; i < loc; i++)
      for (int d = 0; d < dim; d++) {
        double s = 0.0;
        for (int g = 0; g < ngi; g++) s += detwei[g] * (N_(i, g) * DS_(j, g, d));
        R[d + dim * (i + loc * j)] = s;
      }
}

/* dshape_dot_dshape :391-453: gi outer loop, R += (sum_d dsh(i,g,d)*dsh(j,g,d)) * detwei(g) */
static void dshape_dot_dshape(int dim, int loc, int ngi, const double* dshape, const double* detwei,
                              double* R) {
  for (int a = 0; a < loc * loc; a++) R[a] = 0.0;
  for (int g = 0; g < ngi; g++)
    for (int j = 0; j < loc; j++)
      for (int i = 0; i < loc; i++) {
        double t = DS_(i, g, 0) * DS_(j, g, 0);
        for (int d = 1; d < dim; d++) t = t + DS_(i, g, d) * DS_(j, g, d);
        R_(i, j) = R_(i, j) + t * detwei[g];
      }
}

/* dshape_diagtensor_dshape :551-610 */
static void dshape_diagtensor_dshape(int dim, int loc, int ngi, const double* dshape,
                                     const double* tensor /*(dim,dim,ngi)*/, const double* detwei,
                                     double* R) {
  for (int a = 0; a < loc * loc; a++) R[a] = 0.0;
  for (int g = 0; g < ngi; g++)
    for (int j = 0; j < loc; j++)
      for (int i = 0; i < loc; i++) {
        double t = DS_(i, g, 0) * tensor[0 + dim * (0 + dim * g)] * DS_(j, g, 0);
        for (int d = 1; d < dim; d++) t = t + DS_(i, g, d) * tensor[d + dim * (d + dim * g)] * DS_(j, g, d);
        R_(i, j) = R_(i, j) + t * detwei[g];
      }
}

/* dshape_tensor_dshape :668-698:
 * R(i,j) += dot_product(matmul(dshape(i,g,:), tensor(:,:,g)), dshape(j,g,:)) * detwei(g) */
static void dshape_tensor_dshape(int dim, int loc, int ngi, const double* dshape,
                                 const double* tensor, const double* detwei, double* R) {
  for (int a = 0; a < loc * loc; a++) R[a] = 0.0;
  for (int g = 0; g < ngi; g++)
    for (int j = 0; j < loc; j++)
      for (int i = 0; i < loc; i++) {
        double dot = 0.0;
        for (int b = 0; b < dim; b++) {
          double vb = 0.0; /* (dshape(i,g,:) . tensor(:,b,g)) */
          for (int a = 0; a < dim; a++) vb += DS_(i, g, a) * tensor[a + dim * (b + dim * g)];
          dot += vb * DS_(j, g, b);
        }
        R_(i, j) = R_(i, j) + dot * detwei[g];
      }
}

/* dshape_dot_vector_shape :700-721:
 * R(i,j) = dot_product(sum(dshape(i,:,:)*transpose(vector),2)*n(j,:), detwei) */
static void dshape_dot_vector_shape(int dim, int loc, int ngi, const double* dshape,
                                    const double* vector, const double* n, const double* detwei,
                                    double* R) {
  for (int j = 0; j < loc; j++)
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int g = 0; g < ngi; g++) {
        double t = 0.0;
        for (int d = 0; d < dim; d++) t += DS_(i, g, d) * vector[d + dim * g];
        s += (t * N_(j, g)) * detwei[g];
      }
      R_(i, j) = s;
    }
}

/* shape_vector_dot_dshape :749-772:
 * R(i,j) = dot_product(n(i,:)*sum(dshape(j,:,:)*transpose(vector),2), detwei) */
static void shape_vector_dot_dshape(int dim, int loc, int ngi, const double* n, const double* vector,
                                    const double* dshape, const double* detwei, double* R) {
  for (int j = 0; j < loc; j++)
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int g = 0; g < ngi; g++) {
        double t = 0.0;
        for (int d = 0; d < dim; d++) t += DS_(j, g, d) * vector[d + dim * g];
        s += (N_(i, g) * t) * detwei[g];
      }
      R_(i, j) = s;
    }
}

/* shape_rhs :38-54: matmul(n, detwei) */
static void shape_rhs(int loc, int ngi, const double* n, const double* detwei, double* r) {
  for (int i = 0; i < loc; i++) {
    double s = 0.0;
    for (int g = 0; g < ngi; g++) s += N_(i, g) * detwei[g];
    r[i] = s;
  }
}

/* shape_vector_rhs :56-81: r(d,:) = matmul(n, detwei*vector(d,:)) */
static void shape_vector_rhs(int dim, int loc, int ngi, const double* n, const double* vector,
                             const double* detwei, double* r /*(dim,loc)*/) {
  for (int d = 0; d < dim; d++)
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int g = 0; g < ngi; g++) s += N_(i, g) * (detwei[g] * vector[d + dim * g]);
      r[d + dim * i] = s;
    }
}

/* ------------------------------------------------------------------------------------
 * Upwind stabilisation, assemble/Upwind_Stabilisation.F90.
 * ------------------------------------------------------------------------------------ */
static const double su_tolerance = 1.0e-10;             /* :50 */
static const double su_tanh_tolerance = 11.859499013855018; /* :51 */

/* femtools inverse() of a dim x dim matrix (column-major), cofactor/det */
static void small_inverse(int dim, const double* A, double* B) {
  if (dim == 2) {
    double det = A[0] * A[3] - A[2] * A[1];
    B[0] = A[3] / det;
    B[1] = -A[1] / det;
    B[2] = -A[2] / det;
    B[3] = A[0] / det;
  } else {
#define A_(i, j) A[(i) + 3 * (j)]
    double c00 = A_(1, 1) * A_(2, 2) - A_(1, 2) * A_(2, 1);
    double c01 = A_(1, 2) * A_(2, 0) - A_(1, 0) * A_(2, 2);
    double c02 = A_(1, 0) * A_(2, 1) - A_(1, 1) * A_(2, 0);
    double det = A_(0, 0) * c00 + A_(0, 1) * c01 + A_(0, 2) * c02;
    B[0 + 3 * 0] = c00 / det;
    B[1 + 3 * 0] = c01 / det;
    B[2 + 3 * 0] = c02 / det;
    B[0 + 3 * 1] = (A_(0, 2) * A_(2, 1) - A_(0, 1) * A_(2, 2)) / det;
    B[1 + 3 * 1] = (A_(0, 0) * A_(2, 2) - A_(0, 2) * A_(2, 0)) / det;
    B[2 + 3 * 1] = (A_(0, 1) * A_(2, 0) - A_(0, 0) * A_(2, 1)) / det;
    B[0 + 3 * 2] = (A_(0, 1) * A_(1, 2) - A_(0, 2) * A_(1, 1)) / det;
    B[1 + 3 * 2] = (A_(0, 2) * A_(1, 0) - A_(0, 0) * A_(1, 2)) / det;
    B[2 + 3 * 2] = (A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(1, 0)) / det;
#undef A_
  }
}

/* nu_bar_scaled_q :225-320 with xi_optimal :133-162, xi_doubly_asymptotic :164-192,
 * xi_critical_rule :194-223. diff_q may be NULL (=> NU_BAR_UNITY, :248-251). */
static void nu_bar_scaled_q(int dim, int ngi, const double* u_q /*(dim,ngi)*/,
                            const double* j_mat /*(dim,dim,ngi)*/, const double* diff_q,
                            int nu_bar_scheme, double nu_bar_scale, double* out) {
  int scheme = diff_q ? nu_bar_scheme : CGASM_NU_BAR_UNITY;
  for (int g = 0; g < ngi; g++) {
    const double* u = u_q + dim * g;
    const double* Jg = j_mat + dim * dim * g;
    double norm_u = 0.0;
    for (int d = 0; d < dim; d++) norm_u += u[d] * u[d];
    if (norm_u < su_tolerance) {
      out[g] = 0.0;
      continue;
    }
    /* uJ = matmul(u, J): uJ(k) = sum_a u(a) J(a,k) */
    double uJ[MAXDIM];
    for (int k = 0; k < dim; k++) {
      double s = 0.0;
      for (int a = 0; a < dim; a++) s += u[a] * Jg[a + dim * k];
      uJ[k] = s;
    }
    double val = 0.0;
    if (scheme == CGASM_NU_BAR_UNITY) {
      for (int k = 0; k < dim; k++) val += fabs(uJ[k]);
    } else {
      /* pe = 0.5*matmul(u, matmul(J, inverse(diff))) */
      double inv[MAXDIM * MAXDIM], JD[MAXDIM * MAXDIM], pe[MAXDIM], xi[MAXDIM];
      small_inverse(dim, diff_q + dim * dim * g, inv);
      for (int a = 0; a < dim; a++)
        for (int k = 0; k < dim; k++) {
          double s = 0.0;
          for (int b = 0; b < dim; b++) s += Jg[a + dim * b] * inv[b + dim * k];
          JD[a + dim * k] = s;
        }
      for (int k = 0; k < dim; k++) {
        double s = 0.0;
        for (int a = 0; a < dim; a++) s += u[a] * JD[a + dim * k];
        pe[k] = 0.5 * s;
      }
      for (int k = 0; k < dim; k++) {
        double p = pe[k];
        if (scheme == CGASM_NU_BAR_OPTIMAL) {
          if (fabs(p) < su_tolerance) xi[k] = 0.0;
          else if (p > su_tanh_tolerance) xi[k] = 1.0 - (1.0 / p);
          else if (p < -su_tanh_tolerance) xi[k] = -1.0 - (1.0 / p);
          else xi[k] = (1.0 / tanh(p)) - (1.0 / p);
        } else if (scheme == CGASM_NU_BAR_DOUBLY_ASYMPTOTIC) {
          if (fabs(p) <= 3.0) xi[k] = p / 3.0;
          else if (p > 0.0) xi[k] = 1.0;
          else xi[k] = -1.0;
        } else { /* critical rule */
          if (fabs(p) <= 1.0) xi[k] = 0.0;
          else if (p > 0.0) xi[k] = 1.0 - 1.0 / p;
          else xi[k] = -1.0 - 1.0 / p;
        }
      }
      for (int k = 0; k < dim; k++) val += xi[k] * uJ[k];
    }
    out[g] = val / norm_u;
  }
  for (int g = 0; g < ngi; g++) out[g] = out[g] * nu_bar_scale;
}

/* element_upwind_stabilisation :82-131 */
static void element_upwind_stabilisation(int dim, int loc, int ngi, const double* dshape,
                                         const double* u_q, const double* j_mat,
                                         const double* detwei, const double* diff_q,
                                         int nu_bar_scheme, double nu_bar_scale, double* stab) {
  double nu_scaled[MAXNGI], udn[MAXLOC * MAXNGI];
  nu_bar_scaled_q(dim, ngi, u_q, j_mat, diff_q, nu_bar_scheme, nu_bar_scale, nu_scaled);
  for (int g = 0; g < ngi; g++)
    for (int j = 0; j < loc; j++) {
      double s = 0.0;
      for (int d = 0; d < dim; d++) s += u_q[d + dim * g] * DS_(j, g, d);
      udn[j + loc * g] = s;
    }
  for (int j = 0; j < loc; j++)
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int g = 0; g < ngi; g++) s += (udn[i + loc * g] * detwei[g] * nu_scaled[g]) * udn[j + loc * g];
      stab[i + loc * j] = s;
    }
}

/* supg_test_function :418-454 -> make_supg_element: test n(i,g) = n(i,g) +
 * nu_bar_scaled(g) * (u_g . dshape(i,g,:)). Writes the modified n table. */
static void supg_test_function(int dim, int loc, int ngi, const double* n, const double* dshape,
                               const double* u_q, const double* j_mat, const double* diff_q,
                               int nu_bar_scheme, double nu_bar_scale, double* n_test) {
  double nu_scaled[MAXNGI];
  nu_bar_scaled_q(dim, ngi, u_q, j_mat, diff_q, nu_bar_scheme, nu_bar_scale, nu_scaled);
  for (int g = 0; g < ngi; g++)
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int d = 0; d < dim; d++) s += u_q[d + dim * g] * DS_(i, g, d);
      n_test[i + loc * g] = N_(i, g) + nu_scaled[g] * s;
    }
}

/* shape_shape :206-226: R(i,j) = dot_product(n1(i,:)*n2(j,:), detwei); n1 = test function
 * table (differs from the trial table n2 only under SUPG) */
static void shape_shape2(int loc, int ngi, const double* n1, const double* n2, const double* detwei,
                         double* R) {
  for (int j = 0; j < loc; j++)
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int g = 0; g < ngi; g++) s += (n1[i + loc * g] * n2[j + loc * g]) * detwei[g];
      R_(i, j) = s;
    }
}
/* shape_shape_vector :228-254: R(d,i,j) = matmul(vector*spread(n1_i*n2_j), detwei) */
static void shape_shape_vector2(int dim, int loc, int ngi, const double* n1, const double* n2,
                                const double* detwei, const double* vector, double* R) {
  for (int j = 0; j < loc; j++)
    for (int i = 0; i < loc; i++)
      for (int d = 0; d < dim; d++) {
        double s = 0.0;
        for (int g = 0; g < ngi; g++)
          s += (vector[d + dim * g] * (n1[i + loc * g] * n2[j + loc * g])) * detwei[g];
        R[d + dim * (i + loc * j)] = s;
      }
}

/* ------------------------------------------------------------------------------------
 * construct_momentum_element_cg, assemble/Momentum_CG.F90:1193-1490 with
 * add_mass_element_cg :1492-1600, add_advection_element_cg :1602-1715,
 * add_sources_element_cg :1717-1751, add_buoyancy_element_cg :1753-1792,
 * add_absorption_element_cg :1818-1878 + :2036-2073 (plain branch),
 * add_viscosity_element_cg :2079-2133 + :2286-2359 (tensor-form branches).
 *
 * Outputs (all overwritten):
 *   big_m_tensor_addto(dim,dim,loc,loc) with the lumped diagonal folded in (:1462)
 *   rhs_addto(dim,loc)
 *   masslump_addto(dim,loc)   what is added to the masslump field (:1560-1565, :2063-2065)
 *   grad_p_u_mat(dim,loc,loc) (:1401)
 * Returns 0, or CGASM_EUNSUPPORTED for option branches not restated.
 * ------------------------------------------------------------------------------------ */
typedef struct orc_momentum_fields {
  orc_field nu, oldu, density, viscosity, buoyancy, hb_density, gravity, absorption, source;
} orc_momentum_fields;

static int momentum_opts_unsupported(const cgasm_momentum_opts* o) {
  return o->have_les || o->multiphase || o->on_sphere || o->move_mesh || o->have_coriolis ||
         o->have_geostrophic_pressure || o->have_surfacetension ||
         o->have_vertical_stabilization || o->have_swe_bottom_drag || o->have_wd_abs ||
         o->have_temperature_dependent_viscosity || o->stress_form || o->partial_stress_form ||
         o->radial_gravity || o->vel_lump_on_submesh || o->cmc_lump_on_submesh ||
         o->abs_lump_on_submesh; /* assemble_mass_matrix: the `mass` matrix is orc_assemble_momentum_mass below */
}

int orc_momentum_element(const orc_mesh* m, const orc_momentum_fields* f,
                         const cgasm_momentum_opts* o, int ele, double* big_m_tensor_addto,
                         double* rhs_addto, double* masslump_addto, double* grad_p_u_mat) {
  const int dim = m->dim, loc = m->loc, ngi = m->ngi;
  const double* n = m->n;
  if (momentum_opts_unsupported(o)) return CGASM_EUNSUPPORTED;

  double X_val[MAXDIM * MAXLOC], oldu_val[MAXDIM * MAXLOC];
  double du_t[MAXLOC * MAXNGI * MAXDIM], detwei[MAXNGI], J_mat[MAXDIM * MAXDIM * MAXNGI];
  double big_m_diag_addto[MAXDIM * MAXLOC];
  double n_test_buf[MAXLOC * MAXNGI];
  const double* test_n = n; /* test_function = u_shape (:1370) */
  const double* dshape = du_t;

#define T_(a, b, i, j) big_m_tensor_addto[(a) + dim * ((b) + dim * ((i) + loc * (j)))]
#define RH_(d, i) rhs_addto[(d) + dim * (i)]
#define DG_(d, i) big_m_diag_addto[(d) + dim * (i)]
#define OU_(d, i) oldu_val[(d) + dim * (i)]
  for (int a = 0; a < dim * dim * loc * loc; a++) big_m_tensor_addto[a] = 0.0;
  for (int a = 0; a < dim * loc; a++) {
    big_m_diag_addto[a] = 0.0;
    rhs_addto[a] = 0.0;
    masslump_addto[a] = 0.0;
  }
  if (grad_p_u_mat)
    for (int a = 0; a < dim * loc * loc; a++) grad_p_u_mat[a] = 0.0;

  /* oldu_val = ele_val(oldu, ele) (:1308) */
  ele_val_vector(m, &f->oldu, ele, oldu_val);

  /* Step 1: transform (:1313-1320) */
  {
    orc_field Xf = {m->X, CGASM_FIELD_NORMAL};
    ele_val_vector(m, &Xf, ele, X_val);
    orc_transform_to_physical(dim, ngi, X_val, m->dn, m->weight, du_t, detwei,
                              o->stabilisation_scheme == CGASM_STAB_NONE ? NULL : J_mat);
  }

  double density_gi[MAXNGI], ev_s[MAXLOC];
  ele_val_scalar(m, &f->density, ele, ev_s);
  at_quad_scalar(m, ev_s, density_gi);

  double nu_val[MAXDIM * MAXLOC], relu_gi[MAXDIM * MAXNGI];
  double visc_val[MAXDIM * MAXDIM * MAXLOC], viscosity_gi[MAXDIM * MAXDIM * MAXNGI];
  ele_val_vector(m, &f->nu, ele, nu_val);
  at_quad_multi(m, dim, nu_val, relu_gi);
  if (o->have_viscosity) {
    ele_val_tensor(m, &f->viscosity, ele, visc_val);
    at_quad_multi(m, dim * dim, visc_val, viscosity_gi);
  }

  /* Step 2: test function (:1345-1372) */
  if (o->stabilisation_scheme == CGASM_STAB_SUPG) {
    double diff_q[MAXDIM * MAXDIM * MAXNGI];
    if (o->have_viscosity) {
      memcpy(diff_q, viscosity_gi, sizeof(double) * dim * dim * ngi);
      for (int g = 0; g < ngi; g++)
        for (int a = 0; a < dim; a++)
          for (int b = 0; b < dim; b++)
            if (a != b) diff_q[a + dim * (b + dim * g)] = 0.0; /* :1355-1360 */
    }
    supg_test_function(dim, loc, ngi, n, dshape, relu_gi, J_mat, o->have_viscosity ? diff_q : NULL,
                       o->nu_bar_scheme, o->nu_bar_scale, n_test_buf);
    test_n = n_test_buf;
  }

  /* ct_m block (:1377-1404), P1-P1: p_shape == u_shape tables */
  if (o->assemble_ct_matrix_here && grad_p_u_mat) {
    if (o->integrate_continuity_by_parts) {
      /* :1377-1383: grad_p_u_mat = -dshape_shape(dp_t, u_shape, detwei), femtools/FETools.F90:364-389:
       * (d, i, j) = -sum_g detwei_g dshape(i,g,d) n(j,g). Restated ahead of the device path (which still
       * refuses the option): the boundary half is orc_momentum_face_ct below. */
      for (int j = 0; j < loc; j++)
        for (int i = 0; i < loc; i++)
          for (int d = 0; d < dim; d++) {
            double sacc = 0.0;
            for (int g = 0; g < ngi; g++) sacc += detwei[g] * (DS_(i, g, d) * N_(j, g));
            grad_p_u_mat[d + dim * (i + loc * j)] = -sacc;
          }
    } else {
      shape_dshape(dim, loc, ngi, n, dshape, detwei, grad_p_u_mat);
    }
  }

  /* Mass terms (:1411 -> :1492-1600) */
  if (o->assemble_inverse_masslump || !o->exclude_mass) {
    double coefficient_detwei[MAXNGI], mass_mat[MAXLOC * MAXLOC], mass_lump[MAXLOC];
    for (int g = 0; g < ngi; g++) coefficient_detwei[g] = density_gi[g] * detwei[g];
    shape_shape2(loc, ngi, test_n, n, coefficient_detwei, mass_mat);
    for (int i = 0; i < loc; i++) { /* mass_lump = sum(mass_mat, 2) */
      double s = 0.0;
      for (int j = 0; j < loc; j++) s += mass_mat[i + loc * j];
      mass_lump[i] = s;
    }
    if (!o->exclude_mass) {
      if (o->lump_mass) {
        for (int d = 0; d < dim; d++)
          for (int i = 0; i < loc; i++) DG_(d, i) = DG_(d, i) + mass_lump[i];
      } else {
        for (int d = 0; d < dim; d++)
          for (int j = 0; j < loc; j++)
            for (int i = 0; i < loc; i++) T_(d, d, i, j) = T_(d, d, i, j) + mass_mat[i + loc * j];
      }
    }
    if (o->assemble_inverse_masslump)
      for (int d = 0; d < dim; d++)
        for (int i = 0; i < loc; i++) masslump_addto[d + dim * i] += mass_lump[i];
  }

  /* Advection terms (:1416 -> :1602-1715) */
  if (!o->exclude_advection) {
    double div_relu_gi[MAXNGI], advection_mat[MAXLOC * MAXLOC], tmp[MAXLOC * MAXLOC], cd[MAXNGI];
    div_at_quad(m, nu_val, dshape, div_relu_gi);
    if (o->integrate_advection_by_parts) {
      /* -dshape_dot_vector_shape(du_t, relu_gi, u_shape, detwei*density_gi)
         -(1-beta)*shape_shape(test, u_shape, div_relu_gi*detwei*density_gi)  (:1667-1668) */
      for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * density_gi[g];
      dshape_dot_vector_shape(dim, loc, ngi, dshape, relu_gi, n, cd, advection_mat);
      for (int g = 0; g < ngi; g++) cd[g] = div_relu_gi[g] * detwei[g] * density_gi[g];
      shape_shape2(loc, ngi, test_n, n, cd, tmp);
      for (int a = 0; a < loc * loc; a++) advection_mat[a] = -advection_mat[a] - (1. - o->beta) * tmp[a];
    } else {
      /* shape_vector_dot_dshape(test, relu_gi, du_t, density_gi*detwei)
         + beta*shape_shape(test, u_shape, div_relu_gi*detwei*density_gi)   (:1675-1680) */
      for (int g = 0; g < ngi; g++) cd[g] = density_gi[g] * detwei[g];
      shape_vector_dot_dshape(dim, loc, ngi, test_n, relu_gi, dshape, cd, advection_mat);
      for (int g = 0; g < ngi; g++) cd[g] = div_relu_gi[g] * detwei[g] * density_gi[g];
      shape_shape2(loc, ngi, test_n, n, cd, tmp);
      for (int a = 0; a < loc * loc; a++) advection_mat[a] = advection_mat[a] + o->beta * tmp[a];
    }
    if (o->stabilisation_scheme == CGASM_STAB_STREAMLINE_UPWIND) { /* :1686-1708 */
      double diff_q[MAXDIM * MAXDIM * MAXNGI], stab[MAXLOC * MAXLOC];
      if (o->have_viscosity) {
        memcpy(diff_q, viscosity_gi, sizeof(double) * dim * dim * ngi);
        for (int g = 0; g < ngi; g++)
          for (int a = 0; a < dim; a++)
            for (int b = 0; b < dim; b++)
              if (a != b) diff_q[a + dim * (b + dim * g)] = 0.0;
      }
      element_upwind_stabilisation(dim, loc, ngi, dshape, relu_gi, J_mat, detwei,
                                   o->have_viscosity ? diff_q : NULL, o->nu_bar_scheme,
                                   o->nu_bar_scale, stab);
      for (int a = 0; a < loc * loc; a++) advection_mat[a] = advection_mat[a] + stab[a];
    }
    /* :1710-1713 */
    for (int d = 0; d < dim; d++) {
      for (int j = 0; j < loc; j++)
        for (int i = 0; i < loc; i++)
          T_(d, d, i, j) = T_(d, d, i, j) + o->dt * o->theta * advection_mat[i + loc * j];
      for (int i = 0; i < loc; i++) {
        double s = 0.0;
        for (int j = 0; j < loc; j++) s += advection_mat[i + loc * j] * OU_(d, j);
        RH_(d, i) = RH_(d, i) - s;
      }
    }
  }

  /* Source terms (:1421 -> :1717-1751) */
  if (o->have_source) {
    double src_val[MAXDIM * MAXLOC], source_mat[MAXLOC * MAXLOC], cd[MAXNGI];
    ele_val_vector(m, &f->source, ele, src_val);
    for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * density_gi[g];
    shape_shape2(loc, ngi, test_n, n, cd, source_mat);
    if (o->lump_source) {
      for (int i = 0; i < loc; i++) {
        double sl = 0.0;
        for (int j = 0; j < loc; j++) sl += source_mat[i + loc * j];
        for (int d = 0; d < dim; d++) RH_(d, i) = RH_(d, i) + sl * src_val[d + dim * i];
      }
    } else {
      for (int d = 0; d < dim; d++)
        for (int i = 0; i < loc; i++) {
          double s = 0.0;
          for (int j = 0; j < loc; j++) s += source_mat[i + loc * j] * src_val[d + dim * j];
          RH_(d, i) = RH_(d, i) + s;
        }
    }
  }

  /* Buoyancy (:1426 -> :1753-1792) */
  if (o->have_gravity) {
    double b_gi[MAXNGI], hb_gi[MAXNGI], cd[MAXNGI], g_val[MAXDIM * MAXLOC], g_gi[MAXDIM * MAXNGI],
        r[MAXDIM * MAXLOC];
    ele_val_scalar(m, &f->buoyancy, ele, ev_s);
    at_quad_scalar(m, ev_s, b_gi);
    if (o->subtract_out_reference_profile) {
      ele_val_scalar(m, &f->hb_density, ele, ev_s);
      at_quad_scalar(m, ev_s, hb_gi);
      for (int g = 0; g < ngi; g++) cd[g] = o->gravity_magnitude * (b_gi[g] - hb_gi[g]) * detwei[g];
    } else {
      for (int g = 0; g < ngi; g++) cd[g] = o->gravity_magnitude * b_gi[g] * detwei[g];
    }
    ele_val_vector(m, &f->gravity, ele, g_val);
    at_quad_multi(m, dim, g_val, g_gi);
    shape_vector_rhs(dim, loc, ngi, test_n, g_gi, cd, r);
    for (int a = 0; a < dim * loc; a++) rhs_addto[a] = rhs_addto[a] + r[a];
  }

  /* Absorption (:1436 -> :1818-1878, :2036-2073) */
  if (o->have_absorption) {
    double a_val[MAXDIM * MAXLOC], absorption_gi[MAXDIM * MAXNGI], cd[MAXNGI];
    double absorption_mat[MAXDIM * MAXLOC * MAXLOC], absorption_lump[MAXDIM * MAXLOC];
    ele_val_vector(m, &f->absorption, ele, a_val);
    at_quad_multi(m, dim, a_val, absorption_gi);
    for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * density_gi[g];
    shape_shape_vector2(dim, loc, ngi, test_n, n, cd, absorption_gi, absorption_mat);
#define AM_(d, i, j) absorption_mat[(d) + dim * ((i) + loc * (j))]
    if (o->lump_absorption) {
      for (int d = 0; d < dim; d++)
        for (int i = 0; i < loc; i++) {
          double s = 0.0;
          for (int j = 0; j < loc; j++) s += AM_(d, i, j);
          absorption_lump[d + dim * i] = s;
          DG_(d, i) = DG_(d, i) + o->dt * o->theta * s;
          RH_(d, i) = RH_(d, i) - s * OU_(d, i);
        }
    } else {
      for (int d = 0; d < dim; d++) {
        for (int j = 0; j < loc; j++)
          for (int i = 0; i < loc; i++)
            T_(d, d, i, j) = T_(d, d, i, j) + o->dt * o->theta * AM_(d, i, j);
        for (int i = 0; i < loc; i++) {
          double s = 0.0;
          for (int j = 0; j < loc; j++) s += AM_(d, i, j) * OU_(d, j);
          RH_(d, i) = RH_(d, i) - s;
        }
      }
      for (int a = 0; a < dim * loc; a++) absorption_lump[a] = 0.0;
    }
    if (o->pressure_corrected_absorption && o->assemble_inverse_masslump)
      for (int a = 0; a < dim * loc; a++) masslump_addto[a] += o->dt * o->theta * absorption_lump[a];
#undef AM_
  }

  /* Viscosity (:1445 -> :2079-2133, :2286-2359) */
  if (o->have_viscosity) {
    double vm[MAXLOC * MAXLOC], cd[MAXNGI];
    if (o->viscosity_shape == CGASM_TENSOR_ISOTROPIC) {
      for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * viscosity_gi[0 + dim * (0 + dim * g)];
      dshape_dot_dshape(dim, loc, ngi, dshape, cd, vm);
    } else if (o->viscosity_shape == CGASM_TENSOR_DIAGONAL) {
      dshape_diagtensor_dshape(dim, loc, ngi, dshape, viscosity_gi, detwei, vm);
    } else {
      dshape_tensor_dshape(dim, loc, ngi, dshape, viscosity_gi, detwei, vm);
    }
    for (int d = 0; d < dim; d++) {
      for (int j = 0; j < loc; j++)
        for (int i = 0; i < loc; i++)
          T_(d, d, i, j) = T_(d, d, i, j) + o->dt * o->theta * vm[i + loc * j];
      for (int i = 0; i < loc; i++) {
        double s = 0.0;
        for (int j = 0; j < loc; j++) s += vm[i + loc * j] * OU_(d, j);
        RH_(d, i) = RH_(d, i) - s;
      }
    }
  }

  /* add_diagonal_to_tensor (:1462, :1478-1488) */
  for (int d = 0; d < dim; d++)
    for (int i = 0; i < loc; i++) T_(d, d, i, i) = T_(d, d, i, i) + DG_(d, i);
#undef T_
#undef RH_
#undef DG_
#undef OU_
  return 0;
}

/* ------------------------------------------------------------------------------------
 * assemble_advection_diffusion_element_cg, assemble/Advection_Diffusion_CG.F90:702-865
 * with add_mass :867-941, add_advection :943-1127 (default equation type), add_source
 * :1129-1141, add_absorption :1143-1162, add_diffusivity :1164-1202.
 * Outputs matrix_addto(loc,loc), rhs_addto(loc).
 * ------------------------------------------------------------------------------------ */
typedef struct orc_advdiff_fields {
  orc_field t, velocity, source, absorption, diffusivity;
} orc_advdiff_fields;

int orc_advdiff_element(const orc_mesh* m, const orc_advdiff_fields* f, const cgasm_advdiff_opts* o,
                        int ele, double* matrix_addto, double* rhs_addto) {
  const int dim = m->dim, loc = m->loc, ngi = m->ngi;
  const double* n = m->n;
  if (o->move_mesh || o->multiphase || o->equation_type_not_advdiff) return CGASM_EUNSUPPORTED;
  const double dt_theta = o->dt * o->theta; /* :387 */
  const double eps = 2.220446049250313e-16;  /* epsilon(0.0) with -fdefault-real-8 */

  double X_val[MAXDIM * MAXLOC], dt_t[MAXLOC * MAXNGI * MAXDIM], detwei[MAXNGI],
      j_mat[MAXDIM * MAXDIM * MAXNGI], t_val[MAXLOC];
  double n_test_buf[MAXLOC * MAXNGI];
  const double* test_n = n;
  const double* dshape = dt_t;
  const int stab = o->stabilisation_scheme;

  for (int a = 0; a < loc * loc; a++) matrix_addto[a] = 0.0;
  for (int i = 0; i < loc; i++) rhs_addto[i] = 0.0;

  /* Step 1 (:768-776) */
  {
    orc_field Xf = {m->X, CGASM_FIELD_NORMAL};
    ele_val_vector(m, &Xf, ele, X_val);
    orc_transform_to_physical(dim, ngi, X_val, m->dn, m->weight, dt_t, detwei,
                              stab == CGASM_STAB_NONE ? NULL : j_mat);
  }
  ele_val_scalar(m, &f->t, ele, t_val);

  double u_val[MAXDIM * MAXLOC], velocity_at_quad[MAXDIM * MAXNGI];
  double diff_val[MAXDIM * MAXDIM * MAXLOC], diffusivity_gi[MAXDIM * MAXDIM * MAXNGI];
  if (o->have_advection || stab != CGASM_STAB_NONE) {
    ele_val_vector(m, &f->velocity, ele, u_val);
    at_quad_multi(m, dim, u_val, velocity_at_quad);
  }
  if (o->have_diffusivity) {
    ele_val_tensor(m, &f->diffusivity, ele, diff_val);
    at_quad_multi(m, dim * dim, diff_val, diffusivity_gi);
  }

  /* Step 2 (:813-826) */
  if (stab == CGASM_STAB_SUPG) {
    supg_test_function(dim, loc, ngi, n, dshape, velocity_at_quad, j_mat,
                       o->have_diffusivity ? diffusivity_gi : NULL, o->nu_bar_scheme,
                       o->nu_bar_scale, n_test_buf);
    test_n = n_test_buf;
  }

  /* Mass (:834 -> :867-941, default equation type :902-909) */
  if (o->have_mass) {
    double mass_matrix[MAXLOC * MAXLOC];
    shape_shape2(loc, ngi, test_n, n, detwei, mass_matrix);
    if (o->lump_mass) {
      for (int i = 0; i < loc; i++) {
        double s = 0.0;
        for (int j = 0; j < loc; j++) s += mass_matrix[i + loc * j];
        matrix_addto[i + loc * i] = matrix_addto[i + loc * i] + s;
      }
    } else {
      for (int a = 0; a < loc * loc; a++) matrix_addto[a] = matrix_addto[a] + mass_matrix[a];
    }
  }

  /* Advection (:837 -> :943-1127) */
  if (o->have_advection) {
    double advection_mat[MAXLOC * MAXLOC], tmp[MAXLOC * MAXLOC], div_q[MAXNGI], cd[MAXNGI];
    if (o->integrate_advection_by_parts) {
      dshape_dot_vector_shape(dim, loc, ngi, dshape, velocity_at_quad, n, detwei, advection_mat);
      for (int a = 0; a < loc * loc; a++) advection_mat[a] = -advection_mat[a];
      if (fabs(1.0 - o->beta) > eps) { /* :1044-1048 */
        div_at_quad(m, u_val, dshape, div_q);
        for (int g = 0; g < ngi; g++) cd[g] = div_q[g] * detwei[g];
        shape_shape2(loc, ngi, test_n, n, cd, tmp);
        for (int a = 0; a < loc * loc; a++) advection_mat[a] = advection_mat[a] - (1.0 - o->beta) * tmp[a];
      }
    } else {
      shape_vector_dot_dshape(dim, loc, ngi, test_n, velocity_at_quad, dshape, detwei, advection_mat);
      if (fabs(o->beta) > eps) { /* :1093-1098 */
        div_at_quad(m, u_val, dshape, div_q);
        for (int g = 0; g < ngi; g++) cd[g] = div_q[g] * detwei[g];
        shape_shape2(loc, ngi, test_n, n, cd, tmp);
        for (int a = 0; a < loc * loc; a++) advection_mat[a] = advection_mat[a] + o->beta * tmp[a];
      }
    }
    if (stab == CGASM_STAB_STREAMLINE_UPWIND) { /* :1107-1119 */
      double st[MAXLOC * MAXLOC];
      element_upwind_stabilisation(dim, loc, ngi, dshape, velocity_at_quad, j_mat, detwei,
                                   o->have_diffusivity ? diffusivity_gi : NULL, o->nu_bar_scheme,
                                   o->nu_bar_scale, st);
      for (int a = 0; a < loc * loc; a++) advection_mat[a] = advection_mat[a] + st[a];
    }
    if (fabs(dt_theta) > eps)
      for (int a = 0; a < loc * loc; a++) matrix_addto[a] = matrix_addto[a] + dt_theta * advection_mat[a];
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int j = 0; j < loc; j++) s += advection_mat[i + loc * j] * t_val[j];
      rhs_addto[i] = rhs_addto[i] - s;
    }
  }

  /* Absorption (:843 -> :1143-1162) */
  if (o->have_absorption) {
    double ev[MAXLOC], a_gi[MAXNGI], cd[MAXNGI], am[MAXLOC * MAXLOC];
    ele_val_scalar(m, &f->absorption, ele, ev);
    at_quad_scalar(m, ev, a_gi);
    for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * a_gi[g];
    shape_shape2(loc, ngi, test_n, n, cd, am);
    if (fabs(dt_theta) > eps)
      for (int a = 0; a < loc * loc; a++) matrix_addto[a] = matrix_addto[a] + dt_theta * am[a];
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int j = 0; j < loc; j++) s += am[i + loc * j] * t_val[j];
      rhs_addto[i] = rhs_addto[i] - s;
    }
  }

  /* Diffusivity (:846 -> :1164-1202) */
  if (o->have_diffusivity) {
    double dm[MAXLOC * MAXLOC], cd[MAXNGI];
    if (o->diffusivity_shape == CGASM_TENSOR_ISOTROPIC) {
      for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * diffusivity_gi[0 + dim * (0 + dim * g)];
      dshape_dot_dshape(dim, loc, ngi, dshape, cd, dm);
    } else {
      dshape_tensor_dshape(dim, loc, ngi, dshape, diffusivity_gi, detwei, dm);
    }
    if (fabs(dt_theta) > eps)
      for (int a = 0; a < loc * loc; a++) matrix_addto[a] = matrix_addto[a] + dt_theta * dm[a];
    for (int i = 0; i < loc; i++) {
      double s = 0.0;
      for (int j = 0; j < loc; j++) s += dm[i + loc * j] * t_val[j];
      rhs_addto[i] = rhs_addto[i] - s;
    }
  }

  /* Source (:849 -> :1129-1141) */
  if (o->have_source) {
    double ev[MAXLOC], s_gi[MAXNGI], cd[MAXNGI], r[MAXLOC];
    ele_val_scalar(m, &f->source, ele, ev);
    at_quad_scalar(m, ev, s_gi);
    for (int g = 0; g < ngi; g++) cd[g] = detwei[g] * s_gi[g];
    shape_rhs(loc, ngi, test_n, cd, r);
    for (int i = 0; i < loc; i++) rhs_addto[i] = rhs_addto[i] + r[i];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Sparsity: make_sparsity_lists + lists2csr_sparsity, femtools/Sparsity_Patterns.F90
 * :299-379, :381-428, with femtools/Linked_Lists.F90 insert_ascending (sorted singly linked
 * list without duplicates). Outputs are malloc'ed by the oracle; caller frees with orc_free.
 * ------------------------------------------------------------------------------------ */
typedef struct inode {
  int value;
  int next; /* index into pool, -1 = end */
} inode;

typedef struct ilist_pool {
  inode* pool;
  size_t used, cap;
} ilist_pool;

static int pool_new(ilist_pool* p, int value, int next) {
  if (p->used == p->cap) {
    p->cap = p->cap ? p->cap * 2 : 1024;
    p->pool = (inode*)realloc(p->pool, p->cap * sizeof(inode));
  }
  p->pool[p->used].value = value;
  p->pool[p->used].next = next;
  return (int)(p->used++);
}

/* insert_ascending: walk from the head, insert before the first larger value, skip if equal */
static void insert_ascending(ilist_pool* p, int* head, int* length, int value) {
  int prev = -1, cur = *head;
  while (cur >= 0 && p->pool[cur].value < value) {
    prev = cur;
    cur = p->pool[cur].next;
  }
  if (cur >= 0 && p->pool[cur].value == value) return;
  int nn = pool_new(p, value, cur);
  if (prev < 0) *head = nn;
  else p->pool[prev].next = nn;
  (*length)++;
}

int orc_make_sparsity(int n_nodes, int n_elements, int loc, const int* ndglno, int** findrm_out,
                      int** colm_out, int** centrm_out) {
  ilist_pool p = {0, 0, 0};
  int* head = (int*)malloc(sizeof(int) * (size_t)n_nodes);
  int* length = (int*)calloc((size_t)n_nodes, sizeof(int));
  for (int i = 0; i < n_nodes; i++) head[i] = -1;
  /* :323-336 */
  for (int ele = 0; ele < n_elements; ele++) {
    const int* nd = ndglno + (size_t)loc * ele;
    for (int i = 0; i < loc; i++)
      for (int j = 0; j < loc; j++) insert_ascending(&p, &head[nd[i] - 1], &length[nd[i] - 1], nd[j]);
  }
  /* lists2csr_sparsity :399-426 */
  int* findrm = (int*)malloc(sizeof(int) * ((size_t)n_nodes + 1));
  int count = 1;
  for (int i = 0; i < n_nodes; i++) {
    findrm[i] = count;
    count += length[i];
  }
  findrm[n_nodes] = count;
  int* colm = (int*)malloc(sizeof(int) * (size_t)(count - 1 > 0 ? count - 1 : 1));
  int* centrm = (int*)malloc(sizeof(int) * (size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    int k = findrm[i] - 1;
    centrm[i] = 0;
    for (int cur = head[i]; cur >= 0; cur = p.pool[cur].next) {
      colm[k] = p.pool[cur].value;
      if (colm[k] == i + 1) centrm[i] = k + 1;
      k++;
    }
  }
  free(p.pool);
  free(head);
  free(length);
  *findrm_out = findrm;
  *colm_out = colm;
  *centrm_out = centrm;
  return count - 1;
}

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------
 * Colouring: get_mesh_colouring(COLOURING_CG1), femtools/Colouring.F90:85-154: graph =
 * make_sparsity_transpose(P0 mesh, topology) (Sparsity_Patterns.F90:87-148: elements are
 * adjacent iff they share a node, self included), then colour_sparsity :159-199 (greedy,
 * element order, lowest colour not used by lower-numbered neighbours) and colour_sets
 * :250-262 (ascending element ids inside a colour).
 * colour_of(n_elements): 1-based colour per element. Returns number of colours.
 * ------------------------------------------------------------------------------------ */
int orc_colour_elements(int n_nodes, int n_elements, int loc, const int* ndglno, int* colour_of) {
  /* node -> element lists (CSR) */
  int* cnt = (int*)calloc((size_t)n_nodes + 1, sizeof(int));
  for (size_t k = 0; k < (size_t)loc * n_elements; k++) cnt[ndglno[k]]++;
  for (int i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
  int* n2e = (int*)malloc(sizeof(int) * (size_t)loc * n_elements);
  int* fill = (int*)malloc(sizeof(int) * (size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) fill[i] = cnt[i];
  for (int e = 0; e < n_elements; e++)
    for (int i = 0; i < loc; i++) {
      int nd = ndglno[(size_t)loc * e + i] - 1;
      n2e[fill[nd]++] = e;
    }
  int no_colours = 0;
  int cap = 64;
  unsigned char* used = (unsigned char*)calloc((size_t)cap + 2, 1);
  for (int e = 0; e < n_elements; e++) {
    if (e == 0) { /* :171-173 */
      colour_of[0] = 1;
      no_colours = 1;
      continue;
    }
    memset(used, 0, (size_t)no_colours + 2);
    for (int i = 0; i < loc; i++) {
      int nd = ndglno[(size_t)loc * e + i] - 1;
      for (int k = cnt[nd]; k < cnt[nd + 1]; k++) {
        int e2 = n2e[k];
        if (e2 < e) used[colour_of[e2]] = 1; /* cols(i) < node */
      }
    }
    for (int c = 1; c <= no_colours + 1; c++)
      if (!used[c]) {
        colour_of[e] = c;
        if (c > no_colours) {
          no_colours = c;
          if (no_colours + 2 > cap) {
            cap *= 2;
            used = (unsigned char*)realloc(used, (size_t)cap + 2);
          }
        }
        break;
      }
  }
  free(cnt);
  free(n2e);
  free(fill);
  free(used);
  return no_colours;
}

/* ------------------------------------------------------------------------------------
 * csr_sparsity_pos, femtools/Sparse_Tools.F90:2411-2517 (sorted rows => bisection
 * :2438-2497). i, j 1-based; returns the 1-based position in colm or 0.
 * ------------------------------------------------------------------------------------ */
static int csr_sparsity_pos(const int* findrm, const int* colm, int i, int j) {
  const int* row = colm + (findrm[i - 1] - 1);
  int size_row = findrm[i] - findrm[i - 1];
  int base = findrm[i - 1] - 1;
  int upper_pos = size_row, lower_pos = 1;
  int upper_j = row[upper_pos - 1], lower_j = row[0];
  if (upper_j < j) return 0;
  else if (upper_j == j) return upper_pos + base;
  else if (lower_j > j) return 0;
  else if (lower_j == j) return lower_pos + base;
  while (upper_pos - lower_pos > 1) {
    int this_pos = (upper_pos + lower_pos) / 2;
    int this_j = row[this_pos - 1];
    if (this_j == j) return this_pos + base;
    else if (this_j > j) upper_pos = this_pos;
    else lower_pos = this_pos;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * The element loops.
 *   momentum: Momentum_CG.F90:716-752 + insertion :1462-1470. big_m values accumulate like
 *     PETSc MatSetValues(ADD_VALUES) (Sparse_Tools_Petsc.F90:848-879): plain FP64 adds in
 *     call order into the (d,d) blocks; here held as dim nnz-arrays in colm order.
 *   tracer: Advection_Diffusion_CG.F90:569-598 + csr_vaddto (Sparse_Tools.F90:2680-2703,
 *     row-major over (iloc,jloc), exact zeros skipped :2640).
 * colour_ptr/colour_elements == NULL: serial reference order (single colour of all elements
 * ascending, Colouring.F90:145-151). Otherwise colours are visited in order and, inside a
 * colour, elements are distributed over OpenMP threads exactly like the reference's
 * !$OMP DO SCHEDULE(STATIC) (Momentum_CG.F90:732-751).
 * All outputs are zeroed first (Momentum_Equation.F90:593-606).
 * ------------------------------------------------------------------------------------ */
int orc_assemble_momentum(const orc_mesh* m, const orc_momentum_fields* f,
                          const cgasm_momentum_opts* o, const int* findrm, const int* colm,
                          int ncolours, const int* colour_ptr, const int* colour_elements,
                          double* big_m /*[dim][nnz]*/, double* rhs /*(dim,N)*/,
                          double* masslump /*(dim,N) or NULL*/, double* ct_m /*[dim][nnz] or NULL*/) {
  const int dim = m->dim, loc = m->loc;
  const size_t nnz = (size_t)(findrm[m->n_nodes] - 1);
  if (momentum_opts_unsupported(o)) return CGASM_EUNSUPPORTED;
  memset(big_m, 0, sizeof(double) * dim * nnz);
  memset(rhs, 0, sizeof(double) * (size_t)dim * m->n_nodes);
  if (masslump) memset(masslump, 0, sizeof(double) * (size_t)dim * m->n_nodes);
  if (ct_m) memset(ct_m, 0, sizeof(double) * dim * nnz);
  int one_ptr[2] = {1, m->n_elements + 1};
  int nc = colour_ptr ? ncolours : 1;
  const int* cptr = colour_ptr ? colour_ptr : one_ptr;
  int status = 0;
  for (int clr = 0; clr < nc; clr++) {
    const int len = cptr[clr + 1] - cptr[clr];
#pragma omp parallel for schedule(static) if (colour_ptr != NULL)
    for (int nnid = 0; nnid < len; nnid++) {
      int ele = colour_elements ? colour_elements[cptr[clr] - 1 + nnid] : nnid + 1;
      double T[MAXDIM * MAXDIM * MAXLOC * MAXLOC], r[MAXDIM * MAXLOC], ml[MAXDIM * MAXLOC],
          gp[MAXDIM * MAXLOC * MAXLOC];
      int st = orc_momentum_element(m, f, o, ele, T, r, ml, gp);
      if (st) {
        status = st;
        continue;
      }
      const int* nd = ele_nodes(m, ele);
      for (int i = 0; i < loc; i++)
        for (int j = 0; j < loc; j++) {
          int pos = csr_sparsity_pos(findrm, colm, nd[i], nd[j]);
          for (int d = 0; d < dim; d++) {
            big_m[d * nnz + (size_t)(pos - 1)] += T[d + dim * (d + dim * (i + loc * j))];
            if (ct_m && o->assemble_ct_matrix_here)
              ct_m[d * nnz + (size_t)(pos - 1)] += gp[d + dim * (i + loc * j)];
          }
        }
      for (int i = 0; i < loc; i++)
        for (int d = 0; d < dim; d++) {
          rhs[d + (size_t)dim * (nd[i] - 1)] += r[d + dim * i];
          if (masslump) masslump[d + (size_t)dim * (nd[i] - 1)] += ml[d + dim * i];
        }
    }
  }
  return status;
}

/* The `mass` matrix of construct_momentum_cg (assemble_mass_matrix, Momentum_CG.F90:1567-1571): every element adds its
 * density-weighted consistent mass matrix mass_mat = shape_shape(test_function, u_shape, detwei*density_gi) to each
 * diagonal block -- whatever lump_mass / exclude_mass say about big_m -- and, with pressure_corrected_absorption,
 * dt*theta*absorption_mat(dim,:,:) on top (:2073-2078, the full absorption matrix also when the absorption in big_m is
 * lumped). mass: [dim][nnz] diagonal blocks in colm order. Test function = u_shape (no SUPG). */
int orc_assemble_momentum_mass(const orc_mesh* m, const orc_momentum_fields* f, const cgasm_momentum_opts* o,
                               const int* findrm, const int* colm, double* mass) {
  const int dim = m->dim, loc = m->loc, ngi = m->ngi;
  const size_t nnz = (size_t)(findrm[m->n_nodes] - 1);
  if (momentum_opts_unsupported(o) || o->stabilisation_scheme == CGASM_STAB_SUPG) return CGASM_EUNSUPPORTED;
  memset(mass, 0, sizeof(double) * dim * nnz);
  for (int ele = 1; ele <= m->n_elements; ele++) {
    double X_val[MAXDIM * MAXLOC], du_t[MAXLOC * MAXNGI * MAXDIM], detwei[MAXNGI];
    orc_field Xf = {m->X, CGASM_FIELD_NORMAL};
    ele_val_vector(m, &Xf, ele, X_val);
    orc_transform_to_physical(dim, ngi, X_val, m->dn, m->weight, du_t, detwei, NULL);
    double density_gi[MAXNGI], ev_s[MAXLOC], cd[MAXNGI], mass_mat[MAXLOC * MAXLOC];
    ele_val_scalar(m, &f->density, ele, ev_s);
    at_quad_scalar(m, ev_s, density_gi);
    for (int g = 0; g < ngi; g++) cd[g] = density_gi[g] * detwei[g];
    shape_shape2(loc, ngi, m->n, m->n, cd, mass_mat);
    double absorption_mat[MAXDIM * MAXLOC * MAXLOC];
    const int pc = o->have_absorption && o->pressure_corrected_absorption;
    if (pc) {
      double a_val[MAXDIM * MAXLOC], absorption_gi[MAXDIM * MAXNGI];
      ele_val_vector(m, &f->absorption, ele, a_val);
      at_quad_multi(m, dim, a_val, absorption_gi);
      shape_shape_vector2(dim, loc, ngi, m->n, m->n, cd, absorption_gi, absorption_mat);
    }
    const int* nd = ele_nodes(m, ele);
    for (int i = 0; i < loc; i++)
      for (int j = 0; j < loc; j++) {
        const int pos = csr_sparsity_pos(findrm, colm, nd[i], nd[j]);
        for (int d = 0; d < dim; d++) {
          double v = mass_mat[i + loc * j];
          if (pc) v = v + o->dt * o->theta * absorption_mat[d + dim * (i + loc * j)];
          mass[d * nnz + (size_t)(pos - 1)] += v;
        }
      }
  }
  return 0;
}

int orc_assemble_advdiff(const orc_mesh* m, const orc_advdiff_fields* f, const cgasm_advdiff_opts* o,
                         const int* findrm, const int* colm, int ncolours, const int* colour_ptr,
                         const int* colour_elements, double* matrix_val /*nnz*/, double* rhs /*N*/) {
  const int loc = m->loc;
  const size_t nnz = (size_t)(findrm[m->n_nodes] - 1);
  if (o->move_mesh || o->multiphase || o->equation_type_not_advdiff) return CGASM_EUNSUPPORTED;
  memset(matrix_val, 0, sizeof(double) * nnz);
  memset(rhs, 0, sizeof(double) * (size_t)m->n_nodes);
  int one_ptr[2] = {1, m->n_elements + 1};
  int nc = colour_ptr ? ncolours : 1;
  const int* cptr = colour_ptr ? colour_ptr : one_ptr;
  for (int clr = 0; clr < nc; clr++) {
    const int len = cptr[clr + 1] - cptr[clr];
#pragma omp parallel for schedule(static) if (colour_ptr != NULL)
    for (int nnid = 0; nnid < len; nnid++) {
      int ele = colour_elements ? colour_elements[cptr[clr] - 1 + nnid] : nnid + 1;
      double A[MAXLOC * MAXLOC], r[MAXLOC];
      orc_advdiff_element(m, f, o, ele, A, r);
      const int* nd = ele_nodes(m, ele);
      for (int i = 0; i < loc; i++)
        for (int j = 0; j < loc; j++) {
          double v = A[i + loc * j];
          if (v == 0) continue; /* Sparse_Tools.F90:2640 */
          int pos = csr_sparsity_pos(findrm, colm, nd[i], nd[j]);
          matrix_val[pos - 1] += v;
        }
      for (int i = 0; i < loc; i++) rhs[nd[i] - 1] += r[i];
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * halo_update, femtools/Halos_Communications.F90:320-412: owner -> ghost copy. Restated for a
 * set of ranks held in one address space (the tests build all partitions in one process):
 * for every pair (p -> q), field_q[recvs_q(p)[k]] = field_p[sends_p(q)[k]], block_size reals
 * per node. Called once per ordered pair.
 * ------------------------------------------------------------------------------------ */
void orc_halo_copy(int block_size, const double* field_p, const int* sends_p_to_q, int n,
                   double* field_q, const int* recvs_q_from_p) {
  for (int k = 0; k < n; k++)
    for (int b = 0; b < block_size; b++)
      field_q[b + (size_t)block_size * (recvs_q_from_p[k] - 1)] =
          field_p[b + (size_t)block_size * (sends_p_to_q[k] - 1)];
}

/* block addto known answer, femtools/tests/test_petsc_csr_matrix.F90: addto(A,1,1,rows,cols,
 * vals) twice on a dense 4x4 pattern. Exposed so the golden test can drive the same
 * accumulate path the assembly uses. */
void orc_block_addto(const int* findrm, const int* colm, int nrows, const int* rows, int ncols,
                     const int* cols, const double* vals /*(nrows,ncols) col-major*/, double* val) {
  for (int i = 0; i < nrows; i++)
    for (int j = 0; j < ncols; j++) {
      int pos = csr_sparsity_pos(findrm, colm, rows[i], cols[j]);
      if (pos > 0) val[pos - 1] += vals[i + nrows * j];
    }
}

/* ====================================================================================
 * Surface-element loops and strong Dirichlet conditions: the step right after the element
 * loops. Same status as the rest of this file: TEST
 * INFRASTRUCTURE, restated from the cited lines; values pinned by closed forms and an
 * independent numpy evaluation (tests/test_surface_oracle.py) -- the reference has no unit
 * test at this level ("parity unpinned" against a Fortran run).
 * ==================================================================================== */

/* Face quadrature = rule of the mesh's degree on the (dim-1)-simplex
 * (femtools/Fields_Allocates.F90:1302-1320). dim == 3: the 4-point triangle rule above;
 * dim == 2: the 3-point degree-3 interval rule, femtools/Quadrature.F90:1181-1199 with
 * interval_permutations :1893-1905 (the constants are truncated in the source: kept as
 * written). l is (sngi, sloc) column-major; returns sngi. */
int orc_quadrature_face_degree3(int dim, double* l, double* weight) {
  if (dim == 3) return orc_quadrature_degree3(2, l, weight);
  if (dim != 2) return 0;
  const int ngi = 3;
  double coords[2];
  coords[0] = 0.887298334620742;
  coords[1] = 1.0 - coords[0];
  static const int p2[2][2] = {{1, 2}, {2, 1}};
  for (int k = 0; k < 2; k++) {
    for (int j = 0; j < 2; j++) l[k + ngi * j] = coords[p2[k][j] - 1];
    weight[k] = 0.277777777777777;
  }
  coords[0] = 0.5;
  for (int j = 0; j < 2; j++) l[2 + ngi * j] = coords[0];
  weight[2] = 0.444444444444444;
  return ngi;
}

typedef struct orc_surface {
  int n_faces, sloc, sngi;
  const int* sndgln;      /* sloc*n_faces, 1-based global nodes (face_global_nodes)   */
  const int* face_ele;    /* n_faces, 1-based owning element (face_ele)               */
  const double* n_f;      /* faces%shape%n(sloc,sngi)                                 */
  const double* dn_f;     /* faces%shape%dn(sloc,sngi,dim-1)                          */
  const double* weight_f; /* faces%shape%quadrature%weight(sngi)                      */
} orc_surface;

/* transform_facet_to_physical_full, femtools/Transform_elements.F90:1353-1525, linear simplex
 * branch: Jacobian at gi == 1 only (:1427), outward_vector = facet centroid - element centroid
 * (:1445), facet_normal :1527-1553, detwei_f = detJ*weight (:1498, :1514-1517), normal copied to
 * every gi (:1520-1522). X_f(dim,sloc), X_val(dim,loc). */
void orc_transform_facet_to_physical(int dim, int sloc, int sngi, const double* X_f, const double* X_val,
                                     const double* dn_f, const double* weight_f, double* detwei_f,
                                     double* normal /*(dim,sngi)*/) {
  const int loc = dim + 1;
  double outward[MAXDIM], J[MAXDIM * 2], nrm[MAXDIM];
  for (int a = 0; a < dim; a++) {
    double sf = 0.0, sv = 0.0;
    for (int i = 0; i < sloc; i++) sf += X_f[a + dim * i];
    for (int i = 0; i < loc; i++) sv += X_val[a + dim * i];
    outward[a] = sf / sloc - sv / loc;
  }
  /* J = matmul(X_f, dn(:,1,:)): J(a,k) */
  for (int k = 0; k < dim - 1; k++)
    for (int a = 0; a < dim; a++) {
      double s = 0.0;
      for (int i = 0; i < sloc; i++) s += X_f[a + dim * i] * dn_f[i + sloc * (0 + sngi * k)];
      J[a + dim * k] = s;
    }
#define J_(a, k) J[((a) - 1) + dim * ((k) - 1)]
  double detJ = 0.0;
  if (dim == 2) {
    detJ = sqrt(J_(1, 1) * J_(1, 1) + J_(2, 1) * J_(2, 1));
    nrm[0] = -J_(2, 1);
    nrm[1] = J_(1, 1);
  } else {
    for (int i = 1; i <= 3; i++) {
      double c = J_(cyc3(i + 2), 1) * J_(cyc3(i + 1), 2) - J_(cyc3(i + 2), 2) * J_(cyc3(i + 1), 1);
      detJ = detJ + c * c;
    }
    detJ = sqrt(detJ);
    nrm[0] = J_(2, 1) * J_(3, 2) - J_(3, 1) * J_(2, 2);
    nrm[1] = J_(3, 1) * J_(1, 2) - J_(1, 1) * J_(3, 2);
    nrm[2] = J_(1, 1) * J_(2, 2) - J_(2, 1) * J_(1, 2);
  }
#undef J_
  double dotp = 0.0, nn = 0.0;
  for (int a = 0; a < dim; a++) dotp += nrm[a] * outward[a];
  for (int a = 0; a < dim; a++) nrm[a] = nrm[a] * dotp;
  for (int a = 0; a < dim; a++) nn += nrm[a] * nrm[a];
  nn = sqrt(nn);
  for (int g = 0; g < sngi; g++) {
    if (detwei_f) detwei_f[g] = detJ * weight_f[g];
    for (int a = 0; a < dim; a++) normal[a + dim * g] = nrm[a] / nn;
  }
}

#define MAXSLOC 3
#define MAXSNGI 4

static void face_geometry(const orc_mesh* m, const orc_surface* s, int face, double* detwei, double* normal) {
  const int dim = m->dim;
  const int* fn = s->sndgln + (size_t)s->sloc * (size_t)(face - 1);
  const int* en = ele_nodes(m, s->face_ele[face - 1]);
  double X_f[MAXDIM * MAXSLOC], X_val[MAXDIM * MAXLOC];
  for (int i = 0; i < s->sloc; i++)
    for (int a = 0; a < dim; a++) X_f[a + dim * i] = m->X[a + (size_t)dim * (size_t)(fn[i] - 1)];
  for (int i = 0; i < m->loc; i++)
    for (int a = 0; a < dim; a++) X_val[a + dim * i] = m->X[a + (size_t)dim * (size_t)(en[i] - 1)];
  orc_transform_facet_to_physical(dim, s->sloc, s->sngi, X_f, X_val, s->dn_f, s->weight_f, detwei, normal);
}

/* face_val (Fields_Base.F90:2172-2254) of a nodal field with ncomp components: out(ncomp,sloc) */
static void face_val_multi(const orc_surface* s, const orc_field* f, int ncomp, int face, double* out) {
  const int* fn = s->sndgln + (size_t)s->sloc * (size_t)(face - 1);
  for (int i = 0; i < s->sloc; i++)
    for (int c = 0; c < ncomp; c++)
      out[c + ncomp * i] = (f->field_type == CGASM_FIELD_CONSTANT) ? f->val[c] : f->val[c + (size_t)ncomp * (size_t)(fn[i] - 1)];
}
/* face_val_at_quad (Fields_Base.F90:2374-2400): matmul(face_val, faces%shape%n) */
static void face_at_quad(const orc_surface* s, int ncomp, const double* fv, double* q /*(ncomp,sngi)*/) {
  for (int g = 0; g < s->sngi; g++)
    for (int c = 0; c < ncomp; c++) {
      double v = 0.0;
      for (int i = 0; i < s->sloc; i++) v += fv[c + ncomp * i] * s->n_f[i + s->sloc * g];
      q[c + ncomp * g] = v;
    }
}

/* assemble_advection_diffusion_face_cg, assemble/Advection_Diffusion_CG.F90:1228-1283 with
 * add_advection_face_cg :1285-1340 (default equation type) and add_diffusivity_face_cg :1342-1379.
 * bc_type: 0, BC_TYPE_NEUMANN = 1, BC_TYPE_WEAKDIRICHLET = 2, BC_TYPE_ROBIN = 4 (:74-75; internal = 3 faces are
 * skipped by the caller, :629). t_bc / t_bc_2 (sloc): ele_val of the surface fields on this face.
 * Outputs overwritten: matrix_addto(sloc,sloc), rhs_addto(sloc). */
int orc_advdiff_face(const orc_mesh* m, const orc_surface* s, const orc_advdiff_fields* f, const cgasm_advdiff_opts* o,
                     int face, int bc_type, const double* t_bc, const double* t_bc_2, double* matrix_addto,
                     double* rhs_addto) {
  const int dim = m->dim, sloc = s->sloc, sngi = s->sngi;
  const double dt_theta = o->dt * o->theta;
  const double eps = 2.220446049250313e-16;
  if (o->move_mesh || o->multiphase || o->equation_type_not_advdiff) return CGASM_EUNSUPPORTED;
  if (!(bc_type == 0 || bc_type == 1 || bc_type == 2 || bc_type == 4)) return CGASM_EARG; /* assert :1253 */
  double detwei[MAXSNGI], normal[MAXDIM * MAXSNGI], t_face[MAXSLOC], c_g[MAXSNGI], mat[MAXSLOC * MAXSLOC], q[MAXSNGI];
  for (int k = 0; k < sloc * sloc; k++) matrix_addto[k] = 0.0;
  for (int i = 0; i < sloc; i++) rhs_addto[i] = 0.0;
  const int by_parts_adv = o->have_advection && o->integrate_advection_by_parts;
  if (by_parts_adv || (o->have_diffusivity && (bc_type == 1 || bc_type == 4))) face_geometry(m, s, face, detwei, normal);
  face_val_multi(s, &f->t, 1, face, t_face);
  if (by_parts_adv) { /* :1285-1340 */
    double u_f[MAXDIM * MAXSLOC], u_q[MAXDIM * MAXSNGI];
    face_val_multi(s, &f->velocity, dim, face, u_f);
    face_at_quad(s, dim, u_f, u_q);
    for (int g = 0; g < sngi; g++) {
      double un = 0.0;
      for (int a = 0; a < dim; a++) un += u_q[a + dim * g] * normal[a + dim * g];
      c_g[g] = detwei[g] * un;
    }
    shape_shape2(sloc, sngi, s->n_f, s->n_f, c_g, mat);
    if (fabs(dt_theta) > eps) {
      if (bc_type == 2) {
        for (int i = 0; i < sloc; i++) {
          double v = 0.0;
          for (int j = 0; j < sloc; j++) v += mat[i + sloc * j] * (t_bc[j] - t_face[j]);
          rhs_addto[i] = rhs_addto[i] - o->theta * v;
        }
      } else {
        for (int k = 0; k < sloc * sloc; k++) matrix_addto[k] = matrix_addto[k] + dt_theta * mat[k];
      }
    }
    for (int i = 0; i < sloc; i++) {
      double v = 0.0;
      for (int j = 0; j < sloc; j++) v += mat[i + sloc * j] * t_face[j];
      rhs_addto[i] = rhs_addto[i] - v;
    }
  }
  if (o->have_diffusivity) { /* :1342-1379 */
    if (bc_type == 1 || bc_type == 4) {
      double r[MAXSLOC];
      face_at_quad(s, 1, t_bc, q);
      for (int g = 0; g < sngi; g++) c_g[g] = detwei[g] * q[g];
      shape_rhs(sloc, sngi, s->n_f, c_g, r);
      for (int i = 0; i < sloc; i++) rhs_addto[i] = rhs_addto[i] + r[i];
    }
    if (bc_type == 4) {
      face_at_quad(s, 1, t_bc_2, q);
      for (int g = 0; g < sngi; g++) c_g[g] = detwei[g] * q[g];
      shape_shape2(sloc, sngi, s->n_f, s->n_f, c_g, mat);
      if (fabs(dt_theta) > eps)
        for (int k = 0; k < sloc * sloc; k++) matrix_addto[k] = matrix_addto[k] + dt_theta * mat[k];
      for (int i = 0; i < sloc; i++) {
        double v = 0.0;
        for (int j = 0; j < sloc; j++) v += mat[i + sloc * j] * t_face[j];
        rhs_addto[i] = rhs_addto[i] - v;
      }
    } else if (bc_type == 2) {
      return CGASM_EUNSUPPORTED; /* FLExit :1375: weak Dirichlet with diffusivity */
    }
  }
  return 0;
}

/* The face loop of assemble_advection_diffusion_cg, assemble/Advection_Diffusion_CG.F90:609-643: runs only
 * with by-parts advection or diffusivity; internal faces skipped; ascending faces; csr addto of
 * (face_nodes, face_nodes) (exact zeros skipped, Sparse_Tools.F90:2640) and rhs addto. ADDS to
 * matrix_val / rhs (they hold the element-loop result). bc arrays: bc_type(n_faces), t_bc and t_bc_2
 * (sloc, n_faces) (either may be NULL when no face needs it). */
int orc_assemble_advdiff_surface(const orc_mesh* m, const orc_surface* s, const orc_advdiff_fields* f,
                                 const cgasm_advdiff_opts* o, const int* findrm, const int* colm, const int* bc_type,
                                 const double* t_bc, const double* t_bc_2, double* matrix_val, double* rhs) {
  if (!((o->integrate_advection_by_parts && o->have_advection) || o->have_diffusivity)) return 0;
  const int sloc = s->sloc;
  const double zero[MAXSLOC] = {0, 0, 0};
  for (int face = 1; face <= s->n_faces; face++) {
    if (bc_type[face - 1] == 3) continue;
    double A[MAXSLOC * MAXSLOC], r[MAXSLOC];
    int st = orc_advdiff_face(m, s, f, o, face, bc_type[face - 1], t_bc ? t_bc + (size_t)sloc * (face - 1) : zero,
                              t_bc_2 ? t_bc_2 + (size_t)sloc * (face - 1) : zero, A, r);
    if (st) return st;
    const int* fn = s->sndgln + (size_t)sloc * (size_t)(face - 1);
    for (int i = 0; i < sloc; i++)
      for (int j = 0; j < sloc; j++) {
        double v = A[i + sloc * j];
        if (v == 0) continue;
        matrix_val[csr_sparsity_pos(findrm, colm, fn[i], fn[j]) - 1] += v;
      }
    for (int i = 0; i < sloc; i++) rhs[fn[i] - 1] += r[i];
  }
  return 0;
}

/* apply_dirichlet_conditions_scalar, femtools/Boundary_Conditions.F90:1982-2024, for one boundary condition:
 * rhs(node_j) = (value_j - field(node_j))/dt with dt given ("rate of change form"), else value_j. The matrix
 * rows are only flagged inactive there (set_inactive, :2006): `inactive` (n_nodes, may be NULL) receives
 * the flags. nodes 1-based. */
void orc_apply_dirichlet_scalar(int n, const int* nodes, const double* values, const double* field,
                                int have_dt, double dt, double* rhs, int* inactive) {
  for (int j = 0; j < n; j++) {
    const int node = nodes[j] - 1;
    if (inactive) inactive[node] = 1;
    rhs[node] = have_dt ? (values[j] - field[node]) / dt : values[j];
  }
}

/* construct_momentum_surface_element_cg, assemble/Momentum_CG.F90:959-1191, the branches inside the device
 * path's guard (no free-surface stabilisation, continuity not by parts, single phase, static mesh):
 *   by-parts advection boundary term :1029-1071, flux boundary condition :1180-1187.
 * velocity_bc_type(dim): 0, BC_TYPE_WEAKDIRICHLET = 1, NO_NORMAL_FLOW = 2, INTERNAL = 3, FREE_SURFACE = 4,
 * FLUX = 5 (:138-140). velocity_bc(dim, sloc): ele_val of the surface field. Outputs overwritten:
 * big_m_addto(dim, sloc, sloc) (diagonal blocks only), rhs_addto(dim, sloc). */
int orc_momentum_face_ml(const orc_mesh* m, const orc_surface* s, const orc_momentum_fields* f,
                         const cgasm_momentum_opts* o, int face, const int* velocity_bc_type, const double* velocity_bc,
                         double* big_m_addto, double* rhs_addto, double* masslump_addto);

int orc_momentum_face(const orc_mesh* m, const orc_surface* s, const orc_momentum_fields* f,
                      const cgasm_momentum_opts* o, int face, const int* velocity_bc_type, const double* velocity_bc,
                      double* big_m_addto, double* rhs_addto) {
  return orc_momentum_face_ml(m, s, f, o, face, velocity_bc_type, velocity_bc, big_m_addto, rhs_addto, NULL);
}

/* masslump_addto(dim, sloc) (may be NULL): what the free-surface stabilisation adds to masslump (:1167-1173) */
int orc_momentum_face_ml(const orc_mesh* m, const orc_surface* s, const orc_momentum_fields* f,
                         const cgasm_momentum_opts* o, int face, const int* velocity_bc_type, const double* velocity_bc,
                         double* big_m_addto, double* rhs_addto, double* masslump_addto) {
  const int dim = m->dim, sloc = s->sloc, sngi = s->sngi;
  if (momentum_opts_unsupported(o)) return CGASM_EUNSUPPORTED;
  double detwei[MAXSNGI], normal[MAXDIM * MAXSNGI], c_g[MAXSNGI], mat[MAXSLOC * MAXSLOC];
  for (int k = 0; k < dim * sloc * sloc; k++) big_m_addto[k] = 0.0;
  for (int k = 0; k < dim * sloc; k++) rhs_addto[k] = 0.0;
  if (masslump_addto)
    for (int k = 0; k < dim * sloc; k++) masslump_addto[k] = 0.0;
  face_geometry(m, s, face, detwei, normal);
  double oldu_val[MAXDIM * MAXSLOC];
  face_val_multi(s, &f->oldu, dim, face, oldu_val);
  if (velocity_bc_type[0] != 2) {
    if (o->integrate_advection_by_parts && !o->exclude_advection) {
      double nu_f[MAXDIM * MAXSLOC], relu_gi[MAXDIM * MAXSNGI], rho_f[MAXSLOC], rho_q[MAXSNGI];
      face_val_multi(s, &f->nu, dim, face, nu_f);
      face_at_quad(s, dim, nu_f, relu_gi);
      face_val_multi(s, &f->density, 1, face, rho_f);
      face_at_quad(s, 1, rho_f, rho_q);
      for (int g = 0; g < sngi; g++) {
        double un = 0.0;
        for (int a = 0; a < dim; a++) un += relu_gi[a + dim * g] * normal[a + dim * g];
        c_g[g] = detwei[g] * un * rho_q[g];
      }
      shape_shape2(sloc, sngi, s->n_f, s->n_f, c_g, mat);
      for (int d = 0; d < dim; d++) {
        if (velocity_bc_type[d] == 1) {
          for (int i = 0; i < sloc; i++) {
            double v = 0.0;
            for (int j = 0; j < sloc; j++) v += mat[i + sloc * j] * velocity_bc[d + dim * j];
            rhs_addto[d + dim * i] += -v;
          }
        } else {
          for (int k = 0; k < sloc * sloc; k++) big_m_addto[d + dim * k] += o->dt * o->theta * mat[k];
          for (int i = 0; i < sloc; i++) {
            double v = 0.0;
            for (int j = 0; j < sloc; j++) v += mat[i + sloc * j] * oldu_val[d + dim * j];
            rhs_addto[d + dim * i] += -v;
          }
        }
      }
    }
  }
  /* Free-surface stabilisation, :1108-1176 (not on the sphere): upwards = -gravity direction at the quadrature points,
   * ndotk_k(:,g) = fs_sf (normal . upwards) upwards, fs_surfacestab = shape_shape_vector(u_shape, u_shape,
   * detwei_bdy*density_gi, dt*gravity_magnitude*ndotk_k); lumped (lump_mass) or full (no pressure-corrected absorption). */
  if (velocity_bc_type[0] == 4 && o->have_surface_fs_stabilisation) {
    double grav_f[MAXDIM * MAXSLOC], up_gi[MAXDIM * MAXSNGI], rho_f[MAXSLOC], rho_q[MAXSNGI], ndotk[MAXDIM * MAXSNGI];
    double fs[MAXDIM * MAXSLOC * MAXSLOC];
    if (o->on_sphere) return CGASM_EUNSUPPORTED;
    if (!o->lump_mass && o->pressure_corrected_absorption) return CGASM_EUNSUPPORTED; /* FLExit :1161-1163 */
    face_val_multi(s, &f->gravity, dim, face, grav_f);
    face_at_quad(s, dim, grav_f, up_gi);
    for (int k = 0; k < dim * sngi; k++) up_gi[k] = -up_gi[k];
    face_val_multi(s, &f->density, 1, face, rho_f);
    face_at_quad(s, 1, rho_f, rho_q);
    for (int g = 0; g < sngi; g++) {
      double nk = 0.0;
      for (int a = 0; a < dim; a++) nk += normal[a + dim * g] * up_gi[a + dim * g];
      for (int a = 0; a < dim; a++) ndotk[a + dim * g] = o->fs_sf * nk * up_gi[a + dim * g];
      c_g[g] = detwei[g] * rho_q[g];
    }
    /* shape_shape_vector: R(d,i,j) = sum_g n_i n_j detwei_g vector(d,g) */
    for (int d = 0; d < dim; d++)
      for (int i = 0; i < sloc; i++)
        for (int j = 0; j < sloc; j++) {
          double v = 0.0;
          for (int g = 0; g < sngi; g++)
            v += s->n_f[i + sloc * g] * s->n_f[j + sloc * g] * c_g[g] * (o->dt * o->gravity_magnitude * ndotk[d + dim * g]);
          fs[d + dim * (i + sloc * j)] = v;
        }
    if (o->lump_mass) {
      for (int d = 0; d < dim; d++)
        for (int i = 0; i < sloc; i++) {
          double l = 0.0;
          for (int j = 0; j < sloc; j++) l += fs[d + dim * (i + sloc * j)];
          big_m_addto[d + dim * (i + sloc * i)] += o->dt * o->theta * l;
          rhs_addto[d + dim * i] += -l * oldu_val[d + dim * i];
          if (o->pressure_corrected_absorption && masslump_addto) masslump_addto[d + dim * i] += o->dt * o->theta * l;
        }
    } else {
      for (int d = 0; d < dim; d++)
        for (int i = 0; i < sloc; i++) {
          double v = 0.0;
          for (int j = 0; j < sloc; j++) {
            big_m_addto[d + dim * (i + sloc * j)] += o->dt * o->theta * fs[d + dim * (i + sloc * j)];
            v += fs[d + dim * (i + sloc * j)] * oldu_val[d + dim * j];
          }
          rhs_addto[d + dim * i] += -v;
        }
    }
  }
  for (int d = 0; d < dim; d++)
    if (velocity_bc_type[d] == 5) { /* shape_rhs(u_shape, ele_val_at_quad(velocity_bc, sele, dim)*detwei_bdy) */
      double bc_d[MAXSLOC], q[MAXSNGI] = {0, 0, 0, 0}, r[MAXSLOC];
      for (int i = 0; i < sloc; i++) bc_d[i] = velocity_bc[d + dim * i];
      face_at_quad(s, 1, bc_d, q);
      for (int g = 0; g < sngi; g++) c_g[g] = q[g] * detwei[g];
      shape_rhs(sloc, sngi, s->n_f, c_g, r);
      for (int i = 0; i < sloc; i++) rhs_addto[d + dim * i] += r[i];
    }
  return 0;
}

/* surface_element_loop of construct_momentum_cg, assemble/Momentum_CG.F90:795-812: faces whose only
 * condition is no-normal-flow, or that are internal, are skipped unless they carry a pressure condition
 * (:799-803). pressure_bc_type may be NULL (= all zero). ADDS to big_m [dim][nnz] and rhs (dim, N). */
int orc_assemble_momentum_surface_ml(const orc_mesh* m, const orc_surface* s, const orc_momentum_fields* f,
                                     const cgasm_momentum_opts* o, const int* findrm, const int* colm,
                                     const int* velocity_bc_type, const double* velocity_bc, const int* pressure_bc_type,
                                     double* big_m, double* rhs, double* masslump);

int orc_assemble_momentum_surface(const orc_mesh* m, const orc_surface* s, const orc_momentum_fields* f,
                                  const cgasm_momentum_opts* o, const int* findrm, const int* colm,
                                  const int* velocity_bc_type /*(dim,n_faces)*/, const double* velocity_bc /*(dim,sloc,n_faces)*/,
                                  const int* pressure_bc_type, double* big_m, double* rhs) {
  return orc_assemble_momentum_surface_ml(m, s, f, o, findrm, colm, velocity_bc_type, velocity_bc, pressure_bc_type, big_m, rhs, NULL);
}

/* the same with masslump(dim, N) (may be NULL), which the free-surface stabilisation adds to (:1167-1173) */
int orc_assemble_momentum_surface_ml(const orc_mesh* m, const orc_surface* s, const orc_momentum_fields* f,
                                     const cgasm_momentum_opts* o, const int* findrm, const int* colm,
                                     const int* velocity_bc_type, const double* velocity_bc, const int* pressure_bc_type,
                                     double* big_m, double* rhs, double* masslump) {
  const int dim = m->dim, sloc = s->sloc;
  const size_t nnz = (size_t)(findrm[m->n_nodes] - 1);
  for (int face = 1; face <= s->n_faces; face++) {
    const int* bt = velocity_bc_type + (size_t)dim * (face - 1);
    int sum = 0, any_internal = 0;
    for (int d = 0; d < dim; d++) {
      sum += bt[d];
      any_internal |= bt[d] == 3;
    }
    if (((bt[0] == 2 && sum == 2) || any_internal) && (!pressure_bc_type || pressure_bc_type[face - 1] == 0)) continue;
    double B[MAXDIM * MAXSLOC * MAXSLOC], r[MAXDIM * MAXSLOC], ml[MAXDIM * MAXSLOC];
    int st = orc_momentum_face_ml(m, s, f, o, face, bt, velocity_bc + (size_t)dim * sloc * (face - 1), B, r, ml);
    if (st) return st;
    const int* fn = s->sndgln + (size_t)sloc * (size_t)(face - 1);
    for (int i = 0; i < sloc; i++)
      for (int j = 0; j < sloc; j++) {
        int pos = csr_sparsity_pos(findrm, colm, fn[i], fn[j]);
        for (int d = 0; d < dim; d++) big_m[d * nnz + (size_t)(pos - 1)] += B[d + dim * (i + sloc * j)];
      }
    for (int i = 0; i < sloc; i++)
      for (int d = 0; d < dim; d++) {
        rhs[d + (size_t)dim * (fn[i] - 1)] += r[d + dim * i];
        if (masslump) masslump[d + (size_t)dim * (fn[i] - 1)] += ml[d + dim * i];
      }
  }
  return 0;
}

/* ====================================================================================
 * Lumped-mass pressure matrix C M_L^-1 C^T next to the path.
 * ==================================================================================== */

/* make_sparsity_mult, femtools/Sparsity_Patterns.F90:150-210, for mesh1 = mesh2 = mesh3 (P1-P1): the
 * second-order sparsity get_csr_sparsity_secondorder hands to cmc_m
 * (femtools/Sparsity_Patterns_Meshes.F90:122-148). Row i = ascending union of the first-order rows of
 * the nodes in first-order row i (insert_ascending :182). 1-based in and out; caller frees with orc_free. */
int orc_make_sparsity_mult(int n_nodes, const int* findrm, const int* colm, int** findrm_out, int** colm_out) {
  ilist_pool pool = {0};
  int* head = (int*)malloc(sizeof(int) * (size_t)n_nodes);
  int* length = (int*)calloc((size_t)n_nodes, sizeof(int));
  for (int i = 0; i < n_nodes; i++) head[i] = -1;
  /* do i = 1, count_2: row_1 = row_3 = first-order row of node i (lists of mesh2 x mesh1 / mesh3) */
  for (int i = 0; i < n_nodes; i++)
    for (int j = findrm[i] - 1; j < findrm[i + 1] - 1; j++)
      for (int k = findrm[i] - 1; k < findrm[i + 1] - 1; k++)
        insert_ascending(&pool, &head[colm[k] - 1], &length[colm[k] - 1], colm[j]);
  int* fr = (int*)malloc(sizeof(int) * (size_t)(n_nodes + 1));
  fr[0] = 1;
  for (int i = 0; i < n_nodes; i++) fr[i + 1] = fr[i] + length[i];
  const int nnz = fr[n_nodes] - 1;
  int* cm = (int*)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
  for (int i = 0; i < n_nodes; i++) {
    int p = fr[i] - 1;
    for (int node = head[i]; node >= 0; node = pool.pool[node].next) cm[p++] = pool.pool[node].value;
  }
  free(pool.pool);
  free(head);
  free(length);
  *findrm_out = fr;
  *colm_out = cm;
  return nnz;
}

/* mult_div_vector_div_T, femtools/Sparse_Matrices_Fields.F90:590-671, called by assemble_masslumped_cmc
 * (assemble/Assemble_CMC.F90:119-135) with matrix1 = ctp_m, matrix2 = ct_m (1 x dim blocks on the
 * first-order sparsity, sorted rows) and vfield = inverse_masslump(dim, N):
 *   product_ij = sum_k sum_d A_d(i,k) B_d(j,k) v(d,k), both rows walked left to right (:643-656).
 * ct1 / ct2: [dim][nnz]. product: values on the second-order sparsity, overwritten. */
void orc_mult_div_vector_div_T(int dim, int n_nodes, const int* findrm, const int* colm, const double* ct1,
                               const double* ct2, const double* vfield, const int* findrm2, const int* colm2,
                               double* product) {
  const size_t nnz = (size_t)(findrm[n_nodes] - 1);
  size_t nentry0 = 0;
  for (int i = 1; i <= n_nodes; i++) {
    const int r0 = findrm[i - 1] - 1, rn = findrm[i] - findrm[i - 1];
    for (int jcol = findrm2[i - 1] - 1; jcol < findrm2[i] - 1; jcol++) {
      const int j = colm2[jcol];
      const int c0 = findrm[j - 1] - 1, cn = findrm[j] - findrm[j - 1];
      double entry0 = 0.0;
      int k1 = 1, k2 = 1;
      while (k1 <= rn && k2 <= cn) {
        const int a = colm[r0 + k1 - 1], b = colm[c0 + k2 - 1];
        if (a < b) {
          k1 = k1 + 1;
        } else if (a == b) {
          for (int d = 0; d < dim; d++)
            entry0 = entry0 + ct1[d * nnz + (size_t)(r0 + k1 - 1)] * ct2[d * nnz + (size_t)(c0 + k2 - 1)] *
                                  vfield[d + (size_t)dim * (size_t)(a - 1)];
          k1 = k1 + 1;
          k2 = k2 + 1;
        } else {
          k2 = k2 + 1;
        }
      }
      product[nentry0++] = entry0;
    }
  }
}


/* ------------------------------------------------------------------------------------
 * P1-P1 pressure stabilisation, assemble_kmk_matrix (assemble/Momentum_CG.F90:2707-2766):
 *   kt = sum_e 0.5 dshape_tensor_dshape(dp_t, h_bar, dp_t, detwei)      on the first-order pressure sparsity,
 *   h_bar = edge_length_from_eigenvalue(simplex_tensor(X, ele))  (error_measures/Edge_lengths.F90:68-79),
 *   kmk = kt diag(1 / (theta_pg p_masslump)) kt^T  (mult_div_invscalar_div_T, femtools/Sparse_Matrices_Fields.F90:673-748).
 * The reference solves the metric's linear system with LAPACK DGESV (femtools/Vector_Tools.F90 solve) and takes the
 * eigen-decomposition with DSPEV (:401-440). LAPACK is an external library absent from the reference tree: both are
 * restated by their textbook algorithms (LU with partial pivoting; cyclic Jacobi rotations) and pinned by
 * tests/test_kmk.py against numpy.linalg (LAPACK itself) and the reference's own known answers
 * (error_measures/tests/test_simplex_tensor.F90, test_simplex_tensor_edgelens.F90).
 * ------------------------------------------------------------------------------------ */

/* simplex_tensor without `power`, femtools/Metric_tools.F90:852-941: the symmetric M with e^T M e = 1 for every edge e.
 * pos_ele(dim, loc); m(dim, dim). Returns 0, or 1 if the system is singular (degenerate element). */
int orc_simplex_tensor(int dim, const double* pos_ele, double* m) {
  const int loc = dim + 1, d = dim * (dim + 1) / 2;
  double A[36], x[6];
  /* idx(k,l) :919-933 (1-based) */
#define IDX_(k, l) ((((k) < (l) ? (k) : (l)) == 1) ? ((k) > (l) ? (k) : (l)) : ((k) > (l) ? (k) : (l)) + ((k) < (l) ? (k) : (l)) - (dim == 3 ? 0 : 1))
  int n = 0;
  for (int i = 0; i < loc; i++)
    for (int j = i + 1; j < loc; j++) {
      double diff[MAXDIM];
      for (int a = 0; a < dim; a++) diff[a] = pos_ele[a + dim * j] - pos_ele[a + dim * i];
      for (int k = 1; k <= dim; k++)
        for (int l = 1; l <= dim; l++) A[n + d * (IDX_(k, l) - 1)] = diff[k - 1] * diff[l - 1] * (k == l ? 1.0 : 2.0);
      n++;
    }
  for (int i = 0; i < d; i++) x[i] = 1.0;
  /* solve(A, x): LU with partial pivoting (DGESV) */
  for (int c = 0; c < d; c++) {
    int piv = c;
    for (int r = c + 1; r < d; r++)
      if (fabs(A[r + d * c]) > fabs(A[piv + d * c])) piv = r;
    if (A[piv + d * c] == 0.0) return 1;
    if (piv != c) {
      for (int q = 0; q < d; q++) {
        const double t = A[c + d * q];
        A[c + d * q] = A[piv + d * q];
        A[piv + d * q] = t;
      }
      const double t = x[c];
      x[c] = x[piv];
      x[piv] = t;
    }
    for (int r = c + 1; r < d; r++) {
      const double f = A[r + d * c] / A[c + d * c];
      for (int q = c + 1; q < d; q++) A[r + d * q] -= f * A[c + d * q];
      x[r] -= f * x[c];
    }
  }
  for (int c = d - 1; c >= 0; c--) {
    double t = x[c];
    for (int q = c + 1; q < d; q++) t -= A[c + d * q] * x[q];
    x[c] = t / A[c + d * c];
  }
  for (int i = 1; i <= dim; i++)
    for (int j = 1; j <= dim; j++) m[(i - 1) + dim * (j - 1)] = x[IDX_(i, j) - 1];
#undef IDX_
  return 0;
}

/* eigendecomposition_symmetric (Vector_Tools.F90:401-440: DSPEV) by cyclic Jacobi rotations. V: columns = eigenvectors. */
void orc_eig_symmetric(int dim, const double* M, double* V, double* evals) {
  double A[9];
  for (int i = 0; i < dim * dim; i++) {
    A[i] = M[i];
    V[i] = 0.0;
  }
  for (int i = 0; i < dim; i++) V[i + dim * i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < dim; i++)
      for (int j = 0; j < dim; j++) {
        if (i != j) off += A[i + dim * j] * A[i + dim * j];
        else diag += A[i + dim * i] * A[i + dim * i];
      }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < dim; p++)
      for (int q = p + 1; q < dim; q++) {
        const double apq = A[p + dim * q];
        if (apq == 0.0) continue;
        const double theta = (A[q + dim * q] - A[p + dim * p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < dim; k++) { /* A <- A J */
          const double akp = A[k + dim * p], akq = A[k + dim * q];
          A[k + dim * p] = c * akp - sn * akq;
          A[k + dim * q] = sn * akp + c * akq;
        }
        for (int k = 0; k < dim; k++) { /* A <- J^T A */
          const double apk = A[p + dim * k], aqk = A[q + dim * k];
          A[p + dim * k] = c * apk - sn * aqk;
          A[q + dim * k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < dim; k++) {
          const double vkp = V[k + dim * p], vkq = V[k + dim * q];
          V[k + dim * p] = c * vkp - sn * vkq;
          V[k + dim * q] = sn * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < dim; i++) evals[i] = A[i + dim * i];
}

/* edge_length_from_eigenvalue_metric, femtools/Metric_tools.F90:157-164: V diag(1/sqrt|lambda|) V^T */
void orc_edge_length_from_metric(int dim, const double* metric, double* edge) {
  double V[9], ev[3];
  orc_eig_symmetric(dim, metric, V, ev);
  for (int i = 0; i < dim; i++) ev[i] = 1.0 / sqrt(fabs(ev[i])); /* :145-155 */
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) { /* eigenrecomposition: M = V A V^T */
      double s = 0.0;
      for (int k = 0; k < dim; k++) s += V[i + dim * k] * ev[k] * V[j + dim * k];
      edge[i + dim * j] = s;
    }
}

/* The element loop of assemble_kmk_matrix (:2748-2753) on the pressure (= coordinate, P1) mesh, plus the lumped
 * pressure mass it divides by (get_lumped_mass -> compute_lumped_mass: row sums of shape_shape). kt(nnz), p_masslump(N)
 * are overwritten. */
int orc_assemble_kt(const orc_mesh* m, const int* findrm, const int* colm, double* kt, double* p_masslump) {
  const int dim = m->dim, loc = m->loc, ngi = m->ngi;
  const size_t nnz = (size_t)(findrm[m->n_nodes] - 1);
  memset(kt, 0, sizeof(double) * nnz);
  memset(p_masslump, 0, sizeof(double) * (size_t)m->n_nodes);
  for (int ele = 1; ele <= m->n_elements; ele++) {
    const int* nd = ele_nodes(m, ele);
    double X_val[MAXDIM * MAXLOC], dp_t[MAXLOC * MAXNGI * MAXDIM], detwei[MAXNGI], h_bar[MAXDIM * MAXDIM * MAXNGI],
        ele_tensor[MAXDIM * MAXDIM], edge[MAXDIM * MAXDIM], little[MAXLOC * MAXLOC];
    for (int i = 0; i < loc; i++)
      for (int a = 0; a < dim; a++) X_val[a + dim * i] = m->X[a + dim * (size_t)(nd[i] - 1)];
    orc_transform_to_physical(dim, ngi, X_val, m->dn, m->weight, dp_t, detwei, NULL);
    if (orc_simplex_tensor(dim, X_val, ele_tensor)) return CGASM_EARG;
    orc_edge_length_from_metric(dim, ele_tensor, edge);
    for (int g = 0; g < ngi; g++) /* spread(..., 3, ngi) */
      for (int a = 0; a < dim * dim; a++) h_bar[a + dim * dim * g] = edge[a];
    dshape_tensor_dshape(dim, loc, ngi, dp_t, h_bar, detwei, little);
    for (int i = 0; i < loc; i++)
      for (int j = 0; j < loc; j++) {
        const double v = 0.5 * little[i + loc * j];
        if (v == 0) continue;
        kt[csr_sparsity_pos(findrm, colm, nd[i], nd[j]) - 1] += v;
      }
    for (int i = 0; i < loc; i++) {
      double rowsum = 0.0;
      for (int j = 0; j < loc; j++) {
        double mij = 0.0;
        for (int g = 0; g < ngi; g++) mij += m->n[i + loc * g] * m->n[j + loc * g] * detwei[g];
        rowsum += mij;
      }
      p_masslump[nd[i] - 1] += rowsum;
    }
  }
  return 0;
}

/* mult_div_invscalar_div_T, femtools/Sparse_Matrices_Fields.F90:673-748: product = m1 diag(1/s) m2^T on the
 * second-order sparsity; entry0 + row_val(k1)*col_val(k2)/s(row(k1)) (:729-730), rows walked left to right. */
void orc_mult_div_invscalar_div_T(int n_nodes, const int* findrm, const int* colm, const double* m1, const double* sfield,
                                  const double* m2, const int* findrm2, const int* colm2, double* product) {
  size_t nentry0 = 0;
  for (int i = 1; i <= n_nodes; i++) {
    const int r0 = findrm[i - 1] - 1, rn = findrm[i] - findrm[i - 1];
    for (int jcol = findrm2[i - 1] - 1; jcol < findrm2[i] - 1; jcol++) {
      const int j = colm2[jcol];
      const int c0 = findrm[j - 1] - 1, cn = findrm[j] - findrm[j - 1];
      double entry0 = 0.0;
      int k1 = 1, k2 = 1;
      while (k1 <= rn && k2 <= cn) {
        const int a = colm[r0 + k1 - 1], b = colm[c0 + k2 - 1];
        if (a < b) {
          k1 = k1 + 1;
        } else if (a == b) {
          entry0 = entry0 + m1[r0 + k1 - 1] * m2[c0 + k2 - 1] / sfield[a - 1];
          k1 = k1 + 1;
          k2 = k2 + 1;
        } else {
          k2 = k2 + 1;
        }
      }
      product[nentry0++] = entry0;
    }
  }
}


/* The continuity half of construct_momentum_surface_element_cg, assemble/Momentum_CG.F90:1073-1111, taken when
 * integrate_continuity_by_parts and (assemble_ct_matrix_here or include_pressure_and_continuity_bcs): on faces that are
 * neither no-normal-flow nor free-surface (:1075) ct_mat_bdy = shape_shape_vector(p_shape, u_shape, detwei_bdy,
 * normal_bdy) (:1080); per component: a weak Dirichlet velocity (type 1) with include_pressure_and_continuity_bcs moves
 * -ct_mat_bdy . velocity_bc to ct_rhs (:1084-1086), otherwise the block is added to ct_m (:1087-1088); a pressure
 * condition adds -(pressure_bc [- hb_pressure]) . ct_mat_bdy to the momentum rhs (:1090-1098). Outputs overwritten:
 * ct_addto(dim, sloc, sloc) [p node, u node], ct_rhs_addto(sloc), rhs_addto(dim, sloc). pressure_bc / hb_pressure:
 * ele_val on the face (sloc), may be NULL (= 0). */
int orc_momentum_face_ct(const orc_mesh* m, const orc_surface* s, const cgasm_momentum_opts* o, int face,
                         const int* velocity_bc_type, const double* velocity_bc, int pressure_bc_type,
                         const double* pressure_bc, const double* hb_pressure, int include_pressure_and_continuity_bcs,
                         double* ct_addto, double* ct_rhs_addto, double* rhs_addto) {
  const int dim = m->dim, sloc = s->sloc, sngi = s->sngi;
  for (int k = 0; k < dim * sloc * sloc; k++) ct_addto[k] = 0.0;
  for (int k = 0; k < sloc; k++) ct_rhs_addto[k] = 0.0;
  for (int k = 0; k < dim * sloc; k++) rhs_addto[k] = 0.0;
  if (!(o->integrate_continuity_by_parts && (o->assemble_ct_matrix_here || include_pressure_and_continuity_bcs))) return 0;
  if (velocity_bc_type[0] == 2 || velocity_bc_type[0] == 4) return 0;
  double detwei[MAXSNGI], normal[MAXDIM * MAXSNGI], ct_mat_bdy[MAXDIM * MAXSLOC * MAXSLOC];
  face_geometry(m, s, face, detwei, normal);
  shape_shape_vector2(dim, sloc, sngi, s->n_f, s->n_f, detwei, normal, ct_mat_bdy);
#define CB_(d, i, j) ct_mat_bdy[(d) + dim * ((i) + sloc * (j))]
  for (int d = 0; d < dim; d++) {
    if (include_pressure_and_continuity_bcs && velocity_bc_type[d] == 1) {
      for (int i = 0; i < sloc; i++) {
        double v = 0.0;
        for (int j = 0; j < sloc; j++) v += CB_(d, i, j) * velocity_bc[d + dim * j];
        ct_rhs_addto[i] += -v;
      }
    } else if (o->assemble_ct_matrix_here) {
      for (int j = 0; j < sloc; j++)
        for (int i = 0; i < sloc; i++) ct_addto[d + dim * (i + sloc * j)] += CB_(d, i, j);
    }
    if (pressure_bc_type > 0) {
      for (int j = 0; j < sloc; j++) { /* matmul(vector(ploc), ct_mat_bdy(dim,:,:)) -> u node j */
        double v = 0.0;
        for (int i = 0; i < sloc; i++) {
          double pv = pressure_bc ? pressure_bc[i] : 0.0;
          if (o->subtract_out_reference_profile && hb_pressure) pv -= hb_pressure[i];
          v += pv * CB_(d, i, j);
        }
        rhs_addto[d + dim * j] += -v;
      }
    }
  }
#undef CB_
  return 0;
}

/* Adds the continuity boundary blocks of every face to ct_m [dim][nnz] (rows = pressure nodes): the loop
 * :795-812 restricted to the ct_m part, with the skip rule :799-803. */
int orc_assemble_ct_surface(const orc_mesh* m, const orc_surface* s, const cgasm_momentum_opts* o, const int* findrm,
                            const int* colm, const int* velocity_bc_type, const int* pressure_bc_type, double* ct_m) {
  const int dim = m->dim, sloc = s->sloc;
  const size_t nnz = (size_t)(findrm[m->n_nodes] - 1);
  const double zero[MAXDIM * MAXSLOC] = {0};
  for (int face = 1; face <= s->n_faces; face++) {
    const int* bt = velocity_bc_type + (size_t)dim * (face - 1);
    int sum = 0, any_internal = 0;
    for (int d = 0; d < dim; d++) {
      sum += bt[d];
      any_internal |= bt[d] == 3;
    }
    const int pt = pressure_bc_type ? pressure_bc_type[face - 1] : 0;
    if (((bt[0] == 2 && sum == 2) || any_internal) && pt == 0) continue;
    double C[MAXDIM * MAXSLOC * MAXSLOC], cr[MAXSLOC], r[MAXDIM * MAXSLOC];
    int st = orc_momentum_face_ct(m, s, o, face, bt, zero, pt, NULL, NULL, 0, C, cr, r);
    if (st) return st;
    const int* fn = s->sndgln + (size_t)sloc * (size_t)(face - 1);
    for (int i = 0; i < sloc; i++)
      for (int j = 0; j < sloc; j++) {
        int pos = csr_sparsity_pos(findrm, colm, fn[i], fn[j]);
        for (int d = 0; d < dim; d++) ct_m[d * nnz + (size_t)(pos - 1)] += C[d + dim * (i + sloc * j)];
      }
  }
  return 0;
}

/* Thread count of the OpenMP loops above (bench.py's CPU legs: all cores and one core). */
#ifdef _OPENMP
#include <omp.h>
void orc_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
#else
void orc_set_threads(int n) { (void)n; }
#endif

/* correct_masslumped_velocity, assemble/Momentum_CG.F90:2544-2575: per component d
 *   delta_u = mult_T(block(ct_m, 1, d), delta_p)      femtools/Sparse_Tools.F90:3889-3916: rows of ct_m ascending,
 *                                                      vector_out(colm(j)) += val(j) * vector_in(i)
 *   delta_u = delta_u * inverse_masslump(d, :)         scale
 *   u(d, :) = u(d, :) + delta_u                        addto
 * ct_m [dim][nnz] on the first-order sparsity (rows = pressure nodes), inverse_masslump and u (dim, n_nodes). */
void orc_correct_masslumped_velocity(int dim, int n_nodes, const int* findrm, const int* colm, const double* ct_m,
                                     const double* inverse_masslump, const double* delta_p, double* u) {
  const size_t nnz = (size_t)(findrm[n_nodes] - 1);
  double* delta_u = (double*)malloc(sizeof(double) * (size_t)n_nodes);
  for (int d = 0; d < dim; d++) {
    for (int k = 0; k < n_nodes; k++) delta_u[k] = 0.0;
    for (int i = 1; i <= n_nodes; i++)
      for (int j = findrm[i - 1]; j <= findrm[i] - 1; j++) {
        const int k = colm[j - 1];
        delta_u[k - 1] = delta_u[k - 1] + ct_m[d * nnz + (size_t)(j - 1)] * delta_p[i - 1];
      }
    for (int k = 0; k < n_nodes; k++) {
      delta_u[k] = delta_u[k] * inverse_masslump[d + (size_t)dim * k];
      u[d + (size_t)dim * k] = u[d + (size_t)dim * k] + delta_u[k];
    }
  }
  free(delta_u);
}

/* Strong Dirichlet conditions on big_m: apply_dirichlet_conditions_vector_petsc_csr, femtools/Boundary_Conditions.F90
 * :2198-2218 = collect_vector_dirichlet_conditions (:2125-2178: rhs(d, node) = value) + lift_boundary_conditions,
 * femtools/Sparse_Tools_Petsc.F90:1139-1254. The latter is PETSc's MatZeroRowsColumns(A, rows, pivot = 1.0, x, b)
 * (PETSc is an un-vendored dependency, debian/control names 3.8.3; its documented algorithm: for every listed row r
 * b_r = pivot * x_r, for every other row i b_i -= A_ir x_r, row r and column r are zeroed, A_rr = pivot; x is a copy of b
 * taken beforehand, :1185-1190) followed by fix_scaling (:1213-1235): A_rr = its old diagonal value and
 * b_r = old diagonal * b_r. big_m is held as its dim diagonal blocks [dim][nnz] (block_mask, Momentum_CG.F90:1293-1300),
 * rhs (dim, n_nodes). bc_nodes / bc_comps (1-based) list the (node, component) pairs with a strong condition. */
void orc_lift_boundary_conditions(int dim, int n_nodes, const int* findrm, const int* colm, double* big_m, double* rhs,
                                  int nbc, const int* bc_nodes, const int* bc_comps) {
  const size_t nnz = (size_t)(findrm[n_nodes] - 1);
  char* flag = (char*)calloc((size_t)dim * (size_t)n_nodes, 1);
  double* x = (double*)malloc(sizeof(double) * (size_t)dim * (size_t)n_nodes);
  for (size_t k = 0; k < (size_t)dim * (size_t)n_nodes; k++) x[k] = rhs[k]; /* xvec = copy of bvec */
  for (int b = 0; b < nbc; b++) flag[(bc_comps[b] - 1) + (size_t)dim * (bc_nodes[b] - 1)] = 1;
  for (int d = 0; d < dim; d++)
    for (int i = 1; i <= n_nodes; i++) {
      const int mine = flag[d + (size_t)dim * (i - 1)];
      double old_diag = 0.0;
      for (int j = findrm[i - 1]; j <= findrm[i] - 1; j++) {
        const int c = colm[j - 1];
        double* a = &big_m[d * nnz + (size_t)(j - 1)];
        if (mine) {
          if (c == i) old_diag = *a; /* kept: MatSetValue(j, j, old_diagonal_values) */
          else *a = 0.0;
        } else if (flag[d + (size_t)dim * (c - 1)]) {
          rhs[d + (size_t)dim * (i - 1)] = rhs[d + (size_t)dim * (i - 1)] - *a * x[d + (size_t)dim * (c - 1)];
          *a = 0.0;
        }
      }
      if (mine) rhs[d + (size_t)dim * (i - 1)] = old_diag * (1.0 * x[d + (size_t)dim * (i - 1)]);
    }
  free(flag);
  free(x);
}
