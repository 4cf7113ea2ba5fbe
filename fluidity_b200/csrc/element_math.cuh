// element_math.cuh -- per-element FP64 math of the two CG element routines, written for one
// thread = one element with everything in registers.
//
// What is computed (formulae: SURVEY.md 8(a); reference lines cited per block) is the same
// local matrix / rhs as construct_momentum_element_cg (assemble/Momentum_CG.F90:1193-1490)
// and assemble_advection_diffusion_element_cg (assemble/Advection_Diffusion_CG.F90:702-865),
// but NOT evaluated the way the Fortran does it: for P1 simplices grad N is constant over
// the element, so every "shape x dshape" contraction collapses from loc*loc*ngi*dim MACs
// to a few rank-1 updates:
//     A_ij = sum_g N_ig (u_g . gradN_j) c_g      = C_i . gradN_j,  C_i = sum_g N_ig c_g u_g
//     K_ij = sum_g (gradN_i . gradN_j) mu_g dw_g = (mubar gradN_i) . gradN_j
//     sum_j M_ij = sum_g N_ig c_g                  (partition of unity)
// Summation order therefore differs from the reference; parity is 1e-12 relative
// (tests/test_parity_gpu.py), which is what the north-star asks for.
#pragma once
#include <cuda_runtime.h>
#include "../../include/cgasm.h"

namespace cgasm {

template <int DIM>
struct Shape {
  static constexpr int LOC = DIM + 1;
  static constexpr int NGI = (DIM == 3) ? 5 : 4;  // degree-3 rules, Quadrature.F90:690-708,951-970
};

// Reference-element tables, passed by value inside the kernel parameter block so that every
// N[i][g] / w[g] is a constant-bank operand of the DFMA that uses it.
struct Tables {
  double N[4 * 5];  // n(i,g) at [i*NGI + g]
  double w[5];
  // Moments of a node-symmetric rule (filled and verified at cgasm_create, sym = 1 if they hold):
  //   P_ik  = sum_g N_ig N_kg w_g       = Pd (i == k) | Po (i != k)
  //   Q_ikl = sum_g N_ig N_kg N_lg w_g  = Qaaa | Qaab (two equal) | Qabc (all distinct)
  //   W1    = sum_g N_ig w_g (same for every i),  Wsum = sum_g w_g
  double Pd, Po, Qaaa, Qaab, Qabc, W1, Wsum;
  int sym;
};

// Device view of one nodal field (femtools/Fields_Data_Types.F90:154-233).
// stride = 0 for FIELD_TYPE_CONSTANT (every node reads node 0), else components per node.
struct FieldView {
  const double* __restrict__ val;
  int stride;
};

// Packed per-node records of the six fields every element reads (device-private layout; the
// caller's arrays keep the reference layout). One record = one 32-byte sector = one 256-bit
// load, instead of dim+1 scattered 8-byte loads that each drag a whole sector through L2:
//   r0[node] = { X(1..dim), [0], T }     r1[node] = { nu(1..dim), [0], density }
//   r2[node] = { oldu(1..dim), [0], buoyancy }
// CONSTANT fields are broadcast into the records when they are set.
struct NodeRecs {
  const double4* __restrict__ r0;
  const double4* __restrict__ r1;
  const double4* __restrict__ r2;
};

struct MomentumArgs {
  Tables tab;
  cgasm_momentum_opts o;
  const int4* __restrict__ ndglno;  // 0-based, padded to 4 ints
  NodeRecs rec;
  FieldView viscosity, hb_density, gravity, absorption, source;
  int n_elements;
};

struct AdvDiffArgs {
  Tables tab;
  cgasm_advdiff_opts o;
  const int4* __restrict__ ndglno;
  NodeRecs rec;
  FieldView source, absorption, diffusivity;
  int n_elements;
};

__device__ __forceinline__ double4 ld256(const double4* p) {
  double4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

template <int DIM>
__device__ __forceinline__ void unpack(const double4& r, double (&v)[DIM], double& s) {
  v[0] = r.x;
  v[1] = r.y;
  if constexpr (DIM == 3) v[2] = r.z;
  s = r.w;
}

__device__ __forceinline__ int node_of(const int4& nd, int i) {
  return i == 0 ? nd.x : (i == 1 ? nd.y : (i == 2 ? nd.z : nd.w));
}

template <int NC>
__device__ __forceinline__ void gather(const FieldView& f, int node, double (&out)[NC]) {
  const double* p = f.val + (size_t)f.stride * (size_t)node;
#pragma unroll
  for (int c = 0; c < NC; c++) out[c] = __ldg(p + c);
}

// ---- geometry: transform_to_physical, femtools/Transform_elements.F90:807-887 -------------
// J_T(a,k) = x_k[a] - x_loc[a] (dn(i,:,k) = delta_ik, dn(loc,:,k) = -1), cofactor inverse,
// detJ by expanding the first column, gradN_i = invJ(:,i), gradN_loc = -sum_i gradN_i.
template <int DIM>
struct Geom {
  double grad[DIM + 1][DIM];  // dN_i/dx_a
  double absdet;              // |detJ|  (detwei_g = absdet * w_g)
  double JT[DIM][DIM];        // J_T(a,k) = dx_a/dxi_k
};

template <int DIM>
__device__ __forceinline__ void geometry(const double (&X)[DIM + 1][DIM], Geom<DIM>& G) {
#pragma unroll
  for (int a = 0; a < DIM; a++)
#pragma unroll
    for (int k = 0; k < DIM; k++) G.JT[a][k] = X[k][a] - X[DIM][a];
  double C[DIM][DIM];  // cofactors: invJ(a,k)*detJ
  if constexpr (DIM == 2) {
    C[0][0] = G.JT[1][1];
    C[1][0] = -G.JT[0][1];
    C[0][1] = -G.JT[1][0];
    C[1][1] = G.JT[0][0];
  } else {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        C[i][k] = G.JT[i1][k1] * G.JT[i2][k2] - G.JT[i2][k1] * G.JT[i1][k2];
      }
  }
  double det = 0.0;
#pragma unroll
  for (int a = 0; a < DIM; a++) det += G.JT[a][0] * C[a][0];
  const double rdet = 1.0 / det;
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      G.grad[k][a] = C[a][k] * rdet;  // matmul(invJ, e_k)
      s -= G.grad[k][a];
    }
    G.grad[DIM][a] = s;
  }
  G.absdet = fabs(det);
}

// f_g = sum_i f_i N_ig  (ele_val_at_quad, femtools/Fields_Base.F90:2256-2310)
template <int DIM>
__device__ __forceinline__ void at_quad(const Tables& t, const double (&f)[DIM + 1],
                                        double (&q)[Shape<DIM>::NGI]) {
  constexpr int LOC = DIM + 1, NGI = Shape<DIM>::NGI;
#pragma unroll
  for (int g = 0; g < NGI; g++) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < LOC; i++) s += f[i] * t.N[i * NGI + g];
    q[g] = s;
  }
}

// ---- upwind stabilisation (assemble/Upwind_Stabilisation.F90) ------------------------------------
// nu_bar_scaled_q :225-320 with xi_optimal :133-162, xi_doubly_asymptotic :164-192,
// xi_critical_rule :194-223. Jm(a,k) = J(a,k,g) (constant over a P1 element); diff = NULL-like flag
// have_diff false => NU_BAR_UNITY (:248-251). inverse() of the dim x dim diffusivity by cofactors.
template <int DIM>
__device__ __forceinline__ void small_inverse(const double (&A)[DIM * DIM], double (&B)[DIM * DIM]) {
  // cofactors times ONE reciprocal of the determinant (the reference's inverse() divides entry by entry: the results
  // differ in the last bit, inside the 1e-12 of the parity tests; nine divisions per quadrature point were a fifth of
  // the stabilised element kernels' instructions)
  if constexpr (DIM == 2) {
    const double rd = 1.0 / (A[0] * A[3] - A[2] * A[1]);
    B[0] = A[3] * rd;
    B[1] = -A[1] * rd;
    B[2] = -A[2] * rd;
    B[3] = A[0] * rd;
  } else {
#define A_(i, j) A[(i) + 3 * (j)]
    const double c00 = A_(1, 1) * A_(2, 2) - A_(1, 2) * A_(2, 1);
    const double c01 = A_(1, 2) * A_(2, 0) - A_(1, 0) * A_(2, 2);
    const double c02 = A_(1, 0) * A_(2, 1) - A_(1, 1) * A_(2, 0);
    const double rd = 1.0 / (A_(0, 0) * c00 + A_(0, 1) * c01 + A_(0, 2) * c02);
    B[0 + 3 * 0] = c00 * rd;
    B[1 + 3 * 0] = c01 * rd;
    B[2 + 3 * 0] = c02 * rd;
    B[0 + 3 * 1] = (A_(0, 2) * A_(2, 1) - A_(0, 1) * A_(2, 2)) * rd;
    B[1 + 3 * 1] = (A_(0, 0) * A_(2, 2) - A_(0, 2) * A_(2, 0)) * rd;
    B[2 + 3 * 1] = (A_(0, 1) * A_(2, 0) - A_(0, 0) * A_(2, 1)) * rd;
    B[0 + 3 * 2] = (A_(0, 1) * A_(1, 2) - A_(0, 2) * A_(1, 1)) * rd;
    B[1 + 3 * 2] = (A_(0, 2) * A_(1, 0) - A_(0, 0) * A_(1, 2)) * rd;
    B[2 + 3 * 2] = (A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(1, 0)) * rd;
#undef A_
  }
}

// JD = J . inverse(diff): constant over the element when the diffusivity is (P1: J is), so hoisted out of the
// quadrature loop by Stabilisation::setup
template <int DIM>
__device__ __forceinline__ void j_inverse_diff(const double (&Jm)[DIM][DIM], const double (&diff)[DIM * DIM], double (&JD)[DIM][DIM]) {
  double inv[DIM * DIM];
  small_inverse<DIM>(diff, inv);
#pragma unroll
  for (int a = 0; a < DIM; a++)
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double jd = 0.0;
#pragma unroll
      for (int b = 0; b < DIM; b++) jd += Jm[a][b] * inv[b + DIM * k];
      JD[a][k] = jd;
    }
}

// xi_optimal (Upwind_Stabilisation.F90:133-162) between its cut-offs: coth(p) - 1/p. ncu (source page of the stabilised
// tracer element kernel): tanh + the divisions behind it were 45 % of the kernel's instructions, fifteen evaluations per
// element. |p| < 1/2: the Laurent series of coth minus its pole (Bernoulli numbers; the next term is below 1e-17
// relative), which has none of the cancellation of 1/tanh(p) - 1/p; otherwise e = exp(-2|p|) and
// ((1 + e)|p| - (1 - e)) / (|p| (1 - e)): one exp and one reciprocal. Against the reference's formula evaluated in FP64
// the difference is the reference's own cancellation error, <= 3e-16 / p^2 relative to xi ~ p/3: far below 1e-12 of any
// assembled entry (the stabilisation term is O(p^2) of the diffusion term there).
__device__ __forceinline__ double xi_optimal_mid(double p) {
  const double a = fabs(p);
  if (a < 0.5) {
    const double q = p * p;
    double s = -349222.0 / 1531329465290625.0;
    s = fma(s, q, 87734.0 / 38979295480125.0);
    s = fma(s, q, -3617.0 / 162820783125.0);
    s = fma(s, q, 4.0 / 18243225.0);
    s = fma(s, q, -1382.0 / 638512875.0);
    s = fma(s, q, 2.0 / 93555.0);
    s = fma(s, q, -1.0 / 4725.0);
    s = fma(s, q, 2.0 / 945.0);
    s = fma(s, q, -1.0 / 45.0);
    s = fma(s, q, 1.0 / 3.0);
    return p * s;
  }
  const double e = exp(-2.0 * a);
  const double den = a * (1.0 - e);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));  // den >= 0.31: a normal number; two Newton steps
  const double e0 = fma(-den, r, 1.0);
  r = fma(r, e0, r);
  r = fma(r, e0 * e0, r);
  return copysign(fma(a, 1.0 + e, e - 1.0) * r, p);
}

template <int DIM>
__device__ __forceinline__ double nu_bar_scaled(const double (&u)[DIM], const double (&Jm)[DIM][DIM], bool have_diff,
                                                const double (&JD)[DIM][DIM], int scheme, double scale) {
  const double tolerance = 1.0e-10, tanh_tolerance = 11.859499013855018;  // :50-51
  double norm_u = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; d++) norm_u += u[d] * u[d];
  if (norm_u < tolerance) return 0.0;
  double uJ[DIM];
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) s += u[a] * Jm[a][k];
    uJ[k] = s;
  }
  double val = 0.0;
  if (!have_diff || scheme == CGASM_NU_BAR_UNITY) {
#pragma unroll
    for (int k = 0; k < DIM; k++) val += fabs(uJ[k]);
  } else {
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double p = 0.0;  // pe = 0.5 * u . (J . inverse(diff))
#pragma unroll
      for (int a = 0; a < DIM; a++) p += u[a] * JD[a][k];
      p *= 0.5;
      double xi;
      if (scheme == CGASM_NU_BAR_OPTIMAL) {
        if (fabs(p) < tolerance) xi = 0.0;
        else if (p > tanh_tolerance) xi = 1.0 - (1.0 / p);
        else if (p < -tanh_tolerance) xi = -1.0 - (1.0 / p);
        else xi = xi_optimal_mid(p);
      } else if (scheme == CGASM_NU_BAR_DOUBLY_ASYMPTOTIC) {
        if (fabs(p) <= 3.0) xi = p / 3.0;
        else xi = p > 0.0 ? 1.0 : -1.0;
      } else {
        if (fabs(p) <= 1.0) xi = 0.0;
        else xi = p > 0.0 ? 1.0 - 1.0 / p : -1.0 - 1.0 / p;
      }
      val += xi * uJ[k];
    }
  }
  return val / norm_u * scale;
}

// Everything the stabilised element routines need beyond the Galerkin terms:
//   STAB 1 (SU):   stab_ij = sum_g (u_g.gradN_i)(u_g.gradN_j) nubar_g detwei_g      (:82-131)
//   STAB 2 (SUPG): test function n(i,g) + nubar_g (u_g.gradN_i)                      (:418-454)
template <int DIM, int STAB>
struct Stabilisation {
  static constexpr int LOC = DIM + 1, NGI = Shape<DIM>::NGI;
  double nt[STAB == 2 ? LOC * NGI : 1];    // SUPG test function
  double udn[STAB == 1 ? LOC * NGI : 1];   // u_g . gradN_i
  double wq[STAB == 1 ? NGI : 1];          // nubar_g * detwei_g
  __device__ __forceinline__ double test(const Tables& t, int i, int g) const {
    if constexpr (STAB == 2) return nt[i * NGI + g];
    else return t.N[i * NGI + g];
  }
  // ug: u at quadrature points; diffq(g): diffusivity at g (dim x dim, column-major) if have_diff
  // diff_const: the diffusivity does not vary over the element (a constant field): J . inverse(diff) once per element
  template <class DiffAt>
  __device__ __forceinline__ void setup(const Tables& t, const Geom<DIM>& G, const double (&ug)[NGI][DIM], bool have_diff,
                                        bool diff_const, DiffAt diff_at, int scheme, double scale) {
    if constexpr (STAB != 0) {
      double Jm[DIM][DIM];  // J(:,:,gi) = transpose(J_local_T), Transform_elements.F90:878-882
#pragma unroll
      for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int k = 0; k < DIM; k++) Jm[a][k] = G.JT[k][a];
      const bool need_inv = have_diff && scheme != CGASM_NU_BAR_UNITY;
      double JD[DIM][DIM];
#pragma unroll
      for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int k = 0; k < DIM; k++) JD[a][k] = 0.0;
      if (need_inv && diff_const) {
        double dq[DIM * DIM];
#pragma unroll
        for (int ab = 0; ab < DIM * DIM; ab++) dq[ab] = 0.0;
        diff_at(0, dq);
        j_inverse_diff<DIM>(Jm, dq, JD);
      }
#pragma unroll
      for (int g = 0; g < NGI; g++) {
        if (need_inv && !diff_const) {
          double dq[DIM * DIM];
#pragma unroll
          for (int ab = 0; ab < DIM * DIM; ab++) dq[ab] = 0.0;
          diff_at(g, dq);
          j_inverse_diff<DIM>(Jm, dq, JD);
        }
        const double nb = nu_bar_scaled<DIM>(ug[g], Jm, have_diff, JD, scheme, scale);
#pragma unroll
        for (int i = 0; i < LOC; i++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += ug[g][a] * G.grad[i][a];
          if constexpr (STAB == 2) nt[i * NGI + g] = t.N[i * NGI + g] + nb * s;
          if constexpr (STAB == 1) udn[i * NGI + g] = s;
        }
        if constexpr (STAB == 1) wq[g] = nb * G.absdet * t.w[g];
      }
    }
  }
  __device__ __forceinline__ double su(int i, int j) const {
    double s = 0.0;
    if constexpr (STAB == 1) {
#pragma unroll
      for (int g = 0; g < NGI; g++) s += (udn[i * NGI + g] * wq[g]) * udn[j * NGI + g];
    }
    return s;
  }
};

// ---- momentum -----------------------------------------------------------------------------
// Result of one element. L is the part of the diagonal blocks common to every velocity
// component; Labs[d] is added to block d only (non-lumped absorption); diag[d][i] is the
// lumped diagonal (Momentum_CG.F90:1462, add_diagonal_to_tensor).
template <int DIM, bool LABS>
struct MomentumLocal {
  static constexpr int LOC = DIM + 1;
  double L[LOC][LOC];
  double Labs[LABS ? DIM : 1][LOC][LOC];  // only kernels built with LABS carry it
  double diag[DIM][LOC];
  double rhs[DIM][LOC];
  double ml[DIM][LOC];
};

// LABS must be true iff (have_absorption && !lump_absorption); the launcher picks the
// instantiation, so kernels without that option never hold the dim*loc*loc extra block.
// STAB: CGASM_STAB_* (0 none, 1 streamline upwind, 2 SUPG), Momentum_CG.F90:1345-1372, :1686-1708.
template <int DIM, bool LABS, int STAB = 0>
__device__ __forceinline__ void momentum_element(const MomentumArgs& A, const int4 nd,
                                                 MomentumLocal<DIM, LABS>& R, Geom<DIM>& G) {
  constexpr int LOC = DIM + 1, NGI = Shape<DIM>::NGI;
  const cgasm_momentum_opts& o = A.o;
  const Tables& t = A.tab;
  const double dtt = o.dt * o.theta;

  double X[LOC][DIM], nu[LOC][DIM], oldu[LOC][DIM], rho[LOC], buoy[LOC];
#pragma unroll
  for (int i = 0; i < LOC; i++) {
    const int node = node_of(nd, i);
    double unused;
    unpack<DIM>(ld256(A.rec.r0 + node), X[i], unused);
    unpack<DIM>(ld256(A.rec.r1 + node), nu[i], rho[i]);
    unpack<DIM>(ld256(A.rec.r2 + node), oldu[i], buoy[i]);
  }
  geometry<DIM>(X, G);
  double wsum = 0.0;  // sum_g w_g (compile-time foldable: tables sit in the constant bank)
#pragma unroll
  for (int g = 0; g < NGI; g++) wsum += t.w[g];

  // c_g = rho_g * detwei_g  (coefficient_detwei, :1536, :1674)
  double c[NGI];
  at_quad<DIM>(t, rho, c);
#pragma unroll
  for (int g = 0; g < NGI; g++) c[g] *= G.absdet * t.w[g];

#pragma unroll
  for (int i = 0; i < LOC; i++)
#pragma unroll
    for (int j = 0; j < LOC; j++) R.L[i][j] = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; d++)
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      R.diag[d][i] = 0.0;
      R.rhs[d][i] = 0.0;
      R.ml[d][i] = 0.0;
    }

  // relu_gi = ele_val_at_quad(nu) (:1347, :1639)
  double ug[NGI][DIM];
#pragma unroll
  for (int g = 0; g < NGI; g++)
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < LOC; i++) s += nu[i][a] * t.N[i * NGI + g];
      ug[g][a] = s;
    }
  Stabilisation<DIM, STAB> ST;
  if constexpr (STAB != 0) {
    // diff_q = viscosity at the quadrature points with the off-diagonal entries zeroed (:1351-1360)
    auto diff_at = [&](int g, double (&dq)[DIM * DIM]) {
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        double vv = 0.0;
        if (A.viscosity.stride == 0) {
          vv = __ldg(A.viscosity.val + a + DIM * a);
        } else {
#pragma unroll
          for (int i = 0; i < LOC; i++)
            vv += __ldg(A.viscosity.val + (size_t)A.viscosity.stride * node_of(nd, i) + a + DIM * a) * t.N[i * NGI + g];
        }
        dq[a + DIM * a] = vv;
      }
    };
    ST.setup(t, G, ug, o.have_viscosity != 0, A.viscosity.stride == 0, diff_at, o.nu_bar_scheme, o.nu_bar_scale);
  }

  // v[i] accumulates the vector such that (A + K)_ij = v[i] . gradN_j
  double v[LOC][DIM];
#pragma unroll
  for (int i = 0; i < LOC; i++)
#pragma unroll
    for (int a = 0; a < DIM; a++) v[i][a] = 0.0;
  bool have_v = false;

  // Mass (add_mass_element_cg, :1492-1600)
  if (o.assemble_inverse_masslump || !o.exclude_mass) {
    if (!o.exclude_mass && !o.lump_mass) {
#pragma unroll
      for (int i = 0; i < LOC; i++)
#pragma unroll
        for (int j = 0; j < LOC; j++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += (ST.test(t, i, g) * t.N[j * NGI + g]) * c[g];
          R.L[i][j] += s;
        }
    }
    double m[LOC];  // sum(mass_mat,2) = sum_g test_ig c_g
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double s = 0.0;
#pragma unroll
      for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * c[g];
      m[i] = s;
    }
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        if (!o.exclude_mass && o.lump_mass) R.diag[d][i] += m[i];
        if (o.assemble_inverse_masslump) R.ml[d][i] += m[i];
      }
  }

  // Advection (add_advection_element_cg, :1602-1715)
  if (!o.exclude_advection) {
    double divu = 0.0;  // ele_div_at_quad: constant over the element for P1
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int a = 0; a < DIM; a++) divu += nu[i][a] * G.grad[i][a];

    if (o.integrate_advection_by_parts) {
      // A_ij = -sum_g (u_g.gradN_i) N_jg c_g - (1-beta) divu sum_g N_ig N_jg c_g  (:1667-1668)
      //      = -gradN_i . C_j - (1-beta) divu M_ij ; needs the transposed rank-1 form.
      double Cj[LOC][DIM];
#pragma unroll
      for (int j = 0; j < LOC; j++)
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += t.N[j * NGI + g] * (c[g] * ug[g][a]);
          Cj[j][a] = s;
        }
      const double f = (1.0 - o.beta) * divu;
#pragma unroll
      for (int i = 0; i < LOC; i++)
#pragma unroll
        for (int j = 0; j < LOC; j++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += G.grad[i][a] * Cj[j][a];
          double mm = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) mm += (ST.test(t, i, g) * t.N[j * NGI + g]) * c[g];
          const double aij = -s - f * mm + ST.su(i, j);
          R.L[i][j] += dtt * aij;
#pragma unroll
          for (int d = 0; d < DIM; d++) R.rhs[d][i] -= aij * oldu[j][d];
        }
    } else {
      // A_ij = C_i . gradN_j + beta divu M_ij   (:1675-1680)
#pragma unroll
      for (int i = 0; i < LOC; i++)
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * (c[g] * ug[g][a]);
          v[i][a] += s;
        }
      have_v = true;
      if (o.beta != 0.0 || STAB == 1) {
        const double f = o.beta * divu;
#pragma unroll
        for (int i = 0; i < LOC; i++)
#pragma unroll
          for (int j = 0; j < LOC; j++) {
            double mm = 0.0;
#pragma unroll
            for (int g = 0; g < NGI; g++) mm += (ST.test(t, i, g) * t.N[j * NGI + g]) * c[g];
            const double aij = f * mm + ST.su(i, j);
            R.L[i][j] += dtt * aij;
#pragma unroll
            for (int d = 0; d < DIM; d++) R.rhs[d][i] -= aij * oldu[j][d];
          }
      }
    }
  }

  // Viscosity, tensor form (add_viscosity_element_cg, :2286-2359)
  if (o.have_viscosity) {
    // Vbar(a,b) = sum_g (sum_i visc_i(a,b) N_ig) detwei_g ; memory index a + DIM*b
    double Vbar[DIM * DIM];
    if (A.viscosity.stride == 0) {
      // CONSTANT field: sum_i N_ig = 1  =>  Vbar = visc * |detJ| * sum_g w_g
      gather<DIM * DIM>(A.viscosity, 0, Vbar);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] *= G.absdet * wsum;
    } else {
      double visc[LOC][DIM * DIM];
#pragma unroll
      for (int i = 0; i < LOC; i++) gather<DIM * DIM>(A.viscosity, node_of(nd, i), visc[i]);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < LOC; i++) {
          double ni = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) ni += t.N[i * NGI + g] * t.w[g];
          s += visc[i][ab] * ni;
        }
        Vbar[ab] = s * G.absdet;
      }
    }
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      if (o.viscosity_shape == CGASM_TENSOR_ISOTROPIC) {
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += Vbar[0] * G.grad[i][a];
      } else if (o.viscosity_shape == CGASM_TENSOR_DIAGONAL) {
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += Vbar[a + DIM * a] * G.grad[i][a];
      } else {
        // K_ij = gradN_i^T V gradN_j  => v[i](b) += sum_a gradN_i(a) V(a,b)   (FETools :668-698)
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += G.grad[i][a] * Vbar[a + DIM * b];
          v[i][b] += s;
        }
      }
    }
    have_v = true;
  }

  if (have_v) {
    // (A+K)_ij = v_i . gradN_j ; big_m += dt theta (A+K) ; rhs(d,i) -= (A+K)_ij oldu(d,j)
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int j = 0; j < LOC; j++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) s += v[i][a] * G.grad[j][a];
        R.L[i][j] += dtt * s;
#pragma unroll
        for (int d = 0; d < DIM; d++) R.rhs[d][i] -= s * oldu[j][d];
      }
  }

  // Sources (add_sources_element_cg, :1717-1751)
  if (o.have_source) {
    double src[LOC][DIM];
#pragma unroll
    for (int i = 0; i < LOC; i++) gather<DIM>(A.source, node_of(nd, i), src[i]);
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      if (o.lump_source) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * c[g];
#pragma unroll
        for (int d = 0; d < DIM; d++) R.rhs[d][i] += s * src[i][d];
      } else {
#pragma unroll
        for (int j = 0; j < LOC; j++) {
          double mm = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) mm += (ST.test(t, i, g) * t.N[j * NGI + g]) * c[g];
#pragma unroll
          for (int d = 0; d < DIM; d++) R.rhs[d][i] += mm * src[j][d];
        }
      }
    }
  }

  // Buoyancy (add_buoyancy_element_cg, :1753-1792)
  if (o.have_gravity) {
    double b[LOC], bq[NGI];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      b[i] = buoy[i];
      if (o.subtract_out_reference_profile) {
        double r1[1];
        gather<1>(A.hb_density, node_of(nd, i), r1);
        b[i] -= r1[0];
      }
    }
    at_quad<DIM>(t, b, bq);
#pragma unroll
    for (int g = 0; g < NGI; g++) bq[g] *= o.gravity_magnitude * G.absdet * t.w[g];
    if (A.gravity.stride == 0) {
      // CONSTANT gravity direction: rhs(d,i) += ghat_d * sum_g N_ig c_g
      double gd[DIM];
      gather<DIM>(A.gravity, 0, gd);
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * bq[g];
#pragma unroll
        for (int d = 0; d < DIM; d++) R.rhs[d][i] += s * gd[d];
      }
    } else {
      double gv[LOC][DIM];
#pragma unroll
      for (int i = 0; i < LOC; i++) gather<DIM>(A.gravity, node_of(nd, i), gv[i]);
#pragma unroll
      for (int g = 0; g < NGI; g++) {
        double gg[DIM];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < LOC; i++) s += gv[i][a] * t.N[i * NGI + g];
          gg[a] = s * bq[g];
        }
#pragma unroll
        for (int i = 0; i < LOC; i++)
#pragma unroll
          for (int d = 0; d < DIM; d++) R.rhs[d][i] += ST.test(t, i, g) * gg[d];
      }
    }
  }

  // Absorption, plain branch (add_absorption_element_cg, :2036-2073)
  if (o.have_absorption) {
    double sg[LOC][DIM], sq[NGI][DIM];
#pragma unroll
    for (int i = 0; i < LOC; i++) gather<DIM>(A.absorption, node_of(nd, i), sg[i]);
#pragma unroll
    for (int g = 0; g < NGI; g++)
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < LOC; i++) s += sg[i][d] * t.N[i * NGI + g];
        sq[g][d] = s * c[g];
      }
    if (o.lump_absorption) {
#pragma unroll
      for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int i = 0; i < LOC; i++) {
          double s = 0.0;  // sum_j Ab_ij = sum_g test_ig sigma_dg c_g
#pragma unroll
          for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * sq[g][d];
          R.diag[d][i] += dtt * s;
          R.rhs[d][i] -= s * oldu[i][d];
          if (o.pressure_corrected_absorption && o.assemble_inverse_masslump) R.ml[d][i] += dtt * s;
        }
    } else if constexpr (LABS) {
#pragma unroll
      for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int i = 0; i < LOC; i++)
#pragma unroll
          for (int j = 0; j < LOC; j++) {
            double s = 0.0;
#pragma unroll
            for (int g = 0; g < NGI; g++) s += (ST.test(t, i, g) * t.N[j * NGI + g]) * sq[g][d];
            R.Labs[d][i][j] = dtt * s;
            R.rhs[d][i] -= s * oldu[j][d];
          }
    }
  }
}

// ct_m block: grad_p_u_mat(d,i,j) = sum_g N_ig dN_j/dx_d detwei_g   (:1401, FETools :332-362); with
// integrate_continuity_by_parts -dshape_shape(dp_t, u_shape, detwei): (d,i,j) = -sum_g dN_i/dx_d N_jg detwei_g
// (:1379-1383, FETools :364-389) -- for P1 minus the transpose of the plain form; the boundary half is the surface loop's
template <int DIM>
__device__ __forceinline__ double grad_p_u(const Tables& t, const Geom<DIM>& G, int d, int i, int j, bool by_parts = false) {
  constexpr int NGI = Shape<DIM>::NGI;
  const int a = by_parts ? j : i, b = by_parts ? i : j;
  double na = 0.0;
#pragma unroll
  for (int g = 0; g < NGI; g++) na += t.N[a * NGI + g] * t.w[g];
  const double v = na * G.absdet * G.grad[b][d];
  return by_parts ? -v : v;
}

// ---- tracer ----------------------------------------------------------------------------------
template <int DIM>
struct AdvDiffLocal {
  static constexpr int LOC = DIM + 1;
  double A[LOC][LOC];
  double rhs[LOC];
};

template <int DIM, int STAB = 0>
__device__ __forceinline__ void advdiff_element(const AdvDiffArgs& P, const int4 nd,
                                                AdvDiffLocal<DIM>& R) {
  constexpr int LOC = DIM + 1, NGI = Shape<DIM>::NGI;
  const cgasm_advdiff_opts& o = P.o;
  const Tables& t = P.tab;
  const double dtt = o.dt * o.theta;
  const double eps = 2.220446049250313e-16;  // epsilon(0.0), :1121
  const bool implicit = fabs(dtt) > eps;

  double X[LOC][DIM], T[LOC];
#pragma unroll
  for (int i = 0; i < LOC; i++) unpack<DIM>(ld256(P.rec.r0 + node_of(nd, i)), X[i], T[i]);
  Geom<DIM> G;
  geometry<DIM>(X, G);
  double wsum = 0.0;
#pragma unroll
  for (int g = 0; g < NGI; g++) wsum += t.w[g];

#pragma unroll
  for (int i = 0; i < LOC; i++) {
    R.rhs[i] = 0.0;
#pragma unroll
    for (int j = 0; j < LOC; j++) R.A[i][j] = 0.0;
  }

  // velocity at the quadrature points + stabilisation (:813-826, :1107-1119)
  double u[LOC][DIM], uq[NGI][DIM];
  Stabilisation<DIM, STAB> ST;
  if (o.have_advection || STAB != 0) {
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double unused;
      unpack<DIM>(ld256(P.rec.r1 + node_of(nd, i)), u[i], unused);
    }
#pragma unroll
    for (int g = 0; g < NGI; g++)
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < LOC; i++) s += u[i][a] * t.N[i * NGI + g];
        uq[g][a] = s;
      }
  }
  if constexpr (STAB != 0) {
    auto diff_at = [&](int g, double (&dq)[DIM * DIM]) {
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) {
        double vv = 0.0;
        if (P.diffusivity.stride == 0) {
          vv = __ldg(P.diffusivity.val + ab);
        } else {
#pragma unroll
          for (int i = 0; i < LOC; i++)
            vv += __ldg(P.diffusivity.val + (size_t)P.diffusivity.stride * node_of(nd, i) + ab) * t.N[i * NGI + g];
        }
        dq[ab] = vv;
      }
    };
    ST.setup(t, G, uq, o.have_diffusivity != 0, P.diffusivity.stride == 0, diff_at, o.nu_bar_scheme, o.nu_bar_scale);
  }

  // Mass (:867-941): M_ij = |detJ| sum_g N_ig N_jg w_g ; lumped -> |detJ| sum_g N_ig w_g
  if (o.have_mass) {
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      if (o.lump_mass) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * t.w[g];
        R.A[i][i] += s * G.absdet;
      } else {
#pragma unroll
        for (int j = 0; j < LOC; j++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += (ST.test(t, i, g) * t.N[j * NGI + g]) * t.w[g];
          R.A[i][j] += s * G.absdet;
        }
      }
    }
  }

  double v[LOC][DIM];  // (A + D)_ij = v_i . gradN_j
#pragma unroll
  for (int i = 0; i < LOC; i++)
#pragma unroll
    for (int a = 0; a < DIM; a++) v[i][a] = 0.0;
  bool have_v = false;

  // Advection (:943-1127, default equation type)
  if (o.have_advection) {
    double ug[NGI][DIM];  // u_g detwei_g
#pragma unroll
    for (int g = 0; g < NGI; g++)
#pragma unroll
      for (int a = 0; a < DIM; a++) ug[g][a] = uq[g][a] * (G.absdet * t.w[g]);
    double divu = 0.0;
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int a = 0; a < DIM; a++) divu += u[i][a] * G.grad[i][a];
    if (o.integrate_advection_by_parts) {
      const bool with_div = fabs(1.0 - o.beta) > eps;  // :1044
      const double f = (1.0 - o.beta) * divu * G.absdet;
#pragma unroll
      for (int j = 0; j < LOC; j++) {
        double Cj[DIM];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += t.N[j * NGI + g] * ug[g][a];
          Cj[a] = s;
        }
#pragma unroll
        for (int i = 0; i < LOC; i++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += G.grad[i][a] * Cj[a];
          double aij = -s;
          if (with_div) {
            double mm = 0.0;
#pragma unroll
            for (int g = 0; g < NGI; g++) mm += (ST.test(t, i, g) * t.N[j * NGI + g]) * t.w[g];
            aij -= f * mm;
          }
          aij += ST.su(i, j);
          if (implicit) R.A[i][j] += dtt * aij;
          R.rhs[i] -= aij * T[j];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < LOC; i++)
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * ug[g][a];
          v[i][a] += s;
        }
      have_v = true;
      const bool with_div = fabs(o.beta) > eps;  // :1093
      if (with_div || STAB == 1) {
        const double f = with_div ? o.beta * divu * G.absdet : 0.0;
#pragma unroll
        for (int i = 0; i < LOC; i++)
#pragma unroll
          for (int j = 0; j < LOC; j++) {
            double mm = 0.0;
#pragma unroll
            for (int g = 0; g < NGI; g++) mm += (ST.test(t, i, g) * t.N[j * NGI + g]) * t.w[g];
            const double aij = f * mm + ST.su(i, j);
            if (implicit) R.A[i][j] += dtt * aij;
            R.rhs[i] -= aij * T[j];
          }
      }
    }
  }

  // Diffusivity (:1164-1202)
  if (o.have_diffusivity) {
    double Kbar[DIM * DIM];
    if (P.diffusivity.stride == 0) {
      gather<DIM * DIM>(P.diffusivity, 0, Kbar);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] *= G.absdet * wsum;
    } else {
      double kap[LOC][DIM * DIM];
#pragma unroll
      for (int i = 0; i < LOC; i++) gather<DIM * DIM>(P.diffusivity, node_of(nd, i), kap[i]);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < LOC; i++) {
          double ni = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) ni += t.N[i * NGI + g] * t.w[g];
          s += kap[i][ab] * ni;
        }
        Kbar[ab] = s * G.absdet;
      }
    }
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      if (o.diffusivity_shape == CGASM_TENSOR_ISOTROPIC) {
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += Kbar[0] * G.grad[i][a];
      } else {
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += G.grad[i][a] * Kbar[a + DIM * b];
          v[i][b] += s;
        }
      }
    }
    have_v = true;
  }

  if (have_v) {
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int j = 0; j < LOC; j++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) s += v[i][a] * G.grad[j][a];
        if (implicit) R.A[i][j] += dtt * s;
        R.rhs[i] -= s * T[j];
      }
  }

  // Absorption (:1143-1162)
  if (o.have_absorption) {
    double sg[LOC], sq[NGI];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double r1[1];
      gather<1>(P.absorption, node_of(nd, i), r1);
      sg[i] = r1[0];
    }
    at_quad<DIM>(t, sg, sq);
#pragma unroll
    for (int g = 0; g < NGI; g++) sq[g] *= G.absdet * t.w[g];
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int j = 0; j < LOC; j++) {
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < NGI; g++) s += (ST.test(t, i, g) * t.N[j * NGI + g]) * sq[g];
        if (implicit) R.A[i][j] += dtt * s;
        R.rhs[i] -= s * T[j];
      }
  }

  // Source (:1129-1141)
  if (o.have_source) {
    double sg[LOC], sq[NGI];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double r1[1];
      gather<1>(P.source, node_of(nd, i), r1);
      sg[i] = r1[0];
    }
    at_quad<DIM>(t, sg, sq);
#pragma unroll
    for (int g = 0; g < NGI; g++) sq[g] *= G.absdet * t.w[g];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double s = 0.0;
#pragma unroll
      for (int g = 0; g < NGI; g++) s += ST.test(t, i, g) * sq[g];
      R.rhs[i] += s;
    }
  }
}

// =====================================================================================
// Register-lean fused fast paths (used by the tiled kernels).
//
// Same formulae as momentum_element / advdiff_element for the option set shared by the
// example configs: mass lumped or excluded, advection excluded or in
// non-conservative form with beta = 0 and not integrated by parts, tensor-form viscosity of any
// shape (CONSTANT or nodal), buoyancy with a CONSTANT gravity direction, lumped absorption or
// none, no source. Instead of materialising the whole local matrix, rows are produced one at a
// time and handed to a sink, and rows whose node the caller does not own are skipped entirely
// (their rhs / lumped-mass contributions belong to that node too), so elements on a tile
// surface cost only the shared part (geometry, quadrature sums).
// =====================================================================================
__host__ __device__ inline bool momentum_fast_ok(const cgasm_momentum_opts& o, int gravity_stride,
                                                 int absorption_stride) {
  (void)absorption_stride;
  if (o.stabilisation_scheme != CGASM_STAB_NONE) return false;
  if (o.have_source) return false;
  if (o.have_absorption && !o.lump_absorption) return false;
  if (!o.exclude_mass && !o.lump_mass) return false;
  if (!o.exclude_advection && (o.integrate_advection_by_parts || o.beta != 0.0)) return false;
  if (o.have_gravity && gravity_stride != 0) return false;
  return true;
}

// Sink concept:  bool owned(int i);  void mat(int i, int j, int d_or_0, double v)  [d only when
// PERD];  void vec(int i, int d, double rhs_v);  void ml(int i, int d_or_0, double v)
template <int DIM, bool PERD, class Sink>
__device__ __forceinline__ void momentum_fast(const MomentumArgs& A, const int4 nd, Sink& sink) {
  constexpr int LOC = DIM + 1, NGI = Shape<DIM>::NGI;
  const cgasm_momentum_opts& o = A.o;
  const Tables& t = A.tab;
  const double dtt = o.dt * o.theta;
  double wsum = 0.0;
#pragma unroll
  for (int g = 0; g < NGI; g++) wsum += t.w[g];

  Geom<DIM> G;
  double c[NGI];      // rho_g detwei_g
  double v[LOC][DIM];  // (A+K)_ij = v_i . gradN_j
  {
    double X[LOC][DIM], nu[LOC][DIM], rho[LOC];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      const int node = node_of(nd, i);
      double unused;
      unpack<DIM>(ld256(A.rec.r0 + node), X[i], unused);
      unpack<DIM>(ld256(A.rec.r1 + node), nu[i], rho[i]);
    }
    geometry<DIM>(X, G);
    at_quad<DIM>(t, rho, c);
#pragma unroll
    for (int g = 0; g < NGI; g++) c[g] *= G.absdet * t.w[g];
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int a = 0; a < DIM; a++) v[i][a] = 0.0;
    if (!o.exclude_advection) {
#pragma unroll
      for (int g = 0; g < NGI; g++) {
        double q[DIM];  // c_g u_g
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < LOC; i++) s += nu[i][a] * t.N[i * NGI + g];
          q[a] = s * c[g];
        }
#pragma unroll
        for (int i = 0; i < LOC; i++)
#pragma unroll
          for (int a = 0; a < DIM; a++) v[i][a] += t.N[i * NGI + g] * q[a];
      }
    }
  }
  if (o.have_viscosity) {
    double Vbar[DIM * DIM];
    if (A.viscosity.stride == 0) {
      gather<DIM * DIM>(A.viscosity, 0, Vbar);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] *= G.absdet * wsum;
    } else {
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] = 0.0;
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        double vi[DIM * DIM], ni = 0.0;
        gather<DIM * DIM>(A.viscosity, node_of(nd, i), vi);
#pragma unroll
        for (int g = 0; g < NGI; g++) ni += t.N[i * NGI + g] * t.w[g];
        ni *= G.absdet;
#pragma unroll
        for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] += vi[ab] * ni;
      }
    }
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      if (o.viscosity_shape == CGASM_TENSOR_ISOTROPIC) {
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += Vbar[0] * G.grad[i][a];
      } else if (o.viscosity_shape == CGASM_TENSOR_DIAGONAL) {
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += Vbar[a + DIM * a] * G.grad[i][a];
      } else {
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += G.grad[i][a] * Vbar[a + DIM * b];
          v[i][b] += s;
        }
      }
    }
  }

  double oldu[LOC][DIM], bq[NGI];
  {
    double b[LOC];
#pragma unroll
    for (int i = 0; i < LOC; i++) unpack<DIM>(ld256(A.rec.r2 + node_of(nd, i)), oldu[i], b[i]);
    if (o.have_gravity) {
      if (o.subtract_out_reference_profile) {
#pragma unroll
        for (int i = 0; i < LOC; i++) {
          double r1[1];
          gather<1>(A.hb_density, node_of(nd, i), r1);
          b[i] -= r1[0];
        }
      }
      at_quad<DIM>(t, b, bq);
#pragma unroll
      for (int g = 0; g < NGI; g++) bq[g] *= o.gravity_magnitude * G.absdet * t.w[g];
    }
  }
  double gd[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) gd[d] = 0.0;
  if (o.have_gravity) gather<DIM>(A.gravity, 0, gd);
  const bool want_m = o.assemble_inverse_masslump || !o.exclude_mass;
  double sabs[PERD ? NGI : 1][DIM];  // sigma_dg c_g (lumped absorption only)
  if constexpr (PERD) {
#pragma unroll
    for (int g = 0; g < NGI; g++)
#pragma unroll
      for (int d = 0; d < DIM; d++) sabs[g][d] = 0.0;
    if (o.have_absorption) {
#pragma unroll
      for (int k = 0; k < LOC; k++) {
        double ak[DIM];
        gather<DIM>(A.absorption, node_of(nd, k), ak);
#pragma unroll
        for (int g = 0; g < NGI; g++)
#pragma unroll
          for (int d = 0; d < DIM; d++) sabs[g][d] += ak[d] * (t.N[k * NGI + g] * c[g]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < LOC; i++) {
    if (!sink.owned(i)) continue;
    double m = 0.0, nb = 0.0;
#pragma unroll
    for (int g = 0; g < NGI; g++) {
      m += t.N[i * NGI + g] * c[g];
      if (o.have_gravity) nb += t.N[i * NGI + g] * bq[g];
    }
    double rhs[DIM], diag[PERD ? DIM : 1], mlv[PERD ? DIM : 1];
#pragma unroll
    for (int d = 0; d < DIM; d++) rhs[d] = nb * gd[d];
#pragma unroll
    for (int d = 0; d < (PERD ? DIM : 1); d++) {
      diag[d] = (!o.exclude_mass && want_m) ? m : 0.0;
      mlv[d] = o.assemble_inverse_masslump ? m : 0.0;
    }
    if constexpr (PERD) {
      // lumped absorption (:2043-2050): sum_j Ab_ij = sum_g N_ig sigma_dg c_g
      if (o.have_absorption) {
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          double s = 0.0;
#pragma unroll
          for (int g = 0; g < NGI; g++) s += t.N[i * NGI + g] * sabs[g][d];
          diag[d] += dtt * s;
          rhs[d] -= s * oldu[i][d];
          if (o.pressure_corrected_absorption && o.assemble_inverse_masslump) mlv[d] += dtt * s;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < LOC; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) s += v[i][a] * G.grad[j][a];
#pragma unroll
      for (int d = 0; d < DIM; d++) rhs[d] -= s * oldu[j][d];
#pragma unroll
      for (int d = 0; d < (PERD ? DIM : 1); d++) sink.mat(i, j, d, dtt * s + (i == j ? diag[d] : 0.0));
    }
#pragma unroll
    for (int d = 0; d < DIM; d++) sink.vec(i, d, rhs[d]);
#pragma unroll
    for (int d = 0; d < (PERD ? DIM : 1); d++) sink.ml(i, d, mlv[d]);
    sink.row_end(i);
  }
}

__host__ __device__ inline bool advdiff_fast_ok(const cgasm_advdiff_opts& o) {
  if (o.stabilisation_scheme != CGASM_STAB_NONE) return false;
  if (o.have_absorption) return false;
  if (o.have_advection && (o.integrate_advection_by_parts || o.beta != 0.0)) return false;
  return true;
}

// Sink: bool owned(i); void mat(i, j, v); void vec(i, v)
template <int DIM, class Sink>
__device__ __forceinline__ void advdiff_fast(const AdvDiffArgs& P, const int4 nd, Sink& sink) {
  constexpr int LOC = DIM + 1, NGI = Shape<DIM>::NGI;
  const cgasm_advdiff_opts& o = P.o;
  const Tables& t = P.tab;
  const double dtt = o.dt * o.theta;
  const double eps = 2.220446049250313e-16;
  const bool implicit = fabs(dtt) > eps;
  double wsum = 0.0;
#pragma unroll
  for (int g = 0; g < NGI; g++) wsum += t.w[g];

  Geom<DIM> G;
  double T[LOC], v[LOC][DIM];
  {
    double X[LOC][DIM];
#pragma unroll
    for (int i = 0; i < LOC; i++) unpack<DIM>(ld256(P.rec.r0 + node_of(nd, i)), X[i], T[i]);
    geometry<DIM>(X, G);
  }
#pragma unroll
  for (int i = 0; i < LOC; i++)
#pragma unroll
    for (int a = 0; a < DIM; a++) v[i][a] = 0.0;
  if (o.have_advection) {
    double u[LOC][DIM];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double unused;
      unpack<DIM>(ld256(P.rec.r1 + node_of(nd, i)), u[i], unused);
    }
#pragma unroll
    for (int g = 0; g < NGI; g++) {
      double q[DIM];
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < LOC; i++) s += u[i][a] * t.N[i * NGI + g];
        q[a] = s * (G.absdet * t.w[g]);
      }
#pragma unroll
      for (int i = 0; i < LOC; i++)
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += t.N[i * NGI + g] * q[a];
    }
  }
  if (o.have_diffusivity) {
    double Kbar[DIM * DIM];
    if (P.diffusivity.stride == 0) {
      gather<DIM * DIM>(P.diffusivity, 0, Kbar);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] *= G.absdet * wsum;
    } else {
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] = 0.0;
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        double ki[DIM * DIM], ni = 0.0;
        gather<DIM * DIM>(P.diffusivity, node_of(nd, i), ki);
#pragma unroll
        for (int g = 0; g < NGI; g++) ni += t.N[i * NGI + g] * t.w[g];
        ni *= G.absdet;
#pragma unroll
        for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] += ki[ab] * ni;
      }
    }
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      if (o.diffusivity_shape == CGASM_TENSOR_ISOTROPIC) {
#pragma unroll
        for (int a = 0; a < DIM; a++) v[i][a] += Kbar[0] * G.grad[i][a];
      } else {
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) s += G.grad[i][a] * Kbar[a + DIM * b];
          v[i][b] += s;
        }
      }
    }
  }
  double sq[NGI];
  if (o.have_source) {
    double sg[LOC];
#pragma unroll
    for (int i = 0; i < LOC; i++) {
      double r1[1];
      gather<1>(P.source, node_of(nd, i), r1);
      sg[i] = r1[0];
    }
    at_quad<DIM>(t, sg, sq);
#pragma unroll
    for (int g = 0; g < NGI; g++) sq[g] *= G.absdet * t.w[g];
  }
#pragma unroll
  for (int i = 0; i < LOC; i++) {
    if (!sink.owned(i)) continue;
    double rhs = 0.0;
    if (o.have_source) {
#pragma unroll
      for (int g = 0; g < NGI; g++) rhs += t.N[i * NGI + g] * sq[g];
    }
    double mlump = 0.0;
    if (o.have_mass && o.lump_mass) {
#pragma unroll
      for (int g = 0; g < NGI; g++) mlump += t.N[i * NGI + g] * t.w[g];
      mlump *= G.absdet;
    }
#pragma unroll
    for (int j = 0; j < LOC; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) s += v[i][a] * G.grad[j][a];
      rhs -= s * T[j];
      double a_ij = implicit ? dtt * s : 0.0;
      if (o.have_mass) {
        if (o.lump_mass) {
          if (i == j) a_ij += mlump;
        } else {
          double pij = 0.0;  // sum_g N_ig N_jg w_g: constant-bank arithmetic, folded by ptxas
#pragma unroll
          for (int g = 0; g < NGI; g++) pij += (t.N[i * NGI + g] * t.N[j * NGI + g]) * t.w[g];
          a_ij += pij * G.absdet;
        }
      }
      sink.mat(i, j, a_ij);
    }
    sink.vec(i, rhs);
    sink.row_end(i);
  }
}

// =====================================================================================
// Row kernels for the quad mapping (one lane = one local row of one element).
//
// The caller passes the element's nodes rotated so that the lane's own node is local node 0;
// the degree-3 rules are symmetric under node permutations (checked at create: Tables::sym),
// so every quadrature sum collapses to the closed-form moments in Tables:
//   rho-weighted mass row   M_0k = |J| sum_l Q_0kl rho_l
//   advection               v_0  = sum_k M_0k nu_k            (A_0j = v_0 . gradN_j)
//   lumped mass             m_0  = |J| sum_l P_0l rho_l
//   buoyancy                n_0  = g |J| sum_k P_0k b_k
// Option coverage = momentum_fast_ok / advdiff_fast_ok. Sink: mat(jj, d, v), vec(d, v), ml(d, v)
// with jj the ROTATED column index.
// =====================================================================================
// Lean geometry for the row kernels: cofactors + 1/det only. gradN_k = C(:,k)/det for k < dim and
// gradN_loc = -sum_k gradN_k, so v.gradN_j = (v.C(:,j))/det and the last product is minus the sum
// of the others; only gradN_0 (the row's own node) is ever formed explicitly.
template <int DIM>
struct GeomLean {
  double C[DIM][DIM];
  double rdet, absdet;
  __device__ __forceinline__ void grad0(double (&g)[DIM]) const {
#pragma unroll
    for (int a = 0; a < DIM; a++) g[a] = C[a][0] * rdet;
  }
  __device__ __forceinline__ void dots(const double (&v)[DIM], double (&s)[DIM + 1]) const {
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < DIM; j++) {
      double t = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) t += v[a] * C[a][j];
      s[j] = t * rdet;
      tot += s[j];
    }
    s[DIM] = -tot;
  }
};

template <int DIM>
__device__ __forceinline__ void geometry_lean(const double (&X)[DIM + 1][DIM], GeomLean<DIM>& G) {
  double JT[DIM][DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++)
#pragma unroll
    for (int k = 0; k < DIM; k++) JT[a][k] = X[k][a] - X[DIM][a];
  if constexpr (DIM == 2) {
    G.C[0][0] = JT[1][1];
    G.C[1][0] = -JT[0][1];
    G.C[0][1] = -JT[1][0];
    G.C[1][1] = JT[0][0];
  } else {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        G.C[i][k] = JT[i1][k1] * JT[i2][k2] - JT[i2][k1] * JT[i1][k2];
      }
  }
  double det = 0.0;
#pragma unroll
  for (int a = 0; a < DIM; a++) det += JT[a][0] * G.C[a][0];
  G.rdet = 1.0 / det;
  G.absdet = fabs(det);
}

template <int DIM>
__device__ __forceinline__ void load_rot(const double4* __restrict__ rec, const int (&n)[4], double (&v)[DIM + 1][DIM],
                                         double (&s)[DIM + 1]) {
#pragma unroll
  for (int k = 0; k < DIM + 1; k++) unpack<DIM>(ld256(rec + n[k]), v[k], s[k]);
}

// Option views: the runtime one reads the option struct (uniform branches), the "common" ones
// are compile-time constants for the option set shared by the example configs, so the hot
// instantiation carries no option branches at all.
struct MomRuntimeFlags {
  const cgasm_momentum_opts& o;
  int visc_stride;
  __device__ __forceinline__ bool adv() const { return !o.exclude_advection; }
  __device__ __forceinline__ bool visc() const { return o.have_viscosity != 0; }
  __device__ __forceinline__ bool visc_const() const { return visc_stride == 0; }
  __device__ __forceinline__ int visc_shape() const { return o.viscosity_shape; }
  __device__ __forceinline__ bool grav() const { return o.have_gravity != 0; }
  __device__ __forceinline__ bool hb() const { return o.subtract_out_reference_profile != 0; }
  __device__ __forceinline__ bool mass() const { return !o.exclude_mass; }
  __device__ __forceinline__ bool mlump() const { return o.assemble_inverse_masslump != 0; }
};
// lumped mass, advection, CONSTANT isotropic viscosity, buoyancy (no reference profile), inverse
// lumped mass: driven_cavity / lock_exchange / flow_past_sphere(isotropic part) / S3
struct MomCommonFlags {
  __device__ __forceinline__ MomCommonFlags() {}
  __device__ __forceinline__ constexpr bool adv() const { return true; }
  __device__ __forceinline__ constexpr bool visc() const { return true; }
  __device__ __forceinline__ constexpr bool visc_const() const { return true; }
  __device__ __forceinline__ constexpr int visc_shape() const { return CGASM_TENSOR_ISOTROPIC; }
  __device__ __forceinline__ constexpr bool grav() const { return true; }
  __device__ __forceinline__ constexpr bool hb() const { return false; }
  __device__ __forceinline__ constexpr bool mass() const { return true; }
  __device__ __forceinline__ constexpr bool mlump() const { return true; }
};
__host__ __device__ inline bool momentum_common_ok(const cgasm_momentum_opts& o, int visc_stride) {
  return !o.exclude_advection && o.have_viscosity && visc_stride == 0 &&
         o.viscosity_shape == CGASM_TENSOR_ISOTROPIC && o.have_gravity && !o.subtract_out_reference_profile &&
         !o.exclude_mass && o.assemble_inverse_masslump && !o.have_absorption;
}

// Records of the row's own node (rotated local node 0): identical for every pair of the row, so the
// row kernels load them once per row instead of once per pair.
template <int DIM>
struct OwnNode {
  double X[DIM], nu[DIM], oldu[DIM];
  double T, rho, b;
  __device__ __forceinline__ void load(const NodeRecs& rec, int node) {
    unpack<DIM>(ld256(rec.r0 + node), X, T);
    unpack<DIM>(ld256(rec.r1 + node), nu, rho);
    unpack<DIM>(ld256(rec.r2 + node), oldu, b);
  }
  __device__ __forceinline__ void load_tracer(const NodeRecs& rec, int node) {
    unpack<DIM>(ld256(rec.r0 + node), X, T);
    unpack<DIM>(ld256(rec.r1 + node), nu, rho);
  }
};

// loads local nodes 1..LOC-1 from the records, node 0 from `own`
template <int DIM>
__device__ __forceinline__ void load_rot_own(const double4* __restrict__ rec, const int (&n)[4], const double (&own_v)[DIM],
                                             double own_s, double (&v)[DIM + 1][DIM], double (&s)[DIM + 1]) {
#pragma unroll
  for (int a = 0; a < DIM; a++) v[0][a] = own_v[a];
  s[0] = own_s;
#pragma unroll
  for (int k = 1; k < DIM + 1; k++) unpack<DIM>(ld256(rec + n[k]), v[k], s[k]);
}

// Core of the momentum row: node data already in registers. X/nu/oldu are (LOC, DIM) with the row's
// own node first, rho/b (LOC); n[] = node ids (only used by the runtime-flag gathers).
template <int DIM, bool PERD, class Sink, class F>
__device__ __forceinline__ void momentum_row0_data(const MomentumArgs& A, const int (&n)[4],
                                                   const double (&X)[DIM + 1][DIM], const double (&nu)[DIM + 1][DIM],
                                                   const double (&rho)[DIM + 1], const double (&oldu)[DIM + 1][DIM],
                                                   const double (&b_in)[DIM + 1], Sink& sink, const F f) {
  constexpr int LOC = DIM + 1;
  const cgasm_momentum_opts& o = A.o;
  const Tables& t = A.tab;
  const double dtt = o.dt * o.theta;
  GeomLean<DIM> G;
  double g0[DIM];  // gradN of the row's own node
  double M[LOC];  // rho-weighted mass row of local node 0
  double v[DIM];
  double m0;
  geometry_lean<DIM>(X, G);
  G.grad0(g0);
  {
    double S = 0.0;
#pragma unroll
    for (int k = 0; k < LOC; k++) S += rho[k];
    M[0] = G.absdet * ((t.Qaaa - t.Qaab) * rho[0] + t.Qaab * S);
#pragma unroll
    for (int k = 1; k < LOC; k++) M[k] = G.absdet * ((t.Qaab - t.Qabc) * (rho[0] + rho[k]) + t.Qabc * S);
    m0 = G.absdet * ((t.Pd - t.Po) * rho[0] + t.Po * S);
#pragma unroll
    for (int a = 0; a < DIM; a++) v[a] = 0.0;
    if (f.adv()) {
#pragma unroll
      for (int k = 0; k < LOC; k++)
#pragma unroll
        for (int a = 0; a < DIM; a++) v[a] += M[k] * nu[k][a];
    }
  }
  if (f.visc()) {
    double Vbar[DIM * DIM];
    if (f.visc_const()) {
      gather<DIM * DIM>(A.viscosity, 0, Vbar);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] *= G.absdet * t.Wsum;
    } else {
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] = 0.0;
#pragma unroll
      for (int k = 0; k < LOC; k++) {
        double vk[DIM * DIM];
        gather<DIM * DIM>(A.viscosity, n[k], vk);
#pragma unroll
        for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] += vk[ab];
      }
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Vbar[ab] *= G.absdet * t.W1;
    }
    if (f.visc_shape() == CGASM_TENSOR_ISOTROPIC) {
#pragma unroll
      for (int a = 0; a < DIM; a++) v[a] += Vbar[0] * g0[a];
    } else if (f.visc_shape() == CGASM_TENSOR_DIAGONAL) {
#pragma unroll
      for (int a = 0; a < DIM; a++) v[a] += Vbar[a + DIM * a] * g0[a];
    } else {
#pragma unroll
      for (int b = 0; b < DIM; b++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) s += g0[a] * Vbar[a + DIM * b];
        v[b] += s;
      }
    }
  }
  double b[LOC];
#pragma unroll
  for (int k = 0; k < LOC; k++) b[k] = b_in[k];
  double rhs[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rhs[d] = 0.0;
  if (f.grav()) {
    if (f.hb()) {
#pragma unroll
      for (int k = 0; k < LOC; k++) {
        double r1[1];
        gather<1>(A.hb_density, n[k], r1);
        b[k] -= r1[0];
      }
    }
    double S = 0.0;
#pragma unroll
    for (int k = 0; k < LOC; k++) S += b[k];
    const double nb = o.gravity_magnitude * G.absdet * ((t.Pd - t.Po) * b[0] + t.Po * S);
    double gd[DIM];
    gather<DIM>(A.gravity, 0, gd);
#pragma unroll
    for (int d = 0; d < DIM; d++) rhs[d] = nb * gd[d];
  }
  double diag[PERD ? DIM : 1], mlv[PERD ? DIM : 1];
#pragma unroll
  for (int d = 0; d < (PERD ? DIM : 1); d++) {
    diag[d] = f.mass() ? m0 : 0.0;
    mlv[d] = f.mlump() ? m0 : 0.0;
  }
  if constexpr (PERD) {
    if (o.have_absorption) {  // lumped: sum_j Ab_0j(d) = sum_k M_0k sigma_k(d)
      double sd[DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++) sd[d] = 0.0;
#pragma unroll
      for (int k = 0; k < LOC; k++) {
        double ak[DIM];
        gather<DIM>(A.absorption, n[k], ak);
#pragma unroll
        for (int d = 0; d < DIM; d++) sd[d] += M[k] * ak[d];
      }
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        diag[d] += dtt * sd[d];
        rhs[d] -= sd[d] * oldu[0][d];
        if (o.pressure_corrected_absorption && o.assemble_inverse_masslump) mlv[d] += dtt * sd[d];
      }
    }
  }
  double sj[LOC];
  G.dots(v, sj);
#pragma unroll
  for (int j = 0; j < LOC; j++) {
    const double s = sj[j];
#pragma unroll
    for (int d = 0; d < DIM; d++) rhs[d] -= s * oldu[j][d];
#pragma unroll
    for (int d = 0; d < (PERD ? DIM : 1); d++) sink.mat(j, d, dtt * s + (j == 0 ? diag[d] : 0.0));
  }
#pragma unroll
  for (int d = 0; d < DIM; d++) sink.vec(d, rhs[d]);
#pragma unroll
  for (int d = 0; d < (PERD ? DIM : 1); d++) sink.ml(d, mlv[d]);
}

// Sink: mat(jj, v), vec(v)

template <int DIM, bool PERD, class Sink, class F>
__device__ __forceinline__ void momentum_row0(const MomentumArgs& A, const int (&n)[4], const OwnNode<DIM>& own,
                                              Sink& sink, const F f) {
  constexpr int LOC = DIM + 1;
  double X[LOC][DIM], nu[LOC][DIM], oldu[LOC][DIM], T_unused[LOC], rho[LOC], b[LOC];
  load_rot_own<DIM>(A.rec.r0, n, own.X, own.T, X, T_unused);
  load_rot_own<DIM>(A.rec.r1, n, own.nu, own.rho, nu, rho);
  load_rot_own<DIM>(A.rec.r2, n, own.oldu, own.b, oldu, b);
  momentum_row0_data<DIM, PERD>(A, n, X, nu, rho, oldu, b, sink, f);
}

struct AdvRuntimeFlags {
  const cgasm_advdiff_opts& o;
  int diff_stride;
  __device__ __forceinline__ bool adv() const { return o.have_advection != 0; }
  __device__ __forceinline__ bool diff() const { return o.have_diffusivity != 0; }
  __device__ __forceinline__ bool diff_const() const { return diff_stride == 0; }
  __device__ __forceinline__ int diff_shape() const { return o.diffusivity_shape; }
  __device__ __forceinline__ bool source() const { return o.have_source != 0; }
  __device__ __forceinline__ bool mass() const { return o.have_mass != 0; }
  __device__ __forceinline__ bool lump() const { return o.lump_mass != 0; }
};
// consistent mass, advection, CONSTANT isotropic diffusivity, no source: the default CG tracer
struct AdvCommonFlags {
  __device__ __forceinline__ AdvCommonFlags() {}
  __device__ __forceinline__ constexpr bool adv() const { return true; }
  __device__ __forceinline__ constexpr bool diff() const { return true; }
  __device__ __forceinline__ constexpr bool diff_const() const { return true; }
  __device__ __forceinline__ constexpr int diff_shape() const { return CGASM_TENSOR_ISOTROPIC; }
  __device__ __forceinline__ constexpr bool source() const { return false; }
  __device__ __forceinline__ constexpr bool mass() const { return true; }
  __device__ __forceinline__ constexpr bool lump() const { return false; }
};
__host__ __device__ inline bool advdiff_common_ok(const cgasm_advdiff_opts& o, int diff_stride) {
  return o.have_advection && o.have_diffusivity && diff_stride == 0 && o.diffusivity_shape == CGASM_TENSOR_ISOTROPIC &&
         !o.have_source && o.have_mass && !o.lump_mass;
}

template <int DIM, class Sink, class F>
__device__ __forceinline__ void advdiff_row0_data(const AdvDiffArgs& P, const int (&n)[4], const double (&X)[DIM + 1][DIM],
                                                  const double (&T)[DIM + 1], const double (&u)[DIM + 1][DIM], Sink& sink,
                                                  const F f) {
  constexpr int LOC = DIM + 1;
  const cgasm_advdiff_opts& o = P.o;
  const Tables& t = P.tab;
  const double dtt = o.dt * o.theta;
  const bool implicit = fabs(dtt) > 2.220446049250313e-16;
  GeomLean<DIM> G;
  double g0[DIM];
  double v[DIM];
  geometry_lean<DIM>(X, G);
  G.grad0(g0);
#pragma unroll
  for (int a = 0; a < DIM; a++) v[a] = 0.0;
  if (f.adv()) {
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      double S = 0.0;
#pragma unroll
      for (int k = 0; k < LOC; k++) S += u[k][a];
      v[a] = G.absdet * ((t.Pd - t.Po) * u[0][a] + t.Po * S);
    }
  }
  if (f.diff()) {
    double Kbar[DIM * DIM];
    if (f.diff_const()) {
      gather<DIM * DIM>(P.diffusivity, 0, Kbar);
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] *= G.absdet * t.Wsum;
    } else {
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] = 0.0;
#pragma unroll
      for (int k = 0; k < LOC; k++) {
        double kk[DIM * DIM];
        gather<DIM * DIM>(P.diffusivity, n[k], kk);
#pragma unroll
        for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] += kk[ab];
      }
#pragma unroll
      for (int ab = 0; ab < DIM * DIM; ab++) Kbar[ab] *= G.absdet * t.W1;
    }
    if (f.diff_shape() == CGASM_TENSOR_ISOTROPIC) {
#pragma unroll
      for (int a = 0; a < DIM; a++) v[a] += Kbar[0] * g0[a];
    } else {
#pragma unroll
      for (int b = 0; b < DIM; b++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) s += g0[a] * Kbar[a + DIM * b];
        v[b] += s;
      }
    }
  }
  double rhs = 0.0;
  if (f.source()) {
    double S = 0.0, s0 = 0.0;
#pragma unroll
    for (int k = 0; k < LOC; k++) {
      double r1[1];
      gather<1>(P.source, n[k], r1);
      S += r1[0];
      if (k == 0) s0 = r1[0];
    }
    rhs = G.absdet * ((t.Pd - t.Po) * s0 + t.Po * S);
  }
  double sj[LOC];
  G.dots(v, sj);
#pragma unroll
  for (int j = 0; j < LOC; j++) {
    const double s = sj[j];
    rhs -= s * T[j];
    double a_0j = implicit ? dtt * s : 0.0;
    if (f.mass()) {
      if (f.lump()) {
        if (j == 0) a_0j += G.absdet * t.W1;
      } else {
        a_0j += G.absdet * (j == 0 ? t.Pd : t.Po);
      }
    }
    sink.mat(j, a_0j);
  }
  sink.vec(rhs);
}


template <int DIM, class Sink, class F>
__device__ __forceinline__ void advdiff_row0(const AdvDiffArgs& P, const int (&n)[4], const OwnNode<DIM>& own,
                                             Sink& sink, const F f) {
  constexpr int LOC = DIM + 1;
  double X[LOC][DIM], T[LOC], u[LOC][DIM], unused[LOC];
  load_rot_own<DIM>(P.rec.r0, n, own.X, own.T, X, T);
  load_rot_own<DIM>(P.rec.r1, n, own.nu, own.rho, u, unused);
  advdiff_row0_data<DIM>(P, n, X, T, u, sink, f);
}

}  // namespace cgasm
