// strip_common.cuh -- device code shared by the STRIP kernels (strip.cu: node records fetched per entry
// from global memory; strip_staged.cu: node records of the row block staged in shared memory).
#pragma once
#include "gather_plan.h"
#include "strip_plan.h"

namespace cgasm {

#ifndef CGASM_V_SEPARATE_SUMS
#define CGASM_V_SEPARATE_SUMS 0
#endif
constexpr bool kStripSeparateSums = CGASM_V_SEPARATE_SUMS != 0;  // A/B switch (scripts/ab_kernels.py)
// Diagonal by row sum (3-D staged kernels): advection and viscosity / diffusion have zero row sums (sum_j gradN_j = 0), so
// the diagonal entry of a row is minus the sum of its off-diagonal ones -- taken when a column's accumulated entry is
// flushed (one DADD per strip entry) instead of a dot product u . sc per (row, element) pair (four FP64). The reference
// computes A_ii = v_i . gradN_i directly; the difference is rounding (1e-16 of the row's largest entry).
#ifndef CGASM_V_ROWSUM
#define CGASM_V_ROWSUM 1
#endif
template <int DIM>
constexpr bool kStripRowSum = (DIM == 3) && (CGASM_V_ROWSUM != 0);

// Carried cross product (window_geom_carry): measured on the B200 (profiles/r2_ab_carry.txt) -1.1 ... -3.9 % for the tracer
// kernel (118 registers: room for the three carried values), +1.3 ... +3 % for the momentum kernel (at its 128-register
// ceiling): on for the tracer, off for the momentum loop. CGASM_V_CARRY=0/1 forces both (A/B builds).
#ifdef CGASM_V_CARRY
constexpr bool kStripCarryMomentum = CGASM_V_CARRY != 0, kStripCarryTracer = CGASM_V_CARRY != 0;
#else
constexpr bool kStripCarryMomentum = false, kStripCarryTracer = true;
#endif

struct StripConsts {  // passed by value: operands are read straight from the constant bank
  // Option switches are folded into these numbers on the host (consts_momentum / consts_advdiff): a term
  // that is switched off has zero coefficients, so one kernel serves every combination.
  double Qa, Qaab, Qd, Qabc;  // advection: Qa = Qaaa - Qaab, Qd = Qaab - Qabc (Tables); 0 if no advection
  double PdPo, Po;            // momentum: lumped mass / buoyancy moments; tracer: advection moments (0 if none)
  double mPd, mPo;            // tracer mass matrix: (Pd, Po) consistent, (W1, 0) lumped, (0, 0) none
  double mass_on;             // momentum: 1 if the lumped mass goes on the diagonal of big_m (not exclude_mass)
  double V[9];                // constant viscosity / diffusivity tensor * Wsum, [a + dim*b]; V[0] if isotropic
  double grav[3];             // gravity_magnitude * gravity direction (0 if no gravity)
  double dtt;                 // dt*theta (tracer: 0 unless |dt*theta| > epsilon, Advection_Diffusion_CG.F90:1121)
  double sPdPo, sPo;          // tracer source moments (Pd - Po, Po), 0 without a source; with absorption the tracer
                              // kernel reads Qa, Qaab, Qd, Qabc as the absorption moments (0 without absorption)
};

struct StripPlanView {
  const int* __restrict__ rows;
  const long long* __restrict__ ptr;
  const int2* __restrict__ ent;
  const unsigned char* __restrict__ own_slot;
  const int* __restrict__ findrm;
  const int* __restrict__ colm;
  int maxlen, lpr_shift;
};

__device__ __forceinline__ int2 ldg_stream2(const int2* p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

// volatile: the request must be issued HERE (one step ahead of its use); the plain ld256 is sunk by the
// compiler to just before the first use, behind the divergent compute block
__device__ __forceinline__ double4 ld256v(const double4* p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

// 1/x: MUFU.RCP64H seed + two Newton steps (what the compiler's own division starts from, minus
// the special-case branch; det of a valid element is a normal number)
__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  // r1 = r0 (1 + e0), r2 = r1 (1 + e0^2) with e0 = 1 - x r0: the residual of r1 is e0^2 exactly, so the second
  // Newton step needs no second residual evaluation (the dependent chain is one FMA shorter; |e0| < 2^-20)
  const double e = fma(-x, r, 1.0);
  const double r1 = fma(r, e, r);
  return fma(r1, e * e, r1);
}

__device__ __forceinline__ double flip_sign(double v, unsigned sgn) {
  return __hiloint2double(__double2hiint(v) ^ (int)sgn, __double2loint(v));
}

// cofactor vectors of the window: gradN_k = c[k] / det, det = e_0 . c[0], with e_k = X_k - X_r the
// edges from the row's own node (femtools/Transform_elements.F90:807-887 with r as the origin)
#define WQ(k) ((QC + N - (DIM - 1) + (k)) % N)
template <int DIM, int N, int QC>
__device__ __forceinline__ double window_geometry(const double (&X)[N][DIM], double (&c)[DIM][DIM]) {
  if constexpr (DIM == 3) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double(&p)[3] = X[WQ((k + 1) % 3)];
      const double(&q)[3] = X[WQ((k + 2) % 3)];
      c[k][0] = p[1] * q[2] - p[2] * q[1];
      c[k][1] = p[2] * q[0] - p[0] * q[2];
      c[k][2] = p[0] * q[1] - p[1] * q[0];
    }
  } else {
    c[0][0] = X[WQ(1)][1];
    c[0][1] = -X[WQ(1)][0];
    c[1][0] = -X[WQ(0)][1];
    c[1][1] = X[WQ(0)][0];
  }
  double det = 0.0;
#pragma unroll
  for (int a = 0; a < DIM; a++) det = fma(X[WQ(0)][a], c[0][a], det);
  return det;
}

// ---- momentum -----------------------------------------------------------------------------------------
template <int DIM, int N>
struct MomState {
  double X[N][DIM], U[N][DIM], R[N], B[N], A[N];  // edge (after install), nu, density, buoyancy, accumulator
  int meta[N];
  double X0[DIM], U0[DIM], rho0, b0;
  double a0, msum, nbsum;
  double qa_rho0, qd_rho0, pm_rho0, pm_b0;  // row constants of the moments (mom_row_consts)
};

// products of the row's own density / buoyancy with the quadrature moments: once per row instead of once per element
template <int DIM, int N>
__device__ __forceinline__ void mom_row_consts(MomState<DIM, N>& s, const StripConsts& k_) {
  if constexpr (DIM == 3 && !kStripSeparateSums) {  // (2-D keeps the per-element products: measured faster there)
    s.qa_rho0 = k_.Qa * s.rho0;
    s.qd_rho0 = k_.Qd * s.rho0;
    s.pm_rho0 = k_.PdPo * s.rho0;
    s.pm_b0 = k_.PdPo * s.b0;
  }
}

// row 0 (the row's own node) of the element {r, window}: Momentum_CG.F90:1535-1552 (lumped mass),
// :1675-1680 with beta = 0 (advection), :2304-2317 (constant isotropic viscosity), :1770-1789 (buoyancy)
// u = sign(det) (w - (1/det) V^T sc): the viscous/diffusive part of the row vector. FULLV: V is a full constant
// tensor (dshape_tensor_dshape, Momentum_CG.F90:2331-2339), else isotropic (V[0]).
template <int DIM, bool FULLV>
__device__ __forceinline__ void row_vector(const StripConsts& k_, const double (&w)[DIM], const double (&sc)[DIM], double rd,
                                           double det, double (&u)[DIM]) {
  const unsigned sgn = (unsigned)__double2hiint(det) & 0x80000000u;
  if constexpr (FULLV) {
#pragma unroll
    for (int b = 0; b < DIM; b++) {
      double t = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) t = fma(sc[a], k_.V[a + DIM * b], t);
      u[b] = flip_sign(fma(-rd, t, w[b]), sgn);
    }
  } else {
    const double tt = k_.V[0] * rd;
#pragma unroll
    for (int a = 0; a < DIM; a++) u[a] = flip_sign(fma(-tt, sc[a], w[a]), sgn);
  }
}

// cofactors, determinant, its reciprocal and the cofactor sum of the window: shared by every term of the element
template <int DIM>
struct WindowGeom {
  double c[DIM][DIM], sc[DIM], det, rd;
};
template <int DIM, int N, int QC>
__device__ __forceinline__ void window_geom(const double (&X)[N][DIM], WindowGeom<DIM>& g) {
  g.det = window_geometry<DIM, N, QC>(X, g.c);
  g.rd = rcp_nr(g.det);
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    g.sc[a] = g.c[0][a];
#pragma unroll
    for (int k = 1; k < DIM; k++) g.sc[a] += g.c[k][a];
  }
}

// Carried flavour (3-D): the cofactor of the window's OLDEST node, c[0] = e_mid x e_new, is next step's c[2] (the window
// slides by one node), operand for operand -- so every step computes only c[0] (also a non-computing one: its successor
// needs it) and a computing step one more cross product instead of three. Bitwise the same numbers.
template <int N, int QC>
__device__ __forceinline__ void window_cross_new(const double (&X)[N][3], double (&cn)[3]) {
  constexpr int DIM = 3;
  const double(&p)[3] = X[WQ(1)];
  const double(&q)[3] = X[WQ(2)];
  cn[0] = p[1] * q[2] - p[2] * q[1];
  cn[1] = p[2] * q[0] - p[0] * q[2];
  cn[2] = p[0] * q[1] - p[1] * q[0];
}
template <int N, int QC>
__device__ __forceinline__ void window_geom_carry(const double (&X)[N][3], const double (&carry)[3], const double (&cn)[3],
                                                  WindowGeom<3>& g) {
  constexpr int DIM = 3;
  const double(&p)[3] = X[WQ(2)];
  const double(&q)[3] = X[WQ(0)];
  g.c[1][0] = p[1] * q[2] - p[2] * q[1];
  g.c[1][1] = p[2] * q[0] - p[0] * q[2];
  g.c[1][2] = p[0] * q[1] - p[1] * q[0];
  double det = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    g.c[0][a] = cn[a];
    g.c[2][a] = carry[a];
    det = fma(X[WQ(0)][a], cn[a], det);
  }
  g.det = det;
  g.rd = rcp_nr(det);
#pragma unroll
  for (int a = 0; a < 3; a++) g.sc[a] = (g.c[0][a] + g.c[1][a]) + g.c[2][a];
}

template <int DIM, int N, int QC, bool FULLV, bool ROWSUM = false>
__device__ __forceinline__ void mom_terms(MomState<DIM, N>& s, const StripConsts& k_, const WindowGeom<DIM>& g) {
  // density-weighted mass row M_0k = |J| sum_l Q_0kl rho_l (without |J|): M_00 = Qa rho_0 + Qaab S,
  // M_0k = Qd (rho_0 + rho_k) + Qabc S with S = rho_0 + sum_k rho_k; the products with rho_0 are row constants
  double S = s.rho0;
#pragma unroll
  for (int k = 0; k < DIM; k++) S += s.R[WQ(k)];
  double w[DIM];
  if constexpr (DIM == 3 && !kStripSeparateSums) {
    const double tS = fma(k_.Qabc, S, s.qd_rho0);
    const double M0 = fma(k_.Qaab, S, s.qa_rho0);
#pragma unroll
    for (int a = 0; a < DIM; a++) w[a] = M0 * s.U0[a];
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      const double Mk = fma(k_.Qd, s.R[WQ(k)], tS);
#pragma unroll
      for (int a = 0; a < DIM; a++) w[a] = fma(Mk, s.U[WQ(k)][a], w[a]);
    }
  } else {
    const double QS = k_.Qabc * S;
    const double M0 = fma(k_.Qa, s.rho0, k_.Qaab * S);
#pragma unroll
    for (int a = 0; a < DIM; a++) w[a] = M0 * s.U0[a];
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      const double Mk = fma(k_.Qd, s.rho0 + s.R[WQ(k)], QS);
#pragma unroll
      for (int a = 0; a < DIM; a++) w[a] = fma(Mk, s.U[WQ(k)][a], w[a]);
    }
  }
  // v / det with v = |det| (w + Wsum V gradN_0), gradN_0 = -sc / det
  double u[DIM];
  row_vector<DIM, FULLV>(k_, w, g.sc, g.rd, g.det, u);
  // entry (0, k) = u . c_k accumulates straight into the column's register; the diagonal takes -sum_k u . c_k = -u . sc
  // (3-D; in 2-D, where the dot products are two terms long, the separate sum measured faster on the B200)
  if constexpr (DIM == 3 && !kStripSeparateSums) {
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double ak = s.A[WQ(k)];
#pragma unroll
      for (int a = 0; a < DIM; a++) ak = fma(u[a], g.c[k][a], ak);
      s.A[WQ(k)] = ak;
    }
    if constexpr (!ROWSUM) {  // (ROWSUM: the caller subtracts every flushed column from a0)
      double tot = u[0] * g.sc[0];
#pragma unroll
      for (int a = 1; a < DIM; a++) tot = fma(u[a], g.sc[a], tot);
      s.a0 -= tot;
    }
  } else {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double sk = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) sk = fma(u[a], g.c[k][a], sk);
      s.A[WQ(k)] += sk;
      tot += sk;
    }
    s.a0 -= tot;
  }
  const double ad = fabs(g.det);
  double Sb = s.b0;
#pragma unroll
  for (int k = 0; k < DIM; k++) Sb += s.B[WQ(k)];
  if constexpr (DIM == 3 && !kStripSeparateSums) {
    s.msum = fma(ad, fma(k_.Po, S, s.pm_rho0), s.msum);
    s.nbsum = fma(ad, fma(k_.Po, Sb, s.pm_b0), s.nbsum);
  } else {
    s.msum = fma(ad, fma(k_.PdPo, s.rho0, k_.Po * S), s.msum);
    s.nbsum = fma(ad, fma(k_.PdPo, s.b0, k_.Po * Sb), s.nbsum);
  }
}

template <int DIM, int N, int QC, bool FULLV, bool ROWSUM = false>
__device__ __forceinline__ void mom_compute(MomState<DIM, N>& s, const StripConsts& k_) {
  WindowGeom<DIM> g;
  window_geom<DIM, N, QC>(s.X, g);
  mom_terms<DIM, N, QC, FULLV, ROWSUM>(s, k_, g);
}

// rows of the warp -> global memory, LPR = 1 << lpr_shift lanes per row
template <int NOUT>
__device__ __forceinline__ void write_rows(const double* __restrict__ acc, int t, int my_s0, int my_len, int lpr_shift,
                                           size_t nnz, double* __restrict__ out) {
  const int lane = t & 31, wbase = t & ~31;
  const int lpr = 1 << lpr_shift, rpi = 32 >> lpr_shift;
  const int sub = lane >> lpr_shift, sl = lane & (lpr - 1);
  for (int rr = 0; rr < 32; rr += rpi) {
    const int src = rr + sub;
    const int s0r = __shfl_sync(0xffffffffu, my_s0, src);
    const int lr = __shfl_sync(0xffffffffu, my_len, src);
    for (int ss = sl; ss < lr; ss += lpr) {
      const double v = acc[ss * kAS + wbase + src];
#pragma unroll
      for (int d = 0; d < NOUT; d++) __stcs(out + (size_t)d * nnz + s0r + ss, v);
    }
  }
}

// ---- tracer state and element math ----------------------------------------------------------------------
template <int DIM, int N>
struct AdvState {
  double X[N][DIM], U[N][DIM], T[N], A[N], C[N];  // C: sum of |det| over the elements sharing the edge (mass)
  int meta[N];
  double X0[DIM], cU0[DIM], T0;  // cU0 = Pd * nu(row node): adv_row_const
  double a0, c0, rhs;
  double cc[DIM];  // carried cross product (kStripCarry)
};

// the row-constant part of the advecting-velocity moment: (Pd - Po) nu_0 + Po nu_0
template <int DIM>
__device__ __forceinline__ void adv_row_const(const StripConsts& k_, const double (&U0)[DIM], double (&cU0)[DIM]) {
#pragma unroll
  for (int a = 0; a < DIM; a++) cU0[a] = (k_.PdPo + k_.Po) * U0[a];
}

// Advection_Diffusion_CG.F90:909-920 (consistent mass), :1093-1098 with beta = 0, :1192 (constant
// isotropic diffusivity), :1125,1200 (rhs -= (A + D) T)
// the tracer terms of one window on explicit operands (the fused momentum + tracer kernel shares the momentum state's
// velocity buffers and the geometry)
template <int DIM, int N, int QC, bool FULLV, bool ROWSUM = false>
__device__ __forceinline__ void adv_terms(const StripConsts& k_, const WindowGeom<DIM>& g, const double (&U)[N][DIM],
                                          const double (&cU0)[DIM], double (&A)[N], double (&C)[N], double& a0, double& c0) {
  // v = (Pd - Po) nu_0 + Po (nu_0 + sum_k nu_k); cU0 = Pd nu_0 is the same for every element of the row
  double v[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    double Su = U[WQ(0)][a];
#pragma unroll
    for (int k = 1; k < DIM; k++) Su += U[WQ(k)][a];
    v[a] = fma(k_.Po, Su, cU0[a]);
  }
  double u[DIM];
  row_vector<DIM, FULLV>(k_, v, g.sc, g.rd, g.det, u);
  const double ad = fabs(g.det);
  // entry (0, k) = u . c_k accumulates straight into the column's register, the diagonal takes -u . sc; the products with
  // T (rhs -= (A + D) T, Advection_Diffusion_CG.F90:1125,1200) are taken once per column when it leaves the FIFO
  // (adv_evict_rhs) and once per row for the diagonal (adv_finish_rhs): the right-hand side is linear in the entries
  if constexpr (DIM == 3 && !kStripSeparateSums) {
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double ak = A[WQ(k)];
#pragma unroll
      for (int a = 0; a < DIM; a++) ak = fma(u[a], g.c[k][a], ak);
      A[WQ(k)] = ak;
      C[WQ(k)] += ad;
    }
    if constexpr (!ROWSUM) {
      double tot = u[0] * g.sc[0];
#pragma unroll
      for (int a = 1; a < DIM; a++) tot = fma(u[a], g.sc[a], tot);
      a0 -= tot;
    }
  } else {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double sk = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) sk = fma(u[a], g.c[k][a], sk);
      A[WQ(k)] += sk;
      C[WQ(k)] += ad;
      tot += sk;
    }
    a0 -= tot;
  }
  c0 += ad;
}

template <int DIM, int N, int QC, bool FULLV, bool ROWSUM = false>
__device__ __forceinline__ void adv_compute(AdvState<DIM, N>& s, const StripConsts& k_) {
  WindowGeom<DIM> g;
  window_geom<DIM, N, QC>(s.X, g);
  adv_terms<DIM, N, QC, FULLV, ROWSUM>(k_, g, s.U, s.cU0, s.A, s.C, s.a0, s.c0);
}

// rhs -= entry * T(column) for the column leaving the FIFO (a = its accumulated, unscaled entry, tk = its T) ...
__device__ __forceinline__ void adv_evict_rhs(double& rhs, double a, double tk) { rhs = fma(-a, tk, rhs); }
// ... and for the diagonal, after the row's last element
__device__ __forceinline__ void adv_finish_rhs(double& rhs, double a0, double t0) { rhs = fma(-a0, t0, rhs); }

// the row's own absorption / source values (tracer kernel with ABS)
struct AdvOwnExtra {
  double sg0, sq0;
};

// adv_compute plus nodal absorption and source: Ab_0k = |J| [Qa s_0 + Qaab S | Qd (s_0 + s_k) + Qabc S]
// (Advection_Diffusion_CG.F90:1156: shape_shape with the absorption at the quadrature points; scaled by dt*theta
// with the advective / diffusive entries, rhs -= Ab T), source rhs_0 += |J| [(Pd - Po) q_0 + Po sum q] (:1139)
template <int DIM, int N, int QC, bool FULLV>
__device__ __forceinline__ void adv_compute_abs(AdvState<DIM, N>& s, const double (&sg)[N], const double (&sq)[N],
                                                const AdvOwnExtra& ox, const StripConsts& k_) {
  double c[DIM][DIM];
  const double det = window_geometry<DIM, N, QC>(s.X, c);
  const double rd = rcp_nr(det);
  double sc[DIM], v[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    sc[a] = c[0][a];
    double Su = s.U[WQ(0)][a];
#pragma unroll
    for (int k = 1; k < DIM; k++) sc[a] += c[k][a];
#pragma unroll
    for (int k = 1; k < DIM; k++) Su += s.U[WQ(k)][a];
    v[a] = fma(k_.Po, Su, s.cU0[a]);
  }
  double u[DIM];
  row_vector<DIM, FULLV>(k_, v, sc, rd, det, u);
  const double ad = fabs(det);
  double Ss = ox.sg0, Sq = ox.sq0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    Ss += sg[WQ(k)];
    Sq += sq[WQ(k)];
  }
  const double QS = k_.Qabc * Ss;
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    double sk = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) sk = fma(u[a], c[k][a], sk);
    s.A[WQ(k)] += fma(ad, fma(k_.Qd, ox.sg0 + sg[WQ(k)], QS), sk);
    s.C[WQ(k)] += ad;
    tot += sk;
  }
  s.a0 += fma(ad, fma(k_.Qa, ox.sg0, k_.Qaab * Ss), -tot);
  s.c0 += ad;
  s.rhs = fma(ad, fma(k_.sPdPo, ox.sq0, k_.sPo * Sq), s.rhs);  // the source; the products with T: adv_evict_rhs / adv_finish_rhs
}

#undef WQ

// ---- host helpers ---------------------------------------------------------------------------------------
inline StripPlanView plan_view(const Handle* h) {
  const GatherPlan* P = h->gather;
  StripPlanView v;
  v.rows = P->d_rows;
  v.ptr = P->d_strip_ptr;
  v.ent = P->d_strip;
  v.own_slot = P->d_own_slot;
  v.findrm = h->d_findrm;
  v.colm = h->d_colm;
  v.maxlen = P->maxlen;
  int sh = 0;
  while ((1 << sh) < P->maxlen && sh < 5) sh++;
  v.lpr_shift = sh;
  return v;
}

// which option sets the STRIP kernels cover (everything else runs the GATHER kernels)
// (absorption, sources and the reference profile are added by the additive pass, strip_extra.cu: strip_momentum_ok
// asks strip_extra_ok for them)
inline bool strip_momentum_opts_ok(const MomentumArgs& A) {
  const cgasm_momentum_opts& o = A.o;
  if (o.stabilisation_scheme != CGASM_STAB_NONE) return false;
  if (!o.exclude_mass && !o.lump_mass) return false;
  if (!o.exclude_advection && (o.integrate_advection_by_parts || o.beta != 0.0)) return false;
  if (o.have_gravity && A.gravity.stride != 0) return false;
  return A.tab.sym && (!o.have_viscosity || A.viscosity.stride == 0);
}
// (nodal or constant absorption and source: staged kernels only, strip_advdiff_ok checks that)
inline bool strip_advdiff_opts_ok(const AdvDiffArgs& A) {
  const cgasm_advdiff_opts& o = A.o;
  if (o.stabilisation_scheme != CGASM_STAB_NONE) return false;
  if (o.have_advection && (o.integrate_advection_by_parts || o.beta != 0.0)) return false;
  return A.tab.sym && (!o.have_diffusivity || A.diffusivity.stride == 0);
}
inline bool strip_advdiff_needs_extra(const AdvDiffArgs& A) { return A.o.have_absorption || A.o.have_source; }
// full constant tensor needed? (isotropic fields only use V[0])
inline bool strip_full_tensor(int have, int shape) { return have && shape != CGASM_TENSOR_ISOTROPIC; }

inline void consts_tensor(StripConsts& c, int dim, int have, int shape, const double* t, double wsum) {
  for (int q = 0; q < 9; q++) c.V[q] = 0.0;
  if (!have) return;
  if (shape == CGASM_TENSOR_ISOTROPIC) {
    c.V[0] = t[0] * wsum;
  } else {
    for (int b = 0; b < dim; b++)
      for (int a = 0; a < dim; a++)
        c.V[a + dim * b] = (shape == CGASM_TENSOR_DIAGONAL && a != b) ? 0.0 : t[a + dim * b] * wsum;
  }
}

inline StripConsts consts_momentum(const Handle* h, const MomentumArgs& A) {
  const Tables& t = A.tab;
  const cgasm_momentum_opts& o = A.o;
  StripConsts c{};
  const double adv = o.exclude_advection ? 0.0 : 1.0;
  c.Qa = adv * (t.Qaaa - t.Qaab);
  c.Qaab = adv * t.Qaab;
  c.Qd = adv * (t.Qaab - t.Qabc);
  c.Qabc = adv * t.Qabc;
  c.PdPo = t.Pd - t.Po;
  c.Po = t.Po;
  c.mass_on = o.exclude_mass ? 0.0 : 1.0;
  consts_tensor(c, h->dim, o.have_viscosity, o.viscosity_shape, h->fields[CGASM_F_VISCOSITY].h_const, t.Wsum);
  for (int d = 0; d < 3; d++)
    c.grav[d] = (o.have_gravity && d < h->dim) ? o.gravity_magnitude * h->fields[CGASM_F_GRAVITY].h_const[d] : 0.0;
  c.dtt = o.dt * o.theta;
  return c;
}

inline StripConsts consts_advdiff(const Handle* h, const AdvDiffArgs& A) {
  const Tables& t = A.tab;
  const cgasm_advdiff_opts& o = A.o;
  StripConsts c{};
  const double adv = o.have_advection ? 1.0 : 0.0;
  c.PdPo = adv * (t.Pd - t.Po);
  c.Po = adv * t.Po;
  c.mPd = !o.have_mass ? 0.0 : (o.lump_mass ? t.W1 : t.Pd);
  c.mPo = !o.have_mass ? 0.0 : (o.lump_mass ? 0.0 : t.Po);
  consts_tensor(c, h->dim, o.have_diffusivity, o.diffusivity_shape, h->fields[CGASM_F_T_DIFFUSIVITY].h_const, t.Wsum);
  const double dtt = o.dt * o.theta;
  c.dtt = fabs(dtt) > 2.220446049250313e-16 ? dtt : 0.0;
  const double ab = o.have_absorption ? 1.0 : 0.0, sr = o.have_source ? 1.0 : 0.0;
  c.Qa = ab * (t.Qaaa - t.Qaab);
  c.Qaab = ab * t.Qaab;
  c.Qd = ab * (t.Qaab - t.Qabc);
  c.Qabc = ab * t.Qabc;
  c.sPdPo = sr * (t.Pd - t.Po);
  c.sPo = sr * t.Po;
  return c;
}

template <class K>
inline int strip_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CGASM_OK;
}

// strip_staged.cu
bool strip_staged_ok(const Handle* h, bool momentum);
int strip_staged_momentum(Handle* h, const MomentumArgs& A);
int strip_staged_advdiff(Handle* h, const AdvDiffArgs& A);

}  // namespace cgasm
