// strip.cu -- CGASM_SCATTER_STRIP: row-owner assembly with a strip-ordered element stream.
//
// One thread owns one CSR row (node r) and recomputes ITS row of every incident element in closed
// form (element_math.cuh row kernels: P1 simplices + node-symmetric degree-3 rule,
// femtools/Quadrature.F90:690-708,951-970), like the GATHER walk kernels, but restructured around
// what ncu showed to bound them (profiles/r1_kernel_history.md #13: exposed load latency, 3 x 32-byte
// records and 4 shared-memory read-modify-writes per pair, ~160 registers):
//
//  * The elements around r are visited as a STRIP over r's link (strip_plan.h): every entry pushes
//    one node into a FIFO of dim nodes and the oldest node drops out, so all threads replace the
//    same register set at the same step -- the node loop is unrolled over N rotating register
//    buffers (N = dim + prefetch distance) with no selects and no divergence in the load path.
//  * The records of entry j+PD are requested while entry j is being computed (register prefetch).
//  * A contribution to column v accumulates in a REGISTER for as long as v sits in the FIFO and is
//    added to the row's shared-memory slot once, when v is evicted (1.4 RMW per pair instead of 4).
//  * rhs -= (A + K) oldu (Momentum_CG.F90:1712,2346) is linear in the assembled row, so it is one
//    sparse dot product per row after the loop (colm + the {oldu, buoyancy} records) instead of
//    dim*loc FMAs and one more 32-byte record per pair. Likewise the constant gravity direction
//    (:1786-1789) and dt*theta (:1484-1486) are applied once per row.
//  * Momentum reads two records per node, {X, buoyancy} and {nu, density}.
//
// Entries per (row, element) pair: 1.375 on Kuhn meshes (33 pushes for the 24 elements of an
// interior node). Summation order is fixed by the plan: results are bitwise reproducible.
// Option coverage: the common option set only (momentum_common_ok / advdiff_common_ok); everything
// else runs the GATHER kernels (gather.cu).
#include "gather_plan.h"
#include "strip_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace cgasm {

constexpr int kAS = kBR + 1;  // stride (doubles) between slots of the accumulator: odd, so that the
                              // write-out (one row spread over consecutive lanes) is conflict-free

struct StripConsts {  // passed by value: operands are read straight from the constant bank
  double Qa, Qaab, Qd, Qabc;  // Qa = Qaaa - Qaab, Qd = Qaab - Qabc (Tables)
  double PdPo, Po, Pd;        // PdPo = Pd - Po
  double Wsum;
  double dtt;                 // dt*theta (tracer: 0 unless |dt*theta| > epsilon, Advection_Diffusion_CG.F90:1121)
  double gmag;                // gravity_magnitude
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

// 1/x: MUFU.RCP64H seed + two Newton steps (what the compiler's own division starts from, minus
// the special-case branch; det of a valid element is a normal number)
__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
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
};

// row 0 (the row's own node) of the element {r, window}: Momentum_CG.F90:1535-1552 (lumped mass),
// :1675-1680 with beta = 0 (advection), :2304-2317 (constant isotropic viscosity), :1770-1789 (buoyancy)
template <int DIM, int N, int QC>
__device__ __forceinline__ void mom_compute(MomState<DIM, N>& s, const StripConsts& k_, double muW) {
  double c[DIM][DIM];
  const double det = window_geometry<DIM, N, QC>(s.X, c);
  const double rd = rcp_nr(det);
  double sc[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    sc[a] = c[0][a];
#pragma unroll
    for (int k = 1; k < DIM; k++) sc[a] += c[k][a];
  }
  double S = s.rho0;
#pragma unroll
  for (int k = 0; k < DIM; k++) S += s.R[WQ(k)];
  const double QS = k_.Qabc * S;
  const double M0 = fma(k_.Qa, s.rho0, k_.Qaab * S);
  double w[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) w[a] = M0 * s.U0[a];
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    const double Mk = fma(k_.Qd, s.rho0 + s.R[WQ(k)], QS);
#pragma unroll
    for (int a = 0; a < DIM; a++) w[a] = fma(Mk, s.U[WQ(k)][a], w[a]);
  }
  // v / det with v = |det| (w + mu Wsum gradN_0), gradN_0 = -sc / det
  const double tt = muW * rd;
  const unsigned sgn = (unsigned)__double2hiint(det) & 0x80000000u;
  double u[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) u[a] = flip_sign(fma(-tt, sc[a], w[a]), sgn);
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    double sk = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) sk = fma(u[a], c[k][a], sk);
    s.A[WQ(k)] += sk;
    tot += sk;
  }
  s.a0 -= tot;
  const double ad = fabs(det);
  s.msum = fma(ad, fma(k_.PdPo, s.rho0, k_.Po * S), s.msum);
  double Sb = s.b0;
#pragma unroll
  for (int k = 0; k < DIM; k++) Sb += s.B[WQ(k)];
  s.nbsum = fma(ad, fma(k_.PdPo, s.b0, k_.Po * Sb), s.nbsum);
}

template <int DIM, int N, int QC>
__device__ __forceinline__ void mom_step(MomState<DIM, N>& s, const StripConsts& k_, double muW, int j, int deg,
                                         const int2* __restrict__ p, int2& pq0, int2& pq1, const int2 pad,
                                         double* __restrict__ acc_t, const double4* __restrict__ rX,
                                         const double4* __restrict__ rU) {
  constexpr int PD = N - DIM;
  constexpr int QE = (QC + PD) % N;  // buffer of entry j - DIM: evicted now, refilled with entry j + PD
  const int2 en = pq0;
  pq0 = pq1;
  pq1 = (j + PD + 2 < deg) ? ldg_stream2(p + (long long)(j + PD + 2) * kBR) : pad;
  {
    double* sl = acc_t + (s.meta[QE] & 0xff) * kAS;
    *sl += s.A[QE];
    s.A[QE] = 0.0;
  }
  unpack<DIM>(ld256(rX + en.x), s.X[QE], s.B[QE]);
  unpack<DIM>(ld256(rU + en.x), s.U[QE], s.R[QE]);
  s.meta[QE] = en.y;
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if (s.meta[QC] & kStripCompute) mom_compute<DIM, N, QC>(s, k_, muW);
}

template <int DIM, int N, int Q>
struct MomUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, N>& s, const StripConsts& k_, double muW, int j0, Args&&... args) {
    mom_step<DIM, N, Q>(s, k_, muW, j0 + Q, args...);
    if constexpr (Q + 1 < N) MomUnroll<DIM, N, Q + 1>::run(s, k_, muW, j0, args...);
  }
};

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

template <int DIM, int N, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
strip_momentum_kernel(const StripConsts k_, const StripPlanView P, const double4* __restrict__ rX,
                      const double4* __restrict__ rU, const double4* __restrict__ rO,
                      const double* __restrict__ viscosity, const double* __restrict__ gravity, size_t nnz,
                      double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  constexpr int PD = N - DIM;
  extern __shared__ double acc[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = P.rows[b * kBR + t];
  const int r0 = r >= 0 ? r : 0;
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);  // a multiple of N
  const int2* p = P.ent + base + t;
  double* acc_t = acc + t;
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  const int own = P.own_slot[b * kBR + t];
  const int2 pad = make_int2(r0, own);
  const double muW = __ldg(viscosity) * k_.Wsum;
  MomState<DIM, N> s;
  unpack<DIM>(ld256(rX + r0), s.X0, s.b0);
  unpack<DIM>(ld256(rU + r0), s.U0, s.rho0);
  s.a0 = s.msum = s.nbsum = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.R[q] = s.B[q] = s.A[q] = 0.0;
    s.meta[q] = own;
  }
  // prologue: entries 0 .. PD-1 into buffers 0 .. PD-1, entries PD and PD+1 into the queue
#pragma unroll
  for (int q = 0; q < PD; q++) {
    const int2 en = q < deg ? ldg_stream2(p + (long long)q * kBR) : pad;
    unpack<DIM>(ld256(rX + en.x), s.X[q], s.B[q]);
    unpack<DIM>(ld256(rU + en.x), s.U[q], s.R[q]);
    s.meta[q] = en.y;
  }
  int2 pq0 = PD < deg ? ldg_stream2(p + (long long)PD * kBR) : pad;
  int2 pq1 = PD + 1 < deg ? ldg_stream2(p + (long long)(PD + 1) * kBR) : pad;
  for (int j0 = 0; j0 < deg; j0 += N) MomUnroll<DIM, N, 0>::run(s, k_, muW, j0, deg, p, pq0, pq1, pad, acc_t, rX, rU);
#pragma unroll
  for (int q = 0; q < N; q++) acc_t[(s.meta[q] & 0xff) * kAS] += s.A[q];
  acc_t[own * kAS] += s.a0;
  // row epilogue: rhs = gravity term - sum_s (A+K)_s oldu(col_s); big_m = dt theta (A+K) + lumped mass on the diagonal
  int my_s0 = 0, my_len = 0;
  if (r >= 0) {
    my_s0 = P.findrm[r];
    my_len = P.findrm[r + 1] - my_s0;
    double rh[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = (k_.gmag * __ldg(gravity + d)) * s.nbsum;
    for (int q = 0; q < my_len; q++) {
      const int col = __ldg(P.colm + my_s0 + q);
      const double4 o = ld256(rO + col);
      const double v = acc_t[q * kAS];
      rh[0] = fma(-v, o.x, rh[0]);
      rh[1] = fma(-v, o.y, rh[1]);
      if constexpr (DIM == 3) rh[2] = fma(-v, o.z, rh[2]);
      acc_t[q * kAS] = fma(k_.dtt, v, q == own ? s.msum : 0.0);
    }
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      rhs[(size_t)DIM * r + d] = rh[d];
      if (masslump) masslump[(size_t)DIM * r + d] = s.msum;
    }
  }
  __syncwarp();
  write_rows<DIM>(acc, t, my_s0, my_len, P.lpr_shift, nnz, big_m);
}

// ---- tracer -------------------------------------------------------------------------------------------
template <int DIM, int N>
struct AdvState {
  double X[N][DIM], U[N][DIM], T[N], A[N], C[N];  // C: sum of |det| over the elements sharing the edge (mass)
  int meta[N];
  double X0[DIM], U0[DIM], T0;
  double a0, c0, rhs;
};

// Advection_Diffusion_CG.F90:909-920 (consistent mass), :1093-1098 with beta = 0, :1192 (constant
// isotropic diffusivity), :1125,1200 (rhs -= (A + D) T)
template <int DIM, int N, int QC>
__device__ __forceinline__ void adv_compute(AdvState<DIM, N>& s, const StripConsts& k_, double kW) {
  double c[DIM][DIM];
  const double det = window_geometry<DIM, N, QC>(s.X, c);
  const double rd = rcp_nr(det);
  double sc[DIM], v[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    sc[a] = c[0][a];
    double Su = s.U0[a];
#pragma unroll
    for (int k = 1; k < DIM; k++) sc[a] += c[k][a];
#pragma unroll
    for (int k = 0; k < DIM; k++) Su += s.U[WQ(k)][a];
    v[a] = fma(k_.PdPo, s.U0[a], k_.Po * Su);
  }
  const double tt = kW * rd;
  const unsigned sgn = (unsigned)__double2hiint(det) & 0x80000000u;
  double u[DIM];
#pragma unroll
  for (int a = 0; a < DIM; a++) u[a] = flip_sign(fma(-tt, sc[a], v[a]), sgn);
  const double ad = fabs(det);
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    double sk = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) sk = fma(u[a], c[k][a], sk);
    s.A[WQ(k)] += sk;
    s.C[WQ(k)] += ad;
    s.rhs = fma(-sk, s.T[WQ(k)], s.rhs);
    tot += sk;
  }
  s.a0 -= tot;
  s.c0 += ad;
  s.rhs = fma(tot, s.T0, s.rhs);
}

template <int DIM, int N, int QC>
__device__ __forceinline__ void adv_step(AdvState<DIM, N>& s, const StripConsts& k_, double kW, int j, int deg,
                                         const int2* __restrict__ p, int2& pq0, int2& pq1, const int2 pad,
                                         double* __restrict__ acc_t, const double4* __restrict__ rX,
                                         const double4* __restrict__ rU) {
  constexpr int PD = N - DIM;
  constexpr int QE = (QC + PD) % N;
  const int2 en = pq0;
  pq0 = pq1;
  pq1 = (j + PD + 2 < deg) ? ldg_stream2(p + (long long)(j + PD + 2) * kBR) : pad;
  {
    double* sl = acc_t + (s.meta[QE] & 0xff) * kAS;
    *sl += fma(k_.dtt, s.A[QE], k_.Po * s.C[QE]);
    s.A[QE] = 0.0;
    s.C[QE] = 0.0;
  }
  double unused;
  unpack<DIM>(ld256(rX + en.x), s.X[QE], s.T[QE]);
  unpack<DIM>(ld256(rU + en.x), s.U[QE], unused);
  s.meta[QE] = en.y;
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if (s.meta[QC] & kStripCompute) adv_compute<DIM, N, QC>(s, k_, kW);
}

template <int DIM, int N, int Q>
struct AdvUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(AdvState<DIM, N>& s, const StripConsts& k_, double kW, int j0, Args&&... args) {
    adv_step<DIM, N, Q>(s, k_, kW, j0 + Q, args...);
    if constexpr (Q + 1 < N) AdvUnroll<DIM, N, Q + 1>::run(s, k_, kW, j0, args...);
  }
};

template <int DIM, int N, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
strip_advdiff_kernel(const StripConsts k_, const StripPlanView P, const double4* __restrict__ rX,
                     const double4* __restrict__ rU, const double* __restrict__ diffusivity,
                     double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int PD = N - DIM;
  extern __shared__ double acc[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = P.rows[b * kBR + t];
  const int r0 = r >= 0 ? r : 0;
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);
  const int2* p = P.ent + base + t;
  double* acc_t = acc + t;
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  const int own = P.own_slot[b * kBR + t];
  const int2 pad = make_int2(r0, own);
  const double kW = __ldg(diffusivity) * k_.Wsum;
  AdvState<DIM, N> s;
  double unused;
  unpack<DIM>(ld256(rX + r0), s.X0, s.T0);
  unpack<DIM>(ld256(rU + r0), s.U0, unused);
  s.a0 = s.c0 = s.rhs = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.T[q] = s.A[q] = s.C[q] = 0.0;
    s.meta[q] = own;
  }
#pragma unroll
  for (int q = 0; q < PD; q++) {
    const int2 en = q < deg ? ldg_stream2(p + (long long)q * kBR) : pad;
    unpack<DIM>(ld256(rX + en.x), s.X[q], s.T[q]);
    unpack<DIM>(ld256(rU + en.x), s.U[q], unused);
    s.meta[q] = en.y;
  }
  int2 pq0 = PD < deg ? ldg_stream2(p + (long long)PD * kBR) : pad;
  int2 pq1 = PD + 1 < deg ? ldg_stream2(p + (long long)(PD + 1) * kBR) : pad;
  for (int j0 = 0; j0 < deg; j0 += N) AdvUnroll<DIM, N, 0>::run(s, k_, kW, j0, deg, p, pq0, pq1, pad, acc_t, rX, rU);
#pragma unroll
  for (int q = 0; q < N; q++) acc_t[(s.meta[q] & 0xff) * kAS] += fma(k_.dtt, s.A[q], k_.Po * s.C[q]);
  acc_t[own * kAS] += fma(k_.dtt, s.a0, k_.Pd * s.c0);
  int my_s0 = 0, my_len = 0;
  if (r >= 0) {
    my_s0 = P.findrm[r];
    my_len = P.findrm[r + 1] - my_s0;
    rhs[r] = s.rhs;
  }
  __syncwarp();
  write_rows<1>(acc, t, my_s0, my_len, P.lpr_shift, 0, matrix);
}
#undef WQ

// ---- plan ---------------------------------------------------------------------------------------------
static int strip_nbuf(int dim) {
  const char* e = getenv("CGASM_STRIP_NBUF");
  int n = e ? atoi(e) : dim + 1;
  return std::min(std::max(n, dim + 1), dim + 2);
}

int strip_build(Handle* h) {
  GatherPlan* P = h->gather;
  if (!P) CG_FAIL(CGASM_ESTATE, "strip scatter: the gather row blocks must exist first");
  if (P->d_strip) return CGASM_OK;
  const int nb = P->nblocks, loc = h->loc;
  const std::vector<int>& rows = P->h_rows;
  std::vector<std::vector<StripEntry>> rowplans(rows.size());
  std::vector<int> block_deg((size_t)nb, 0);
  long long total_real = 0;
#pragma omp parallel
  {
    std::vector<StripEntry> tmp;
#pragma omp for schedule(dynamic, 8) reduction(+ : total_real)
    for (int b = 0; b < nb; b++) {
      int deg = 0;
      for (int t = 0; t < kBR; t++) {
        const size_t q = (size_t)b * kBR + t;
        const int r = rows[q];
        if (r < 0) continue;
        build_strip_row(loc, h->h_nd0.data(), h->n2e_ptr.data(), h->n2e.data(), h->h_findrm.data(), h->h_colm.data(), r, tmp);
        rowplans[q] = tmp;
        deg = std::max(deg, (int)tmp.size());
        total_real += (long long)tmp.size();
      }
      block_deg[b] = deg;
    }
  }
  std::vector<long long> ptr((size_t)nb + 1, 0);
  const int mult = strip_nbuf(h->dim);
  for (int b = 0; b < nb; b++) {
    const int deg = (block_deg[b] + mult - 1) / mult * mult;
    ptr[b + 1] = ptr[b] + (long long)deg * kBR;
  }
  P->strip_mult = mult;
  P->n_strip = ptr[nb];
  std::vector<int2> ent((size_t)std::max<long long>(P->n_strip, 1));
#pragma omp parallel for schedule(dynamic, 8)
  for (int b = 0; b < nb; b++) {
    const int deg = (int)((ptr[b + 1] - ptr[b]) / kBR);
    for (int t = 0; t < kBR; t++) {
      const size_t q = (size_t)b * kBR + t;
      const int r = rows[q];
      const auto& rp = rowplans[q];
      int own = 0;
      if (r >= 0) {
        const int* cb = h->h_colm.data() + h->h_findrm[r];
        own = (int)(std::lower_bound(cb, (const int*)h->h_colm.data() + h->h_findrm[r + 1], r) - cb);
      }
      for (int k = 0; k < deg; k++) {
        int2 v = make_int2(r >= 0 ? r : 0, own);  // padding: re-push the own node, nothing computed
        if (k < (int)rp.size()) v = make_int2(rp[k].node, rp[k].meta);
        ent[(size_t)(ptr[b] + (long long)k * kBR + t)] = v;
      }
    }
  }
  P->strip_entries_per_pair = h->n2e.empty() ? 0.0 : (double)total_real / (double)h->n2e.size();
  if (getenv("CGASM_DEBUG"))
    fprintf(stderr, "[cgasm] strip plan: %.3f entries per (row, element) pair, %lld padded entries\n",
            P->strip_entries_per_pair, P->n_strip);
  CG_CUDA(cudaMalloc(&P->d_strip_ptr, sizeof(long long) * ptr.size()));
  CG_CUDA(cudaMemcpy(P->d_strip_ptr, ptr.data(), sizeof(long long) * ptr.size(), cudaMemcpyHostToDevice));
  CG_CUDA(cudaMalloc(&P->d_strip, sizeof(int2) * ent.size()));
  CG_CUDA(cudaMemcpy(P->d_strip, ent.data(), sizeof(int2) * ent.size(), cudaMemcpyHostToDevice));
  if (!P->d_own_slot) {
    std::vector<unsigned char> own_slot(rows.size(), 0);
    for (size_t q = 0; q < rows.size(); q++) {
      const int r = rows[q];
      if (r < 0) continue;
      const int* cb = h->h_colm.data() + h->h_findrm[r];
      own_slot[q] = (unsigned char)(std::lower_bound(cb, (const int*)h->h_colm.data() + h->h_findrm[r + 1], r) - cb);
    }
    CG_CUDA(cudaMalloc(&P->d_own_slot, own_slot.size()));
    CG_CUDA(cudaMemcpy(P->d_own_slot, own_slot.data(), own_slot.size(), cudaMemcpyHostToDevice));
  }
  return CGASM_OK;
}

void strip_free(GatherPlan* P) {
  if (P->d_strip_ptr) cudaFree(P->d_strip_ptr);
  if (P->d_strip) cudaFree(P->d_strip);
  P->d_strip_ptr = nullptr;
  P->d_strip = nullptr;
}

// ---- launch -------------------------------------------------------------------------------------------
static StripPlanView plan_view(const Handle* h) {
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

static StripConsts consts_of(const Tables& t, double dtt, double gmag) {
  StripConsts c;
  c.Qa = t.Qaaa - t.Qaab;
  c.Qaab = t.Qaab;
  c.Qd = t.Qaab - t.Qabc;
  c.Qabc = t.Qabc;
  c.PdPo = t.Pd - t.Po;
  c.Po = t.Po;
  c.Pd = t.Pd;
  c.Wsum = t.Wsum;
  c.dtt = dtt;
  c.gmag = gmag;
  return c;
}

template <class K>
static int strip_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CGASM_OK;
}

bool strip_momentum_ok(const Handle* h, const MomentumArgs& A, bool want_ml) {
  const GatherPlan* P = h->gather;
  return P && P->d_strip && A.tab.sym && want_ml && !A.o.have_absorption &&
         momentum_fast_ok(A.o, A.gravity.stride, A.absorption.stride) && momentum_common_ok(A.o, A.viscosity.stride);
}

bool strip_advdiff_ok(const Handle* h, const AdvDiffArgs& A) {
  const GatherPlan* P = h->gather;
  return P && P->d_strip && A.tab.sym && advdiff_fast_ok(A.o) && advdiff_common_ok(A.o, A.diffusivity.stride);
}

template <int DIM>
static int strip_momentum_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = sizeof(double) * (size_t)P->maxlen * kAS;
  if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "strip scatter: CSR rows too long for the shared-memory accumulator");
  const StripConsts c = consts_of(A.tab, A.o.dt * A.o.theta, A.o.gravity_magnitude);
  const StripPlanView v = plan_view(h);
  const int minb = getenv("CGASM_STRIP_MINB") ? atoi(getenv("CGASM_STRIP_MINB")) : 3;
  int st;
#define LAUNCH(N_, MINB_)                                                                                      \
  do {                                                                                                         \
    if ((st = strip_smem(strip_momentum_kernel<DIM, N_, MINB_>, smem))) return st;                             \
    strip_momentum_kernel<DIM, N_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(                               \
        c, v, h->d_rec3, h->d_rec1, h->d_rec2, A.viscosity.val, A.gravity.val, (size_t)h->nnz, h->d_big_m,      \
        h->d_mom_rhs, h->d_masslump);                                                                          \
  } while (0)
  if (P->strip_mult == DIM + 2) {
    if (minb >= 3) LAUNCH(DIM + 2, 3);
    else LAUNCH(DIM + 2, 2);
  } else {
    if (minb >= 4) LAUNCH(DIM + 1, 4);
    else if (minb == 3) LAUNCH(DIM + 1, 3);
    else LAUNCH(DIM + 1, 2);
  }
#undef LAUNCH
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int strip_momentum(Handle* h, const MomentumArgs& A) {
  return h->dim == 3 ? strip_momentum_dim<3>(h, A) : strip_momentum_dim<2>(h, A);
}

template <int DIM>
static int strip_advdiff_dim(Handle* h, const AdvDiffArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = sizeof(double) * (size_t)P->maxlen * kAS;
  if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "strip scatter: CSR rows too long for the shared-memory accumulator");
  const double dtt = A.o.dt * A.o.theta;
  const StripConsts c = consts_of(A.tab, fabs(dtt) > 2.220446049250313e-16 ? dtt : 0.0, 0.0);
  const StripPlanView v = plan_view(h);
  const int minb = getenv("CGASM_STRIP_MINB_ADV") ? atoi(getenv("CGASM_STRIP_MINB_ADV")) : 4;
  int st;
#define LAUNCH(N_, MINB_)                                                                                      \
  do {                                                                                                         \
    if ((st = strip_smem(strip_advdiff_kernel<DIM, N_, MINB_>, smem))) return st;                              \
    strip_advdiff_kernel<DIM, N_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(                                \
        c, v, h->d_rec0, h->d_rec1, A.diffusivity.val, h->d_adv_matrix, h->d_adv_rhs);                          \
  } while (0)
  if (P->strip_mult == DIM + 2) {
    if (minb >= 4) LAUNCH(DIM + 2, 4);
    else LAUNCH(DIM + 2, 3);
  } else {
    if (minb >= 5) LAUNCH(DIM + 1, 5);
    else if (minb == 4) LAUNCH(DIM + 1, 4);
    else LAUNCH(DIM + 1, 3);
  }
#undef LAUNCH
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int strip_advdiff(Handle* h, const AdvDiffArgs& A) {
  return h->dim == 3 ? strip_advdiff_dim<3>(h, A) : strip_advdiff_dim<2>(h, A);
}

}  // namespace cgasm
