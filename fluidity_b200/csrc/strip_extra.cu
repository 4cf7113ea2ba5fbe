// strip_extra.cu -- the ADDITIVE momentum pass of the STRIP variant: everything construct_momentum_element_cg adds on
// top of the common terms that is linear in fields the common kernel never reads, for a CONSTANT density field
// (Boussinesq: the four example configs). It runs after staged_momentum_kernel on the same strip plan and
// adds to its results in place:
//   absorption       add_absorption_element_cg (assemble/Momentum_CG.F90:2036-2073): shape_shape_vector(test, shape,
//                    detwei*density_gi, absorption_gi). With constant density and a P1 vector absorption field s_d:
//                      Ab^d_0k = rho |J| sum_l Q_0kl s_dl = rho |J| [Qa s_d0 + Qaab S_d | Qd (s_d0 + s_dk) + Qabc S_d]
//                    (the density-weighted mass row of the common kernel with s_d in the place of rho). Full matrix:
//                    big_m(d,d) += dt theta Ab^d, rhs_d -= Ab^d oldu_d (:2060-2066). Lumped (:2047-2056):
//                      sum_k Ab^d_0k = rho |J| [(Pd - Po) s_d0 + Po S_d]  on the diagonal and against oldu_d(row);
//                    pressure-corrected: dt theta times the lumped rows into masslump too (:2068-2072).
//   sources          add_sources_element_cg (:1717-1751): rhs_d += rho |J| [(Pd - Po) q_d0 + Po sum q_d], or lumped:
//                    rhs_d += (rho sum_e |J| W1) q_d(row).
//   reference        subtract_out_reference_profile (:1767-1771): the buoyancy moments with hb_density, subtracted.
//   profile
// tests/strip_emulation.py holds the same closed forms in numpy on the same plan (CPU suite, against the oracle).
//
// Cost per (row, element) pair: one cross product + one dot (|J|) instead of the full cofactor geometry; without a full
// absorption matrix nothing is accumulated per column except |J| itself, and the fields are read once per strip entry
// when their node leaves the FIFO:  sum_e |J_e| sum_{k in e} f_k  =  sum_k f_k (sum_{e has k} |J_e|).
#include "strip_staged.cuh"

namespace cgasm {

struct ExtraConsts {
  double rho;                 // the constant density
  double Qa, Qaab, Qd, Qabc;  // full absorption matrix (0 otherwise)
  double lPdPo, lPo;          // lumped absorption moments (Pd - Po, Po), 0 otherwise
  double sPdPo, sPo;          // consistent source moments, 0 otherwise
  double sW1;                 // lumped source: W1 = sum_k P_0k, 0 otherwise
  double hPdPo, hPo;          // reference-profile buoyancy moments, 0 otherwise
  double grav[3];             // gravity_magnitude * gravity direction
  double dtt;                 // dt * theta
  double ml_on;               // 1: dt theta * lumped absorption goes into masslump (pressure_corrected_absorption)
};

// staged chunks (16 bytes each, stride NL): 0,1 = {X | z, -}   2,3 = {s_x, s_y | s_z, hb}   4,5 = {q_x, q_y | q_z, -}
//                                           6 = oldu {x, y}, then the plain double array oldu z
template <int DIM, int NL>
__device__ __forceinline__ void stage_extra(const StagedView& P, int b, int t, unsigned nsa, const double4* __restrict__ rX,
                                            const double4* __restrict__ rS, const double4* __restrict__ rQ,
                                            const double4* __restrict__ rO) {
  const int* ids = P.blk_nodes + (size_t)b * NL;
  for (int i = t; i < NL; i += kBR) {
    const int node = __ldg(ids + i);
    if (node < 0) continue;
    const unsigned d = nsa + (unsigned)i * 16u;
    const double2* sx = reinterpret_cast<const double2*>(rX + node);
    const double2* ss = reinterpret_cast<const double2*>(rS + node);
    const double2* sq = reinterpret_cast<const double2*>(rQ + node);
    const double2* so = reinterpret_cast<const double2*>(rO + node);
    cp_async16(d + 0 * NL * 16, sx);
    cp_async16(d + 1 * NL * 16, sx + 1);
    cp_async16(d + 2 * NL * 16, ss);
    cp_async16(d + 3 * NL * 16, ss + 1);
    cp_async16(d + 4 * NL * 16, sq);
    cp_async16(d + 5 * NL * 16, sq + 1);
    cp_async16(d + 6 * NL * 16, so);
    if constexpr (DIM == 3) cp_async8(nsa + 7 * NL * 16 + (unsigned)i * 8u, so + 1);
  }
}

template <int DIM, int NL>
__device__ __forceinline__ void load_oldu_extra(unsigned nsa, unsigned noff, double (&o)[DIM]) {
  const double2 a = lds128(nsa + noff + (unsigned)(6 * NL * 16));
  o[0] = a.x;
  o[1] = a.y;
  if constexpr (DIM == 3) o[2] = lds64(nsa + (noff >> 1) + (unsigned)(7 * NL * 16));
}

template <int DIM>
struct ExtraState {
  double X[DIM][DIM];                         // edges of the FIFO nodes (after install)
  double C[DIM];                              // sum of |J| over the computed windows the node was part of
  int meta[DIM];
  double X0[DIM];
  double csum;                                // sum of |J| over the row's elements
  double fs[DIM], fq[DIM], fh;                // sum_k C_k f_k for absorption, source, hb_density
};

// the node behind entry m leaves the FIFO: its fields times the |J| it has seen
template <int DIM, int NL>
__device__ __forceinline__ void extra_evict(ExtraState<DIM>& s, unsigned nsa, unsigned m, double c) {
  const unsigned nb = nsa + (m & 0xfff0u);
  const double2 sa_ = lds128(nb + (unsigned)(2 * NL * 16)), sb_ = lds128(nb + (unsigned)(3 * NL * 16));
  const double2 qa_ = lds128(nb + (unsigned)(4 * NL * 16)), qb_ = lds128(nb + (unsigned)(5 * NL * 16));
  s.fs[0] = fma(c, sa_.x, s.fs[0]);
  s.fs[1] = fma(c, sa_.y, s.fs[1]);
  s.fq[0] = fma(c, qa_.x, s.fq[0]);
  s.fq[1] = fma(c, qa_.y, s.fq[1]);
  if constexpr (DIM == 3) {
    s.fs[2] = fma(c, sb_.x, s.fs[2]);
    s.fq[2] = fma(c, qb_.x, s.fq[2]);
  }
  s.fh = fma(c, sb_.y, s.fh);
}

#define WQ(k) ((QC + DIM - (DIM - 1) + (k)) % DIM)
template <int DIM, int QC, int NL>
__device__ __forceinline__ void extra_step(ExtraState<DIM>& s, const unsigned* __restrict__ p, unsigned& pq0, unsigned& pq1,
                                           unsigned& pq2, unsigned nsa) {
  const unsigned en = pq0;
  pq0 = pq1;
  pq1 = pq2;
  extra_evict<DIM, NL>(s, nsa, (unsigned)s.meta[QC], s.C[QC]);
  s.C[QC] = 0.0;
  const unsigned nb = nsa + (en & 0xfff0u);
  {
    const double2 a = lds128(nb), b = lds128(nb + (unsigned)(NL * 16));
    s.X[QC][0] = a.x - s.X0[0];
    s.X[QC][1] = a.y - s.X0[1];
    if constexpr (DIM == 3) s.X[QC][2] = b.x - s.X0[2];
  }
  s.meta[QC] = (int)en;
  pq2 = ldg_stream1(p + (QC + 3) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
  if (en & kStagedCompute) {
    // |J| = |e_0 . (e_1 x e_2)| (2-D: |e_0 x e_1|)
    double det;
    if constexpr (DIM == 3) {
      const double(&u)[3] = s.X[WQ(1)];
      const double(&v)[3] = s.X[WQ(2)];
      const double c0 = u[1] * v[2] - u[2] * v[1], c1 = u[2] * v[0] - u[0] * v[2], c2 = u[0] * v[1] - u[1] * v[0];
      det = fma(s.X[WQ(0)][0], c0, fma(s.X[WQ(0)][1], c1, s.X[WQ(0)][2] * c2));
    } else {
      det = s.X[WQ(0)][0] * s.X[WQ(1)][1] - s.X[WQ(0)][1] * s.X[WQ(1)][0];
    }
    const double ad = fabs(det);
    s.csum += ad;
#pragma unroll
    for (int k = 0; k < DIM; k++) s.C[WQ(k)] += ad;
  }
}
#undef WQ

template <int DIM, int Q, int NL>
struct ExtraUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(ExtraState<DIM>& s, Args&&... args) {
    extra_step<DIM, Q, NL>(s, args...);
    if constexpr (Q + 1 < DIM) ExtraUnroll<DIM, Q + 1, NL>::run(s, args...);
  }
};

// Per-row quantities only (lumped absorption, sources, reference profile): no per-column accumulator, the results are
// added to rhs / the diagonal of big_m / masslump of the common kernel.
template <int DIM, int NL>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 4 : 2))
staged_momentum_extra_kernel(const ExtraConsts k_, const StagedView P, const double4* __restrict__ rX,
                             const double4* __restrict__ rS, const double4* __restrict__ rQ,
                             const double4* __restrict__ rO, size_t nnz, double* __restrict__ big_m,
                             double* __restrict__ rhs, double* __restrict__ masslump) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  stage_extra<DIM, NL>(P, b, t, nsa, rX, rS, rQ, rO);
  const int r = P.rows[b * kBR + t];
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);
  const unsigned* p = P.ent + base + t;
  const unsigned pad = P.own_local[b * kBR + t];
  const unsigned own_off = pad & 0xfff0u;
  const int own = (int)(pad >> 16) / kAS;
  unsigned pq0 = ldg_stream1(p);
  unsigned pq1 = ldg_stream1(p + kBR);
  unsigned pq2 = ldg_stream1(p + 2 * kBR);
#pragma unroll
  for (int q = 3; q < kPlanAhead; q++) prefetch_l2(p + q * kBR);
  cp_async_commit_wait_all();
  __syncthreads();
  ExtraState<DIM> s;
  double s0[DIM], q0[DIM], hb0;
  {
    const unsigned nb = nsa + own_off;
    const double2 xa = lds128(nb), xb = lds128(nb + (unsigned)(NL * 16));
    const double2 sa_ = lds128(nb + (unsigned)(2 * NL * 16)), sb_ = lds128(nb + (unsigned)(3 * NL * 16));
    const double2 qa_ = lds128(nb + (unsigned)(4 * NL * 16)), qb_ = lds128(nb + (unsigned)(5 * NL * 16));
    s.X0[0] = xa.x;
    s.X0[1] = xa.y;
    s0[0] = sa_.x;
    s0[1] = sa_.y;
    q0[0] = qa_.x;
    q0[1] = qa_.y;
    if constexpr (DIM == 3) {
      s.X0[2] = xb.x;
      s0[2] = sb_.x;
      q0[2] = qb_.x;
    }
    hb0 = sb_.y;
  }
  s.csum = s.fh = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; d++) s.fs[d] = s.fq[d] = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = 0.0;
    s.C[q] = 0.0;
    s.meta[q] = (int)pad;
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR) ExtraUnroll<DIM, 0, NL>::run(s, p, pq0, pq1, pq2, nsa);
#pragma unroll
  for (int q = 0; q < DIM; q++) extra_evict<DIM, NL>(s, nsa, (unsigned)s.meta[q], s.C[q]);
  // row epilogue. sum_e |J_e| [(Pd - Po) f_0 + Po (f_0 + sum_k f_k)] = Pd f_0 csum + Po sum_k C_k f_k
  if (r >= 0) {
    const size_t diag = (size_t)P.findrm[r] + own;
    double ou[DIM];
    load_oldu_extra<DIM, NL>(nsa, own_off, ou);
    const double nbh = fma(k_.hPdPo + k_.hPo, hb0 * s.csum, k_.hPo * s.fh);
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double lump = k_.rho * fma(k_.lPdPo + k_.lPo, s0[d] * s.csum, k_.lPo * s.fs[d]);
      const double src = k_.rho * (fma(k_.sPdPo + k_.sPo, q0[d] * s.csum, k_.sPo * s.fq[d]) + k_.sW1 * s.csum * q0[d]);
      rhs[(size_t)DIM * r + d] += src - lump * ou[d] - k_.grav[d] * nbh;
      if (k_.lPo != 0.0) big_m[(size_t)d * nnz + diag] += k_.dtt * lump;
      if (masslump && k_.ml_on != 0.0) masslump[(size_t)DIM * r + d] += k_.dtt * lump;
    }
  }
}

// ---- full absorption matrix, one diagonal block per blockIdx.y --------------------------------------------
// The dim diagonal blocks differ only through s_d: a block of threads assembles ONE component d = blockIdx.y of its
// 128 rows with a single accumulator column per thread and scalar FIFO state (measured on the B200: the first version
// kept all dim blocks per thread -- 156 registers, 92 KB of shared memory, 2 blocks per SM -- and took 4x the common
// kernel's time). Staged per node: {X | z, -} as two chunks, then plain double arrays s_d and oldu_d.
template <int DIM>
struct AbsState {
  double X[DIM][DIM], S[DIM], A[DIM];
  int meta[DIM];
  double X0[DIM], s0, a0;
};

#define WQ(k) ((QC + DIM - (DIM - 1) + (k)) % DIM)
template <int DIM, int QC, int NL>
__device__ __forceinline__ void abs_step(AbsState<DIM>& s, double& rh, const ExtraConsts& k_, const unsigned* __restrict__ p,
                                         unsigned& pq0, unsigned& pq1, unsigned& pq2, unsigned acc_sa, unsigned nsa) {
  const unsigned en = pq0;
  pq0 = pq1;
  pq1 = pq2;
  {
    const unsigned m = (unsigned)s.meta[QC];
    const double on = lds64(nsa + (unsigned)(2 * NL * 16 + NL * 8) + ((m & 0xfff0u) >> 1));
    const unsigned sa = acc_sa + ((m >> 16) << 3);
    const double a = s.A[QC];
    sts64(sa, lds64(sa) + a);
    rh = fma(-a, on, rh);
    s.A[QC] = 0.0;
  }
  const unsigned noff = en & 0xfff0u;
  {
    const double2 a = lds128(nsa + noff), b = lds128(nsa + noff + (unsigned)(NL * 16));
    s.S[QC] = lds64(nsa + (unsigned)(2 * NL * 16) + (noff >> 1));
    s.X[QC][0] = a.x - s.X0[0];
    s.X[QC][1] = a.y - s.X0[1];
    if constexpr (DIM == 3) s.X[QC][2] = b.x - s.X0[2];
  }
  s.meta[QC] = (int)en;
  pq2 = ldg_stream1(p + (QC + 3) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
  if (en & kStagedCompute) {
    double det;
    if constexpr (DIM == 3) {
      const double(&u)[3] = s.X[WQ(1)];
      const double(&v)[3] = s.X[WQ(2)];
      const double c0 = u[1] * v[2] - u[2] * v[1], c1 = u[2] * v[0] - u[0] * v[2], c2 = u[0] * v[1] - u[1] * v[0];
      det = fma(s.X[WQ(0)][0], c0, fma(s.X[WQ(0)][1], c1, s.X[WQ(0)][2] * c2));
    } else {
      det = s.X[WQ(0)][0] * s.X[WQ(1)][1] - s.X[WQ(0)][1] * s.X[WQ(1)][0];
    }
    const double adr = fabs(det) * k_.rho;
    double Sd = s.s0;
#pragma unroll
    for (int k = 0; k < DIM; k++) Sd += s.S[WQ(k)];
    const double QS = k_.Qabc * Sd;
    s.a0 = fma(adr, fma(k_.Qa, s.s0, k_.Qaab * Sd), s.a0);
#pragma unroll
    for (int k = 0; k < DIM; k++) s.A[WQ(k)] = fma(adr, fma(k_.Qd, s.s0 + s.S[WQ(k)], QS), s.A[WQ(k)]);
  }
}
#undef WQ

template <int DIM, int Q, int NL>
struct AbsUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(AbsState<DIM>& s, Args&&... args) {
    abs_step<DIM, Q, NL>(s, args...);
    if constexpr (Q + 1 < DIM) AbsUnroll<DIM, Q + 1, NL>::run(s, args...);
  }
};

template <int DIM, int NL>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 5 : 2))
staged_momentum_abs_kernel(const ExtraConsts k_, const StagedView P, const double4* __restrict__ rX,
                           const double4* __restrict__ rS, const double4* __restrict__ rO, size_t nnz,
                           double* __restrict__ big_m, double* __restrict__ rhs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x, d = blockIdx.y;
  {
    const int* ids = P.blk_nodes + (size_t)b * NL;
    for (int i = t; i < NL; i += kBR) {
      const int node = __ldg(ids + i);
      if (node < 0) continue;
      const double2* sx = reinterpret_cast<const double2*>(rX + node);
      cp_async16(nsa + (unsigned)i * 16u, sx);
      cp_async16(nsa + (unsigned)(NL * 16) + (unsigned)i * 16u, sx + 1);
      cp_async8(nsa + (unsigned)(2 * NL * 16) + (unsigned)i * 8u, reinterpret_cast<const double*>(rS + node) + d);
      cp_async8(nsa + (unsigned)(2 * NL * 16 + NL * 8) + (unsigned)i * 8u, reinterpret_cast<const double*>(rO + node) + d);
    }
  }
  const int r = P.rows[b * kBR + t];
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);
  const unsigned* p = P.ent + base + t;
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  const unsigned pad = P.own_local[b * kBR + t];
  const unsigned own_off = pad & 0xfff0u;
  const int own = (int)(pad >> 16) / kAS;
  unsigned pq0 = ldg_stream1(p);
  unsigned pq1 = ldg_stream1(p + kBR);
  unsigned pq2 = ldg_stream1(p + 2 * kBR);
#pragma unroll
  for (int q = 3; q < kPlanAhead; q++) prefetch_l2(p + q * kBR);
  cp_async_commit_wait_all();
  __syncthreads();
  AbsState<DIM> s;
  {
    const double2 a = lds128(nsa + own_off), bb = lds128(nsa + own_off + (unsigned)(NL * 16));
    s.X0[0] = a.x;
    s.X0[1] = a.y;
    if constexpr (DIM == 3) s.X0[2] = bb.x;
    s.s0 = lds64(nsa + (unsigned)(2 * NL * 16) + (own_off >> 1));
  }
  s.a0 = 0.0;
  double rh = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = 0.0;
    s.S[q] = s.A[q] = 0.0;
    s.meta[q] = (int)pad;
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR) AbsUnroll<DIM, 0, NL>::run(s, rh, k_, p, pq0, pq1, pq2, acc_sa, nsa);
#pragma unroll
  for (int q = 0; q < DIM; q++) {
    const unsigned m = (unsigned)s.meta[q];
    acc_t[m >> 16] += s.A[q];
    rh = fma(-s.A[q], lds64(nsa + (unsigned)(2 * NL * 16 + NL * 8) + ((m & 0xfff0u) >> 1)), rh);
  }
  acc_t[own * kAS] += s.a0;
  int my_s0 = 0, my_len = 0;
  if (r >= 0) {
    my_s0 = P.findrm[r];
    my_len = P.findrm[r + 1] - my_s0;
    const double ou = lds64(nsa + (unsigned)(2 * NL * 16 + NL * 8) + (own_off >> 1));
    rhs[(size_t)DIM * r + d] += fma(-s.a0, ou, rh);
  }
  __syncwarp();
  // big_m(d,d) += dt theta Ab^d
  const int lane = t & 31, wbase = t & ~31;
  const int lpr = 1 << P.lpr_shift, rpi = 32 >> P.lpr_shift;
  const int sub = lane >> P.lpr_shift, sl = lane & (lpr - 1);
  double* out = big_m + (size_t)d * nnz;
  for (int rr = 0; rr < 32; rr += rpi) {
    const int src = rr + sub;
    const int s0r = __shfl_sync(0xffffffffu, my_s0, src);
    const int lr = __shfl_sync(0xffffffffu, my_len, src);
    for (int ss = sl; ss < lr; ss += lpr) out[s0r + ss] = fma(k_.dtt, acc[ss * kAS + wbase + src], out[s0r + ss]);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
// what the extra pass is needed for / can do
bool strip_extra_needed(const MomentumArgs& A) {
  const cgasm_momentum_opts& o = A.o;
  return o.have_absorption || o.have_source || (o.have_gravity && o.subtract_out_reference_profile);
}

bool strip_extra_ok(const Handle* h, const MomentumArgs& A) {
  const cgasm_momentum_opts& o = A.o;
  const GatherPlan* P = h->gather;
  if (!P || !P->staged_ok || !P->d_strip_local || getenv("CGASM_STRIP_GLOBAL") || getenv("CGASM_STRIP_NO_EXTRA")) return false;
  // constant density only: with a nodal density the absorption / source rows are 4-index moments
  if (h->fields[CGASM_F_DENSITY].field_type != CGASM_FIELD_CONSTANT) return false;
  if (o.have_absorption && h->fields[CGASM_F_ABSORPTION].field_type != CGASM_FIELD_NORMAL) return false;
  if (o.have_source && h->fields[CGASM_F_SOURCE].field_type != CGASM_FIELD_NORMAL) return false;
  if (o.have_gravity && o.subtract_out_reference_profile && h->fields[CGASM_F_HB_DENSITY].field_type != CGASM_FIELD_NORMAL)
    return false;
  const size_t smem = std::max(staged_acc_bytes(P, 0) + (size_t)P->nl * 120, staged_acc_bytes(P, 1) + (size_t)P->nl * 48);
  return smem <= 110 * 1024;
}

template <int DIM>
static int strip_extra_dim(Handle* h, const MomentumArgs& A, bool skip_full) {
  GatherPlan* P = h->gather;
  const cgasm_momentum_opts& o = A.o;
  const Tables& t = A.tab;
  const bool full = o.have_absorption && !o.lump_absorption && !skip_full;
  const bool light = (o.have_absorption && o.lump_absorption) || o.have_source ||
                     (o.have_gravity && o.subtract_out_reference_profile);
  ExtraConsts c{};
  c.rho = h->fields[CGASM_F_DENSITY].h_const[0];
  c.Qa = t.Qaaa - t.Qaab;
  c.Qaab = t.Qaab;
  c.Qd = t.Qaab - t.Qabc;
  c.Qabc = t.Qabc;
  if (o.have_absorption && o.lump_absorption) {
    c.lPdPo = t.Pd - t.Po;
    c.lPo = t.Po;
  }
  if (o.have_source && !o.lump_source) {
    c.sPdPo = t.Pd - t.Po;
    c.sPo = t.Po;
  }
  if (o.have_source && o.lump_source) c.sW1 = t.W1;
  if (o.have_gravity && o.subtract_out_reference_profile) {
    c.hPdPo = t.Pd - t.Po;
    c.hPo = t.Po;
    for (int d = 0; d < DIM; d++) c.grav[d] = o.gravity_magnitude * h->fields[CGASM_F_GRAVITY].h_const[d];
  }
  c.dtt = o.dt * o.theta;
  c.ml_on = (o.have_absorption && o.lump_absorption && o.pressure_corrected_absorption) ? 1.0 : 0.0;
  int st = ensure_extra_records(h);
  if (st) return st;
  if (h->d_perm && (!h->d_prec[5] || !h->d_prec[6]) && (st = refresh_permuted(h, 1u << 5 | 1u << 6, nullptr, 0, h->stream))) return st;
  double* ml = o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  if (light) {
    const StagedView v = staged_view(h, 0);
    const size_t smem = (size_t)v.acc_bytes + (size_t)P->nl * 120;
#define LAUNCH_NL(NL_)                                                                                          \
  do {                                                                                                          \
    if ((st = strip_smem(staged_momentum_extra_kernel<DIM, NL_>, smem))) return st;                      \
    staged_momentum_extra_kernel<DIM, NL_><<<P->nblocks, kBR, smem, h->stream>>>(                        \
        c, v, (const double4*)staged_rec(h, 3), (const double4*)staged_rec(h, 5), (const double4*)staged_rec(h, 6),                 \
        (const double4*)staged_rec(h, 2), (size_t)h->nnz, h->d_big_m, h->d_mom_rhs, ml);         \
    h->launches++;                                                                                              \
  } while (0)
    CGASM_FOR_NL(LAUNCH_NL);
#undef LAUNCH_NL
  }
  if (full) {
    const StagedView v = staged_view(h, 1);
    const size_t smem = (size_t)v.acc_bytes + (size_t)P->nl * 48;
#define LAUNCH_NL(NL_)                                                                                          \
  do {                                                                                                          \
    if ((st = strip_smem(staged_momentum_abs_kernel<DIM, NL_>, smem))) return st;                               \
    staged_momentum_abs_kernel<DIM, NL_><<<dim3(P->nblocks, DIM), kBR, smem, h->stream>>>(                      \
        c, v, (const double4*)staged_rec(h, 3), (const double4*)staged_rec(h, 5), (const double4*)staged_rec(h, 2),                 \
        (size_t)h->nnz, h->d_big_m, h->d_mom_rhs);                        \
    h->launches++;                                                                                              \
  } while (0)
    CGASM_FOR_NL(LAUNCH_NL);
#undef LAUNCH_NL
  }
  CG_CUDA(cudaGetLastError());
  return st;
}

int strip_extra(Handle* h, const MomentumArgs& A, bool skip_full_absorption) {
  return h->dim == 3 ? strip_extra_dim<3>(h, A, skip_full_absorption) : strip_extra_dim<2>(h, A, skip_full_absorption);
}

}  // namespace cgasm
