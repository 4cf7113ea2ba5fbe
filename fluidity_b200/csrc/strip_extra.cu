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

template <int DIM, bool FULLABS>
struct ExtraState {
  double X[DIM][DIM];                         // edges of the FIFO nodes (after install)
  double C[DIM];                              // sum of |J| over the computed windows the node was part of
  double S[FULLABS ? DIM : 1][DIM];           // absorption of the FIFO nodes (full matrix only)
  double A[FULLABS ? DIM : 1][DIM];           // accumulated entries of the dim diagonal blocks (full matrix only)
  int meta[DIM];
  double X0[DIM], s0[DIM];
  double csum;                                // sum of |J| over the row's elements
  double a0[DIM];                             // diagonal of the full absorption rows
  double fs[DIM], fq[DIM], fh;                // sum_k C_k f_k for absorption, source, hb_density
};

#define WQ(k) ((QC + DIM - (DIM - 1) + (k)) % DIM)
template <int DIM, int QC, bool FULLABS>
__device__ __forceinline__ void extra_compute(ExtraState<DIM, FULLABS>& s, const ExtraConsts& k_) {
  // |J| = |e_0 . (e_1 x e_2)| (2-D: |e_0 x e_1|)
  double det;
  if constexpr (DIM == 3) {
    const double(&p)[3] = s.X[WQ(1)];
    const double(&q)[3] = s.X[WQ(2)];
    const double c0 = p[1] * q[2] - p[2] * q[1], c1 = p[2] * q[0] - p[0] * q[2], c2 = p[0] * q[1] - p[1] * q[0];
    det = fma(s.X[WQ(0)][0], c0, fma(s.X[WQ(0)][1], c1, s.X[WQ(0)][2] * c2));
  } else {
    det = s.X[WQ(0)][0] * s.X[WQ(1)][1] - s.X[WQ(0)][1] * s.X[WQ(1)][0];
  }
  const double ad = fabs(det);
  s.csum += ad;
#pragma unroll
  for (int k = 0; k < DIM; k++) s.C[WQ(k)] += ad;
  if constexpr (FULLABS) {
    const double adr = ad * k_.rho;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      double Sd = s.s0[d];
#pragma unroll
      for (int k = 0; k < DIM; k++) Sd += s.S[WQ(k)][d];
      const double QS = k_.Qabc * Sd;
      s.a0[d] = fma(adr, fma(k_.Qa, s.s0[d], k_.Qaab * Sd), s.a0[d]);
#pragma unroll
      for (int k = 0; k < DIM; k++) s.A[WQ(k)][d] = fma(adr, fma(k_.Qd, s.s0[d] + s.S[WQ(k)][d], QS), s.A[WQ(k)][d]);
    }
  }
}
#undef WQ

template <int DIM, int QC, int NL, bool FULLABS>
__device__ __forceinline__ void extra_step(ExtraState<DIM, FULLABS>& s, double (&rh)[DIM], const ExtraConsts& k_,
                                           const unsigned* __restrict__ p, unsigned& pq0, unsigned& pq1, unsigned& pq2,
                                           unsigned acc_sa, unsigned nsa, int acc_block_bytes) {
  const unsigned en = pq0;
  pq0 = pq1;
  pq1 = pq2;
  {
    // the node in buffer QC leaves the FIFO: its fields times the |J| it has seen; full matrix: its entries
    const unsigned m = (unsigned)s.meta[QC];
    const unsigned nb = nsa + (m & 0xfff0u);
    const double2 sa_ = lds128(nb + (unsigned)(2 * NL * 16)), sb_ = lds128(nb + (unsigned)(3 * NL * 16));
    const double2 qa_ = lds128(nb + (unsigned)(4 * NL * 16)), qb_ = lds128(nb + (unsigned)(5 * NL * 16));
    const double c = s.C[QC];
    s.fs[0] = fma(c, sa_.x, s.fs[0]);
    s.fs[1] = fma(c, sa_.y, s.fs[1]);
    s.fq[0] = fma(c, qa_.x, s.fq[0]);
    s.fq[1] = fma(c, qa_.y, s.fq[1]);
    if constexpr (DIM == 3) {
      s.fs[2] = fma(c, sb_.x, s.fs[2]);
      s.fq[2] = fma(c, qb_.x, s.fq[2]);
    }
    s.fh = fma(c, sb_.y, s.fh);
    s.C[QC] = 0.0;
    if constexpr (FULLABS) {
      double on[DIM];
      load_oldu_extra<DIM, NL>(nsa, m & 0xfff0u, on);
      const unsigned sa = acc_sa + ((m >> 16) << 3);
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const double a = s.A[QC][d];
        sts64(sa + (unsigned)(d * acc_block_bytes), lds64(sa + (unsigned)(d * acc_block_bytes)) + a);
        rh[d] = fma(-a, on[d], rh[d]);
        s.A[QC][d] = 0.0;
      }
    }
  }
  const unsigned nb = nsa + (en & 0xfff0u);
  {
    const double2 a = lds128(nb), b = lds128(nb + (unsigned)(NL * 16));
    s.X[QC][0] = a.x - s.X0[0];
    s.X[QC][1] = a.y - s.X0[1];
    if constexpr (DIM == 3) s.X[QC][2] = b.x - s.X0[2];
  }
  if constexpr (FULLABS) {
    const double2 a = lds128(nb + (unsigned)(2 * NL * 16)), b = lds128(nb + (unsigned)(3 * NL * 16));
    s.S[QC][0] = a.x;
    s.S[QC][1] = a.y;
    if constexpr (DIM == 3) s.S[QC][2] = b.x;
  }
  s.meta[QC] = (int)en;
  pq2 = ldg_stream1(p + (QC + 3) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
  if (en & kStagedCompute) extra_compute<DIM, QC, FULLABS>(s, k_);
}

template <int DIM, int Q, int NL, bool FULLABS>
struct ExtraUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(ExtraState<DIM, FULLABS>& s, Args&&... args) {
    extra_step<DIM, Q, NL, FULLABS>(s, args...);
    if constexpr (Q + 1 < DIM) ExtraUnroll<DIM, Q + 1, NL, FULLABS>::run(s, args...);
  }
};

template <int DIM, int NL, bool FULLABS>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? (FULLABS ? 2 : 4) : 2))
staged_momentum_extra_kernel(const ExtraConsts k_, const StagedView P, const double4* __restrict__ rX,
                             const double4* __restrict__ rS, const double4* __restrict__ rQ,
                             const double4* __restrict__ rO, size_t nnz, double* __restrict__ big_m,
                             double* __restrict__ rhs, double* __restrict__ masslump) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  stage_extra<DIM, NL>(P, b, t, nsa, rX, rS, rQ, rO);
  const int r = P.rows[b * kBR + t];
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);
  const unsigned* p = P.ent + base + t;
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  const int acc_block = P.maxlen * kAS;  // doubles per diagonal block of the accumulator
  if constexpr (FULLABS)
    for (int q = 0; q < DIM * P.maxlen; q++) acc_t[q * kAS] = 0.0;
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
  ExtraState<DIM, FULLABS> s;
  double q0[DIM], hb0;
  {
    const unsigned nb = nsa + own_off;
    const double2 xa = lds128(nb), xb = lds128(nb + (unsigned)(NL * 16));
    const double2 sa_ = lds128(nb + (unsigned)(2 * NL * 16)), sb_ = lds128(nb + (unsigned)(3 * NL * 16));
    const double2 qa_ = lds128(nb + (unsigned)(4 * NL * 16)), qb_ = lds128(nb + (unsigned)(5 * NL * 16));
    s.X0[0] = xa.x;
    s.X0[1] = xa.y;
    s.s0[0] = sa_.x;
    s.s0[1] = sa_.y;
    q0[0] = qa_.x;
    q0[1] = qa_.y;
    if constexpr (DIM == 3) {
      s.X0[2] = xb.x;
      s.s0[2] = sb_.x;
      q0[2] = qb_.x;
    }
    hb0 = sb_.y;
  }
  s.csum = s.fh = 0.0;
  double rh[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rh[d] = s.a0[d] = s.fs[d] = s.fq[d] = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = 0.0;
    s.C[q] = 0.0;
    s.meta[q] = (int)pad;
    if constexpr (FULLABS) {
#pragma unroll
      for (int a = 0; a < DIM; a++) s.S[q][a] = s.A[q][a] = 0.0;
    }
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR)
    ExtraUnroll<DIM, 0, NL, FULLABS>::run(s, rh, k_, p, pq0, pq1, pq2, acc_sa, nsa, acc_block * 8);
  // drain the FIFO
#pragma unroll
  for (int q = 0; q < DIM; q++) {
    const unsigned m = (unsigned)s.meta[q];
    const unsigned nb = nsa + (m & 0xfff0u);
    const double2 sa_ = lds128(nb + (unsigned)(2 * NL * 16)), sb_ = lds128(nb + (unsigned)(3 * NL * 16));
    const double2 qa_ = lds128(nb + (unsigned)(4 * NL * 16)), qb_ = lds128(nb + (unsigned)(5 * NL * 16));
    const double c = s.C[q];
    s.fs[0] = fma(c, sa_.x, s.fs[0]);
    s.fs[1] = fma(c, sa_.y, s.fs[1]);
    s.fq[0] = fma(c, qa_.x, s.fq[0]);
    s.fq[1] = fma(c, qa_.y, s.fq[1]);
    if constexpr (DIM == 3) {
      s.fs[2] = fma(c, sb_.x, s.fs[2]);
      s.fq[2] = fma(c, qb_.x, s.fq[2]);
    }
    s.fh = fma(c, sb_.y, s.fh);
    if constexpr (FULLABS) {
      double o[DIM];
      load_oldu_extra<DIM, NL>(nsa, m & 0xfff0u, o);
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        acc_t[d * acc_block + (m >> 16)] += s.A[q][d];
        rh[d] = fma(-s.A[q][d], o[d], rh[d]);
      }
    }
  }
  // row epilogue. sum_e |J_e| [(Pd - Po) f_0 + Po (f_0 + sum_k f_k)] = Pd f_0 csum + Po sum_k C_k f_k
  int my_s0 = 0, my_len = 0;
  double lump[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) lump[d] = 0.0;
  if (r >= 0) {
    my_s0 = P.findrm[r];
    my_len = P.findrm[r + 1] - my_s0;
    double ou[DIM];
    load_oldu_extra<DIM, NL>(nsa, own_off, ou);
    const double nbh = fma(k_.hPdPo + k_.hPo, hb0 * s.csum, k_.hPo * s.fh);
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      lump[d] = k_.rho * fma(k_.lPdPo + k_.lPo, s.s0[d] * s.csum, k_.lPo * s.fs[d]);
      const double src = k_.rho * (fma(k_.sPdPo + k_.sPo, q0[d] * s.csum, k_.sPo * s.fq[d]) + k_.sW1 * s.csum * q0[d]);
      double v = rh[d] + src - lump[d] * ou[d] - k_.grav[d] * nbh;
      if constexpr (FULLABS) {
        acc_t[d * acc_block + own * kAS] += s.a0[d];
        v = fma(-s.a0[d], ou[d], v);
      }
      rhs[(size_t)DIM * r + d] += v;
      if (masslump && k_.ml_on != 0.0) masslump[(size_t)DIM * r + d] += k_.dtt * lump[d];
    }
  }
  __syncwarp();
  // big_m(d,d) += dt theta (Ab^d row + lumped absorption on the diagonal)
  const int lane = t & 31, wbase = t & ~31;
  const int lpr = 1 << P.lpr_shift, rpi = 32 >> P.lpr_shift;
  const int sub = lane >> P.lpr_shift, sl = lane & (lpr - 1);
  for (int rr = 0; rr < 32; rr += rpi) {
    const int src = rr + sub;
    const int s0r = __shfl_sync(0xffffffffu, my_s0, src);
    const int lr = __shfl_sync(0xffffffffu, my_len, src);
    const int ownr = __shfl_sync(0xffffffffu, own, src);
    double lm[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) lm[d] = __shfl_sync(0xffffffffu, lump[d], src);
    if constexpr (FULLABS) {
      for (int ss = sl; ss < lr; ss += lpr)
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          double* o = big_m + (size_t)d * nnz + s0r + ss;
          *o += k_.dtt * (acc[d * acc_block + ss * kAS + wbase + src] + (ss == ownr ? lm[d] : 0.0));
        }
    } else {
      if (sl == 0 && lr > 0)
#pragma unroll
        for (int d = 0; d < DIM; d++) big_m[(size_t)d * nnz + s0r + ownr] += k_.dtt * lm[d];
    }
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
  const bool full = o.have_absorption && !o.lump_absorption;
  const size_t smem = staged_acc_bytes(P, full ? h->dim : 0) + (size_t)P->nl * 120;
  return smem <= 110 * 1024;
}

template <int DIM>
static int strip_extra_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  const cgasm_momentum_opts& o = A.o;
  const Tables& t = A.tab;
  const bool full = o.have_absorption && !o.lump_absorption;
  ExtraConsts c{};
  c.rho = h->fields[CGASM_F_DENSITY].h_const[0];
  if (full) {
    c.Qa = t.Qaaa - t.Qaab;
    c.Qaab = t.Qaab;
    c.Qd = t.Qaab - t.Qabc;
    c.Qabc = t.Qabc;
  }
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
  int est = ensure_extra_records(h);
  if (est) return est;
  StagedView v = staged_view(h, full ? DIM : 0);
  const size_t smem = (size_t)v.acc_bytes + (size_t)P->nl * 120;
  double* ml = o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  int st = CGASM_OK;
#define LAUNCH(NL_, FULL_)                                                                                      \
  do {                                                                                                          \
    if ((st = strip_smem(staged_momentum_extra_kernel<DIM, NL_, FULL_>, smem))) return st;                      \
    staged_momentum_extra_kernel<DIM, NL_, FULL_><<<P->nblocks, kBR, smem, h->stream>>>(                        \
        c, v, h->d_rec3, h->d_rec5, h->d_rec6, h->d_rec2, (size_t)h->nnz, h->d_big_m, h->d_mom_rhs, ml);         \
    h->launches++;                                                                                              \
  } while (0)
#define LAUNCH_NL(NL_)                   \
  do {                                   \
    if (full) LAUNCH(NL_, true);         \
    else LAUNCH(NL_, false);             \
  } while (0)
  CGASM_FOR_NL(LAUNCH_NL);
#undef LAUNCH_NL
#undef LAUNCH
  CG_CUDA(cudaGetLastError());
  return st;
}

int strip_extra(Handle* h, const MomentumArgs& A) {
  return h->dim == 3 ? strip_extra_dim<3>(h, A) : strip_extra_dim<2>(h, A);
}

}  // namespace cgasm
