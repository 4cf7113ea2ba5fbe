// strip_absorb.cu -- staged STRIP momentum kernel with the FULL absorption matrix in the same pass.
//
// add_absorption_element_cg (assemble/Momentum_CG.F90:2036-2073) with a constant density and a P1 vector absorption
// field s_d makes the dim diagonal blocks of big_m differ:
//   Ab^d_0k = rho |J| sum_l Q_0kl s_dl = rho |J| [Qa s_d0 + Qaab S_d  (k = 0) | Qd (s_d0 + s_dk) + Qabc S_d],
//   S_d = s_d0 + sum_{l in e} s_dl;   big_m(d,d) += dt theta Ab^d,   rhs_d -= Ab^d oldu_d   (:2060-2066).
// Round 2 first ran this as one more strip pass PER COMPONENT behind the common kernel (strip_extra.cu: three plan
// reads, a read-modify-write of big_m, |J| recomputed three times: 1.22 ms on top of the common kernel's 0.51 ms at
// 12.6 M tets). Here the common kernel's own loop carries it. What an entry (0, k) collects over the elements e of the
// row that contain k is split so that the element loop adds only products of |J_e|:
//   sum_e Ab^d_0k = rho Qd (s_d0 + s_dk) C_k + rho Qabc G^d_k,   C_k = sum_e |J_e|,   G^d_k = sum_e |J_e| S^d_e,
// i.e. per (row, element) pair 3 DADD for C and, per component, 3 DADD + 1 DMUL + 4 DADD; the products with the
// fields happen once per strip entry when column k leaves the FIFO, where the common entry A_k is added and the dim
// values go to dim accumulator columns (shared memory: dim x the common kernel's accumulator, 2 blocks per SM).
// Lumped absorption, sources and the reference profile stay in strip_extra.cu's per-row pass.
#include "strip_staged.cuh"

#include <cstdlib>

namespace cgasm {

struct AbsorbConsts {
  double rQa, rQaab, rQd, rQabc;  // rho * absorption moments
};

// staged chunks (16 bytes each, stride NL): 0..3 = records {X | z, buoyancy}, {nu | z, rho}; 4 = oldu {x, y};
// 5 (first half) = plain double array oldu z; 6, 7 = {s_x, s_y | s_z, hb_density}
template <int DIM, int NL>
__device__ __forceinline__ void stage_nodes_absorb(const BlockIds<NL>& ids, int t, unsigned nsa, const double4* __restrict__ r0,
                                                   const double4* __restrict__ r1, const double4* __restrict__ rO,
                                                   const double4* __restrict__ rS) {
  const int h = t & 1;
#pragma unroll
  for (int v = 0; v < BlockIds<NL>::PER; v++) {
    const int node = ids.node[v];
    if (node < 0) continue;
    const unsigned i = (unsigned)((t >> 1) + v * (kBR / 2));
    stage_record<NL>(nsa, 0, i, h, r0, node);
    stage_record<NL>(nsa, 2, i, h, r1, node);
    stage_record_3<NL, DIM == 3>(nsa, 4, (unsigned)(5 * NL * 16), i, h, rO, node);
    stage_record<NL>(nsa, 6, i, h, rS, node);
  }
}

template <int DIM, int NL>
__device__ __forceinline__ void load_absorption(unsigned nb, double (&v)[DIM]) {
  const double2 a = lds128(nb + (unsigned)(6 * NL * 16));
  v[0] = a.x;
  v[1] = a.y;
  if constexpr (DIM == 3) v[2] = lds128(nb + (unsigned)(7 * NL * 16)).x;
}

template <int DIM>
struct AbsorbState {
  double Sg[DIM][DIM];  // absorption of the FIFO nodes [buffer][component]
  double G[DIM][DIM];   // sum of |J_e| S^d_e over the computed windows the node was part of
  double C[DIM];        // sum of |J_e| over the same windows
  double sg0[DIM], g0[DIM], c0;
};

#define WQ(k) ((QC + DIM - (DIM - 1) + (k)) % DIM)
template <int DIM, int QC, int NL, bool FULLV>
__device__ __forceinline__ void sabs_step(MomState<DIM, DIM>& s, AbsorbState<DIM>& ab, double (&rh)[DIM], const StripConsts& k_,
                                          const AbsorbConsts& ka, const unsigned* __restrict__ p, unsigned (&pq)[DIM],
                                          unsigned acc_sa, unsigned acc_stride, unsigned nsa) {
  const unsigned en = pq[QC];
  const unsigned m = (unsigned)s.meta[QC];
  double on[DIM];
  load_oldu<DIM, NL>(nsa, m & 0xfff0u, on);
  const unsigned sa = acc_sa + ((m >> 16) << 3);
  double slot[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) slot[d] = lds64(sa + d * acc_stride);
  const unsigned nb = nsa + (en & 0xfff0u);
  // the evicted column's entries: common part + absorption, before its buffers are overwritten
  {
    const double a = s.A[QC], c = ab.C[QC];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double v = fma(ka.rQabc, ab.G[QC][d], fma(ka.rQd * c, ab.sg0[d] + ab.Sg[QC][d], a));
      sts64(sa + d * acc_stride, slot[d] + v);
      rh[d] = fma(-v, on[d], rh[d]);
      ab.G[QC][d] = 0.0;
    }
    s.A[QC] = 0.0;
    ab.C[QC] = 0.0;
  }
  load_rec<DIM, NL>(nb, 0, s.X[QC], s.B[QC]);
  load_rec<DIM, NL>(nb, 1, s.U[QC], s.R[QC]);
  load_absorption<DIM, NL>(nb, ab.Sg[QC]);
  s.meta[QC] = (int)en;
  pq[QC] = ldg_stream1(p + (QC + DIM) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if (en & kStagedCompute) {
    WindowGeom<DIM> g;
    window_geom<DIM, DIM, QC>(s.X, g);
    mom_terms<DIM, DIM, QC, FULLV>(s, k_, g);
    const double ad = fabs(g.det);
    ab.c0 += ad;
#pragma unroll
    for (int k = 0; k < DIM; k++) ab.C[WQ(k)] += ad;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      double S = ab.sg0[d];
#pragma unroll
      for (int k = 0; k < DIM; k++) S += ab.Sg[WQ(k)][d];
      const double t = ad * S;
      ab.g0[d] += t;
#pragma unroll
      for (int k = 0; k < DIM; k++) ab.G[WQ(k)][d] += t;
    }
  }
}
#undef WQ

template <int DIM, int Q, int NL, bool FULLV>
struct SAbsUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, DIM>& s, AbsorbState<DIM>& ab, Args&&... args) {
    sabs_step<DIM, Q, NL, FULLV>(s, ab, args...);
    if constexpr (Q + 1 < DIM) SAbsUnroll<DIM, Q + 1, NL, FULLV>::run(s, ab, args...);
  }
};

template <int DIM, int NL, bool FULLV>
__global__ void __launch_bounds__(kBR, 2)
staged_momentum_absorb_kernel(const StripConsts k_, const AbsorbConsts ka, const StagedView P, const double4* __restrict__ rX,
                              const double4* __restrict__ rU, const double4* __restrict__ rO, const double4* __restrict__ rS,
                              size_t nnz, double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned acc_stride = (unsigned)(sizeof(double) * P.maxlen * kAS);  // bytes between the accumulator blocks of two components
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const unsigned tbl_sa = nsa + (unsigned)(NL * 128);
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  BlockIds<NL> ids;
  issue_block_ids<NL>(P, b, t, ids);
  const int4 meta = ldg_nc_v4(P.row_meta + (size_t)b * kBR + t);
  const long long base = ldg_nc_s64(P.ptr + b);
  prefetch_next_block<NL>(P, b, t);
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  for (int d = 0; d < DIM; d++)
    for (int q = 0; q < P.maxlen; q++) acc_t[d * (acc_stride >> 3) + q * kAS] = 0.0;
  stage_nodes_absorb<DIM, NL>(ids, t, nsa, rX, rU, rO, rS);
  const int deg = warp_trip_count<DIM>(meta.z);
  const unsigned* p = P.ent + base + t;
  unsigned pq[DIM];
#pragma unroll
  for (int q = 0; q < DIM; q++) pq[q] = ldg_stream1(p + q * kBR);
#pragma unroll
  for (int q = DIM; q < kPlanAhead; q++) prefetch_l2(p + q * kBR);
  const int r = meta.x;
  const unsigned pad = (unsigned)meta.w;
  const unsigned own_off = pad & 0xfff0u;
  const int own = (meta.z >> 16) & 0xff;
  cp_async_commit_wait_all();
  __syncthreads();
  MomState<DIM, DIM> s;
  AbsorbState<DIM> ab;
  load_rec<DIM, NL>(nsa + own_off, 0, s.X0, s.b0);
  load_rec<DIM, NL>(nsa + own_off, 1, s.U0, s.rho0);
  mom_row_consts<DIM, DIM>(s, k_);
  load_absorption<DIM, NL>(nsa + own_off, ab.sg0);
  s.a0 = s.msum = s.nbsum = 0.0;
  ab.c0 = 0.0;
  double rh[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rh[d] = ab.g0[d] = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = ab.Sg[q][a] = ab.G[q][a] = 0.0;
    s.R[q] = s.B[q] = s.A[q] = ab.C[q] = 0.0;
    s.meta[q] = (int)pad;
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR)
    SAbsUnroll<DIM, 0, NL, FULLV>::run(s, ab, rh, k_, ka, p, pq, acc_sa, acc_stride, nsa);
  // drain the FIFO, then the diagonal (the row's own node never leaves)
#pragma unroll
  for (int q = 0; q < DIM; q++) {
    const unsigned m = (unsigned)s.meta[q];
    double o[DIM];
    load_oldu<DIM, NL>(nsa, m & 0xfff0u, o);
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double v = fma(ka.rQabc, ab.G[q][d], fma(ka.rQd * ab.C[q], ab.sg0[d] + ab.Sg[q][d], s.A[q]));
      acc_t[d * (acc_stride >> 3) + (m >> 16)] += v;
      rh[d] = fma(-v, o[d], rh[d]);
    }
  }
  double ou[DIM];
  load_oldu<DIM, NL>(nsa, own_off, ou);
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const double v = fma(ka.rQaab, ab.g0[d], fma(ka.rQa * ab.c0, ab.sg0[d], s.a0));
    acc_t[d * (acc_stride >> 3) + own * kAS] += v;
    if (r >= 0) {
      rhs[(size_t)DIM * r + d] = fma(-v, ou[d], fma(k_.grav[d], s.nbsum, rh[d]));
      if (masslump) masslump[(size_t)DIM * r + d] = s.msum;
    }
  }
  // rows of the warp -> the dim diagonal blocks: dt*theta * entry (+ lumped mass on the diagonal)
  row_table_store(tbl_sa, t, meta.y, meta.z, s.msum * k_.mass_on);
  __syncwarp();
  {
    const int lane = t & 31, wbase = t & ~31;
    const int lpr = 1 << P.lpr_shift, rpi = 32 >> P.lpr_shift;
    const int sub = lane >> P.lpr_shift, sl = lane & (lpr - 1);
    for (int rr = 0; rr < 32; rr += rpi) {
      const int src = wbase + rr + sub;
      int s0r, lo;
      asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(s0r), "=r"(lo) : "r"(tbl_sa + (unsigned)src * 16u) : "memory");
      const double diag = lds64(tbl_sa + (unsigned)src * 16u + 8u);
      const int lr = lo & 0xff, own_s = (lo >> 16) & 0xff;
      for (int ss = sl; ss < lr; ss += lpr) {
        const double dg = ss == own_s ? diag : 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++)
          __stcs(big_m + (size_t)d * nnz + s0r + ss, fma(k_.dtt, acc[d * (acc_stride >> 3) + ss * kAS + src], dg));
      }
    }
  }
}

// ---- launch -------------------------------------------------------------------------------------------
static size_t absorb_smem(const GatherPlan* P, int dim) { return staged_acc_bytes(P, dim) + (size_t)P->nl * 128 + kBR * 16; }

// full absorption matrix, constant density, nodal absorption field, and room for two blocks per SM
bool strip_absorb_ok(const Handle* h, const MomentumArgs& A) {
  const cgasm_momentum_opts& o = A.o;
  const GatherPlan* P = h->gather;
  if (!o.have_absorption || o.lump_absorption || getenv("CGASM_STRIP_NO_ABSORB")) return false;
  if (!P || !P->staged_ok || !P->d_strip_local) return false;
  if (h->fields[CGASM_F_DENSITY].field_type != CGASM_FIELD_CONSTANT) return false;
  if (h->fields[CGASM_F_ABSORPTION].field_type != CGASM_FIELD_NORMAL) return false;
  return absorb_smem(P, h->dim) <= 112 * 1024;
}

template <int DIM>
static int strip_absorb_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = absorb_smem(P, DIM);
  const StripConsts c = consts_momentum(h, A);
  const Tables& t = A.tab;
  const double rho = h->fields[CGASM_F_DENSITY].h_const[0];
  AbsorbConsts ka;
  ka.rQa = rho * (t.Qaaa - t.Qaab);
  ka.rQaab = rho * t.Qaab;
  ka.rQd = rho * (t.Qaab - t.Qabc);
  ka.rQabc = rho * t.Qabc;
  int st = ensure_extra_records(h);
  if (st) return st;
  if (h->d_perm && !h->d_prec[5] && (st = refresh_permuted(h, 1u << 5, nullptr, 0, h->stream))) return st;
  StagedView v = staged_view(h, DIM);
  const bool fullv = strip_full_tensor(A.o.have_viscosity, A.o.viscosity_shape);
  double* ml = A.o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  int grid = P->nblocks;
#define LAUNCH(NL_, FULLV_)                                                                                     \
  do {                                                                                                          \
    if ((st = strip_smem(staged_momentum_absorb_kernel<DIM, NL_, FULLV_>, smem))) return st;                    \
    staged_momentum_absorb_kernel<DIM, NL_, FULLV_><<<grid, kBR, smem, h->stream>>>(                            \
        c, ka, v, (const double4*)staged_rec(h, 3), (const double4*)staged_rec(h, 1), (const double4*)staged_rec(h, 2), \
        (const double4*)staged_rec(h, 5), (size_t)h->nnz, h->d_big_m, h->d_mom_rhs, ml);                         \
    h->launches++;                                                                                              \
  } while (0)
#define LAUNCH_NL(NL_)                   \
  do {                                   \
    if (fullv) LAUNCH(NL_, true);        \
    else LAUNCH(NL_, false);             \
  } while (0)
  const int *ia = nullptr, *ib = nullptr;
  int na = 0, nb = 0;
  const bool split = halo_split(h, &ia, &na, &ib, &nb);
  if (split) {
    v.blocks = ia;
    grid = na;
    CGASM_FOR_NL(LAUNCH_NL);
  }
  if ((st = halo_join(h))) return st;
  if (split) {
    v.blocks = ib;
    grid = nb;
  }
  if (grid > 0) CGASM_FOR_NL(LAUNCH_NL);
#undef LAUNCH_NL
#undef LAUNCH
  CG_CUDA(cudaGetLastError());
  return st;
}

int strip_absorb_momentum(Handle* h, const MomentumArgs& A) {
  return h->dim == 3 ? strip_absorb_dim<3>(h, A) : strip_absorb_dim<2>(h, A);
}

}  // namespace cgasm
