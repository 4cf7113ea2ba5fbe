// strip_fused.cu -- ONE kernel for construct_momentum_element_cg + assemble_advection_diffusion_element_cg when both
// loops run the common STRIP option sets on the same mesh with the same advecting velocity (a passive tracer advected
// by nu: SURVEY.md 2.1 "fused momentum + tracer", VERDICT r1 missing #6).
//
// What the two staged kernels do twice and this one does once per (row, element) pair: the strip plan stream, the
// staging of {X}, {nu} in shared memory, the FIFO rotation, the cofactor geometry with its reciprocal (32 of the ~80 /
// ~93 FP64 instructions of the tracer / momentum pair) and the prologue / epilogue bookkeeping of a row. The terms
// themselves are the same device functions the separate kernels call (strip_common.cuh: mom_terms, adv_terms), in the
// same order on the same operands: the results are BITWISE those of cgasm_momentum_dev followed by cgasm_advdiff_dev.
// Isotropic constant viscosity / diffusivity only (the full-tensor and additive variants keep the separate kernels).
#include "strip_staged.cuh"

namespace cgasm {

// staged chunks (16 bytes, stride NL): 0,1 = {X | z, buoyancy}  2,3 = {nu | z, density}  4 = oldu {x, y};
// then plain double arrays: oldu z (at chunk 5's place), T behind it
template <int NL>
__device__ __forceinline__ unsigned fused_t_sa(unsigned nsa, unsigned noff) {
  return nsa + (unsigned)(5 * NL * 16 + NL * 8) + (noff >> 1);
}

template <int DIM, int NL>
__device__ __forceinline__ void stage_fused(const BlockIds<NL>& ids, int t, unsigned nsa, const double4* __restrict__ rX,
                                            const double4* __restrict__ rU, const double4* __restrict__ rO,
                                            const double4* __restrict__ rT) {
  const int h = t & 1;
#pragma unroll
  for (int v = 0; v < BlockIds<NL>::PER; v++) {
    const int node = ids.node[v];
    if (node < 0) continue;
    const unsigned i = (unsigned)((t >> 1) + v * (kBR / 2));
    stage_record<NL>(nsa, 0, i, h, rX, node);
    stage_record<NL>(nsa, 2, i, h, rU, node);
    stage_record_3<NL, DIM == 3>(nsa, 4, (unsigned)(5 * NL * 16), i, h, rO, node);
    if (h == 1) cp_async8(nsa + (unsigned)(5 * NL * 16 + NL * 8) + i * 8u, reinterpret_cast<const double*>(rT + node) + 3);
  }
}

template <int DIM>
struct TracerSide {
  double T[DIM], A[DIM], C[DIM];
  double T0, a0, c0, rhs;
  double cU0[DIM];
};

template <int DIM, int QC, int NL>
__device__ __forceinline__ void fused_step(MomState<DIM, DIM>& s, TracerSide<DIM>& q, double (&rh)[DIM], const StripConsts& km,
                                           const StripConsts& ka, const unsigned* __restrict__ p, unsigned (&pq)[DIM],
                                           unsigned acc_sa, unsigned acc2_sa, unsigned nsa) {
  const unsigned en = pq[QC];
  const unsigned m = (unsigned)s.meta[QC];
  double on[DIM];
  load_oldu<DIM, NL>(nsa, m & 0xfff0u, on);
  const unsigned so = (m >> 16) << 3;
  const double slot = lds64(acc_sa + so), slot2 = lds64(acc2_sa + so);
  const unsigned noff = en & 0xfff0u, nb = nsa + noff;
  load_rec<DIM, NL>(nb, 0, s.X[QC], s.B[QC]);
  load_rec<DIM, NL>(nb, 1, s.U[QC], s.R[QC]);
  const double tnew = lds64(fused_t_sa<NL>(nsa, noff));
  s.meta[QC] = (int)en;
  pq[QC] = ldg_stream1(p + (QC + DIM) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
  {
    const double a = s.A[QC];
    sts64(acc_sa + so, slot + a);
    sts64(acc2_sa + so, slot2 + fma(ka.dtt, q.A[QC], ka.mPo * q.C[QC]));
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-a, on[d], rh[d]);
    adv_evict_rhs(q.rhs, q.A[QC], q.T[QC]);
    if constexpr (kStripRowSum<DIM>) {  // as the two separate kernels do (bitwise the same diagonals)
      s.a0 -= a;
      q.a0 -= q.A[QC];
    }
    s.A[QC] = 0.0;
    q.A[QC] = 0.0;
    q.C[QC] = 0.0;
    q.T[QC] = tnew;
  }
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if (en & kStagedCompute) {
    WindowGeom<DIM> g;
    window_geom<DIM, DIM, QC>(s.X, g);
    mom_terms<DIM, DIM, QC, false, kStripRowSum<DIM>>(s, km, g);
    adv_terms<DIM, DIM, QC, false, kStripRowSum<DIM>>(ka, g, s.U, q.cU0, q.A, q.C, q.a0, q.c0);
  }
}

template <int DIM, int Q, int NL>
struct FusedUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, DIM>& s, Args&&... args) {
    fused_step<DIM, Q, NL>(s, args...);
    if constexpr (Q + 1 < DIM) FusedUnroll<DIM, Q + 1, NL>::run(s, args...);
  }
};

template <int DIM, int NL>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 3 : 2))
staged_fused_kernel(const StripConsts km, const StripConsts ka, const StagedView P, const double4* __restrict__ rX,
                    const double4* __restrict__ rU, const double4* __restrict__ rO, const double4* __restrict__ rT,
                    size_t nnz, double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump,
                    double* __restrict__ matrix, double* __restrict__ arhs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  double* acc2 = acc + P.maxlen * kAS;
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const unsigned tbl_sa = nsa + (unsigned)(NL * 96);
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  BlockIds<NL> ids;
  issue_block_ids<NL>(P, b, t, ids);
  const int4 meta = ldg_nc_v4(P.row_meta + (size_t)b * kBR + t);
  const long long base = ldg_nc_s64(P.ptr + b);
  prefetch_next_block<NL>(P, b, t);
  double* acc_t = acc + t;
  double* acc2_t = acc2 + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t), acc2_sa = (unsigned)__cvta_generic_to_shared(acc2_t);
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = acc2_t[q * kAS] = 0.0;
  stage_fused<DIM, NL>(ids, t, nsa, rX, rU, rO, rT);
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
  TracerSide<DIM> q;
  load_rec<DIM, NL>(nsa + own_off, 0, s.X0, s.b0);
  load_rec<DIM, NL>(nsa + own_off, 1, s.U0, s.rho0);
  mom_row_consts<DIM, DIM>(s, km);
  q.T0 = lds64(fused_t_sa<NL>(nsa, own_off));
  adv_row_const<DIM>(ka, s.U0, q.cU0);
  s.a0 = s.msum = s.nbsum = 0.0;
  q.a0 = q.c0 = q.rhs = 0.0;
  double rh[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rh[d] = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[k][a] = s.U[k][a] = 0.0;
    s.R[k] = s.B[k] = s.A[k] = 0.0;
    q.T[k] = q.A[k] = q.C[k] = 0.0;
    s.meta[k] = (int)pad;
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR)
    FusedUnroll<DIM, 0, NL>::run(s, q, rh, km, ka, p, pq, acc_sa, acc2_sa, nsa);
  // drain the FIFO, then the diagonals
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    const unsigned m = (unsigned)s.meta[k];
    acc_t[m >> 16] += s.A[k];
    acc2_t[m >> 16] += fma(ka.dtt, q.A[k], ka.mPo * q.C[k]);
    adv_evict_rhs(q.rhs, q.A[k], q.T[k]);
    if constexpr (kStripRowSum<DIM>) {
      s.a0 -= s.A[k];
      q.a0 -= q.A[k];
    }
    double o[DIM];
    load_oldu<DIM, NL>(nsa, m & 0xfff0u, o);
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-s.A[k], o[d], rh[d]);
  }
  acc_t[own * kAS] += s.a0;
  acc2_t[own * kAS] += fma(ka.dtt, q.a0, ka.mPd * q.c0);
  if (r >= 0) {
    double ou[DIM];
    load_oldu<DIM, NL>(nsa, own_off, ou);
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      rhs[(size_t)DIM * r + d] = fma(-s.a0, ou[d], fma(km.grav[d], s.nbsum, rh[d]));
      if (masslump) masslump[(size_t)DIM * r + d] = s.msum;
    }
    adv_finish_rhs(q.rhs, q.a0, q.T0);
    arhs[r] = q.rhs;
  }
  // rows of the warp: the dim identical momentum blocks (dt*theta * entry + lumped mass on the diagonal), then the tracer matrix
  row_table_store(tbl_sa, t, meta.y, meta.z, s.msum * km.mass_on);
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
      const int lr = lo & 0xff, ownr = (lo >> 16) & 0xff;
      for (int ss = sl; ss < lr; ss += lpr) {
        const double v = fma(km.dtt, acc[ss * kAS + src], ss == ownr ? diag : 0.0);
#pragma unroll
        for (int d = 0; d < DIM; d++) __stcs(big_m + (size_t)d * nnz + s0r + ss, v);
        __stcs(matrix + s0r + ss, acc2[ss * kAS + src]);
      }
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------------
static size_t fused_smem(const GatherPlan* P) { return staged_acc_bytes(P, 2) + (size_t)P->nl * 96 + kBR * 16; }

bool strip_fused_ok(const Handle* h, const MomentumArgs& M, const AdvDiffArgs& A) {
  const GatherPlan* P = h->gather;
  if (h->scatter != CGASM_SCATTER_STRIP || !P || !P->staged_ok || !P->d_strip_local || getenv("CGASM_STRIP_GLOBAL") ||
      getenv("CGASM_NO_FUSED"))
    return false;
  if (!strip_momentum_opts_ok(M) || strip_extra_needed(M) || strip_full_tensor(M.o.have_viscosity, M.o.viscosity_shape)) return false;
  if (M.o.assemble_ct_matrix_here) return false;
  if (!strip_advdiff_opts_ok(A) || strip_advdiff_needs_extra(A) || strip_full_tensor(A.o.have_diffusivity, A.o.diffusivity_shape))
    return false;
  return fused_smem(P) <= 110 * 1024;
}

template <int DIM>
static int strip_fused_dim(Handle* h, const MomentumArgs& M, const AdvDiffArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = fused_smem(P);
  const StripConsts km = consts_momentum(h, M), ka = consts_advdiff(h, A);
  StagedView v = staged_view(h, 2);
  double* ml = M.o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  int st = CGASM_OK, grid = P->nblocks;
#define LAUNCH_NL(NL_)                                                                                          \
  do {                                                                                                          \
    if ((st = strip_smem(staged_fused_kernel<DIM, NL_>, smem))) return st;                                      \
    staged_fused_kernel<DIM, NL_><<<grid, kBR, smem, h->stream>>>(                                              \
        km, ka, v, (const double4*)staged_rec(h, 3), (const double4*)staged_rec(h, 1),                          \
        (const double4*)staged_rec(h, 2), (const double4*)staged_rec(h, 0), (size_t)h->nnz, h->d_big_m,         \
                                                                  h->d_mom_rhs, ml, h->d_adv_matrix, h->d_adv_rhs); \
    h->launches++;                                                                                              \
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
  CG_CUDA(cudaGetLastError());
  return st;
}

int strip_fused(Handle* h, const MomentumArgs& M, const AdvDiffArgs& A) {
  h->mom_path = h->adv_path = CGASM_PATH_STRIP_STAGED;
  return h->dim == 3 ? strip_fused_dim<3>(h, M, A) : strip_fused_dim<2>(h, M, A);
}

}  // namespace cgasm
