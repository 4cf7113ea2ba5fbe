// strip_pipe.cu -- staged STRIP kernels with the shared-memory reads of a strip step issued ONE STEP AHEAD.
//
// ncu source page of staged_momentum_kernel (profiles/r2_kernel_history.md #41): 23 % of a warp's time in the loop is
// short-scoreboard stall -- the step's own LDS (records of the node it pushes, oldu and slot of the node it evicts) are
// issued at its top and consumed a few instructions later, and with the L1 data pipe 67 % busy a shared-memory round
// trip takes ~200 cycles. The FIFO kernels cannot issue them earlier: the register buffer an entry is loaded into is the
// one its predecessor's element still reads.
// Here the FIFO has dim + 1 rotating register buffers. In step j (buffer QC = j mod (dim + 1)):
//   install   node j, whose records were requested during step j - 1;
//   flush     node j - dim (buffer QC + 1), with the oldu / slot values requested during step j - 1;
//   request   the records of node j + 1 into the buffer just flushed, and oldu / slot of node j + 1 - dim (the next flush);
//   compute   the element {row, j, j-1, .., j-dim+1} if the plan says so.
// Every LDS has a full step (~1000 cycles) to land. Price: 26 more registers (152: three blocks per SM instead of four).
// Same plan, same arithmetic per element, same order of additions per slot as the FIFO kernels: bitwise the same results.
#include "strip_staged.cuh"

#include <cstdlib>

namespace cgasm {

// (same as strip_staged.cu) records of every node of the block -> shared memory
template <int DIM, int NL, int EXTRA>
__device__ __forceinline__ void stage_nodes_p(const BlockIds<NL>& ids, int t, unsigned nsa, const double4* __restrict__ r0,
                                              const double4* __restrict__ r1, const void* __restrict__ rE) {
  const int h = t & 1;
#pragma unroll
  for (int v = 0; v < BlockIds<NL>::PER; v++) {
    const int node = ids.node[v];
    if (node < 0) continue;
    const unsigned i = (unsigned)((t >> 1) + v * (kBR / 2));
    stage_record<NL>(nsa, 0, i, h, r0, node);
    stage_record<NL>(nsa, 2, i, h, r1, node);
    if constexpr (EXTRA == 1) {
      stage_record_3<NL, DIM == 3>(nsa, 4, (unsigned)(5 * NL * 16), i, h, reinterpret_cast<const double4*>(rE), node);
    } else if constexpr (EXTRA == 2) {
      if (h == 0) cp_async16(nsa + (unsigned)(4 * NL * 16) + i * 16u, reinterpret_cast<const double2*>(rE) + node);
    }
  }
}

// ---- momentum -----------------------------------------------------------------------------------------
template <int DIM>
struct FlushAhead {
  double on[DIM];  // oldu of the node the next step flushes
  double slot;     // its accumulator slot
};

template <int DIM, int QC, int NL, bool FULLV>
__device__ __forceinline__ void pmom_step(MomState<DIM, DIM + 1>& s, FlushAhead<DIM>& fa, double (&rh)[DIM], const StripConsts& k_,
                                          const unsigned* __restrict__ p, unsigned (&pq)[DIM + 1], unsigned acc_sa, unsigned nsa) {
  constexpr int N = DIM + 1;
  constexpr int QF = (QC + 1) % N;  // buffer of node j - dim: flushed now, refilled with node j + 1
  constexpr int QN = (QC + 2) % N;  // buffer of node j + 1 - dim: the next step's flush
  // install node j (requested one step ago)
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  // flush node j - dim with the values requested one step ago
  {
    const unsigned m = (unsigned)s.meta[QF];
    const double a = s.A[QF];
    sts64(acc_sa + ((m >> 16) << 3), fa.slot + a);
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-a, fa.on[d], rh[d]);
    s.A[QF] = 0.0;
  }
  // requests for step j + 1 (program order = issue order: the slot read follows the slot write above)
  {
    const unsigned en = pq[QF];
    const unsigned nb = nsa + (en & 0xfff0u);
    load_rec<DIM, NL>(nb, 0, s.X[QF], s.B[QF]);
    load_rec<DIM, NL>(nb, 1, s.U[QF], s.R[QF]);
    s.meta[QF] = (int)en;
    const unsigned mn = (unsigned)s.meta[QN];
    load_oldu<DIM, NL>(nsa, mn & 0xfff0u, fa.on);
    fa.slot = lds64(acc_sa + ((mn >> 16) << 3));
    pq[QF] = ldg_stream1(p + (QC + 1 + N) * kBR);  // entry j + 1 + N takes the place of entry j + 1
    prefetch_l2(p + (QC + kPlanAhead) * kBR);
  }
  if ((unsigned)s.meta[QC] & kStagedCompute) mom_compute<DIM, N, QC, FULLV>(s, k_);
}

template <int DIM, int Q, int NL, bool FULLV>
struct PMomUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, DIM + 1>& s, Args&&... args) {
    pmom_step<DIM, Q, NL, FULLV>(s, args...);
    if constexpr (Q + 1 < DIM + 1) PMomUnroll<DIM, Q + 1, NL, FULLV>::run(s, args...);
  }
};

template <int DIM, int NL, bool FULLV>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 3 : 2))
piped_momentum_kernel(const StripConsts k_, const StagedView P, const double4* __restrict__ rX, const double4* __restrict__ rU,
                      const double4* __restrict__ rO, size_t nnz, double* __restrict__ big_m, double* __restrict__ rhs,
                      double* __restrict__ masslump) {
  constexpr int N = DIM + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const unsigned tbl_sa = nsa + (unsigned)(NL * 88);
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  BlockIds<NL> ids;
  issue_block_ids<NL>(P, b, t, ids);
  const int4 meta = ldg_nc_v4(P.row_meta + (size_t)b * kBR + t);
  const long long base = ldg_nc_s64(P.ptr + b);
  prefetch_next_block<NL>(P, b, t);
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  stage_nodes_p<DIM, NL, 1>(ids, t, nsa, rX, rU, rO);
  const int deg = warp_trip_count<N>(meta.z);
  const unsigned* p = P.ent + base + t;
  // plan queue: pq[q] holds entry j with j mod N == q; entry 0 goes straight into buffer 0
  unsigned pq[N];
  const unsigned en0 = ldg_stream1(p);
#pragma unroll
  for (int q = 1; q < N; q++) pq[q] = ldg_stream1(p + q * kBR);
  pq[0] = ldg_stream1(p + N * kBR);
#pragma unroll
  for (int q = N + 1; q < kPlanAhead; q++) prefetch_l2(p + q * kBR);
  const int r = meta.x;
  const unsigned pad = (unsigned)meta.w;
  const unsigned own_off = pad & 0xfff0u;
  const int own = (meta.z >> 16) & 0xff;
  cp_async_commit_wait_all();
  __syncthreads();
  MomState<DIM, N> s;
  load_rec<DIM, NL>(nsa + own_off, 0, s.X0, s.b0);
  load_rec<DIM, NL>(nsa + own_off, 1, s.U0, s.rho0);
  s.a0 = s.msum = s.nbsum = 0.0;
  double rh[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rh[d] = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.R[q] = s.B[q] = s.A[q] = 0.0;
    s.meta[q] = (int)pad;  // the own node, compute flag clear: a flush of it adds 0 to the diagonal slot
  }
  // step 0's inputs: node 0 in buffer 0, the first flush (buffer 1: the padding node, A = 0)
  load_rec<DIM, NL>(nsa + (en0 & 0xfff0u), 0, s.X[0], s.B[0]);
  load_rec<DIM, NL>(nsa + (en0 & 0xfff0u), 1, s.U[0], s.R[0]);
  s.meta[0] = (int)en0;
  FlushAhead<DIM> fa;
  load_oldu<DIM, NL>(nsa, own_off, fa.on);
  fa.slot = 0.0;
  for (int j0 = 0; j0 < deg; j0 += N, p += N * kBR) PMomUnroll<DIM, 0, NL, FULLV>::run(s, fa, rh, k_, p, pq, acc_sa, nsa);
  // drain: after deg (a multiple of N) steps the window is buffers N-1, .., 1; buffer 0 holds the read-ahead of entry deg
#pragma unroll
  for (int q = 1; q < N; q++) {
    const unsigned m = (unsigned)s.meta[q];
    acc_t[m >> 16] += s.A[q];
    double o[DIM];
    load_oldu<DIM, NL>(nsa, m & 0xfff0u, o);
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-s.A[q], o[d], rh[d]);
  }
  acc_t[own * kAS] += s.a0;
  if (r >= 0) {
    double ou[DIM];
    load_oldu<DIM, NL>(nsa, own_off, ou);
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      rhs[(size_t)DIM * r + d] = fma(-s.a0, ou[d], fma(k_.grav[d], s.nbsum, rh[d]));
      if (masslump) masslump[(size_t)DIM * r + d] = s.msum;
    }
  }
  row_table_store(tbl_sa, t, meta.y, meta.z, s.msum * k_.mass_on);
  __syncwarp();
  write_rows_table<DIM>(acc, tbl_sa, t, k_.dtt, P.lpr_shift, nnz, big_m);
}

// ---- tracer -------------------------------------------------------------------------------------------
template <int DIM, int QC, int NL, bool FULLV>
__device__ __forceinline__ void padv_step(AdvState<DIM, DIM + 1>& s, double& fslot, const StripConsts& k_,
                                          const unsigned* __restrict__ p, unsigned (&pq)[DIM + 1], unsigned acc_sa, unsigned nsa) {
  constexpr int N = DIM + 1;
  constexpr int QF = (QC + 1) % N, QN = (QC + 2) % N;
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  {
    const unsigned m = (unsigned)s.meta[QF];
    sts64(acc_sa + ((m >> 16) << 3), fslot + fma(k_.dtt, s.A[QF], k_.mPo * s.C[QF]));
    s.A[QF] = 0.0;
    s.C[QF] = 0.0;
  }
  {
    const unsigned en = pq[QF];
    const unsigned nb = nsa + (en & 0xfff0u);
    double unused;
    load_rec<DIM, NL>(nb, 0, s.X[QF], s.T[QF]);
    load_rec<DIM, NL>(nb, 1, s.U[QF], unused);
    s.meta[QF] = (int)en;
    fslot = lds64(acc_sa + (((unsigned)s.meta[QN] >> 16) << 3));
    pq[QF] = ldg_stream1(p + (QC + 1 + N) * kBR);  // entry j + 1 + N takes the place of entry j + 1
    prefetch_l2(p + (QC + kPlanAhead) * kBR);
  }
  if ((unsigned)s.meta[QC] & kStagedCompute) adv_compute<DIM, N, QC, FULLV>(s, k_);
}

template <int DIM, int Q, int NL, bool FULLV>
struct PAdvUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(AdvState<DIM, DIM + 1>& s, Args&&... args) {
    padv_step<DIM, Q, NL, FULLV>(s, args...);
    if constexpr (Q + 1 < DIM + 1) PAdvUnroll<DIM, Q + 1, NL, FULLV>::run(s, args...);
  }
};

template <int DIM, int NL, bool FULLV>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 3 : 2))
piped_advdiff_kernel(const StripConsts k_, const StagedView P, const double4* __restrict__ rX, const double4* __restrict__ rU,
                     double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int N = DIM + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const unsigned tbl_sa = nsa + (unsigned)(NL * 64);
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  BlockIds<NL> ids;
  issue_block_ids<NL>(P, b, t, ids);
  const int4 meta = ldg_nc_v4(P.row_meta + (size_t)b * kBR + t);
  const long long base = ldg_nc_s64(P.ptr + b);
  prefetch_next_block<NL>(P, b, t);
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  stage_nodes_p<DIM, NL, 0>(ids, t, nsa, rX, rU, nullptr);
  const int deg = warp_trip_count<N>(meta.z);
  const unsigned* p = P.ent + base + t;
  unsigned pq[N];
  const unsigned en0 = ldg_stream1(p);
#pragma unroll
  for (int q = 1; q < N; q++) pq[q] = ldg_stream1(p + q * kBR);
  pq[0] = ldg_stream1(p + N * kBR);
#pragma unroll
  for (int q = N + 1; q < kPlanAhead; q++) prefetch_l2(p + q * kBR);
  const int r = meta.x;
  const unsigned pad = (unsigned)meta.w;
  const unsigned own_off = pad & 0xfff0u;
  const int own = (meta.z >> 16) & 0xff;
  cp_async_commit_wait_all();
  __syncthreads();
  AdvState<DIM, N> s;
  double unused;
  {
    double U0[DIM];
    load_rec<DIM, NL>(nsa + own_off, 0, s.X0, s.T0);
    load_rec<DIM, NL>(nsa + own_off, 1, U0, unused);
    adv_row_const<DIM>(k_, U0, s.cU0);
  }
  s.a0 = s.c0 = s.rhs = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.T[q] = s.A[q] = s.C[q] = 0.0;
    s.meta[q] = (int)pad;
  }
  load_rec<DIM, NL>(nsa + (en0 & 0xfff0u), 0, s.X[0], s.T[0]);
  load_rec<DIM, NL>(nsa + (en0 & 0xfff0u), 1, s.U[0], unused);
  s.meta[0] = (int)en0;
  double fslot = 0.0;
  for (int j0 = 0; j0 < deg; j0 += N, p += N * kBR) PAdvUnroll<DIM, 0, NL, FULLV>::run(s, fslot, k_, p, pq, acc_sa, nsa);
#pragma unroll
  for (int q = 1; q < N; q++) acc_t[(unsigned)s.meta[q] >> 16] += fma(k_.dtt, s.A[q], k_.mPo * s.C[q]);
  acc_t[own * kAS] += fma(k_.dtt, s.a0, k_.mPd * s.c0);
  if (r >= 0) rhs[r] = s.rhs;
  row_table_store(tbl_sa, t, meta.y, meta.z, 0.0);
  __syncwarp();
  write_rows_table<1>(acc, tbl_sa, t, 1.0, P.lpr_shift, 0, matrix);
}

// ---- launch -------------------------------------------------------------------------------------------
static size_t piped_smem(const GatherPlan* P, bool momentum) {
  return staged_acc_bytes(P, 1) + (size_t)P->nl * (momentum ? 88 : 64) + kBR * 16;
}

// CGASM_STRIP_PIPE=1 selects the pipelined kernels (momentum without the in-loop absorption, tracer without absorption /
// source); 0 or unset: the FIFO kernels of strip_staged.cu
bool strip_piped_ok(const Handle* h, bool momentum) {
  const char* e = getenv("CGASM_STRIP_PIPE");
  if (!e || atoi(e) == 0) return false;
  const GatherPlan* P = h->gather;
  return P && P->staged_ok && P->d_strip_local && piped_smem(P, momentum) <= 110 * 1024;
}

template <int DIM>
static int piped_momentum_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = piped_smem(P, true);
  const StripConsts c = consts_momentum(h, A);
  StagedView v = staged_view(h);
  const bool fullv = strip_full_tensor(A.o.have_viscosity, A.o.viscosity_shape);
  double* ml = A.o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  int st = CGASM_OK, grid = P->nblocks;
#define LAUNCH(NL_, FULLV_)                                                                                     \
  do {                                                                                                          \
    if ((st = strip_smem(piped_momentum_kernel<DIM, NL_, FULLV_>, smem))) return st;                            \
    piped_momentum_kernel<DIM, NL_, FULLV_><<<grid, kBR, smem, h->stream>>>(                                    \
        c, v, (const double4*)staged_rec(h, 3), (const double4*)staged_rec(h, 1), (const double4*)staged_rec(h, 2), \
        (size_t)h->nnz, h->d_big_m, h->d_mom_rhs, ml);                                                          \
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

int strip_piped_momentum(Handle* h, const MomentumArgs& A) {
  return h->dim == 3 ? piped_momentum_dim<3>(h, A) : piped_momentum_dim<2>(h, A);
}

template <int DIM>
static int piped_advdiff_dim(Handle* h, const AdvDiffArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = piped_smem(P, false);
  const StripConsts c = consts_advdiff(h, A);
  StagedView v = staged_view(h);
  const bool fullv = strip_full_tensor(A.o.have_diffusivity, A.o.diffusivity_shape);
  int st = CGASM_OK, grid = P->nblocks;
#define LAUNCH(NL_, FULLV_)                                                                                     \
  do {                                                                                                          \
    if ((st = strip_smem(piped_advdiff_kernel<DIM, NL_, FULLV_>, smem))) return st;                             \
    piped_advdiff_kernel<DIM, NL_, FULLV_><<<grid, kBR, smem, h->stream>>>(                                     \
        c, v, (const double4*)staged_rec(h, 0), (const double4*)staged_rec(h, 1), h->d_adv_matrix, h->d_adv_rhs); \
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

int strip_piped_advdiff(Handle* h, const AdvDiffArgs& A) {
  return h->dim == 3 ? piped_advdiff_dim<3>(h, A) : piped_advdiff_dim<2>(h, A);
}

}  // namespace cgasm
