// strip_staged.cu -- STRIP kernels with the node records of a row block staged in shared memory.
//
// ncu on the first STRIP kernels (strip.cu; profiles/r1_kernel_history.md #14): every lane fetches its
// own 2 x 32-byte records per entry, so a warp-level load waits for its slowest lane (L1 hit rate
// 68 %, L2 39 %: usually a DRAM round trip, more than one loop body of prefetch distance) and the
// L1 data pipe runs at 57 % moving 1 KB per LDG.256 through 64-byte wavefronts.
// A block of 128 Morton-adjacent rows touches only ~320-370 distinct nodes (its own 8x4x4 brick plus one
// layer), each of them ~14 x 2.4 times. So:
//   phase 0  the block copies the records of its distinct nodes (sorted list in the plan) into
//            shared memory with cp.async, coalesced, bypassing L1 and registers; one barrier;
//   loop     N = dim register buffers (the FIFO itself: shared-memory latency needs no prefetch buffer);
//            plan entries are 32 bits and carry BYTE OFFSETS (local node index << 4, CSR slot * kAS << 3):
//            the chunk stride NL of the staged records is a template parameter, so a FIFO buffer is filled by
//            LDS.128 [base + immediate] from a structure-of-16-byte-chunks layout (neighbouring rows read
//            neighbouring indices: conflict-free) after ONE mask and ONE add; nothing in the loop waits on DRAM
//            except the plan stream (three entries ahead in registers, its lines pulled into L2 ten ahead;
//            both unguarded: the plan is padded behind its last block, a block that reads ahead of its own
//            entries sees its successor's and never uses them);
//   flush    when a node leaves the FIFO its accumulated entry goes to the row's slot AND into
//            rhs -= entry * oldu(node) (Momentum_CG.F90:1712,2346 is linear in the entries), with
//            oldu read from the staged records: no epilogue pass over colm;
//   write    dt*theta and the lumped mass on the diagonal (:1550, :1484-1486) are applied while the
//            warp streams its rows out.
// Round 2 (profiles/r2_kernel_history.md): byte-offset entries + compile-time chunk stride + unguarded plan
// loads removed ~15 of the 48 non-FP64 instructions per entry; the reciprocal's Newton chain is one step shorter;
// the tracer kernel takes nodal absorption and source (Advection_Diffusion_CG.F90:1129-1162) in the same pass.
#include "strip_staged.cuh"

#include <cstdlib>

namespace cgasm {

// Copies the records of every node of the block (ids already in registers, issue_block_ids). EXTRA: 0 none, 1 oldu
// (momentum), 2 one more 16-byte chunk from a double2 array (tracer absorption / source).
template <int DIM, int NL, int EXTRA>
__device__ __forceinline__ void stage_nodes(const BlockIds<NL>& ids, int t, unsigned nsa, const double4* __restrict__ r0,
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
// One strip entry. Program order = issue order (all memory asm is volatile): flush the evicted buffer
// (slot accumulator and rhs -= entry * oldu of the evicted node), request the records of entry j and
// plan entry j + 3, then install and compute entry j.
template <int DIM, int QC, int NL, bool FULLV>
__device__ __forceinline__ void smom_step(MomState<DIM, DIM>& s, double (&rh)[DIM], double (&cc)[DIM], const StripConsts& k_,
                                          const unsigned* __restrict__ p, unsigned (&pq)[DIM], unsigned acc_sa, unsigned nsa) {
  // plan queue: slot QC holds entry j, refilled with entry j + DIM (static indices: the unroll factor is the queue
  // length, so nothing is moved between registers at the loop's back edge)
  const unsigned en = pq[QC];
  // every shared-memory read of the step is issued before the first one is consumed (one exposed LDS latency per
  // step instead of three): the evicted node's oldu and slot, then the records of the node that takes its buffer
  const unsigned m = (unsigned)s.meta[QC];
  double on[DIM];
  load_oldu<DIM, NL>(nsa, m & 0xfff0u, on);
  const unsigned sa = acc_sa + ((m >> 16) << 3);
  const double slot = lds64(sa);
  const unsigned nb = nsa + (en & 0xfff0u);
  load_rec<DIM, NL>(nb, 0, s.X[QC], s.B[QC]);
  load_rec<DIM, NL>(nb, 1, s.U[QC], s.R[QC]);
  s.meta[QC] = (int)en;
  pq[QC] = ldg_stream1(p + (QC + DIM) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
  {
    const double a = s.A[QC];
    sts64(sa, slot + a);
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-a, on[d], rh[d]);
    if constexpr (kStripRowSum<DIM>) s.a0 -= a;
    s.A[QC] = 0.0;
  }
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if constexpr (kStripCarryMomentum && DIM == 3) {
    double cn[3];
    window_cross_new<DIM, QC>(s.X, cn);
    if (en & kStagedCompute) {
      WindowGeom<3> g;
      window_geom_carry<DIM, QC>(s.X, cc, cn, g);
      mom_terms<DIM, DIM, QC, FULLV, kStripRowSum<DIM>>(s, k_, g);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) cc[a] = cn[a];
  } else {
    if (en & kStagedCompute) mom_compute<DIM, DIM, QC, FULLV, kStripRowSum<DIM>>(s, k_);
  }
}

template <int DIM, int Q, int NL, bool FULLV>
struct SMomUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, DIM>& s, double (&rh)[DIM], double (&cc)[DIM], const StripConsts& k_, Args&&... args) {
    smom_step<DIM, Q, NL, FULLV>(s, rh, cc, k_, args...);
    if constexpr (Q + 1 < DIM) SMomUnroll<DIM, Q + 1, NL, FULLV>::run(s, rh, cc, k_, args...);
  }
};

template <int DIM, int NL, bool FULLV>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 4 : (NL <= 768 ? 3 : 2)))
staged_momentum_kernel(const StripConsts k_, const StagedView P, const double4* __restrict__ rX,
                       const double4* __restrict__ rU, const double4* __restrict__ rO, size_t nnz,
                       double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const unsigned tbl_sa = nsa + (unsigned)(NL * 88);
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  // every independent load of the block first
  BlockIds<NL> ids;
  issue_block_ids<NL>(P, b, t, ids);
  const int4 meta = ldg_nc_v4(P.row_meta + (size_t)b * kBR + t);
  const long long base = ldg_nc_s64(P.ptr + b);
  prefetch_next_block<NL>(P, b, t);
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  stage_nodes<DIM, NL, 1>(ids, t, nsa, rX, rU, rO);
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
  load_rec<DIM, NL>(nsa + own_off, 0, s.X0, s.b0);
  load_rec<DIM, NL>(nsa + own_off, 1, s.U0, s.rho0);
  mom_row_consts<DIM, DIM>(s, k_);
  s.a0 = s.msum = s.nbsum = 0.0;
  double rh[DIM], cc[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rh[d] = cc[d] = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.R[q] = s.B[q] = s.A[q] = 0.0;
    s.meta[q] = (int)pad;
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR) SMomUnroll<DIM, 0, NL, FULLV>::run(s, rh, cc, k_, p, pq, acc_sa, nsa);
  // drain the FIFO, then the diagonal (the row's own node never leaves)
#pragma unroll
  for (int q = 0; q < DIM; q++) {
    const unsigned m = (unsigned)s.meta[q];
    acc_t[m >> 16] += s.A[q];
    if constexpr (kStripRowSum<DIM>) s.a0 -= s.A[q];
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
  // rows of the warp -> the dim identical diagonal blocks: dt*theta * entry (+ lumped mass on the diagonal)
  row_table_store(tbl_sa, t, meta.y, meta.z, s.msum * k_.mass_on);
  __syncwarp();
  write_rows_table<DIM>(acc, tbl_sa, t, k_.dtt, P.lpr_shift, nnz, big_m);
}

// ---- tracer -------------------------------------------------------------------------------------------
// ABS: nodal absorption and source staged as one more 16-byte chunk {absorption, source} per node
// (assemble_advection_diffusion_element_cg: add_absorption_element_cg :1144-1162, add_source_element_cg :1129-1142)
template <int DIM, int QC, int NL, bool FULLV, bool ABS>
__device__ __forceinline__ void sadv_step(AdvState<DIM, DIM>& s, double (&sg)[ABS ? DIM : 1], double (&sq)[ABS ? DIM : 1],
                                          AdvOwnExtra& ox, const StripConsts& k_, const unsigned* __restrict__ p,
                                          unsigned (&pq)[DIM], unsigned acc_sa, unsigned nsa) {
  const unsigned en = pq[QC];
  const unsigned sa = acc_sa + (((unsigned)s.meta[QC] >> 16) << 3);
  const double slot = lds64(sa);
  const unsigned nb = nsa + (en & 0xfff0u);
  double unused;
  adv_evict_rhs(s.rhs, s.A[QC], s.T[QC]);
  load_rec<DIM, NL>(nb, 0, s.X[QC], s.T[QC]);
  load_rec<DIM, NL>(nb, 1, s.U[QC], unused);
  sts64(sa, slot + fma(k_.dtt, s.A[QC], k_.mPo * s.C[QC]));
  if constexpr (kStripRowSum<DIM> && !ABS) s.a0 -= s.A[QC];
  s.A[QC] = 0.0;
  s.C[QC] = 0.0;
  if constexpr (ABS) {
    const double2 e = lds128(nb + (unsigned)(4 * NL * 16));
    sg[QC] = e.x;
    sq[QC] = e.y;
  }
  s.meta[QC] = (int)en;
  pq[QC] = ldg_stream1(p + (QC + DIM) * kBR);
  prefetch_l2(p + (QC + kPlanAhead) * kBR);
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if constexpr (kStripCarryTracer && DIM == 3 && !ABS) {
    double cn[3];
    window_cross_new<DIM, QC>(s.X, cn);
    if (en & kStagedCompute) {
      WindowGeom<3> g;
      window_geom_carry<DIM, QC>(s.X, s.cc, cn, g);
      adv_terms<DIM, DIM, QC, FULLV, kStripRowSum<DIM>>(k_, g, s.U, s.cU0, s.A, s.C, s.a0, s.c0);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) s.cc[a] = cn[a];
  } else {
    if (en & kStagedCompute) {
      if constexpr (ABS) adv_compute_abs<DIM, DIM, QC, FULLV>(s, sg, sq, ox, k_);
      else adv_compute<DIM, DIM, QC, FULLV, kStripRowSum<DIM>>(s, k_);
    }
  }
}

template <int DIM, int Q, int NL, bool FULLV, bool ABS>
struct SAdvUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(AdvState<DIM, DIM>& s, Args&&... args) {
    sadv_step<DIM, Q, NL, FULLV, ABS>(s, args...);
    if constexpr (Q + 1 < DIM) SAdvUnroll<DIM, Q + 1, NL, FULLV, ABS>::run(s, args...);
  }
};

template <int DIM, int NL, bool FULLV, bool ABS>
__global__ void __launch_bounds__(kBR, (NL <= 512 ? 4 : (NL <= 768 ? 3 : 2)))
staged_advdiff_kernel(const StripConsts k_, const StagedView P, const double4* __restrict__ rX,
                      const double4* __restrict__ rU, const double2* __restrict__ rE, double* __restrict__ matrix,
                      double* __restrict__ rhs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)P.acc_bytes;
  const unsigned tbl_sa = nsa + (unsigned)(NL * (ABS ? 80 : 64));
  const int b = P.blocks ? P.blocks[blockIdx.x] : (int)blockIdx.x, t = threadIdx.x;
  BlockIds<NL> ids;
  issue_block_ids<NL>(P, b, t, ids);
  const int4 meta = ldg_nc_v4(P.row_meta + (size_t)b * kBR + t);
  const long long base = ldg_nc_s64(P.ptr + b);
  prefetch_next_block<NL>(P, b, t);
  double* acc_t = acc + t;
  const unsigned acc_sa = (unsigned)__cvta_generic_to_shared(acc_t);
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  stage_nodes<DIM, NL, ABS ? 2 : 0>(ids, t, nsa, rX, rU, rE);
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
  AdvState<DIM, DIM> s;
  double unused;
  {
    double U0[DIM];
    load_rec<DIM, NL>(nsa + own_off, 0, s.X0, s.T0);
    load_rec<DIM, NL>(nsa + own_off, 1, U0, unused);
    adv_row_const<DIM>(k_, U0, s.cU0);
  }
  s.a0 = s.c0 = s.rhs = 0.0;
  double sg[ABS ? DIM : 1], sq[ABS ? DIM : 1];
  AdvOwnExtra ox;
  ox.sg0 = ox.sq0 = 0.0;
  if constexpr (ABS) {
    const double2 e = lds128(nsa + own_off + (unsigned)(4 * NL * 16));
    ox.sg0 = e.x;
    ox.sq0 = e.y;
  }
#pragma unroll
  for (int q = 0; q < DIM; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.T[q] = s.A[q] = s.C[q] = 0.0;
    s.cc[q] = 0.0;
    s.meta[q] = (int)pad;
    if constexpr (ABS) sg[q] = sq[q] = 0.0;
  }
  for (int j0 = 0; j0 < deg; j0 += DIM, p += DIM * kBR)
    SAdvUnroll<DIM, 0, NL, FULLV, ABS>::run(s, sg, sq, ox, k_, p, pq, acc_sa, nsa);
#pragma unroll
  for (int q = 0; q < DIM; q++) {
    acc_t[(unsigned)s.meta[q] >> 16] += fma(k_.dtt, s.A[q], k_.mPo * s.C[q]);
    adv_evict_rhs(s.rhs, s.A[q], s.T[q]);
    if constexpr (kStripRowSum<DIM> && !ABS) s.a0 -= s.A[q];
  }
  adv_finish_rhs(s.rhs, s.a0, s.T0);
  acc_t[own * kAS] += fma(k_.dtt, s.a0, k_.mPd * s.c0);
  if (r >= 0) rhs[r] = s.rhs;
  row_table_store(tbl_sa, t, meta.y, meta.z, 0.0);
  __syncwarp();
  write_rows_table<1>(acc, tbl_sa, t, 1.0, P.lpr_shift, 0, matrix);
}

// ---- launch -------------------------------------------------------------------------------------------
// bytes per staged node: 4 chunks (two records) + momentum: oldu (24) / tracer with absorption+source: one chunk
static size_t staged_smem(const GatherPlan* P, bool momentum, bool extra) {
  return staged_acc_bytes(P, 1) + (size_t)P->nl * (momentum ? 88 : (extra ? 80 : 64)) + kBR * 16;  // + the row table
}

bool strip_staged_ok(const Handle* h, bool momentum) {
  const GatherPlan* P = h->gather;
  if (!P || !P->d_strip_local || !P->staged_ok || getenv("CGASM_STRIP_GLOBAL")) return false;
  return staged_smem(P, momentum, !momentum) <= 110 * 1024;  // at least two blocks per SM, else the per-entry kernels
}

template <int DIM>
static int staged_momentum_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  size_t smem = staged_smem(P, true, false);
  const StripConsts c = consts_momentum(h, A);
  StagedView v = staged_view(h);
  const bool fullv = strip_full_tensor(A.o.have_viscosity, A.o.viscosity_shape);
  double* ml = A.o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  int st = CGASM_OK, grid = P->nblocks;
#define LAUNCH(NL_, FULLV_)                                                                                     \
  do {                                                                                                          \
    if ((st = strip_smem(staged_momentum_kernel<DIM, NL_, FULLV_>, smem))) return st;                           \
    staged_momentum_kernel<DIM, NL_, FULLV_><<<grid, kBR, smem, h->stream>>>(                                   \
        c, v, (const double4*)staged_rec(h, 3), (const double4*)staged_rec(h, 1), (const double4*)staged_rec(h, 2), (size_t)h->nnz,    \
        h->d_big_m, h->d_mom_rhs, ml);                    \
    h->launches++;                                                                                              \
  } while (0)
#define LAUNCH_NL(NL_)                   \
  do {                                   \
    if (fullv) LAUNCH(NL_, true);        \
    else LAUNCH(NL_, false);             \
  } while (0)
  // A pending halo exchange (cgasm_halo_set_overlap): the blocks that read no received node run beside it
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
  } else if (const GatherPlan::StagedClass* cls = staged_classes(h, 88, 1)) {
    // the blocks that fit a smaller chunk stride and accumulator first, at more blocks per SM; then the crowded rest
    const StagedView whole = v;
    const size_t smem_whole = smem;
    v.blocks = cls->d_small;
    v.maxlen = cls->ml_small;
    v.acc_bytes = (int)staged_acc_bytes_of(cls->ml_small, 1);
    grid = cls->n_small;
    smem = (size_t)v.acc_bytes + (size_t)cls->nl_small * 88 + kBR * 16;
    if (grid > 0) CGASM_FOR_NL_OF(cls->nl_small, LAUNCH_NL);
    v = whole;
    smem = smem_whole;
    v.blocks = cls->d_large;
    grid = cls->n_large;
  }
  if (grid > 0) CGASM_FOR_NL(LAUNCH_NL);
#undef LAUNCH_NL
#undef LAUNCH
  CG_CUDA(cudaGetLastError());
  return st;
}

int strip_staged_momentum(Handle* h, const MomentumArgs& A) {
  return h->dim == 3 ? staged_momentum_dim<3>(h, A) : staged_momentum_dim<2>(h, A);
}

template <int DIM>
static int staged_advdiff_dim(Handle* h, const AdvDiffArgs& A) {
  GatherPlan* P = h->gather;
  const bool abs = A.o.have_absorption || A.o.have_source;
  size_t smem = staged_smem(P, false, abs);
  const StripConsts c = consts_advdiff(h, A);
  StagedView v = staged_view(h);
  const bool fullv = strip_full_tensor(A.o.have_diffusivity, A.o.diffusivity_shape);
  int st = CGASM_OK, grid = P->nblocks;
#define LAUNCH(NL_, FULLV_, ABS_)                                                                               \
  do {                                                                                                          \
    if ((st = strip_smem(staged_advdiff_kernel<DIM, NL_, FULLV_, ABS_>, smem))) return st;                      \
    staged_advdiff_kernel<DIM, NL_, FULLV_, ABS_><<<grid, kBR, smem, h->stream>>>(                              \
        c, v, (const double4*)staged_rec(h, 0), (const double4*)staged_rec(h, 1), (const double2*)staged_rec(h, 4),                  \
        h->d_adv_matrix, h->d_adv_rhs);                                   \
    h->launches++;                                                                                              \
  } while (0)
#define LAUNCH_NL(NL_)                              \
  do {                                              \
    if (abs) {                                      \
      if (fullv) LAUNCH(NL_, true, true);           \
      else LAUNCH(NL_, false, true);                \
    } else {                                        \
      if (fullv) LAUNCH(NL_, true, false);          \
      else LAUNCH(NL_, false, false);               \
    }                                               \
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
  } else if (const GatherPlan::StagedClass* cls = staged_classes(h, abs ? 80 : 64, 1)) {
    const StagedView whole = v;
    const size_t smem_whole = smem;
    v.blocks = cls->d_small;
    v.maxlen = cls->ml_small;
    v.acc_bytes = (int)staged_acc_bytes_of(cls->ml_small, 1);
    grid = cls->n_small;
    smem = (size_t)v.acc_bytes + (size_t)cls->nl_small * (abs ? 80 : 64) + kBR * 16;
    if (grid > 0) CGASM_FOR_NL_OF(cls->nl_small, LAUNCH_NL);
    v = whole;
    smem = smem_whole;
    v.blocks = cls->d_large;
    grid = cls->n_large;
  }
  if (grid > 0) CGASM_FOR_NL(LAUNCH_NL);
#undef LAUNCH_NL
#undef LAUNCH
  CG_CUDA(cudaGetLastError());
  return st;
}

int strip_staged_advdiff(Handle* h, const AdvDiffArgs& A) {
  return h->dim == 3 ? staged_advdiff_dim<3>(h, A) : staged_advdiff_dim<2>(h, A);
}

}  // namespace cgasm
