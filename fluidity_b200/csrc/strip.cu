// strip.cu -- CGASM_SCATTER_STRIP: row-owner assembly with a strip-ordered element stream.
// This file: the plan (strip_build), the dispatch, and the kernels that fetch each node's records from
// global memory per entry. They are the FALLBACK: the default kernels stage a block's node records in
// shared memory (strip_staged.cu) and run whenever that staging fits (strip_staged_ok).
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
//  * The records of entry j+1 are requested while entry j is being computed (register prefetch; the
//    staged kernels read shared memory instead and need no prefetch buffer: N = dim there).
//  * A contribution to column v accumulates in a REGISTER for as long as v sits in the FIFO and is
//    added to the row's shared-memory slot once, when v is evicted (1.4 RMW per pair instead of 4).
//  * rhs -= (A + K) oldu (Momentum_CG.F90:1712,2346) is linear in the assembled row, so it is one
//    sparse dot product per row after the loop (colm + the {oldu, buoyancy} records; the staged kernels
//    fold it into the eviction instead) rather than dim*loc FMAs and one more 32-byte record per pair.
//    Likewise the constant gravity direction (:1786-1789) and dt*theta (:1484-1486) are applied once per row.
//  * Momentum reads two records per node, {X, buoyancy} and {nu, density}.
//
// Entries per (row, element) pair: 1.375 on Kuhn meshes (33 pushes for the 24 elements of an
// interior node). Summation order is fixed by the plan: results are bitwise reproducible.
// Option coverage (strip_momentum_opts_ok / strip_advdiff_opts_ok): lumped or excluded mass, advection
// on/off (not by parts, beta = 0), CONSTANT viscosity / diffusivity of any tensor shape or none, constant
// gravity direction with nodal buoyancy or no gravity, tracer mass consistent / lumped / none -- the switches
// are folded into the coefficients on the host (strip_common.cuh), so one kernel (two with the full-tensor
// variant) serves them all: driven_cavity, lock_exchange, flow_past_sphere_Re100 and S3 take this path.
// Absorption, sources, SU/SUPG, by-parts advection, nodal viscosity run the GATHER kernels (gather.cu).
#include "strip_common.cuh"
#include "strip_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace cgasm {

template <int DIM, int N, int QC>
__device__ __forceinline__ void mom_step(MomState<DIM, N>& s, const StripConsts& k_, int j, int deg,
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
  unpack<DIM>(ld256v(rX + en.x), s.X[QE], s.B[QE]);
  unpack<DIM>(ld256v(rU + en.x), s.U[QE], s.R[QE]);
  s.meta[QE] = en.y;
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if (s.meta[QC] & kStripCompute) mom_compute<DIM, N, QC, false>(s, k_);
}

template <int DIM, int N, int Q>
struct MomUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, N>& s, const StripConsts& k_, int j0, Args&&... args) {
    mom_step<DIM, N, Q>(s, k_, j0 + Q, args...);
    if constexpr (Q + 1 < N) MomUnroll<DIM, N, Q + 1>::run(s, k_, j0, args...);
  }
};

template <int DIM, int N, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
strip_momentum_kernel(const StripConsts k_, const StripPlanView P, const double4* __restrict__ rX,
                      const double4* __restrict__ rU, const double4* __restrict__ rO, size_t nnz,
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
  MomState<DIM, N> s;
  unpack<DIM>(ld256(rX + r0), s.X0, s.b0);
  unpack<DIM>(ld256(rU + r0), s.U0, s.rho0);
  mom_row_consts<DIM, N>(s, k_);
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
  for (int j0 = 0; j0 < deg; j0 += N) MomUnroll<DIM, N, 0>::run(s, k_, j0, deg, p, pq0, pq1, pad, acc_t, rX, rU);
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
    for (int d = 0; d < DIM; d++) rh[d] = k_.grav[d] * s.nbsum;
    for (int q = 0; q < my_len; q++) {
      const int col = __ldg(P.colm + my_s0 + q);
      const double4 o = ld256(rO + col);
      const double v = acc_t[q * kAS];
      rh[0] = fma(-v, o.x, rh[0]);
      rh[1] = fma(-v, o.y, rh[1]);
      if constexpr (DIM == 3) rh[2] = fma(-v, o.z, rh[2]);
      acc_t[q * kAS] = fma(k_.dtt, v, q == own ? s.msum * k_.mass_on : 0.0);
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
template <int DIM, int N, int QC>
__device__ __forceinline__ void adv_step(AdvState<DIM, N>& s, const StripConsts& k_, int j, int deg,
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
    *sl += fma(k_.dtt, s.A[QE], k_.mPo * s.C[QE]);
    adv_evict_rhs(s.rhs, s.A[QE], s.T[QE]);
    s.A[QE] = 0.0;
    s.C[QE] = 0.0;
  }
  double unused;
  unpack<DIM>(ld256v(rX + en.x), s.X[QE], s.T[QE]);
  unpack<DIM>(ld256v(rU + en.x), s.U[QE], unused);
  s.meta[QE] = en.y;
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if (s.meta[QC] & kStripCompute) adv_compute<DIM, N, QC, false>(s, k_);
}

template <int DIM, int N, int Q>
struct AdvUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(AdvState<DIM, N>& s, const StripConsts& k_, int j0, Args&&... args) {
    adv_step<DIM, N, Q>(s, k_, j0 + Q, args...);
    if constexpr (Q + 1 < N) AdvUnroll<DIM, N, Q + 1>::run(s, k_, j0, args...);
  }
};

template <int DIM, int N, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
strip_advdiff_kernel(const StripConsts k_, const StripPlanView P, const double4* __restrict__ rX,
                     const double4* __restrict__ rU, double* __restrict__ matrix, double* __restrict__ rhs) {
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
  AdvState<DIM, N> s;
  double unused;
  unpack<DIM>(ld256(rX + r0), s.X0, s.T0);
  {
    double U0[DIM];
    unpack<DIM>(ld256(rU + r0), U0, unused);
    adv_row_const<DIM>(k_, U0, s.cU0);
  }
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
  for (int j0 = 0; j0 < deg; j0 += N) AdvUnroll<DIM, N, 0>::run(s, k_, j0, deg, p, pq0, pq1, pad, acc_t, rX, rU);
#pragma unroll
  for (int q = 0; q < N; q++) {
    acc_t[(s.meta[q] & 0xff) * kAS] += fma(k_.dtt, s.A[q], k_.mPo * s.C[q]);
    adv_evict_rhs(s.rhs, s.A[q], s.T[q]);
  }
  adv_finish_rhs(s.rhs, s.a0, s.T0);
  acc_t[own * kAS] += fma(k_.dtt, s.a0, k_.mPd * s.c0);
  int my_s0 = 0, my_len = 0;
  if (r >= 0) {
    my_s0 = P.findrm[r];
    my_len = P.findrm[r + 1] - my_s0;
    rhs[r] = s.rhs;
  }
  __syncwarp();
  write_rows<1>(acc, t, my_s0, my_len, P.lpr_shift, 0, matrix);
}

// ---- plan ---------------------------------------------------------------------------------------------
// Block degrees are padded to a multiple of the kernels' unroll factor so the loops need no tail: dim + 1
// buffers for the per-entry-fetch kernels (one entry of prefetch), dim for the staged ones (FIFO only).
static int pad_to(int deg, int mult) { return (deg + mult - 1) / mult * mult; }

// per-entry-fetch plan (64-bit entries, global node ids): only for meshes the staged kernels cannot take
static int strip_build_global(Handle* h) {
  GatherPlan* P = h->gather;
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
  for (int b = 0; b < nb; b++) ptr[b + 1] = ptr[b] + (long long)pad_to(block_deg[b], h->dim + 1) * kBR;
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
  CG_CUDA(cudaMalloc(&P->d_strip_ptr, sizeof(long long) * ptr.size()));
  CG_CUDA(cg_upload(P->d_strip_ptr, ptr.data(), sizeof(long long) * ptr.size()));
  CG_CUDA(cudaMalloc(&P->d_strip, sizeof(int2) * ent.size()));
  CG_CUDA(cg_upload(P->d_strip, ent.data(), sizeof(int2) * ent.size()));
  if (!P->d_own_slot) {
    std::vector<unsigned char> own_slot(rows.size(), 0);
    for (size_t q = 0; q < rows.size(); q++) {
      const int r = rows[q];
      if (r < 0) continue;
      const int* cb = h->h_colm.data() + h->h_findrm[r];
      own_slot[q] = (unsigned char)(std::lower_bound(cb, (const int*)h->h_colm.data() + h->h_findrm[r + 1], r) - cb);
    }
    CG_CUDA(cudaMalloc(&P->d_own_slot, own_slot.size()));
    CG_CUDA(cg_upload(P->d_own_slot, own_slot.data(), own_slot.size()));
  }
  return CGASM_OK;
}

// The staged plan is built first (strip_plan.cpp, host only); the per-entry-fetch plan only if the mesh does not fit
// the staged kernels (more than 1024 distinct nodes around a 128-row block, very long rows) or CGASM_STRIP_GLOBAL asks.
int strip_build(Handle* h) {
  GatherPlan* P = h->gather;
  if (!P) CG_FAIL(CGASM_ESTATE, "strip scatter: the gather row blocks must exist first");
  if (P->d_strip || P->d_strip_local) return CGASM_OK;
  // Scattered numbering (a gmsh / adapted mesh without renumbering; tests: synthetic.shuffled): the rows of a block, sorted
  // by node id, are hardly ever consecutive. The staged kernels then read permuted mirrors of the node records, laid out in
  // the order of the row blocks. CGASM_STRIP_PERMUTE=0/1 overrides the choice.
  std::vector<int> perm;
  {
    long long pairs = 0, consecutive = 0;
    const std::vector<int>& rows = P->h_rows;
    std::vector<int> ids;
    for (size_t b0 = 0; b0 < rows.size(); b0 += kBR) {  // (the rows of a block are ordered by node degree first)
      ids.clear();
      for (size_t q = b0; q < b0 + kBR && q < rows.size(); q++)
        if (rows[q] >= 0) ids.push_back(rows[q]);
      std::sort(ids.begin(), ids.end());
      for (size_t q = 1; q < ids.size(); q++) {
        pairs++;
        consecutive += ids[q] == ids[q - 1] + 1;
      }
    }
    bool scattered = pairs > 0 && (double)consecutive < 0.25 * (double)pairs;
    if (const char* e = getenv("CGASM_STRIP_PERMUTE")) scattered = atoi(e) != 0;
    if (scattered) {
      perm.assign((size_t)h->n_nodes, 0);
      int next = 0;
      for (size_t q = 0; q < rows.size(); q++)
        if (rows[q] >= 0) perm[rows[q]] = next++;
      if (getenv("CGASM_DEBUG"))
        fprintf(stderr, "[cgasm] scattered numbering (%.1f %% of a block's rows consecutive): node records mirrored in row-block order\n",
                pairs ? 100.0 * consecutive / pairs : 0.0);
    }
  }
  {
    int pst = set_permutation(h, perm);
    if (pst) return pst;
  }
  StagedPlanHost sp;
  build_staged_plan_host(h, P->h_rows, P->nblocks, P->maxlen, perm, sp);
  P->blk_nodes_max = sp.blk_nodes_max;
  P->strip_entries_per_pair = h->n2e.empty() ? 0.0 : (double)sp.total_real / (double)h->n2e.size();
  if (getenv("CGASM_DEBUG"))
    fprintf(stderr, "[cgasm] strip plan: %.3f entries per (row, element) pair, %lld padded entries, <= %d nodes per block (stride %d)%s\n",
            P->strip_entries_per_pair, sp.ptr[P->nblocks], sp.blk_nodes_max, sp.nl, sp.ok ? "" : " -- staged kernels not applicable");
  if (sp.ok) {
    const size_t total = (size_t)sp.ptr[P->nblocks], tail = (size_t)kStagedTailRows * kBR;
    CG_CUDA(cudaMalloc(&P->d_strip_local, sizeof(unsigned) * (total + tail)));
    CG_CUDA(cudaMemsetAsync(P->d_strip_local + total, 0, sizeof(unsigned) * tail, h->stream));
    for (size_t k = 0; k < sp.ent.size(); k++) {
      const int b0 = (int)k * sp.task_blocks;
      if (b0 >= P->nblocks || sp.ent[k].empty()) continue;
      CG_CUDA(cudaMemcpyAsync(P->d_strip_local + sp.ptr[b0], sp.ent[k].data(), sizeof(unsigned) * sp.ent[k].size(),
                              cudaMemcpyHostToDevice, h->stream));
    }
    CG_CUDA(cudaMalloc(&P->d_blk_nodes, sizeof(int) * sp.blk_nodes.size()));
    CG_CUDA(cudaMemcpyAsync(P->d_blk_nodes, sp.blk_nodes.data(), sizeof(int) * sp.blk_nodes.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaMalloc(&P->d_strip_local_ptr, sizeof(long long) * sp.ptr.size()));
    CG_CUDA(cudaMemcpyAsync(P->d_strip_local_ptr, sp.ptr.data(), sizeof(long long) * sp.ptr.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaMalloc(&P->d_own_local, sizeof(unsigned) * sp.own_local.size()));
    CG_CUDA(cudaMemcpyAsync(P->d_own_local, sp.own_local.data(), sizeof(unsigned) * sp.own_local.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaMalloc(&P->d_row_meta, sizeof(int) * sp.row_meta.size()));
    CG_CUDA(cudaMemcpyAsync(P->d_row_meta, sp.row_meta.data(), sizeof(int) * sp.row_meta.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
    P->nl = sp.nl;
    P->h_blk_nn = std::move(sp.blk_nn);
    P->h_blk_ml = std::move(sp.blk_ml);
    P->staged_ok = true;
  }
  if (!strip_staged_ok(h, true) || !strip_staged_ok(h, false)) return strip_build_global(h);
  return CGASM_OK;
}

void strip_free(GatherPlan* P) {
  for (GatherPlan::StagedClass& c : P->classes) {
    if (c.d_small) cudaFree(c.d_small);
    if (c.d_large) cudaFree(c.d_large);
  }
  P->classes.clear();
  P->h_blk_nn.clear();
  P->h_blk_ml.clear();
  if (P->d_strip_ptr) cudaFree(P->d_strip_ptr);
  if (P->d_strip) cudaFree(P->d_strip);
  if (P->d_strip_local_ptr) cudaFree(P->d_strip_local_ptr);
  if (P->d_blk_nodes) cudaFree(P->d_blk_nodes);
  if (P->d_strip_local) cudaFree(P->d_strip_local);
  if (P->d_own_local) cudaFree(P->d_own_local);
  if (P->d_row_meta) cudaFree(P->d_row_meta);
  P->d_row_meta = nullptr;
  P->d_strip_ptr = nullptr;
  P->d_strip = nullptr;
  P->d_strip_local_ptr = nullptr;
  P->d_blk_nodes = nullptr;
  P->d_strip_local = P->d_own_local = nullptr;
}

// ---- launch -------------------------------------------------------------------------------------------
bool strip_momentum_ok(const Handle* h, const MomentumArgs& A, bool want_ml) {
  (void)want_ml;
  const GatherPlan* P = h->gather;
  if (!P || !(P->d_strip || P->d_strip_local) || !strip_momentum_opts_ok(A)) return false;
  if (strip_extra_needed(A) && !(strip_staged_ok(h, true) && strip_extra_ok(h, A))) return false;
  // a full constant tensor needs the staged kernels
  return strip_staged_ok(h, true) || (P->d_strip && !strip_full_tensor(A.o.have_viscosity, A.o.viscosity_shape));
}

bool strip_advdiff_ok(const Handle* h, const AdvDiffArgs& A) {
  const GatherPlan* P = h->gather;
  if (!P || !(P->d_strip || P->d_strip_local) || !strip_advdiff_opts_ok(A)) return false;
  // a full constant tensor, absorption and sources need the staged kernels
  return strip_staged_ok(h, false) ||
         (P->d_strip && !strip_full_tensor(A.o.have_diffusivity, A.o.diffusivity_shape) && !strip_advdiff_needs_extra(A));
}

template <int DIM>
static int strip_momentum_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = sizeof(double) * (size_t)P->maxlen * kAS;
  if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "strip scatter: CSR rows too long for the shared-memory accumulator");
  const StripConsts c = consts_momentum(h, A);
  const StripPlanView v = plan_view(h);
  const int minb = getenv("CGASM_STRIP_MINB") ? atoi(getenv("CGASM_STRIP_MINB")) : 4;
  int st;
#define LAUNCH(N_, MINB_)                                                                                      \
  do {                                                                                                         \
    if ((st = strip_smem(strip_momentum_kernel<DIM, N_, MINB_>, smem))) return st;                             \
    strip_momentum_kernel<DIM, N_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(                               \
        c, v, h->d_rec3, h->d_rec1, h->d_rec2, (size_t)h->nnz, h->d_big_m, h->d_mom_rhs,                        \
        A.o.assemble_inverse_masslump ? h->d_masslump : nullptr);                                                                        \
  } while (0)
  if (minb >= 4) LAUNCH(DIM + 1, 4);
  else if (minb == 3) LAUNCH(DIM + 1, 3);
  else LAUNCH(DIM + 1, 2);
#undef LAUNCH
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int strip_momentum(Handle* h, const MomentumArgs& A) {
  h->mom_path = CGASM_PATH_STRIP;
  if (strip_staged_ok(h, true)) {
    h->mom_path = CGASM_PATH_STRIP_STAGED;
    // full absorption matrix: carried by the common kernel's own loop (strip_absorb.cu) where it fits
    const bool absorb = strip_extra_needed(A) && strip_absorb_ok(h, A);
    int st = absorb ? strip_absorb_momentum(h, A) : strip_staged_momentum(h, A);
    if (st == CGASM_OK && strip_extra_needed(A)) st = strip_extra(h, A, absorb);  // adds to the common result in place
    return st;
  }
  if (int js = halo_join(h)) return js;
  return h->dim == 3 ? strip_momentum_dim<3>(h, A) : strip_momentum_dim<2>(h, A);
}

template <int DIM>
static int strip_advdiff_dim(Handle* h, const AdvDiffArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = sizeof(double) * (size_t)P->maxlen * kAS;
  if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "strip scatter: CSR rows too long for the shared-memory accumulator");
  const StripConsts c = consts_advdiff(h, A);
  const StripPlanView v = plan_view(h);
  const int minb = getenv("CGASM_STRIP_MINB_ADV") ? atoi(getenv("CGASM_STRIP_MINB_ADV")) : 4;
  int st;
#define LAUNCH(N_, MINB_)                                                                                      \
  do {                                                                                                         \
    if ((st = strip_smem(strip_advdiff_kernel<DIM, N_, MINB_>, smem))) return st;                              \
    strip_advdiff_kernel<DIM, N_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(                                \
        c, v, h->d_rec0, h->d_rec1, h->d_adv_matrix, h->d_adv_rhs);                         \
  } while (0)
  if (minb >= 5) LAUNCH(DIM + 1, 5);
  else if (minb == 4) LAUNCH(DIM + 1, 4);
  else LAUNCH(DIM + 1, 3);
#undef LAUNCH
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int strip_advdiff(Handle* h, const AdvDiffArgs& A) {
  h->adv_path = CGASM_PATH_STRIP;
  if (strip_staged_ok(h, false)) {
    h->adv_path = CGASM_PATH_STRIP_STAGED;
    return strip_staged_advdiff(h, A);
  }
  if (int js = halo_join(h)) return js;
  return h->dim == 3 ? strip_advdiff_dim<3>(h, A) : strip_advdiff_dim<2>(h, A);
}

}  // namespace cgasm
