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
//            entries carry block-LOCAL node indices (32-bit entries instead of 64-bit): the FIFO
//            buffers are filled by LDS.128 from a structure-of-16-byte-chunks layout (neighbouring
//            rows read neighbouring indices: conflict-free), nothing in the loop waits on DRAM
//            except the plan stream (three entries ahead in registers, its lines pulled into L2 ten ahead);
//   flush    when a node leaves the FIFO its accumulated entry goes to the row's slot AND into
//            rhs -= entry * oldu(node) (Momentum_CG.F90:1712,2346 is linear in the entries), with
//            oldu read from the staged records: no epilogue pass over colm;
//   write    dt*theta and the lumped mass on the diagonal (:1550, :1484-1486) are applied while the
//            warp streams its rows out.
#include "strip_common.cuh"

#include <cstdlib>

namespace cgasm {

struct StagedView {
  const int* __restrict__ rows;
  const long long* __restrict__ ptr;      // strip entries of the block (block-interleaved)
  const unsigned* __restrict__ ent;       // local index | slot << 16 | compute << 24
  const unsigned* __restrict__ own_local; // own node: local index | own slot << 16
  const int* __restrict__ blk_nodes;      // [nblocks][nl], -1 padded
  const int* __restrict__ findrm;
  int maxlen, lpr_shift;
  int nl;         // chunk stride (nodes) of the staged records
  int acc_bytes;  // bytes of the accumulator in front of them (multiple of 16)
};

constexpr unsigned kLocalCompute = 1u << 24;
// ptxas sinks the plan loads to about one step before their use whatever the source order says (it
// shortens the live range), which exposes a DRAM round trip per step: so the plan line of step
// j + kPlanAhead is pulled into L2 by a prefetch (no destination register, nothing to sink) and the
// sunk load then hits L2.
constexpr int kPlanAhead = 10;

__device__ __forceinline__ unsigned ldg_stream1(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}

// Copies the records of every node of the block: 16-byte chunk c of local node i at nodes[c*nl + i]
// (chunks 2q, 2q+1 = record q). If OLDU, chunk 4 = oldu(x, y) and the z components follow as a plain
// double array (an 8-byte read from a 16-byte-strided chunk would be a 2-way bank conflict).
template <int DIM, bool OLDU>
__device__ __forceinline__ void stage_nodes(const StagedView& P, int b, int t, double2* __restrict__ nodes,
                                            const double4* __restrict__ r0, const double4* __restrict__ r1,
                                            const double4* __restrict__ rO) {
  const int* ids = P.blk_nodes + (size_t)b * P.nl;  // fixed stride: no pointer load in front of the id loads
  double* oz = reinterpret_cast<double*>(nodes + 5 * P.nl);
  constexpr int U = 4;  // node ids of U rounds are requested together, then their copies are issued
  for (int i0 = t; i0 < P.nl; i0 += U * kBR) {
    int node[U];
#pragma unroll
    for (int u = 0; u < U; u++) node[u] = i0 + u * kBR < P.nl ? __ldg(ids + i0 + u * kBR) : -1;
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (node[u] < 0) continue;
      const int i = i0 + u * kBR;
      const double2* s0 = reinterpret_cast<const double2*>(r0 + node[u]);
      const double2* s1 = reinterpret_cast<const double2*>(r1 + node[u]);
      cp_async16(nodes + 0 * P.nl + i, s0);
      cp_async16(nodes + 1 * P.nl + i, s0 + 1);
      cp_async16(nodes + 2 * P.nl + i, s1);
      cp_async16(nodes + 3 * P.nl + i, s1 + 1);
      if constexpr (OLDU) {
        const double2* s2 = reinterpret_cast<const double2*>(rO + node[u]);
        cp_async16(nodes + 4 * P.nl + i, s2);
        if constexpr (DIM == 3) cp_async8(oz + i, s2 + 1);
      }
    }
  }
}

// Shared-memory reads of the staged records are volatile asm with a memory clobber: they must be ISSUED
// where they are written (one step ahead of their use) -- left to the compiler they sink to the first
// use and every step pays the LDS latency in its prologue (ncu: short_scoreboard on the install DADDs).
__device__ __forceinline__ double2 lds128(unsigned sa) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sa) : "memory");
  return v;
}
__device__ __forceinline__ double lds64(unsigned sa) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa) : "memory");
  return v;
}

// chunk c of local node li (16-byte chunks; nsa = shared address of the staged records)
__device__ __forceinline__ unsigned chunk_sa(unsigned nsa, int nl, int c, int li) { return nsa + (unsigned)(c * nl + li) * 16u; }

template <int DIM>
__device__ __forceinline__ void load_rec(unsigned nsa, int nl, int rec, int li, double (&v)[DIM], double& s) {
  const double2 a = lds128(chunk_sa(nsa, nl, 2 * rec, li));
  const double2 b = lds128(chunk_sa(nsa, nl, 2 * rec + 1, li));
  v[0] = a.x;
  v[1] = a.y;
  if constexpr (DIM == 3) v[2] = b.x;
  s = b.y;
}

// oldu of local node li: chunk 4 = {x, y}; the z components sit behind chunk 4 as a plain double array
template <int DIM>
__device__ __forceinline__ void load_oldu(unsigned nsa, int nl, int li, double (&o)[DIM]) {
  const double2 a = lds128(chunk_sa(nsa, nl, 4, li));
  o[0] = a.x;
  o[1] = a.y;
  if constexpr (DIM == 3) o[2] = lds64(nsa + (unsigned)(5 * nl) * 16u + (unsigned)li * 8u);
}

// ---- momentum -----------------------------------------------------------------------------------------
// One strip entry. Program order = issue order (all memory asm is volatile): flush the evicted buffer
// (slot accumulator and rhs -= entry * oldu of the evicted node), request the records of entry j + PD and
// plan entry j + PD + 3, then install and compute entry j.
template <int DIM, int N, int QC, bool FULLV, bool PF>
__device__ __forceinline__ void smom_step(MomState<DIM, N>& s, double (&rh)[DIM], const StripConsts& k_, int j, int deg,
                                          const unsigned* __restrict__ p, unsigned& pq0, unsigned& pq1, unsigned& pq2,
                                          const unsigned pad, double* __restrict__ acc_t, unsigned nsa, int nl) {
  constexpr int PD = N - DIM;
  constexpr int QE = (QC + PD) % N;  // holds entry j - DIM: evicted now, refilled with entry j + PD
  const unsigned en = pq0;
  pq0 = pq1;
  pq1 = pq2;
  {
    const unsigned m = (unsigned)s.meta[QE];
    double on[DIM];
    load_oldu<DIM>(nsa, nl, (int)(m & 0xffffu), on);
    double* sl = acc_t + ((m >> 16) & 0xffu) * kAS;
    const double a = s.A[QE];
    *sl += a;
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-a, on[d], rh[d]);
    s.A[QE] = 0.0;
  }
  const int li = (int)(en & 0xffffu);
  load_rec<DIM>(nsa, nl, 0, li, s.X[QE], s.B[QE]);
  load_rec<DIM>(nsa, nl, 1, li, s.U[QE], s.R[QE]);
  s.meta[QE] = (int)en;
  pq2 = (j + PD + 3 < deg) ? ldg_stream1(p + (long long)(j + PD + 3) * kBR) : pad;
  if (PF && j + kPlanAhead < deg) prefetch_l2(p + (long long)(j + kPlanAhead) * kBR);
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if ((unsigned)s.meta[QC] & kLocalCompute) mom_compute<DIM, N, QC, FULLV>(s, k_);
}

template <int DIM, int N, int Q, bool FULLV, bool PF>
struct SMomUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(MomState<DIM, N>& s, double (&rh)[DIM], const StripConsts& k_, int j0,
                                             Args&&... args) {
    smom_step<DIM, N, Q, FULLV, PF>(s, rh, k_, j0 + Q, args...);
    if constexpr (Q + 1 < N) SMomUnroll<DIM, N, Q + 1, FULLV, PF>::run(s, rh, k_, j0, args...);
  }
};

// rows of the warp -> the dim identical diagonal blocks: dt*theta * entry (+ lumped mass on the diagonal)
template <int DIM>
__device__ __forceinline__ void write_rows_scaled(const double* __restrict__ acc, int t, int my_s0, int my_len, int my_own,
                                                  double my_mass, double dtt, int lpr_shift, size_t nnz,
                                                  double* __restrict__ out) {
  const int lane = t & 31, wbase = t & ~31;
  const int lpr = 1 << lpr_shift, rpi = 32 >> lpr_shift;
  const int sub = lane >> lpr_shift, sl = lane & (lpr - 1);
  for (int rr = 0; rr < 32; rr += rpi) {
    const int src = rr + sub;
    const int s0r = __shfl_sync(0xffffffffu, my_s0, src);
    const int lr = __shfl_sync(0xffffffffu, my_len, src);
    const int own = __shfl_sync(0xffffffffu, my_own, src);
    const double mass = __shfl_sync(0xffffffffu, my_mass, src);
    for (int ss = sl; ss < lr; ss += lpr) {
      const double v = fma(dtt, acc[ss * kAS + wbase + src], ss == own ? mass : 0.0);
#pragma unroll
      for (int d = 0; d < DIM; d++) __stcs(out + (size_t)d * nnz + s0r + ss, v);
    }
  }
}

template <int DIM, int N, int MINB, bool FULLV, bool PF>
__global__ void __launch_bounds__(kBR, MINB)
staged_momentum_kernel(const StripConsts k_, const StagedView P, const double4* __restrict__ rX,
                       const double4* __restrict__ rU, const double4* __restrict__ rO, size_t nnz,
                       double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  constexpr int PD = N - DIM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  double2* nodes = reinterpret_cast<double2*>(smem_raw + P.acc_bytes);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(nodes);
  const int b = blockIdx.x, t = threadIdx.x, nl = P.nl;
  stage_nodes<DIM, true>(P, b, t, nodes, rX, rU, rO);
  const int r = P.rows[b * kBR + t];
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);  // a multiple of N
  const unsigned* p = P.ent + base + t;
  double* acc_t = acc + t;
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  const unsigned pad = P.own_local[b * kBR + t];
  const int own = (int)((pad >> 16) & 0xffu), own_li = (int)(pad & 0xffffu);
  unsigned first[PD > 0 ? PD : 1];
#pragma unroll
  for (int q = 0; q < PD; q++) first[q] = q < deg ? ldg_stream1(p + (long long)q * kBR) : pad;
  unsigned pq0 = PD < deg ? ldg_stream1(p + (long long)PD * kBR) : pad;
  unsigned pq1 = PD + 1 < deg ? ldg_stream1(p + (long long)(PD + 1) * kBR) : pad;
  unsigned pq2 = PD + 2 < deg ? ldg_stream1(p + (long long)(PD + 2) * kBR) : pad;
  if constexpr (PF) {
#pragma unroll
    for (int q = PD + 3; q < kPlanAhead; q++)
      if (q < deg) prefetch_l2(p + (long long)q * kBR);
  }
  cp_async_commit_wait_all();
  __syncthreads();
  MomState<DIM, N> s;
  load_rec<DIM>(nsa, nl, 0, own_li, s.X0, s.b0);
  load_rec<DIM>(nsa, nl, 1, own_li, s.U0, s.rho0);
  s.a0 = s.msum = s.nbsum = 0.0;
  double rh[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) rh[d] = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.R[q] = s.B[q] = s.A[q] = 0.0;
    s.meta[q] = (int)pad;
  }
#pragma unroll
  for (int q = 0; q < PD; q++) {
    const int li = (int)(first[q] & 0xffffu);
    load_rec<DIM>(nsa, nl, 0, li, s.X[q], s.B[q]);
    load_rec<DIM>(nsa, nl, 1, li, s.U[q], s.R[q]);
    s.meta[q] = (int)first[q];
  }
  for (int j0 = 0; j0 < deg; j0 += N)
    SMomUnroll<DIM, N, 0, FULLV, PF>::run(s, rh, k_, j0, deg, p, pq0, pq1, pq2, pad, acc_t, nsa, nl);
  // drain the FIFO, then the diagonal (the row's own node never leaves)
#pragma unroll
  for (int q = 0; q < N; q++) {
    const unsigned m = (unsigned)s.meta[q];
    acc_t[((m >> 16) & 0xffu) * kAS] += s.A[q];
    double o[DIM];
    load_oldu<DIM>(nsa, nl, (int)(m & 0xffffu), o);
#pragma unroll
    for (int d = 0; d < DIM; d++) rh[d] = fma(-s.A[q], o[d], rh[d]);
  }
  acc_t[own * kAS] += s.a0;
  int my_s0 = 0, my_len = 0;
  if (r >= 0) {
    my_s0 = P.findrm[r];
    my_len = P.findrm[r + 1] - my_s0;
    double ou[DIM];
    load_oldu<DIM>(nsa, nl, own_li, ou);
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      rhs[(size_t)DIM * r + d] = fma(-s.a0, ou[d], fma(k_.grav[d], s.nbsum, rh[d]));
      if (masslump) masslump[(size_t)DIM * r + d] = s.msum;
    }
  }
  __syncwarp();
  write_rows_scaled<DIM>(acc, t, my_s0, my_len, own, s.msum * k_.mass_on, k_.dtt, P.lpr_shift, nnz, big_m);
}

// ---- tracer -------------------------------------------------------------------------------------------
template <int DIM, int N, int QC, bool FULLV, bool PF>
__device__ __forceinline__ void sadv_step(AdvState<DIM, N>& s, const StripConsts& k_, int j, int deg,
                                          const unsigned* __restrict__ p, unsigned& pq0, unsigned& pq1, unsigned& pq2,
                                          const unsigned pad, double* __restrict__ acc_t, unsigned nsa, int nl) {
  constexpr int PD = N - DIM;
  constexpr int QE = (QC + PD) % N;
  const unsigned en = pq0;
  pq0 = pq1;
  pq1 = pq2;
  {
    double* sl = acc_t + (((unsigned)s.meta[QE] >> 16) & 0xffu) * kAS;
    *sl += fma(k_.dtt, s.A[QE], k_.mPo * s.C[QE]);
    s.A[QE] = 0.0;
    s.C[QE] = 0.0;
  }
  const int li = (int)(en & 0xffffu);
  double unused;
  load_rec<DIM>(nsa, nl, 0, li, s.X[QE], s.T[QE]);
  load_rec<DIM>(nsa, nl, 1, li, s.U[QE], unused);
  s.meta[QE] = (int)en;
  pq2 = (j + PD + 3 < deg) ? ldg_stream1(p + (long long)(j + PD + 3) * kBR) : pad;
  if (PF && j + kPlanAhead < deg) prefetch_l2(p + (long long)(j + kPlanAhead) * kBR);
#pragma unroll
  for (int a = 0; a < DIM; a++) s.X[QC][a] -= s.X0[a];
  if ((unsigned)s.meta[QC] & kLocalCompute) adv_compute<DIM, N, QC, FULLV>(s, k_);
}

template <int DIM, int N, int Q, bool FULLV, bool PF>
struct SAdvUnroll {
  template <class... Args>
  static __device__ __forceinline__ void run(AdvState<DIM, N>& s, const StripConsts& k_, int j0, Args&&... args) {
    sadv_step<DIM, N, Q, FULLV, PF>(s, k_, j0 + Q, args...);
    if constexpr (Q + 1 < N) SAdvUnroll<DIM, N, Q + 1, FULLV, PF>::run(s, k_, j0, args...);
  }
};

template <int DIM, int N, int MINB, bool FULLV, bool PF>
__global__ void __launch_bounds__(kBR, MINB)
staged_advdiff_kernel(const StripConsts k_, const StagedView P, const double4* __restrict__ rX,
                      const double4* __restrict__ rU, double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int PD = N - DIM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);
  double2* nodes = reinterpret_cast<double2*>(smem_raw + P.acc_bytes);
  const unsigned nsa = (unsigned)__cvta_generic_to_shared(nodes);
  const int b = blockIdx.x, t = threadIdx.x, nl = P.nl;
  stage_nodes<DIM, false>(P, b, t, nodes, rX, rU, nullptr);
  const int r = P.rows[b * kBR + t];
  const long long base = P.ptr[b];
  const int deg = (int)((P.ptr[b + 1] - base) / kBR);
  const unsigned* p = P.ent + base + t;
  double* acc_t = acc + t;
  for (int q = 0; q < P.maxlen; q++) acc_t[q * kAS] = 0.0;
  const unsigned pad = P.own_local[b * kBR + t];
  const int own = (int)((pad >> 16) & 0xffu), own_li = (int)(pad & 0xffffu);
  unsigned first[PD > 0 ? PD : 1];
#pragma unroll
  for (int q = 0; q < PD; q++) first[q] = q < deg ? ldg_stream1(p + (long long)q * kBR) : pad;
  unsigned pq0 = PD < deg ? ldg_stream1(p + (long long)PD * kBR) : pad;
  unsigned pq1 = PD + 1 < deg ? ldg_stream1(p + (long long)(PD + 1) * kBR) : pad;
  unsigned pq2 = PD + 2 < deg ? ldg_stream1(p + (long long)(PD + 2) * kBR) : pad;
  if constexpr (PF) {
#pragma unroll
    for (int q = PD + 3; q < kPlanAhead; q++)
      if (q < deg) prefetch_l2(p + (long long)q * kBR);
  }
  cp_async_commit_wait_all();
  __syncthreads();
  AdvState<DIM, N> s;
  double unused;
  load_rec<DIM>(nsa, nl, 0, own_li, s.X0, s.T0);
  load_rec<DIM>(nsa, nl, 1, own_li, s.U0, unused);
  s.a0 = s.c0 = s.rhs = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) {
#pragma unroll
    for (int a = 0; a < DIM; a++) s.X[q][a] = s.U[q][a] = 0.0;
    s.T[q] = s.A[q] = s.C[q] = 0.0;
    s.meta[q] = (int)pad;
  }
#pragma unroll
  for (int q = 0; q < PD; q++) {
    const int li = (int)(first[q] & 0xffffu);
    load_rec<DIM>(nsa, nl, 0, li, s.X[q], s.T[q]);
    load_rec<DIM>(nsa, nl, 1, li, s.U[q], unused);
    s.meta[q] = (int)first[q];
  }
  for (int j0 = 0; j0 < deg; j0 += N) SAdvUnroll<DIM, N, 0, FULLV, PF>::run(s, k_, j0, deg, p, pq0, pq1, pq2, pad, acc_t, nsa, nl);
#pragma unroll
  for (int q = 0; q < N; q++)
    acc_t[(((unsigned)s.meta[q] >> 16) & 0xffu) * kAS] += fma(k_.dtt, s.A[q], k_.mPo * s.C[q]);
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

// ---- launch -------------------------------------------------------------------------------------------
static size_t acc_bytes_of(const GatherPlan* P) {
  return (sizeof(double) * (size_t)P->maxlen * kAS + 15) & ~(size_t)15;
}
static int nl_of(const GatherPlan* P) { return (P->blk_nodes_max + 7) & ~7; }
static size_t staged_smem(const GatherPlan* P, bool momentum) { return acc_bytes_of(P) + (size_t)nl_of(P) * (momentum ? 88 : 64); }

bool strip_staged_ok(const Handle* h, bool momentum) {
  const GatherPlan* P = h->gather;
  if (!P || !P->d_strip_local || getenv("CGASM_STRIP_GLOBAL")) return false;
  return staged_smem(P, momentum) <= 100 * 1024;  // at least two blocks per SM, else the per-entry kernels
}

static StagedView staged_view(const Handle* h) {
  const GatherPlan* P = h->gather;
  StagedView v;
  v.rows = P->d_rows;
  v.ptr = P->d_strip_local_ptr;
  v.ent = P->d_strip_local;
  v.own_local = P->d_own_local;
  v.blk_nodes = P->d_blk_nodes;
  v.findrm = h->d_findrm;
  v.maxlen = P->maxlen;
  int sh = 0;
  while ((1 << sh) < P->maxlen && sh < 5) sh++;
  v.lpr_shift = sh;
  v.nl = nl_of(P);
  v.acc_bytes = (int)acc_bytes_of(P);
  return v;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int DIM>
static int staged_momentum_dim(Handle* h, const MomentumArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = staged_smem(P, true);
  const StripConsts c = consts_momentum(h, A);
  const StagedView v = staged_view(h);
  // tuning switches (defaults = best measured on S3, profiles/r1_kernel_history.md)
  const int minb = env_int("CGASM_STRIP_MINB", 4);
  const bool pf = env_int("CGASM_STRIP_PF", 1) != 0;
  const bool fullv = strip_full_tensor(A.o.have_viscosity, A.o.viscosity_shape);
  double* ml = A.o.assemble_inverse_masslump ? h->d_masslump : nullptr;
  int st;
#define LAUNCH(N_, MINB_, FULLV_, PF_)                                                                          \
  do {                                                                                                          \
    if ((st = strip_smem(staged_momentum_kernel<DIM, N_, MINB_, FULLV_, PF_>, smem))) return st;                \
    staged_momentum_kernel<DIM, N_, MINB_, FULLV_, PF_><<<P->nblocks, kBR, smem, h->stream>>>(                  \
        c, v, h->d_rec3, h->d_rec1, h->d_rec2, (size_t)h->nnz, h->d_big_m, h->d_mom_rhs, ml);                    \
  } while (0)
#define LAUNCH_V(N_, MINB_, PF_)                    \
  do {                                              \
    if (fullv) LAUNCH(N_, MINB_, true, PF_);        \
    else LAUNCH(N_, MINB_, false, PF_);             \
  } while (0)
  if (minb >= 4) {
    if (pf) LAUNCH_V(DIM, 4, true);
    else LAUNCH_V(DIM, 4, false);
  } else {
    if (pf) LAUNCH_V(DIM, 3, true);
    else LAUNCH_V(DIM, 3, false);
  }
#undef LAUNCH_V
#undef LAUNCH
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int strip_staged_momentum(Handle* h, const MomentumArgs& A) {
  return h->dim == 3 ? staged_momentum_dim<3>(h, A) : staged_momentum_dim<2>(h, A);
}

template <int DIM>
static int staged_advdiff_dim(Handle* h, const AdvDiffArgs& A) {
  GatherPlan* P = h->gather;
  const size_t smem = staged_smem(P, false);
  const StripConsts c = consts_advdiff(h, A);
  const StagedView v = staged_view(h);
  const int minb = env_int("CGASM_STRIP_MINB_ADV", 4);
  const bool pf = env_int("CGASM_STRIP_PF_ADV", 1) != 0;
  const bool fullv = strip_full_tensor(A.o.have_diffusivity, A.o.diffusivity_shape);
  int st;
#define LAUNCH(N_, MINB_, FULLV_, PF_)                                                                          \
  do {                                                                                                          \
    if ((st = strip_smem(staged_advdiff_kernel<DIM, N_, MINB_, FULLV_, PF_>, smem))) return st;                 \
    staged_advdiff_kernel<DIM, N_, MINB_, FULLV_, PF_><<<P->nblocks, kBR, smem, h->stream>>>(                   \
        c, v, h->d_rec0, h->d_rec1, h->d_adv_matrix, h->d_adv_rhs);                                              \
  } while (0)
#define LAUNCH_V(N_, MINB_, PF_)                    \
  do {                                              \
    if (fullv) LAUNCH(N_, MINB_, true, PF_);        \
    else LAUNCH(N_, MINB_, false, PF_);             \
  } while (0)
  if (minb >= 4) {
    if (pf) LAUNCH_V(DIM, 4, true);
    else LAUNCH_V(DIM, 4, false);
  } else {
    if (pf) LAUNCH_V(DIM, 3, true);
    else LAUNCH_V(DIM, 3, false);
  }
#undef LAUNCH_V
#undef LAUNCH
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int strip_staged_advdiff(Handle* h, const AdvDiffArgs& A) {
  return h->dim == 3 ? staged_advdiff_dim<3>(h, A) : staged_advdiff_dim<2>(h, A);
}

}  // namespace cgasm
