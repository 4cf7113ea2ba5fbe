// strip_staged.cuh -- device helpers and launch plumbing shared by the staged STRIP kernels
// (strip_staged.cu: the two element loops; strip_extra.cu: the additive momentum pass).
#pragma once
#include "strip_common.cuh"
#include "strip_plan.h"

namespace cgasm {

struct StagedView {
  const int* __restrict__ rows;
  const long long* __restrict__ ptr;      // strip entries of the block (block-interleaved)
  const unsigned* __restrict__ ent;       // kStagedCompute | local index << 4 | (slot * kAS) << 16
  const unsigned* __restrict__ own_local; // own node in the same encoding
  const int* __restrict__ blk_nodes;      // [nblocks][NL], -1 padded
  const int* __restrict__ findrm;
  const int* __restrict__ blocks;         // the row blocks this launch works on (nullptr: all, in order)
  int maxlen, lpr_shift;
  int acc_bytes;  // bytes of the accumulator in front of the staged records (multiple of 16)
};

// ptxas sinks the plan loads to about one step before their use whatever the source order says (it
// shortens the live range), which exposes a DRAM round trip per step: so the plan line of step
// j + kPlanAhead is pulled into L2 by a prefetch (no destination register, nothing to sink) and the
// sunk load then hits L2.
constexpr int kPlanAhead = 10;
static_assert(kPlanAhead + 3 <= kStagedTailRows, "the plan's tail padding must cover the read-ahead");

__device__ __forceinline__ unsigned ldg_stream1(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Shared-memory accesses of the loop are volatile asm with a memory clobber: they must be ISSUED
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
__device__ __forceinline__ void sts64(unsigned sa, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(sa), "d"(v) : "memory");
}

// Layout of the staged records: 16-byte chunk c of local node i at nsa + (c * NL + i) * 16.
//   chunks 0,1 = record A {x, y | z, s}    chunks 2,3 = record B {x, y | z, s}
//   momentum: chunk 4 = oldu {x, y}, then a plain double array oldu z at chunk 5's place (an 8-byte read
//             from a 16-byte-strided chunk would be a 2-way bank conflict)
//   tracer with absorption / source: chunk 4 = {absorption, source}
template <int DIM, int NL>
__device__ __forceinline__ void load_rec(unsigned nb, int rec, double (&v)[DIM], double& s) {
  const double2 a = lds128(nb + (unsigned)(2 * rec * NL * 16));
  const double2 b = lds128(nb + (unsigned)((2 * rec + 1) * NL * 16));
  v[0] = a.x;
  v[1] = a.y;
  if constexpr (DIM == 3) v[2] = b.x;
  s = b.y;
}

// oldu of the node at byte offset noff (= local index << 4) of the staged chunks
template <int DIM, int NL>
__device__ __forceinline__ void load_oldu(unsigned nsa, unsigned noff, double (&o)[DIM]) {
  const double2 a = lds128(nsa + noff + (unsigned)(4 * NL * 16));
  o[0] = a.x;
  o[1] = a.y;
  if constexpr (DIM == 3) o[2] = lds64(nsa + (noff >> 1) + (unsigned)(5 * NL * 16));
}

static inline size_t staged_acc_bytes(const GatherPlan* P, int nblocks_acc) {
  return (sizeof(double) * (size_t)nblocks_acc * P->maxlen * kAS + 15) & ~(size_t)15;
}

static inline StagedView staged_view(const Handle* h, int nblocks_acc = 1) {
  const GatherPlan* P = h->gather;
  StagedView v;
  v.rows = P->d_rows;
  v.ptr = P->d_strip_local_ptr;
  v.ent = P->d_strip_local;
  v.own_local = P->d_own_local;
  v.blk_nodes = P->d_blk_nodes;
  v.findrm = h->d_findrm;
  v.blocks = nullptr;
  v.maxlen = P->maxlen;
  int sh = 0;
  while ((1 << sh) < P->maxlen && sh < 5) sh++;
  v.lpr_shift = sh;
  v.acc_bytes = (int)staged_acc_bytes(P, nblocks_acc);
  return v;
}

// the chunk strides of 2-D meshes stay small (a 128-row brick of a triangulation touches ~200 nodes)
#define CGASM_FOR_NL(X)                                        \
  do {                                                         \
    switch (P->nl) {                                           \
      case 128: X(128); break;                                 \
      case 256: X(256); break;                                 \
      case 384: X(384); break;                                 \
      case 512: X(512); break;                                 \
      case 768: X(768); break;                                 \
      default: X(1024); break;                                 \
    }                                                          \
  } while (0)

}  // namespace cgasm
