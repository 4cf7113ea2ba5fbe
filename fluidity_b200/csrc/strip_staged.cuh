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
  const int4* __restrict__ row_meta;      // {row node, first CSR entry, length | own slot << 16, own_local}
  int nblocks, ahead;                     // blocks in the plan; how far ahead a block pulls its successor's metadata into L2
  int nl_stride;                          // ints between the node lists of consecutive blocks (>= the kernel's NL)
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


// ---- block prologue / epilogue shared by the staged kernels -----------------------------------------------------
// ncu (profiles/r2_kernel_history.md): a quarter of a warp's life went into FOUR dependent DRAM round trips around the
// loop -- node ids, then (rows, ptr, own_local), then the first plan entries, and findrm in the epilogue. Now every
// independent load of the block is issued back to back at the top (ids, the row's int4 meta record that replaces rows /
// own_local / findrm, the block's plan pointers), the accumulator is zeroed while they fly, and the block pulls the
// same lines of the block `ahead` positions further on into L2, so its successor's first round trip is an L2 hit.
__device__ __forceinline__ int ldg_nc_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int4 ldg_nc_v4(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long ldg_nc_s64(const long long* p) {
  long long v;
  asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Staging map: a PAIR of lanes copies one 32-byte node record -- lane 2i the first 16 bytes of the record of local node
// i, lane 2i+1 the second -- so a warp-wide cp.async reads 16 records = 512 contiguous bytes wherever the block's
// (sorted) node list runs through consecutive ids. ncu (profiles/r2_kernel_history.md #32): with one lane per record and
// two copies per lane every LDGSTS.128 touched 32 half-used sectors and took 16 shared-memory wavefronts where 2-4
// suffice; staging was 14 % of the kernel's wavefronts on the busiest pipe (L1 data, 67 %).
template <int NL>
struct BlockIds {
  static constexpr int PER = 2 * NL / kBR;
  int node[PER];  // node of local index (t >> 1) + v * (kBR / 2)
};

template <int NL>
__device__ __forceinline__ void issue_block_ids(const StagedView& P, int b, int t, BlockIds<NL>& ids) {
  const int* p = P.blk_nodes + (size_t)b * P.nl_stride + (t >> 1);  // no pointer load in front of the id loads
#pragma unroll
  for (int v = 0; v < BlockIds<NL>::PER; v++) ids.node[v] = ldg_nc_s32(p + v * (kBR / 2));
}

// half h of the record of `node` -> chunk (chunk0 + h) of local node i
template <int NL>
__device__ __forceinline__ void stage_record(unsigned nsa, int chunk0, unsigned i, int h, const double4* __restrict__ rec, int node) {
  cp_async16(nsa + (unsigned)((chunk0 + h) * NL * 16) + i * 16u, reinterpret_cast<const double2*>(rec + node) + h);
}
// {a, b | c, -} record: a, b -> chunk `chunk` (16 bytes), c -> the plain double array at byte offset arr_off
template <int NL, bool THIRD>
__device__ __forceinline__ void stage_record_3(unsigned nsa, int chunk, unsigned arr_off, unsigned i, int h,
                                               const double4* __restrict__ rec, int node) {
  if (h == 0) cp_async16(nsa + (unsigned)(chunk * NL * 16) + i * 16u, reinterpret_cast<const double2*>(rec + node));
  else if (THIRD) cp_async8(nsa + arr_off + i * 8u, reinterpret_cast<const double*>(rec + node) + 2);
}

template <int NL>
__device__ __forceinline__ void prefetch_next_block(const StagedView& P, int b, int t) {
  if (P.blocks) return;
  const int nb = b + P.ahead;
  if (nb >= P.nblocks) return;
  if (t < 16) prefetch_l2(P.row_meta + (size_t)nb * kBR + t * 8);                      // 128 rows x 16 bytes
  else if (t < 16 + NL / 32) prefetch_l2(P.blk_nodes + (size_t)nb * P.nl_stride + (t - 16) * 32);  // NL ids
  else if (t == 16 + NL / 32) prefetch_l2(P.ptr + nb);
}

// Strip entries this WARP walks: its longest row's, in the row record's bits 24-31 in units of DIM entries (the plan pads
// every row of the block to the block's longest with no-op entries; a warp whose rows are shorter skips them).
template <int DIM>
__device__ __forceinline__ int warp_trip_count(int meta_z) {
  return (int)__reduce_max_sync(0xffffffffu, (unsigned)meta_z >> 24) * DIM;
}

// per-row table for the write-out: {first CSR entry, length | own slot << 16, value for the diagonal}, 16 bytes per row
__device__ __forceinline__ void row_table_store(unsigned tbl_sa, int t, int s0, int lenown, double diag) {
  asm volatile("st.shared.v2.s32 [%0], {%1,%2};" ::"r"(tbl_sa + (unsigned)t * 16u), "r"(s0), "r"(lenown) : "memory");
  sts64(tbl_sa + (unsigned)t * 16u + 8u, diag);
}

// rows of the warp -> global memory, LPR = 1 << lpr_shift lanes per row: out[d*nnz + s0 + ss] = scale * acc[ss] (+ diag
// on the own slot), for d < NOUT. The row table replaces four shuffles per row by one broadcast LDS.128.
template <int NOUT>
__device__ __forceinline__ void write_rows_table(const double* __restrict__ acc, unsigned tbl_sa, int t, double scale,
                                                 int lpr_shift, size_t nnz, double* __restrict__ out) {
  const int lane = t & 31, wbase = t & ~31;
  const int lpr = 1 << lpr_shift, rpi = 32 >> lpr_shift;
  const int sub = lane >> lpr_shift, sl = lane & (lpr - 1);
  for (int rr = 0; rr < 32; rr += rpi) {
    const int src = wbase + rr + sub;
    int s0r, lo;
    double diag;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(s0r), "=r"(lo) : "r"(tbl_sa + (unsigned)src * 16u) : "memory");
    diag = lds64(tbl_sa + (unsigned)src * 16u + 8u);
    const int lr = lo & 0xff, own = (lo >> 16) & 0xff;
    for (int ss = sl; ss < lr; ss += lpr) {
      const double v = fma(scale, acc[ss * kAS + src], ss == own ? diag : 0.0);
#pragma unroll
      for (int d = 0; d < NOUT; d++) __stcs(out + (size_t)d * nnz + s0r + ss, v);
    }
  }
}

static inline size_t staged_acc_bytes_of(int maxlen, int nblocks_acc) {
  return (sizeof(double) * (size_t)nblocks_acc * maxlen * kAS + 15) & ~(size_t)15;
}
static inline size_t staged_acc_bytes(const GatherPlan* P, int nblocks_acc) { return staged_acc_bytes_of(P->maxlen, nblocks_acc); }

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
  v.row_meta = P->d_row_meta;
  v.nblocks = P->nblocks;
  v.nl_stride = P->nl;
  v.ahead = 4 * 148;  // about the number of blocks resident on the chip
  v.maxlen = P->maxlen;
  int sh = 0;
  while ((1 << sh) < P->maxlen && sh < 5) sh++;
  v.lpr_shift = sh;
  v.acc_bytes = (int)staged_acc_bytes(P, nblocks_acc);
  return v;
}

// Occupancy classes of a kernel family (bytes of staged records per node, accumulator columns per thread): nullptr if the
// whole plan already runs at the blocks per SM a smaller class would reach, or if too few blocks would gain. The small
// class keeps the plan's node lists (stride nl) and entries -- local indices do not depend on the chunk stride -- and
// only needs nl_small >= its blocks' node counts and ml_small >= their longest rows.
static inline const GatherPlan::StagedClass* staged_classes(Handle* h, int bytes_per_node, int nacc) {
  GatherPlan* P = h->gather;
  if (P->h_blk_nn.empty() || getenv("CGASM_STRIP_NOCLASSES")) return nullptr;
  for (const GatherPlan::StagedClass& c : P->classes)
    if (c.bytes_per_node == bytes_per_node && c.nacc == nacc) return c.nl_small ? &c : nullptr;
  GatherPlan::StagedClass c;
  c.bytes_per_node = bytes_per_node;
  c.nacc = nacc;
  const size_t sm_bytes = 227 * 1024, reserved = 1024, table = kBR * 16;
  auto per_sm = [&](int nl, int ml) {
    const size_t b = ((sizeof(double) * (size_t)nacc * ml * kAS + 15) & ~(size_t)15) + (size_t)nl * bytes_per_node + table + reserved;
    return (int)std::min<size_t>(4, sm_bytes / b);  // the kernels are built for at most four blocks per SM
  };
  const int now = per_sm(P->nl, P->maxlen);
  int best_cover = 0;
  for (int nl : kStagedNL) {
    if (nl > P->nl) break;
    for (int k = now + 1; k <= 4; k++) {
      const long long room = (long long)(sm_bytes / k) - (long long)reserved - (long long)table - (long long)nl * bytes_per_node;
      const int ml = (int)std::min<long long>(P->maxlen, room / (long long)(sizeof(double) * nacc * kAS));
      if (ml < 4 || (nl == P->nl && ml == P->maxlen)) continue;
      int cover = 0;
      for (int b = 0; b < P->nblocks; b++) cover += P->h_blk_nn[b] <= nl && P->h_blk_ml[b] <= ml;
      // weigh a class by the blocks it holds times the residency it buys
      const int score = (int)((long long)cover * (k - now) / k);
      if (2 * cover >= P->nblocks && score > best_cover) {
        best_cover = score;
        c.nl_small = nl;
        c.ml_small = ml;
      }
    }
  }
  if (c.nl_small) {
    std::vector<int> small, large;
    for (int b = 0; b < P->nblocks; b++)
      (P->h_blk_nn[b] <= c.nl_small && P->h_blk_ml[b] <= c.ml_small ? small : large).push_back(b);
    c.n_small = (int)small.size();
    c.n_large = (int)large.size();
    if (cudaMalloc(&c.d_small, sizeof(int) * std::max<size_t>(1, small.size())) != cudaSuccess ||
        cudaMalloc(&c.d_large, sizeof(int) * std::max<size_t>(1, large.size())) != cudaSuccess ||
        cg_upload(c.d_small, small.data(), sizeof(int) * small.size()) != cudaSuccess ||
        cg_upload(c.d_large, large.data(), sizeof(int) * large.size()) != cudaSuccess) {
      cudaGetLastError();
      if (c.d_small) cudaFree(c.d_small);
      if (c.d_large) cudaFree(c.d_large);
      c.d_small = c.d_large = nullptr;
      c.nl_small = 0;
    }
    if (getenv("CGASM_VERBOSE"))
      fprintf(stderr, "cgasm: occupancy class (%d B/node, %d accumulators): %d of %d blocks at nl %d / maxlen %d (%d blocks per SM, plan: nl %d / maxlen %d, %d per SM)\n",
              bytes_per_node, nacc, c.n_small, P->nblocks, c.nl_small, c.ml_small, per_sm(c.nl_small, c.ml_small), P->nl, P->maxlen, now);
  }
  P->classes.push_back(c);
  return P->classes.back().nl_small ? &P->classes.back() : nullptr;
}

// the chunk strides of 2-D meshes stay small (a 128-row brick of a triangulation touches ~200 nodes)
#define CGASM_FOR_NL(X) CGASM_FOR_NL_OF(P->nl, X)
#define CGASM_FOR_NL_OF(NLV, X)                                \
  do {                                                         \
    switch (NLV) {                                             \
      case 128: X(128); break;                                 \
      case 256: X(256); break;                                 \
      case 384: X(384); break;                                 \
      case 512: X(512); break;                                 \
      case 768: X(768); break;                                 \
      default: X(1024); break;                                 \
    }                                                          \
  } while (0)

}  // namespace cgasm
