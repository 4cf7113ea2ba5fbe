// gather_plan.h -- row-block plans shared by the row-owner kernels (gather.cu, strip.cu).
#pragma once
#include "cgasm_internal.h"

#include <vector>

namespace cgasm {

constexpr int kBR = 128;  // rows (= threads) per gather block
constexpr int kAS = kBR + 1;  // stride (doubles) between slots of the STRIP accumulator: odd, so that the
                              // write-out (one row spread over consecutive lanes) is conflict-free

struct GatherPlan {
  long long serial = 0;            // unique per built plan (plans derived from this one remember it)
  int nblocks = 0;
  int maxlen = 0;                  // longest CSR row
  int* d_rows = nullptr;           // [nblocks*kBR] node of each row slot, -1 = padding
  long long* d_block_ptr = nullptr;  // [nblocks+1] first plan entry of a block
  uint2* d_pairs = nullptr;        // block-interleaved: entry k of thread t at ptr + k*kBR + t
                                   //   .x = element*4 + local row (0xFFFFFFFF = none), .y = 4 x 8-bit slot
  int4* d_pair_nodes = nullptr;    // same indexing, ONE 16-byte load per pair: {n1, n2, n3, slots} = the element's
                                   //   other nodes in rotated order (the row's own node is rotated node 0;
                                   //   n1 < 0 = none) and the 4 x 8-bit CSR slots of rotated nodes 0..3
  long long n_entries = 0;
  // walk plan (single-pass kernels for the common option set): per row the incident elements are
  // ordered as a face-adjacent walk around the node, so consecutive elements share two of their
  // three other nodes; an entry loads ONE node into one of the three register positions.
  //   .x = node to load (-1 = padding), .y = position (bits 0-1) | compute flag (bit 2) | CSR slot << 8
  long long* d_walk_ptr = nullptr;  // [nblocks+1]
  int2* d_walk = nullptr;           // block-interleaved like d_pairs
  unsigned char* d_own_slot = nullptr;  // [nblocks*kBR] slot of the diagonal inside the row
  long long n_walk = 0;
  double walk_entries_per_pair = 0.0;
  // strip plan (strip.cu / strip_plan.h): block-interleaved {node, slot | compute << 8}, block degree
  // padded to a multiple of dim + 1 (= register buffers of the per-entry-fetch kernels) with no-op entries
  std::vector<int> h_rows;           // host copy of d_rows
  long long* d_strip_ptr = nullptr;  // [nblocks+1]
  int2* d_strip = nullptr;
  long long n_strip = 0;
  double strip_entries_per_pair = 0.0;
  // staged strip plan (strip_staged.cu, built by strip_plan.cpp build_staged_plan_host): the distinct nodes each
  // row block touches (sorted) and the strip entries re-expressed with block-local node indices, pre-scaled to
  // the byte offsets the kernel adds to its shared-memory bases:
  //   bit 0 compute, bits 4-15 local node index (<< 4 = byte offset of its 16-byte chunk), bits 16-31 CSR slot * kAS
  //   (<< 3 = byte offset of the slot in the thread's accumulator column)
  int* d_blk_nodes = nullptr;        // [nblocks][nl], -1 padded
  long long* d_strip_local_ptr = nullptr;  // [nblocks+1] into d_strip_local (degrees padded to a multiple of dim)
  unsigned* d_strip_local = nullptr; // block-interleaved like d_strip, kStagedTailRows rows of padding at the end
  unsigned* d_own_local = nullptr;   // [nblocks*kBR] own node in the same encoding (compute = 0)
  int4* d_row_meta = nullptr;        // [nblocks*kBR] {row node, first CSR entry, length | own slot << 16, own_local}
  int blk_nodes_max = 0;
  int nl = 0;                        // chunk stride of the staged records (one of kStagedNL)
  // Occupancy classes (strip_staged.cuh staged_classes): on an unstructured mesh ONE crowded block sets nl and maxlen --
  // and with them the shared memory of every block. The blocks that fit a smaller chunk stride and a shorter
  // accumulator are launched as a class of their own with more blocks per SM.
  std::vector<int> h_blk_nn, h_blk_ml;  // per block: distinct nodes touched, longest CSR row
  struct StagedClass {
    int bytes_per_node = 0, nacc = 0;  // the kernel family the split was made for
    int nl_small = 0, ml_small = 0;    // 0: no split pays
    int n_small = 0, n_large = 0;
    int* d_small = nullptr;            // block ids
    int* d_large = nullptr;
  };
  std::vector<StagedClass> classes;
  bool staged_ok = false;            // the mesh fits the staged encoding
  double* d_stage = nullptr;       // staging buffer (grown on demand)
  size_t stage_doubles = 0;
};

}  // namespace cgasm
