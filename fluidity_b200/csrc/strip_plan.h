// strip_plan.h -- per-row "strip" orderings for the row-owner kernels in strip.cu (host side).
//
// A CSR row (node r) receives one contribution from every element incident to r (the pairs the
// reference visits through addto, femtools/Sparse_Tools.F90:2680-2703). The elements around r are
// the simplices of r's LINK (each element minus r): triangles in 3-D, edges in 2-D. The strip is a
// node sequence v_0, v_1, ... such that every link simplex appears as a window of `dim` consecutive
// nodes -- a generalised triangle strip over the link. A row thread then keeps a FIFO of `dim`
// nodes in registers, loads ONE node per entry, and the replacement pattern (always drop the
// oldest) is identical for every thread: no position selects, no divergence in the load path.
#pragma once
#include <cstdint>
#include <vector>

namespace cgasm {

struct StripEntry {
  int node;  // 0-based node to push
  int meta;  // bits 0-7: CSR slot of `node` inside row r; bit 8: the window ending here is an
             // element of r that has not been computed yet -> compute it
};
constexpr int kStripCompute = 0x100;

// Strip of row r. nd0: 0-based connectivity with stride 4; n2e_ptr/n2e: node -> element adjacency
// (ascending element ids); findrm/colm: 0-based sorted CSR rows. Deterministic in the relative
// order of the ids, so rows with the same local topology get the same compute pattern.
void build_strip_row(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm,
                     const int* colm, int r, std::vector<StripEntry>& out);
// The same greedy without the fixed-size local tables (O(m^2) per step): what build_strip_row falls back to
// for links that do not fit them, and the reference implementation the CPU tests compare it with.
void build_strip_row_generic(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm,
                             const int* colm, int r, std::vector<StripEntry>& out);

// The same on a canonical relabelling of the link by geometric keys (geokey[node], nullptr = build_strip_row).
void build_strip_row_keyed(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm,
                           const int64_t* geokey, int r, std::vector<StripEntry>& out);

// ---- staged strip plan (host side of strip_staged.cu) ---------------------------------------------------
constexpr unsigned kStagedCompute = 1u;
constexpr int kStagedNL[] = {128, 256, 384, 512, 768, 1024};  // chunk strides the staged kernels are instantiated for
constexpr int kStagedTailRows = 16;  // rows of padding behind the last block: the kernels read / prefetch ahead unguarded

struct StagedPlanHost {
  int nblocks = 0;
  int blk_nodes_max = 0, nl = 0;
  std::vector<int> blk_nn, blk_ml;     // per block: distinct nodes touched, longest CSR row (occupancy classes)
  bool ok = false;                     // false: the mesh does not fit the encoding (use the per-entry-fetch kernels)
  std::vector<long long> ptr;          // [nblocks+1] first entry of a block; degrees padded to a multiple of dim
  int task_blocks = 0;                 // blocks per chunk of `ent`
  std::vector<std::vector<unsigned>> ent;  // entries of blocks [k*task_blocks, (k+1)*task_blocks), block-interleaved
  std::vector<unsigned> own_local;     // [nblocks*block_rows]
  std::vector<int> row_meta;           // [nblocks*block_rows][4] = {row node (-1 = padding), first CSR entry of the row,
                                       //   row length | own slot << 16, own_local}: everything a row thread needs, one load
  std::vector<int> blk_nodes;          // [nblocks][nl] sorted distinct nodes of the block, -1 padded
  long long total_real = 0;            // strip entries before padding
};

struct Handle;
// rows: nblocks*block_rows node ids (-1 = padding), block_rows = kBR. Needs connectivity, adjacency and sparsity
// of the handle (host copies only).
// perm (empty, or n_nodes entries): the staged records are stored at perm[node]; the block node lists then hold and are
// sorted by those positions.
void build_staged_plan_host(const Handle* h, const std::vector<int>& rows, int nblocks, int maxlen,
                            const std::vector<int>& perm, StagedPlanHost& out);

}  // namespace cgasm
