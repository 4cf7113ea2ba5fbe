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

}  // namespace cgasm
