// strip_plan.cpp -- host-side construction of the per-row strips (strip_plan.h). No CUDA here, so
// the CPU test-suite can exercise it through cgasm_debug_strip_plan.
#include "strip_plan.h"

#include <algorithm>
#include <cstdint>
#include <cstring>

#include "cgasm_internal.h"

namespace cgasm {

namespace {

struct Link {
  int m = 0;              // simplices of the link (= elements incident to the row's node)
  int w = 0;              // nodes per link simplex (= dim)
  std::vector<int> oth;   // m x 3, sorted ascending, unused = -1
  std::vector<char> todo;
  bool has(int t, int v) const { return oth[3 * t] == v || oth[3 * t + 1] == v || oth[3 * t + 2] == v; }
};

// index of the not-yet-computed link simplex made of exactly these w nodes, or -1
int find_todo(const Link& L, const int* win) {
  if (L.w == 3 && (win[0] == win[1] || win[1] == win[2] || win[0] == win[2])) return -1;
  if (L.w == 2 && win[0] == win[1]) return -1;
  for (int t = 0; t < L.m; t++) {
    if (!L.todo[t]) continue;
    bool all = true;
    for (int i = 0; i < L.w; i++) all = all && L.has(t, win[i]);
    if (all) return t;
  }
  return -1;
}

struct Seq {
  std::vector<int> node;
  std::vector<char> comp;
  int remaining = 0;
  // push one node; the window ending at it is computed if it is a pending link simplex
  void push(Link& L, int v) {
    node.push_back(v);
    char c = 0;
    const int n = (int)node.size();
    if (n >= L.w) {
      const int t = find_todo(L, node.data() + n - L.w);
      if (t >= 0) {
        L.todo[t] = 0;
        remaining--;
        c = 1;
      }
    }
    comp.push_back(c);
  }
};

int third_node(const Link& L, int t, int x, int y) {
  for (int i = 0; i < 3; i++) {
    const int v = L.oth[3 * t + i];
    if (v != x && v != y) return v;
  }
  return -1;
}

// 3-D: greedy generalised triangle strip over the link, FIFO of three nodes (a, b, c).
void strip3(Link& L, int start, const int* perm, Seq& S) {
  std::fill(L.todo.begin(), L.todo.end(), 1);
  S.node.clear();
  S.comp.clear();
  S.remaining = L.m;
  for (int i = 0; i < 3; i++) S.push(L, L.oth[3 * start + perm[i]]);
  while (S.remaining > 0) {
    const int n = (int)S.node.size();
    const int a = S.node[n - 3], b = S.node[n - 2], c = S.node[n - 1];
    // (1) drop the oldest: a pending triangle on the edge (b, c); prefer the one whose new edge
    //     (c, d) has the most pending triangles left, so the strip can go on
    int best = -1, best_score = -1;
    for (int t = 0; t < L.m; t++) {
      if (!L.todo[t] || !L.has(t, b) || !L.has(t, c)) continue;
      const int d = third_node(L, t, b, c);
      int score = 0;
      for (int u = 0; u < L.m; u++) score += (L.todo[u] && u != t && L.has(u, c) && L.has(u, d));
      if (score > best_score) {
        best_score = score;
        best = t;
      }
    }
    if (best >= 0) {
      S.push(L, third_node(L, best, b, c));
      continue;
    }
    // (2) swap: a pending triangle on (a, c): push a again, then its third node
    int t2 = -1;
    for (int t = 0; t < L.m && t2 < 0; t++)
      if (L.todo[t] && L.has(t, a) && L.has(t, c)) t2 = t;
    if (t2 >= 0) {
      const int d = third_node(L, t2, a, c);
      S.push(L, a);
      S.push(L, d);
      continue;
    }
    // (3) a pending triangle on (a, b): push a, b, then its third node
    for (int t = 0; t < L.m && t2 < 0; t++)
      if (L.todo[t] && L.has(t, a) && L.has(t, b)) t2 = t;
    if (t2 >= 0) {
      const int d = third_node(L, t2, a, b);
      S.push(L, a);
      S.push(L, b);
      S.push(L, d);
      continue;
    }
    // (4) jump to the pending triangle sharing most nodes with the window (at most one here)
    int bt = -1, bs = -1;
    for (int t = 0; t < L.m; t++) {
      if (!L.todo[t]) continue;
      const int sh = (int)L.has(t, a) + (int)L.has(t, b) + (int)L.has(t, c);
      if (sh > bs) {
        bs = sh;
        bt = t;
      }
    }
    const int* o = &L.oth[3 * bt];
    if (bs == 1) {
      const int x = L.has(bt, c) ? c : (L.has(bt, b) ? b : a);
      int rest[2], q = 0;
      for (int i = 0; i < 3; i++)
        if (o[i] != x) rest[q++] = o[i];
      if (x != c) S.push(L, x);
      S.push(L, rest[0]);
      S.push(L, rest[1]);
    } else {
      S.push(L, o[0]);
      S.push(L, o[1]);
      S.push(L, o[2]);
    }
    if (L.todo[bt]) {  // cannot happen (the last push completes the window); guard against loops
      L.todo[bt] = 0;
      S.remaining--;
    }
  }
}

// 2-D: the link is a cycle (interior node) or a set of paths (boundary): walk it edge by edge.
void strip2(Link& L, Seq& S) {
  std::fill(L.todo.begin(), L.todo.end(), 1);
  S.node.clear();
  S.comp.clear();
  S.remaining = L.m;
  auto pending_degree = [&](int v) {
    int d = 0;
    for (int t = 0; t < L.m; t++) d += (L.todo[t] && L.has(t, v));
    return d;
  };
  while (S.remaining > 0) {
    int next = -1;
    if (!S.node.empty()) {
      const int c = S.node.back();
      for (int t = 0; t < L.m && next < 0; t++)
        if (L.todo[t] && L.has(t, c)) next = (L.oth[3 * t] == c) ? L.oth[3 * t + 1] : L.oth[3 * t];
    }
    if (next >= 0) {
      S.push(L, next);
      continue;
    }
    // start a new path at an end (a node with one pending edge) if there is one
    int st = -1, sv = -1;
    for (int t = 0; t < L.m && sv < 0; t++) {
      if (!L.todo[t]) continue;
      if (st < 0) st = t;
      for (int i = 0; i < 2; i++)
        if (pending_degree(L.oth[3 * t + i]) == 1) {
          st = t;
          sv = L.oth[3 * t + i];
          break;
        }
    }
    if (sv < 0) sv = L.oth[3 * st];
    const int other = (L.oth[3 * st] == sv) ? L.oth[3 * st + 1] : L.oth[3 * st];
    S.push(L, sv);
    S.push(L, other);
  }
}

}  // namespace

void build_strip_row(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm,
                     int r, std::vector<StripEntry>& out) {
  out.clear();
  const int64_t k0 = n2e_ptr[r];
  Link L;
  L.m = (int)(n2e_ptr[r + 1] - k0);
  L.w = loc - 1;
  if (L.m == 0) return;
  L.oth.assign((size_t)3 * L.m, -1);
  L.todo.assign((size_t)L.m, 1);
  for (int k = 0; k < L.m; k++) {
    const int* nd = nd0 + (size_t)4 * n2e[k0 + k];
    int q = 0;
    for (int i = 0; i < loc; i++)
      if (nd[i] != r) L.oth[3 * k + q++] = nd[i];
    std::sort(&L.oth[3 * k], &L.oth[3 * k] + L.w);
  }
  Seq best;
  if (L.w == 2) {
    strip2(L, best);
  } else {
    // start from a triangle with the fewest edge neighbours (an end of the fan on a boundary);
    // all six orders of its nodes are tried and the shortest strip kept
    int start = 0, start_nb = 1 << 30;
    for (int t = 0; t < L.m; t++) {
      int nb = 0;
      for (int u = 0; u < L.m; u++)
        if (u != t) nb += ((int)L.has(u, L.oth[3 * t]) + (int)L.has(u, L.oth[3 * t + 1]) + (int)L.has(u, L.oth[3 * t + 2])) == 2;
      if (nb < start_nb) {
        start_nb = nb;
        start = t;
      }
    }
    static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    Seq S;
    for (int p = 0; p < 6; p++) {
      strip3(L, start, perms[p], S);
      if (p == 0 || S.node.size() < best.node.size()) best = S;
    }
  }
  const int s0 = findrm[r], s1 = findrm[r + 1];
  out.resize(best.node.size());
  for (size_t k = 0; k < best.node.size(); k++) {
    const int* cb = colm + s0;
    const int slot = (int)(std::lower_bound(cb, colm + s1, best.node[k]) - cb);
    out[k].node = best.node[k];
    out[k].meta = (slot & 0xff) | (best.comp[k] ? kStripCompute : 0);
  }
}

}  // namespace cgasm

// ---- diagnostics ABI: the strips of every row of a mesh, built on the host (no GPU needed) ----------
extern "C" int cgasm_strip_plan_host(int loc, int n_nodes, int n_elements, const int* ndglno, long long* row_ptr,
                                     int* entries, long long capacity, long long* needed) {
  using namespace cgasm;
  if ((loc != 3 && loc != 4) || n_nodes <= 0 || n_elements <= 0 || !ndglno || !row_ptr || !needed)
    CG_FAIL(CGASM_EARG, "cgasm_strip_plan_host: bad argument");
  std::vector<int> nd0((size_t)4 * n_elements, -1);
  for (int e = 0; e < n_elements; e++)
    for (int i = 0; i < loc; i++) {
      const int v = ndglno[(size_t)loc * e + i] - 1;
      if (v < 0 || v >= n_nodes) CG_FAIL(CGASM_EARG, "cgasm_strip_plan_host: node id out of range");
      nd0[(size_t)4 * e + i] = v;
    }
  std::vector<int64_t> n2e_ptr;
  std::vector<int> n2e, findrm, colm;
  build_node_to_element(n_nodes, n_elements, loc, nd0.data(), n2e_ptr, n2e);
  build_sparsity(n_nodes, n_elements, loc, nd0.data(), n2e_ptr, n2e, findrm, colm);
  std::vector<StripEntry> row;
  long long total = 0;
  row_ptr[0] = 0;
  for (int r = 0; r < n_nodes; r++) {
    build_strip_row(loc, nd0.data(), n2e_ptr.data(), n2e.data(), findrm.data(), colm.data(), r, row);
    for (size_t k = 0; k < row.size(); k++, total++)
      if (entries && total < capacity) {
        entries[2 * total] = row[k].node + 1;
        entries[2 * total + 1] = row[k].meta;
      }
    row_ptr[r + 1] = total;
  }
  *needed = total;
  return CGASM_OK;
}

// ---- diagnostics ABI: the row blocks of the row-owner variants, built on the host (no GPU needed) -----
extern "C" int cgasm_row_blocks_host(int dim, int n_nodes, int n_elements, const int* ndglno, const double* X,
                                     int block_rows, int* rows, int capacity_blocks, int* nblocks, double* lattice_scale) {
  using namespace cgasm;
  const int loc = dim + 1;
  if ((dim != 2 && dim != 3) || n_nodes <= 0 || n_elements <= 0 || !ndglno || !X || !nblocks || block_rows < 1)
    CG_FAIL(CGASM_EARG, "cgasm_row_blocks_host: bad argument");
  Handle h;
  h.dim = dim;
  h.loc = loc;
  h.n_nodes = n_nodes;
  h.n_elements = n_elements;
  h.h_nd0.assign((size_t)4 * n_elements, -1);
  for (int e = 0; e < n_elements; e++)
    for (int i = 0; i < loc; i++) {
      const int v = ndglno[(size_t)loc * e + i] - 1;
      if (v < 0 || v >= n_nodes) CG_FAIL(CGASM_EARG, "cgasm_row_blocks_host: node id out of range");
      h.h_nd0[(size_t)4 * e + i] = v;
    }
  h.h_X.assign(X, X + (size_t)dim * n_nodes);
  build_node_to_element(n_nodes, n_elements, loc, h.h_nd0.data(), h.n2e_ptr, h.n2e);
  build_sparsity(n_nodes, n_elements, loc, h.h_nd0.data(), h.n2e_ptr, h.n2e, h.h_findrm, h.h_colm);
  std::vector<int> order, r;
  MortonFrame F;
  morton_order(&h, order, F);
  *nblocks = form_row_blocks(&h, order, F, block_rows, r);
  if (lattice_scale)
    for (int a = 0; a < dim; a++) lattice_scale[a] = F.scale[a];
  if (rows && *nblocks <= capacity_blocks)
    for (size_t q = 0; q < r.size(); q++) rows[q] = r[q] >= 0 ? r[q] + 1 : 0;
  return CGASM_OK;
}
