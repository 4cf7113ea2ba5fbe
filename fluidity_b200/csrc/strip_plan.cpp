// strip_plan.cpp -- host-side construction of the per-row strips (strip_plan.h). No CUDA here, so
// the CPU test-suite can exercise it through cgasm_debug_strip_plan.
#include "strip_plan.h"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_set>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <omp.h>

#include "cgasm_internal.h"
#include "gather_plan.h"

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


// ---- fast 3-D path ------------------------------------------------------------------------------------
// Same greedy, same tie-breaking (candidates are always scanned in ascending link-triangle order), but on
// local vertex ids with a vertex -> incident-triangle table: "the pending triangles on edge (b, c)" is a
// scan of the ~5 triangles around c instead of all m, so a step costs O(1) and a row ~2 us instead of
// ~30 us (the plan of a 100 M-tet mesh: seconds instead of a minute).
constexpr int kFastMaxTri = 96, kFastMaxVert = 64, kFastMaxInc = 24;

struct FastLink {
  int m = 0, nv = 0;
  int vid[kFastMaxVert];                  // local -> global node id, ascending
  unsigned char tri[kFastMaxTri][3];      // local ids, ascending
  unsigned char ninc[kFastMaxVert];
  unsigned char inc[kFastMaxVert][kFastMaxInc];  // incident triangles of a vertex, ascending
  bool todo[kFastMaxTri];
  bool has(int t, int v) const { return tri[t][0] == v || tri[t][1] == v || tri[t][2] == v; }
  int third(int t, int x, int y) const {
    for (int i = 0; i < 3; i++)
      if (tri[t][i] != x && tri[t][i] != y) return tri[t][i];
    return -1;
  }
};

struct FastSeq {
  unsigned char node[4 * kFastMaxTri + 8];
  bool comp[4 * kFastMaxTri + 8];
  int n = 0, remaining = 0;
  void push(FastLink& L, int v) {
    node[n] = (unsigned char)v;
    bool c = false;
    if (n >= 2) {
      const int a = node[n - 2], b = node[n - 1];
      if (a != b && b != v && a != v)
        for (int q = 0; q < L.ninc[v]; q++) {
          const int t = L.inc[v][q];
          if (L.todo[t] && L.has(t, a) && L.has(t, b)) {
            L.todo[t] = false;
            remaining--;
            c = true;
            break;
          }
        }
    }
    comp[n++] = c;
  }
};

// false if the link does not fit the fixed-size tables (the generic path handles it). The link's vertices are
// the row's columns minus r itself (already sorted and distinct), so local ids come from the CSR row.
bool fast_link(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm, int r,
               FastLink& L, const int64_t* geokey = nullptr) {
  const int64_t k0 = n2e_ptr[r];
  const int m = (int)(n2e_ptr[r + 1] - k0);
  if (m > kFastMaxTri) return false;
  const int len = findrm[r + 1] - findrm[r];
  const int nv = len - 1;
  if (nv > kFastMaxVert || nv < 3) return false;
  const int* row = colm + findrm[r];
  {
    int q = 0;
    for (int k = 0; k < len; k++)
      if (row[k] != r) {
        if (q == nv) return false;  // r is not in its own row
        L.vid[q++] = row[k];
      }
    if (q != nv) return false;
  }
  L.m = m;
  L.nv = nv;
  // Canonical labels (build_strip_row_keyed): vertices by geometric key instead of id, triangles sorted by their labels
  int byid[kFastMaxVert];
  unsigned char label[kFastMaxVert];  // position in the id-sorted row -> label
  if (geokey) {
    std::copy(L.vid, L.vid + nv, byid);
    unsigned char ord[kFastMaxVert];
    for (int k = 0; k < nv; k++) ord[k] = (unsigned char)k;
    std::sort(ord, ord + nv, [&](unsigned char a, unsigned char b) {
      const int64_t ka = geokey[byid[a]], kb = geokey[byid[b]];
      return ka != kb ? ka < kb : a < b;
    });
    for (int k = 0; k < nv; k++) {
      label[ord[k]] = (unsigned char)k;
      L.vid[k] = byid[ord[k]];
    }
  }
  const int* ids = geokey ? byid : L.vid;
  for (int k = 0; k < m; k++) {
    const int* nd = nd0 + (size_t)4 * n2e[k0 + k];
    unsigned char* t = L.tri[k];
    int q = 0;
    for (int i = 0; i < loc; i++) {
      if (nd[i] == r) continue;
      if (q == 3) return false;
      const int* f = std::lower_bound(ids, ids + nv, nd[i]);
      if (f == ids + nv || *f != nd[i]) return false;  // the sparsity does not hold this element pair
      t[q++] = geokey ? label[f - ids] : (unsigned char)(f - ids);
    }
    if (q != 3) return false;  // degenerate element (repeated node)
    if (t[0] > t[1]) std::swap(t[0], t[1]);
    if (t[1] > t[2]) std::swap(t[1], t[2]);
    if (t[0] > t[1]) std::swap(t[0], t[1]);
    if (t[0] == t[1] || t[1] == t[2]) return false;
  }
  if (geokey) {  // m is ~24: insertion sort of the label triples
    for (int a = 1; a < m; a++) {
      unsigned char cur[3] = {L.tri[a][0], L.tri[a][1], L.tri[a][2]};
      int b = a - 1;
      while (b >= 0 && (L.tri[b][0] > cur[0] || (L.tri[b][0] == cur[0] && (L.tri[b][1] > cur[1] ||
                                                (L.tri[b][1] == cur[1] && L.tri[b][2] > cur[2]))))) {
        L.tri[b + 1][0] = L.tri[b][0];
        L.tri[b + 1][1] = L.tri[b][1];
        L.tri[b + 1][2] = L.tri[b][2];
        b--;
      }
      L.tri[b + 1][0] = cur[0];
      L.tri[b + 1][1] = cur[1];
      L.tri[b + 1][2] = cur[2];
    }
  }
  std::fill(L.ninc, L.ninc + nv, 0);
  for (int k = 0; k < m; k++) {
    const unsigned char* t = L.tri[k];
    for (int i = 0; i < 3; i++) {
      if (L.ninc[t[i]] == kFastMaxInc) return false;
      L.inc[t[i]][L.ninc[t[i]]++] = (unsigned char)k;
    }
  }
  return true;
}

void fast_strip3(FastLink& L, int start, const int* perm, FastSeq& S) {
  std::fill(L.todo, L.todo + L.m, true);
  S.n = 0;
  S.remaining = L.m;
  for (int i = 0; i < 3; i++) S.push(L, L.tri[start][perm[i]]);
  while (S.remaining > 0) {
    const int a = S.node[S.n - 3], b = S.node[S.n - 2], c = S.node[S.n - 1];
    // (1) a pending triangle on (b, c), preferring the one whose new edge (c, d) has most pending triangles
    int best = -1, best_score = -1;
    for (int q = 0; q < L.ninc[c]; q++) {
      const int t = L.inc[c][q];
      if (!L.todo[t] || !L.has(t, b)) continue;
      const int d = L.third(t, b, c);
      int score = 0;
      for (int q2 = 0; q2 < L.ninc[c]; q2++) {
        const int u = L.inc[c][q2];
        score += (L.todo[u] && u != t && L.has(u, d));
      }
      if (score > best_score) {
        best_score = score;
        best = t;
      }
    }
    if (best >= 0) {
      S.push(L, L.third(best, b, c));
      continue;
    }
    // (2) swap: a pending triangle on (a, c)
    int t2 = -1;
    for (int q = 0; q < L.ninc[c] && t2 < 0; q++) {
      const int t = L.inc[c][q];
      if (L.todo[t] && L.has(t, a)) t2 = t;
    }
    if (t2 >= 0) {
      const int d = L.third(t2, a, c);
      S.push(L, a);
      S.push(L, d);
      continue;
    }
    // (3) a pending triangle on (a, b)
    for (int q = 0; q < L.ninc[b] && t2 < 0; q++) {
      const int t = L.inc[b][q];
      if (L.todo[t] && L.has(t, a)) t2 = t;
    }
    if (t2 >= 0) {
      const int d = L.third(t2, a, b);
      S.push(L, a);
      S.push(L, b);
      S.push(L, d);
      continue;
    }
    // (4) jump to the pending triangle sharing most nodes with the window
    int bt = -1, bs = -1;
    for (int t = 0; t < L.m; t++) {
      if (!L.todo[t]) continue;
      const int sh = (int)L.has(t, a) + (int)L.has(t, b) + (int)L.has(t, c);
      if (sh > bs) {
        bs = sh;
        bt = t;
      }
    }
    const unsigned char* o = L.tri[bt];
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
    if (L.todo[bt]) {
      L.todo[bt] = false;
      S.remaining--;
    }
  }
}

// ---- bitmask flavour of the fast path (links of at most 64 triangles: every mesh of decent quality) ------
// Triangle sets are 64-bit masks: "pending triangles on edge (b, c)" = todo & vm[b] & vm[c]; candidates are
// visited by ascending index (ctz), so the result is the generic algorithm's, step for step.
struct MaskSeq {
  unsigned char node[4 * 64 + 8];
  bool comp[4 * 64 + 8];
  int n = 0;
};

struct MaskLink {
  int m = 0;
  const unsigned char (*tri)[3] = nullptr;
  uint64_t vm[kFastMaxVert];
  int third(int t, int x, int y) const {
    for (int i = 0; i < 3; i++)
      if (tri[t][i] != x && tri[t][i] != y) return tri[t][i];
    return -1;
  }
  bool has(int t, int v) const { return (vm[v] >> t) & 1; }
};

inline void mask_push(const MaskLink& L, uint64_t& todo, MaskSeq& S, int v) {
  S.node[S.n] = (unsigned char)v;
  bool c = false;
  if (S.n >= 2) {
    const int a = S.node[S.n - 2], b = S.node[S.n - 1];
    if (a != b && b != v && a != v) {
      const uint64_t w = todo & L.vm[a] & L.vm[b] & L.vm[v];
      if (w) {
        todo &= ~(w & (~w + 1));  // lowest set bit = first pending triangle in ascending order
        c = true;
      }
    }
  }
  S.comp[S.n++] = c;
}

void mask_strip3(const MaskLink& L, int start, const int* perm, MaskSeq& S) {
  uint64_t todo = L.m == 64 ? ~0ull : ((1ull << L.m) - 1);
  S.n = 0;
  for (int i = 0; i < 3; i++) mask_push(L, todo, S, L.tri[start][perm[i]]);
  while (todo) {
    const int a = S.node[S.n - 3], b = S.node[S.n - 2], c = S.node[S.n - 1];
    uint64_t cand = todo & L.vm[b] & L.vm[c];
    if (cand) {
      int best = -1, best_score = -1;
      while (cand) {
        const int t = __builtin_ctzll(cand);
        cand &= cand - 1;
        const int d = L.third(t, b, c);
        const int score = __builtin_popcountll(todo & L.vm[c] & L.vm[d] & ~(1ull << t));
        if (score > best_score) {
          best_score = score;
          best = t;
        }
      }
      mask_push(L, todo, S, L.third(best, b, c));
      continue;
    }
    uint64_t w = todo & L.vm[a] & L.vm[c];
    if (w) {
      const int d = L.third(__builtin_ctzll(w), a, c);
      mask_push(L, todo, S, a);
      mask_push(L, todo, S, d);
      continue;
    }
    w = todo & L.vm[a] & L.vm[b];
    if (w) {
      const int d = L.third(__builtin_ctzll(w), a, b);
      mask_push(L, todo, S, a);
      mask_push(L, todo, S, b);
      mask_push(L, todo, S, d);
      continue;
    }
    int bt = -1, bs = -1;
    for (uint64_t q = todo; q; q &= q - 1) {
      const int t = __builtin_ctzll(q);
      const int sh = (int)L.has(t, a) + (int)L.has(t, b) + (int)L.has(t, c);
      if (sh > bs) {
        bs = sh;
        bt = t;
      }
    }
    const unsigned char* o = L.tri[bt];
    if (bs == 1) {
      const int x = L.has(bt, c) ? c : (L.has(bt, b) ? b : a);
      int rest[2], q = 0;
      for (int i = 0; i < 3; i++)
        if (o[i] != x) rest[q++] = o[i];
      if (x != c) mask_push(L, todo, S, x);
      mask_push(L, todo, S, rest[0]);
      mask_push(L, todo, S, rest[1]);
    } else {
      mask_push(L, todo, S, o[0]);
      mask_push(L, todo, S, o[1]);
      mask_push(L, todo, S, o[2]);
    }
    todo &= ~(1ull << bt);
  }
}

// ---- strip search: fewer pushes than the greedy ---------------------------------------------------------------
// The kernels walk a row's strip in trips of `dim` entries, every non-computing push still pays its shared-memory reads
// and the flush, and the greedy needs 33 pushes for the 24 elements around an interior node of a Kuhn mesh where 30
// are enough (three bands of eight triangles, found by exhaustive search: 29 is infeasible). On meshes whose rows fall
// into a few congruence classes (strip_search_prescan) the strip of every class is improved once by a bounded
// branch-and-bound over {computing push | one-push restart that keeps the newest node | fresh triangle in any of its
// six orders}: depth-first from every (first triangle, order) root with a growing per-root budget -- a third of the
// roots of the Kuhn link reach 30 within ~12 000 states while others exhaust millions -- pruned by
// pushes + pending + (components of the pending triangles' dual graph that miss the window edge) > limit and by a
// table of visited (window, pending set) states. The target is the next multiple of 3 below the greedy's length:
// anything in between saves no trip. Results are shared between threads (one search per class per process).
std::atomic<int> g_strip_search{0};
std::atomic<long> g_strip_search_states{0};  // states left for this plan build (all classes together)  // set per plan build by strip_search_prescan (CGASM_STRIP_SEARCH=0/1 overrides)

struct StripSearch {
  const MaskLink* L = nullptr;
  int m = 0, limit = 0;
  long states = 0, budget = 0, spent = 0;
  uint64_t adj[64];
  unsigned char seq[4 * 64 + 8];
  int n = 0;
  unsigned char best[4 * 64 + 8];
  int best_n = 0;
  struct Ent {
    uint64_t todo;
    uint32_t tag;  // epoch << 20 | a << 14 | b << 8 | pushes
  };
  static constexpr int kTabBits = 18;
  std::vector<Ent> tab;
  uint32_t epoch = 0;

  void init(const MaskLink& M) {
    L = &M;
    m = M.m;
    for (int i = 0; i < m; i++) {
      uint64_t w = 0;
      for (int e = 0; e < 3; e++) w |= M.vm[M.tri[i][e]] & M.vm[M.tri[i][(e + 1) % 3]];
      adj[i] = w & ~(1ull << i);
    }
    if (tab.empty()) tab.assign((size_t)1 << kTabBits, Ent{0, 0});
  }
  void next_epoch() {
    if (++epoch >= (1u << 12)) {
      epoch = 1;
      std::fill(tab.begin(), tab.end(), Ent{0, 0});
    }
  }
  // components of the dual graph of the pending triangles that do not touch edge (a, b): one non-computing push each at least
  int stranded(uint64_t todo, int a, int b) const {
    const uint64_t onedge = L->vm[a] & L->vm[b];
    int comps = 0;
    while (todo) {
      uint64_t c = todo & (~todo + 1), f = c;
      while (f) {
        uint64_t nx = 0;
        for (uint64_t q = f; q; q &= q - 1) nx |= adj[__builtin_ctzll(q)];
        nx &= todo & ~c;
        c |= nx;
        f = nx;
      }
      todo &= ~c;
      comps += (c & onedge) == 0;
    }
    return comps;
  }
  bool dfs(int a, int b, uint64_t todo, int used) {
    if (!todo) {
      memcpy(best, seq, (size_t)n);
      best_n = n;
      return true;
    }
    const int rem = __builtin_popcountll(todo);
    if (used + rem > limit || states > budget) return false;
    if (used + rem + stranded(todo, a, b) > limit) return false;
    ++states;
    {
      const uint64_t hsh = (todo ^ (todo >> 29) ^ ((uint64_t)(a * 64 + b) << 40)) * 0x9e3779b97f4a7c15ull;
      Ent& e = tab[hsh >> (64 - kTabBits)];
      const uint32_t key = epoch << 20 | (uint32_t)a << 14 | (uint32_t)b << 8;
      if (e.todo == todo && (e.tag & ~0xffu) == key && (int)(e.tag & 0xffu) <= used) return false;
      e.todo = todo;
      e.tag = key | (uint32_t)used;
    }
    const MaskLink& M = *L;
    {  // computing pushes: the triangle whose new edge has more pending triangles first
      int cand[2], sc[2], nc = 0;
      for (uint64_t c = todo & M.vm[a] & M.vm[b]; c && nc < 2; c &= c - 1) {
        const int t = __builtin_ctzll(c), d = M.third(t, a, b);
        cand[nc] = t;
        sc[nc++] = __builtin_popcountll(todo & M.vm[b] & M.vm[d] & ~(1ull << t));
      }
      if (nc == 2 && sc[1] > sc[0]) std::swap(cand[0], cand[1]);
      for (int i = 0; i < nc; i++) {
        const int t = cand[i], d = M.third(t, a, b);
        seq[n++] = (unsigned char)d;
        if (dfs(b, d, todo & ~(1ull << t), used + 1)) return true;
        n--;
      }
    }
    if (used + rem + 1 > limit) return false;
    for (uint64_t c = todo & M.vm[b] & ~M.vm[a]; c; c &= c - 1) {  // restart keeping b: push x, then d computes {b, x, d}
      const int t = __builtin_ctzll(c);
      int o[2], q = 0;
      for (int i = 0; i < 3; i++)
        if (M.tri[t][i] != b) o[q++] = M.tri[t][i];
      for (int s2 = 0; s2 < 2; s2++) {
        seq[n++] = (unsigned char)o[s2];
        seq[n++] = (unsigned char)o[1 - s2];
        if (dfs(o[s2], o[1 - s2], todo & ~(1ull << t), used + 2)) return true;
        n -= 2;
      }
    }
    if (used + rem + 2 > limit) return false;
    static const int P[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    for (uint64_t c = todo; c; c &= c - 1) {  // fresh triangle
      const int t = __builtin_ctzll(c);
      for (const auto& p : P) {
        const int x = M.tri[t][p[0]], y = M.tri[t][p[1]], z = M.tri[t][p[2]];
        if (rem > 1 && !(todo & ~(1ull << t) & M.vm[y] & M.vm[z])) continue;  // a strip of one triangle: never cheaper than elsewhere
        seq[n++] = (unsigned char)x;
        seq[n++] = (unsigned char)y;
        seq[n++] = (unsigned char)z;
        if (dfs(y, z, todo & ~(1ull << t), used + 3)) return true;
        n -= 3;
      }
    }
    return false;
  }
  // a node sequence of at most `lim` pushes covering every triangle, or false
  bool run(int lim, long total_budget) {
    static const int P[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    const uint64_t full = m == 64 ? ~0ull : ((1ull << m) - 1);
    limit = lim;
    spent = 0;
    for (long slice = 16000; spent < total_budget; slice *= 8) {
      bool exhausted_all = true;
      for (int t = 0; t < m; t++)
        for (const auto& p : P) {
          next_epoch();
          states = 0;
          budget = slice;
          n = 0;
          seq[n++] = L->tri[t][p[0]];
          seq[n++] = L->tri[t][p[1]];
          seq[n++] = L->tri[t][p[2]];
          const bool ok = dfs(L->tri[t][p[1]], L->tri[t][p[2]], full & ~(1ull << t), 3);
          spent += states;
          if (ok) return true;
          exhausted_all = exhausted_all && states <= budget;
          if (spent >= total_budget) return false;
        }
      if (exhausted_all) return false;  // every root searched to the end: infeasible
    }
    return false;
  }
};

struct SearchedStrip {
  uint64_t hash;
  int m, n;
  std::vector<unsigned char> key, node;
  std::vector<bool> comp;
};
std::mutex g_search_mu;
std::vector<SearchedStrip> g_searched;  // a few dozen classes per structured mesh: linear scan under the lock
StripSearch g_search;

// best: the greedy's strip on entry, a shorter one (by whole trips of 3) on return if the search finds it
void improve_strip(const MaskLink& M, const unsigned char* key, uint64_t hk, MaskSeq& best) {
  const int floor_n = M.m + 2;
  int target = (best.n - 1) / 3 * 3;
  if (target < floor_n || M.m > 64 || M.m < 2 || best.n > 250) return;  // (the state table keeps the push count in 8 bits)
  std::lock_guard<std::mutex> lock(g_search_mu);
  for (const SearchedStrip& it : g_searched)
    if (it.hash == hk && it.m == M.m && !memcmp(it.key.data(), key, (size_t)3 * M.m)) {
      if (it.n < best.n) {
        best.n = it.n;
        for (int k = 0; k < it.n; k++) {
          best.node[k] = it.node[k];
          best.comp[k] = it.comp[k];
        }
      }
      return;
    }
  // Whatever happens below is final for this class in this process (the strip must be a function of the link alone: the
  // per-thread memos and the two passes of cgasm_strip_plan_host rely on it), so the outcome is recorded even when the
  // budget is gone; a full table answers with the greedy for every class it does not hold.
  if (g_searched.size() >= 4096) return;
  g_search.init(M);
  long budget = std::min(1500000L, g_strip_search_states.load(std::memory_order_relaxed));
  if (budget <= 0) target = -1;
  const int greedy_n = best.n;
  const double t_log = getenv("CGASM_STRIP_SEARCH_LOG") ? omp_get_wtime() : 0.0;
  auto run = [&](int lim, long b) {
    const bool ok = g_search.run(lim, b);
    g_strip_search_states.fetch_sub(g_search.spent, std::memory_order_relaxed);
    return ok;
  };
  while (target >= floor_n && run(target, budget)) {
    // replay through mask_push: it computes a pending triangle whenever the window forms one, so every triangle of
    // the found cover is computed at the planned push or earlier
    MaskSeq S;
    uint64_t todo = M.m == 64 ? ~0ull : ((1ull << M.m) - 1);
    for (int k = 0; k < g_search.best_n; k++) mask_push(M, todo, S, g_search.best[k]);
    if (todo != 0 || S.n >= best.n) break;
    best = S;
    target -= 3;
    budget = std::min(200000L, g_strip_search_states.load(std::memory_order_relaxed));  // a second trip saved is rare: look, but briefly
  }
  if (t_log != 0.0)
    fprintf(stderr, "strip search: link of %d triangles, greedy %d -> %d pushes, %.3f s\n", M.m, greedy_n, best.n, omp_get_wtime() - t_log);
  {
    SearchedStrip it;
    it.hash = hk;
    it.m = M.m;
    it.n = best.n;
    it.key.assign(key, key + (size_t)3 * M.m);
    it.node.assign(best.node, best.node + best.n);
    it.comp.assign(best.comp, best.comp + best.n);
    g_searched.push_back(std::move(it));
  }
}

// Per-thread memo of finished strips keyed by the link's local triangle list: the strip (as local ids) is a
// function of that list alone, and on structured meshes a few dozen lists cover every row.
struct StripMemo {
  static constexpr int kSlots = 1024, kMaxKey = 3 * 64;
  struct Item {
    uint64_t hash = 0;
    int m = 0, n = 0;
    unsigned char key[kMaxKey];
    unsigned char node[4 * 64 + 8];
    bool comp[4 * 64 + 8];
  };
  std::vector<Item> items;
  int used = 0;
  StripMemo() : items(kSlots) {}
  static uint64_t hash_of(const unsigned char* k, int len) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; i++) h = (h ^ k[i]) * 1099511628211ull;
    return h | 1;  // 0 marks an empty slot
  }
  Item* find(uint64_t h, const unsigned char* k, int m) {
    for (size_t p = h % kSlots, probes = 0; probes < 8; p = (p + 1) % kSlots, probes++) {
      Item& it = items[p];
      if (it.hash == 0) return nullptr;
      if (it.hash == h && it.m == m && !memcmp(it.key, k, (size_t)3 * m)) return &it;
    }
    return nullptr;
  }
  void store(uint64_t h, const unsigned char* k, int m, const MaskSeq& S) {
    if (used > kSlots / 2) return;
    for (size_t p = h % kSlots, probes = 0; probes < 8; p = (p + 1) % kSlots, probes++) {
      Item& it = items[p];
      if (it.hash != 0) continue;
      it.hash = h;
      it.m = m;
      it.n = S.n;
      memcpy(it.key, k, (size_t)3 * m);
      memcpy(it.node, S.node, (size_t)S.n);
      memcpy(it.comp, S.comp, (size_t)S.n * sizeof(bool));
      used++;
      return;
    }
  }
};

// the 3-D strip of row r on the fast path; false = use the generic one
bool fast_strip_row(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm, int r,
                    std::vector<int>& nodes, std::vector<char>& comp, const int64_t* geokey = nullptr) {
  FastLink L;
  if (!fast_link(loc, nd0, n2e_ptr, n2e, findrm, colm, r, L, geokey)) return false;
  static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  if (L.m <= 64) {
    static thread_local StripMemo memo;
    const unsigned char* key = &L.tri[0][0];
    // (a strip memoised while the search was off must not answer for a build that has it on, and vice versa)
    const uint64_t hk = StripMemo::hash_of(key, 3 * L.m) ^ (g_strip_search.load(std::memory_order_relaxed) ? 0xb5297a4d3f84d5b4ull : 0ull);
    if (const StripMemo::Item* it = memo.find(hk, key, L.m)) {
      nodes.resize((size_t)it->n);
      comp.resize((size_t)it->n);
      for (int k = 0; k < it->n; k++) {
        nodes[k] = L.vid[it->node[k]];
        comp[k] = it->comp[k] ? 1 : 0;
      }
      return true;
    }
    MaskLink M;
    M.m = L.m;
    M.tri = L.tri;
    for (int v = 0; v < L.nv; v++) {
      uint64_t w = 0;
      for (int q = 0; q < L.ninc[v]; q++) w |= 1ull << L.inc[v][q];
      M.vm[v] = w;
    }
    int start = 0, start_nb = 1 << 30;
    for (int t = 0; t < L.m; t++) {
      const unsigned char* o = L.tri[t];
      const int nb = __builtin_popcountll(M.vm[o[0]] & M.vm[o[1]]) + __builtin_popcountll(M.vm[o[1]] & M.vm[o[2]]) +
                     __builtin_popcountll(M.vm[o[0]] & M.vm[o[2]]) - 3;
      if (nb < start_nb) {
        start_nb = nb;
        start = t;
      }
    }
    MaskSeq S, best;
    for (int p = 0; p < 6; p++) {
      mask_strip3(M, start, perms[p], S);
      if (p == 0 || S.n < best.n) best = S;
      if (best.n == L.m + 2) break;  // one push per triangle after the first: cannot be shorter
    }
    if (g_strip_search.load(std::memory_order_relaxed)) improve_strip(M, key, hk, best);
    memo.store(hk, key, L.m, best);
    nodes.resize((size_t)best.n);
    comp.resize((size_t)best.n);
    for (int k = 0; k < best.n; k++) {
      nodes[k] = L.vid[best.node[k]];
      comp[k] = best.comp[k] ? 1 : 0;
    }
    return true;
  }
  // start from a triangle with the fewest edge neighbours
  int start = 0, start_nb = 1 << 30;
  for (int t = 0; t < L.m; t++) {
    int nb = 0;
    // triangles sharing exactly two nodes with t: around each edge of t, minus t itself
    for (int e = 0; e < 3; e++) {
      const int x = L.tri[t][e], y = L.tri[t][(e + 1) % 3];
      for (int q = 0; q < L.ninc[x]; q++) {
        const int u = L.inc[x][q];
        nb += (u != t && L.has(u, y));
      }
    }
    if (nb < start_nb) {
      start_nb = nb;
      start = t;
    }
  }
  FastSeq S, best;
  for (int p = 0; p < 6; p++) {
    fast_strip3(L, start, perms[p], S);
    if (p == 0 || S.n < best.n) best = S;
  }
  nodes.resize((size_t)best.n);
  comp.resize((size_t)best.n);
  for (int k = 0; k < best.n; k++) {
    nodes[k] = L.vid[best.node[k]];
    comp[k] = best.comp[k] ? 1 : 0;
  }
  return true;
}

// Decides whether the strips of this mesh are worth searching (improve_strip): the links of up to 2048 evenly spaced
// rows are canonicalised as the builder will see them; a mesh whose sample falls into a few classes is structured
// enough for one search per class to pay. CGASM_STRIP_SEARCH=0 / 1 forces the answer.
void strip_search_prescan(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm,
                          const int64_t* geokey, int n_nodes) {
  int on = 0;
  if (const char* e = getenv("CGASM_STRIP_SEARCH")) {
    on = atoi(e) != 0;
  } else if (loc == 4 && n_nodes > 0) {
    const int samples = std::min(n_nodes, 2048);
    std::unordered_set<uint64_t> classes;
    int ok = 0;
    for (int q = 0; q < samples; q++) {
      const int r = (int)((int64_t)q * n_nodes / samples);
      FastLink L;
      if (n2e_ptr[r + 1] == n2e_ptr[r] || !fast_link(loc, nd0, n2e_ptr, n2e, findrm, colm, r, L, geokey) || L.m > 64) continue;
      ok++;
      classes.insert(StripMemo::hash_of(&L.tri[0][0], 3 * L.m) ^ (uint64_t)L.m << 56);
    }
    on = ok > 0 && (int)classes.size() <= std::max(48, ok / 8);
  }
  g_strip_search.store(on, std::memory_order_relaxed);
  g_strip_search_states.store(4000000, std::memory_order_relaxed);  // ~2 s of search per build at most
}

}  // namespace

static void build_strip_row_impl(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm,
                                 const int* colm, int r, std::vector<StripEntry>& out, bool allow_fast) {
  out.clear();
  const int64_t k0 = n2e_ptr[r];
  const int m = (int)(n2e_ptr[r + 1] - k0);
  if (m == 0) return;
  std::vector<int> nodes;
  std::vector<char> comp;
  if (!(allow_fast && loc == 4 && fast_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, nodes, comp))) {
    Link L;
    L.m = m;
    L.w = loc - 1;
    L.oth.assign((size_t)3 * L.m, -1);
    L.todo.assign((size_t)L.m, 1);
    for (int k = 0; k < L.m; k++) {
      const int* nd = nd0 + (size_t)4 * n2e[k0 + k];
      int q = 0;
      for (int i = 0; i < loc; i++)
        if (nd[i] != r && q < 3) L.oth[3 * k + q++] = nd[i];
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
    nodes = best.node;
    comp = best.comp;
  }
  const int s0 = findrm[r], s1 = findrm[r + 1];
  out.resize(nodes.size());
  for (size_t k = 0; k < nodes.size(); k++) {
    const int* cb = colm + s0;
    const int slot = (int)(std::lower_bound(cb, colm + s1, nodes[k]) - cb);
    out[k].node = nodes[k];
    out[k].meta = (slot & 0xff) | (comp[k] ? kStripCompute : 0);
  }
}

void build_strip_row(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm,
                     int r, std::vector<StripEntry>& out) {
  build_strip_row_impl(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out, true);
}

void build_strip_row_generic(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm,
                             const int* colm, int r, std::vector<StripEntry>& out) {
  build_strip_row_impl(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out, false);
}

// The strip of row r built on a CANONICAL copy of its link: the link's vertices are relabelled by the rank of their
// geometric key (lexicographic lattice coordinates, ties by id) and its simplices sorted by their relabelled vertex
// tuples, so the result depends on the geometry around r and not on how the mesh generator numbered nodes and
// elements. Rows with congruent neighbourhoods then walk their neighbours in the same relative order: on a renumbered
// structured mesh the lanes of a warp read the staged records at a common offset again (shared-memory bank conflicts
// measured on the B200 with id-ordered strips: 2.1x the wavefronts of the lexicographically numbered mesh), and the
// per-thread memo of finished strips hits as on the unshuffled mesh. With a lexicographic numbering the keyed and the
// id-ordered labels coincide.
void build_strip_row_keyed(int loc, const int* nd0, const int64_t* n2e_ptr, const int* n2e, const int* findrm, const int* colm,
                           const int64_t* geokey, int r, std::vector<StripEntry>& out) {
  if (!geokey) return build_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out);
  if (loc == 4) {  // 3-D links that fit the fixed-size tables are relabelled inside fast_link (no copy of the link)
    static thread_local std::vector<int> nodes;
    static thread_local std::vector<char> comp;
    if (n2e_ptr[r + 1] > n2e_ptr[r] && fast_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, nodes, comp, geokey)) {
      const int* cb = colm + findrm[r];
      const int* ce = colm + findrm[r + 1];
      out.resize(nodes.size());
      for (size_t k = 0; k < nodes.size(); k++) {
        out[k].node = nodes[k];
        out[k].meta = ((int)(std::lower_bound(cb, ce, nodes[k]) - cb) & 0xff) | (comp[k] ? kStripCompute : 0);
      }
      return;
    }
  }
  const int s0 = findrm[r], len = findrm[r + 1] - s0;
  const int64_t k0 = n2e_ptr[r];
  const int m = (int)(n2e_ptr[r + 1] - k0);
  out.clear();
  if (m == 0) return;
  constexpr int kMaxV = 512;
  if (len < 2 || len > kMaxV) return build_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out);
  const int nv = len - 1;
  int byid[kMaxV], rank_of[kMaxV], verts[kMaxV];  // byid: the row minus r (ascending ids); verts: the same by key
  {
    int q = 0;
    for (int k = 0; k < len; k++)
      if (colm[s0 + k] != r && q < nv) byid[q++] = colm[s0 + k];
    if (q != nv) return build_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out);
  }
  {
    int ord[kMaxV];
    for (int k = 0; k < nv; k++) ord[k] = k;
    std::sort(ord, ord + nv, [&](int a, int b) {
      const int64_t ka = geokey[byid[a]], kb = geokey[byid[b]];
      return ka != kb ? ka < kb : byid[a] < byid[b];
    });
    for (int k = 0; k < nv; k++) {
      rank_of[ord[k]] = k;
      verts[k] = byid[ord[k]];
    }
  }
  // the link in local labels: element k = {nv (the row's own node), ranks of its other nodes ascending}, sorted
  const int w = loc - 1;
  // per-thread scratch: this runs once per row of the mesh
  static thread_local std::vector<int> tri, tord, nd_l, n2e_l, colm_l, findrm_l;
  static thread_local std::vector<int64_t> ptr_l;
  static thread_local std::vector<StripEntry> loc_out;
  tri.assign((size_t)m * 3, 0);
  for (int k = 0; k < m; k++) {
    const int* nd = nd0 + (size_t)4 * n2e[k0 + k];
    int q = 0;
    for (int i = 0; i < loc; i++) {
      if (nd[i] == r) continue;
      const int* f = std::lower_bound(byid, byid + nv, nd[i]);
      if (q == w || f == byid + nv || *f != nd[i]) return build_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out);
      tri[(size_t)3 * k + q++] = rank_of[f - byid];
    }
    if (q != w) return build_strip_row(loc, nd0, n2e_ptr, n2e, findrm, colm, r, out);
    std::sort(&tri[(size_t)3 * k], &tri[(size_t)3 * k] + w);
  }
  tord.resize((size_t)m);
  for (int k = 0; k < m; k++) tord[k] = k;
  std::sort(tord.begin(), tord.end(), [&](int a, int b) {
    for (int i = 0; i < w; i++)
      if (tri[(size_t)3 * a + i] != tri[(size_t)3 * b + i]) return tri[(size_t)3 * a + i] < tri[(size_t)3 * b + i];
    return a < b;
  });
  nd_l.assign((size_t)m * 4, -1);
  n2e_l.resize((size_t)m);
  colm_l.resize((size_t)nv + 1);
  ptr_l.assign((size_t)nv + 2, 0);
  findrm_l.assign((size_t)nv + 2, 0);
  for (int k = 0; k < m; k++) {
    nd_l[(size_t)4 * k] = nv;
    for (int i = 0; i < w; i++) nd_l[(size_t)4 * k + 1 + i] = tri[(size_t)3 * tord[k] + i];
    n2e_l[k] = k;
  }
  ptr_l[(size_t)nv + 1] = m;
  findrm_l[(size_t)nv + 1] = nv + 1;
  for (int k = 0; k <= nv; k++) colm_l[k] = k;
  build_strip_row(loc, nd_l.data(), ptr_l.data(), n2e_l.data(), findrm_l.data(), colm_l.data(), nv, loc_out);
  out.resize(loc_out.size());
  for (size_t k = 0; k < loc_out.size(); k++) {
    const int node = verts[loc_out[k].node];
    const int slot = (int)(std::lower_bound(colm + s0, colm + s0 + len, node) - (colm + s0));
    out[k].node = node;
    out[k].meta = (slot & 0xff) | (loc_out[k].meta & kStripCompute);
  }
}

// ---- staged plan ------------------------------------------------------------------------------------------
namespace {
// node -> small index, O(1) clear by generation stamps (a block touches a few hundred nodes)
struct NodeIndexMap {
  std::vector<int> key, val;
  std::vector<unsigned> gen;
  unsigned cur = 1;
  size_t mask, count = 0;
  explicit NodeIndexMap(size_t cap = 16384) : key(cap), val(cap), gen(cap, 0), mask(cap - 1) {}
  static size_t hash(int v) { return (size_t)((uint32_t)v * 2654435761u) >> 9; }
  void clear() {
    count = 0;
    if (++cur == 0) {
      std::fill(gen.begin(), gen.end(), 0u);
      cur = 1;
    }
  }
  void grow() {
    std::vector<int> k2, v2;
    for (size_t p = 0; p <= mask; p++)
      if (gen[p] == cur) {
        k2.push_back(key[p]);
        v2.push_back(val[p]);
      }
    const size_t cap = 2 * (mask + 1);
    key.assign(cap, 0);
    val.assign(cap, 0);
    gen.assign(cap, 0);
    mask = cap - 1;
    cur = 1;
    count = 0;
    for (size_t i = 0; i < k2.size(); i++) *slot(k2[i]) = v2[i];
  }
  // pointer to the value of `node`, inserted with value -1 if absent
  int* slot(int node) {
    if (2 * (count + 1) > mask + 1) grow();
    size_t p = hash(node) & mask;
    while (gen[p] == cur) {
      if (key[p] == node) return &val[p];
      p = (p + 1) & mask;
    }
    key[p] = node;
    val[p] = -1;
    gen[p] = cur;
    count++;
    return &val[p];
  }
};
// Bank groups for the staged node records of one row block. groups: 8 local node indices per (quarter-warp, step) read,
// -1 padded, distinct; a read costs max over bank groups of the nodes it touches there. Starts from index mod 8 (the
// geometric order), then moves single nodes to the bank group that lowers the summed cost most, a few sweeps, within
// the capacity of the chunk stride the block needs anyway. Returns false (leave the order alone) when the block is
// conflict-light already or nothing was gained; newidx[i] = 8 * position + group otherwise.
bool assign_banks(int nn, const std::vector<int>& groups, std::vector<int>& newidx) {
  const int ng = (int)(groups.size() / 8);
  if (nn < 16 || ng == 0) return false;
  static thread_local std::vector<unsigned char> cls;
  static thread_local std::vector<int> gptr, glist;
  cls.resize((size_t)nn);
  for (int i = 0; i < nn; i++) cls[i] = (unsigned char)(i & 7);
  auto cost_of = [&](int g) {
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
    for (int m = 0; m < 8; m++) {
      const int i = groups[(size_t)8 * g + m];
      if (i >= 0) mx = std::max(mx, ++cnt[cls[i]]);
    }
    return mx;
  };
  long long cost0 = 0;
  for (int g = 0; g < ng; g++) cost0 += cost_of(g);
  if ((double)cost0 < 1.35 * ng) return false;
  int nl_cap = 0;
  for (int c : kStagedNL)
    if (!nl_cap && c >= nn) nl_cap = c;
  if (!nl_cap) return false;
  const int cap = nl_cap / 8;
  gptr.assign((size_t)nn + 1, 0);
  for (size_t q = 0; q < groups.size(); q++)
    if (groups[q] >= 0) gptr[(size_t)groups[q] + 1]++;
  for (int i = 0; i < nn; i++) gptr[(size_t)i + 1] += gptr[i];
  glist.resize((size_t)gptr[nn]);
  {
    std::vector<int> fill(gptr.begin(), gptr.end() - 1);
    for (int g = 0; g < ng; g++)
      for (int m = 0; m < 8; m++) {
        const int i = groups[(size_t)8 * g + m];
        if (i >= 0) glist[(size_t)fill[i]++] = g;
      }
  }
  int size[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nn; i++) size[cls[i]]++;
  long long cost = cost0;
  for (int sweep = 0; sweep < 4; sweep++) {  // (single-node moves converge after two or three sweeps: 2.54 -> 2.14 wavefronts per read)
    long long gained = 0;
    for (int i = 0; i < nn; i++) {
      const int c0 = cls[i];
      int delta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int q = gptr[i]; q < gptr[(size_t)i + 1]; q++) {
        const int g = glist[q];
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int m = 0; m < 8; m++) {
          const int j = groups[(size_t)8 * g + m];
          if (j >= 0) cnt[cls[j]]++;
        }
        int before = 0;
        for (int c = 0; c < 8; c++) before = std::max(before, cnt[c]);
        cnt[c0]--;
        int rest = 0;
        for (int c = 0; c < 8; c++) rest = std::max(rest, cnt[c]);
        for (int c = 0; c < 8; c++) delta[c] += std::max(rest, cnt[c] + 1) - before;
      }
      int best = c0;
      for (int c = 0; c < 8; c++)
        if (c != c0 && size[c] < cap && delta[c] < delta[best]) best = c;
      if (best != c0) {
        gained -= delta[best] - delta[c0];
        cls[i] = (unsigned char)best;
        size[c0]--;
        size[best]++;
      }
    }
    cost -= gained;
    if (gained * 200 < cost) break;  // less than half a percent: done
  }
  if (cost >= cost0) return false;
  int pos[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  newidx.resize((size_t)nn);
  for (int i = 0; i < nn; i++) newidx[i] = 8 * pos[cls[i]]++ + cls[i];
  return true;
}

}  // namespace

void build_staged_plan_host(const Handle* h, const std::vector<int>& rows, int nblocks, int maxlen,
                            const std::vector<int>& perm, StagedPlanHost& out) {
  const bool permuted = !perm.empty();
  const int loc = h->loc, dim = h->dim;
  const int64_t* gk = (h->geokey.empty() || getenv("CGASM_STRIP_NOKEY")) ? nullptr : h->geokey.data();
  strip_search_prescan(loc, h->h_nd0.data(), h->n2e_ptr.data(), h->n2e.data(), h->h_findrm.data(), h->h_colm.data(), gk, h->n_nodes);
  // bank groups are searched on meshes that are NOT a handful of congruence classes (there the geometric order is
  // conflict-free and keeps the staging copies contiguous); CGASM_STRIP_BANKS=0/1 overrides
  const bool bank_search = getenv("CGASM_STRIP_BANKS") ? atoi(getenv("CGASM_STRIP_BANKS")) != 0
                                                        : !g_strip_search.load(std::memory_order_relaxed);
  constexpr int kTask = 128;  // blocks per task: fixed, so the layout does not depend on the thread count
  const int ntasks = (nblocks + kTask - 1) / kTask;
  out.nblocks = nblocks;
  out.task_blocks = kTask;
  out.ent.assign((size_t)std::max(ntasks, 1), {});
  out.own_local.assign(rows.size(), 0);
  out.row_meta.assign(rows.size() * 4, 0);
  out.ptr.assign((size_t)nblocks + 1, 0);
  std::vector<int> block_ldeg((size_t)nblocks, 0), block_nn((size_t)nblocks, 0);
  std::vector<std::vector<int>> task_nodes((size_t)std::max(ntasks, 1));  // node lists of the task's blocks, concatenated
  long long total_real = 0;
  int overflow = 0;
#pragma omp parallel reduction(+ : total_real, overflow)
  {
    std::vector<StripEntry> rp[kBR];
    NodeIndexMap map;
    std::vector<int> bn, bn2, groups, newidx;
    std::vector<std::pair<int64_t, int>> bk;
#pragma omp for schedule(dynamic, 1)
    for (int task = 0; task < ntasks; task++) {
      std::vector<unsigned>& ent = out.ent[task];
      std::vector<int>& tn = task_nodes[task];
      const int b0 = task * kTask, b1 = std::min(nblocks, b0 + kTask);
      ent.reserve((size_t)(b1 - b0) * kBR * 36);
      for (int b = b0; b < b1; b++) {
        int deg = 0;
        map.clear();
        bn.clear();
        bk.clear();
        // the map and the block list are keyed by the record position (perm[node] when the records are permuted)
        auto key_of = [&](int node) { return permuted ? perm[node] : node; };
        auto note = [&](int node) {
          int* v = map.slot(key_of(node));
          if (*v < 0) {
            *v = 0;
            bk.emplace_back(gk ? gk[node] : 0, key_of(node));
          }
        };
        for (int t = 0; t < kBR; t++) {
          const int r = rows[(size_t)b * kBR + t];
          rp[t].clear();
          note(r >= 0 ? r : 0);
          if (r < 0) continue;
          build_strip_row_keyed(loc, h->h_nd0.data(), h->n2e_ptr.data(), h->n2e.data(), h->h_findrm.data(), h->h_colm.data(),
                                gk, r, rp[t]);
          deg = std::max(deg, (int)rp[t].size());
          total_real += (long long)rp[t].size();
          for (const StripEntry& e : rp[t]) note(e.node);
        }
        // local indices follow the geometric key (the lexicographic order of the touched region: what the id order is on
        // a lexicographically numbered mesh, whose staged reads are bank-conflict free), then the record position
        std::sort(bk.begin(), bk.end());
        bn.resize(bk.size());
        for (size_t i = 0; i < bk.size(); i++) bn[i] = bk[i].second;
        for (size_t i = 0; i < bn.size(); i++) *map.slot(bn[i]) = (int)i;
        if (bank_search && dim == 3) {
          // Unstructured blocks: the 8 lanes of a quarter-warp LDS.128 read 8 unrelated nodes, 2.5 wavefronts per access
          // with indices in any fixed order (8 balls into 8 bank groups). assign_banks moves nodes between the bank
          // groups (index mod 8) to spread every (warp, step, quarter) group of reads; a no-op where the geometric order
          // is conflict-free already (structured bricks).
          groups.clear();
          for (int w8 = 0; w8 < kBR / 8; w8++)
            for (int k = 0; k < deg; k++) {
              int g[8], ng = 0;
              for (int l = 0; l < 8; l++) {
                const int t = 8 * w8 + l;
                if (k >= (int)rp[t].size()) continue;
                const int i = *map.slot(key_of(rp[t][k].node));
                bool dup = false;
                for (int m = 0; m < ng; m++) dup |= g[m] == i;
                if (!dup) g[ng++] = i;
              }
              if (ng > 1) {
                for (int m = 0; m < 8; m++) groups.push_back(m < ng ? g[m] : -1);
              }
            }
          if (assign_banks((int)bn.size(), groups, newidx)) {
            int top = 0;
            for (size_t i = 0; i < bn.size(); i++) top = std::max(top, newidx[i] + 1);
            bn2.assign((size_t)top, -1);
            for (size_t i = 0; i < bn.size(); i++) bn2[(size_t)newidx[i]] = bn[i];
            bn.swap(bn2);
            for (size_t i = 0; i < bn.size(); i++)
              if (bn[i] >= 0) *map.slot(bn[i]) = (int)i;
          }
        }
        block_nn[b] = (int)bn.size();
        tn.insert(tn.end(), bn.begin(), bn.end());
        if (bn.size() > 4095) overflow++;
        const int ldeg = (deg + dim - 1) / dim * dim;
        block_ldeg[b] = ldeg;
        const size_t base = ent.size();
        ent.resize(base + (size_t)ldeg * kBR);
        for (int t = 0; t < kBR; t++) {
          const size_t q = (size_t)b * kBR + t;
          const int r = rows[q];
          int own = 0;
          if (r >= 0) {
            const int* cb = h->h_colm.data() + h->h_findrm[r];
            own = (int)(std::lower_bound(cb, (const int*)h->h_colm.data() + h->h_findrm[r + 1], r) - cb);
          }
          const unsigned ol = (unsigned)(own * kAS) << 16 | (unsigned)(*map.slot(key_of(r >= 0 ? r : 0)) & 0xfff) << 4;
          out.own_local[q] = ol;
          out.row_meta[4 * q + 0] = r;
          out.row_meta[4 * q + 1] = r >= 0 ? h->h_findrm[r] : 0;
          // bits 0-7 row length, 16-23 own slot, 24-31 strip length in units of dim entries (the warp's trip count)
          const int steps = ((int)rp[t].size() + dim - 1) / dim;
          if (steps > 255) overflow++;
          out.row_meta[4 * q + 2] = (r >= 0 ? h->h_findrm[r + 1] - h->h_findrm[r] : 0) | own << 16 | (int)((unsigned)(steps & 0xff) << 24);
          out.row_meta[4 * q + 3] = (int)ol;
          const int n = (int)rp[t].size();
          for (int k = 0; k < ldeg; k++) {
            unsigned lv = ol;  // padding: re-push the own node, nothing computed
            if (k < n) {
              const StripEntry& e = rp[t][k];
              lv = (unsigned)((e.meta & 0xff) * kAS) << 16 | (unsigned)(*map.slot(key_of(e.node)) & 0xfff) << 4 |
                   ((e.meta & kStripCompute) ? kStagedCompute : 0u);
            }
            ent[base + (size_t)k * kBR + t] = lv;
          }
        }
      }
    }
  }
  for (int b = 0; b < nblocks; b++) out.ptr[(size_t)b + 1] = out.ptr[b] + (long long)block_ldeg[b] * kBR;
  out.total_real = total_real;
  out.blk_nn = block_nn;
  out.blk_ml.assign((size_t)nblocks, 0);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; b++) {
    int ml = 0;
    for (int t = 0; t < kBR; t++) {
      const int r = rows[(size_t)b * kBR + t];
      if (r >= 0) ml = std::max(ml, h->h_findrm[r + 1] - h->h_findrm[r]);
    }
    out.blk_ml[b] = ml;
  }
  out.blk_nodes_max = 0;
  for (int b = 0; b < nblocks; b++) out.blk_nodes_max = std::max(out.blk_nodes_max, block_nn[b]);
  out.nl = 0;
  for (int c : kStagedNL)
    if (!out.nl && c >= out.blk_nodes_max) out.nl = c;
  out.ok = !overflow && out.nl > 0 && (long long)maxlen * kAS < 65536;
  if (!out.ok) return;
  const int nl = out.nl;
  out.blk_nodes.assign((size_t)std::max(nblocks, 1) * nl, -1);
#pragma omp parallel for schedule(dynamic, 1)
  for (int task = 0; task < ntasks; task++) {
    const int b0 = task * kTask, b1 = std::min(nblocks, b0 + kTask);
    size_t off = 0;
    for (int b = b0; b < b1; b++) {
      std::copy(task_nodes[task].begin() + off, task_nodes[task].begin() + off + block_nn[b], out.blk_nodes.begin() + (size_t)b * nl);
      off += (size_t)block_nn[b];
    }
  }
}

}  // namespace cgasm

// ---- diagnostics ABI: the strips of every row of a mesh, built on the host (no GPU needed) ----------
extern "C" int cgasm_strip_plan_host(int loc, int n_nodes, int n_elements, const int* ndglno, long long* row_ptr,
                                     int* entries, long long capacity, long long* needed) {
  using namespace cgasm;
  if ((loc != 3 && loc != 4) || n_nodes <= 0 || n_elements <= 0 || !ndglno || !row_ptr || !needed)
    CG_FAIL(CGASM_EARG, "cgasm_strip_plan_host: bad argument");
  IVec nd0((size_t)4 * n_elements, -1);
  for (int e = 0; e < n_elements; e++)
    for (int i = 0; i < loc; i++) {
      const int v = ndglno[(size_t)loc * e + i] - 1;
      if (v < 0 || v >= n_nodes) CG_FAIL(CGASM_EARG, "cgasm_strip_plan_host: node id out of range");
      nd0[(size_t)4 * e + i] = v;
    }
  I64Vec n2e_ptr;
  IVec n2e, findrm, colm;
  build_node_to_element(n_nodes, n_elements, loc, nd0.data(), n2e_ptr, n2e);
  build_sparsity(n_nodes, n_elements, loc, nd0.data(), n2e_ptr, n2e, findrm, colm);
  // CGASM_STRIP_GENERIC=1: the O(m^2) reference implementation of the same greedy (equivalence tests)
  const bool generic = getenv("CGASM_STRIP_GENERIC") != nullptr;
  strip_search_prescan(generic ? 0 : loc, nd0.data(), n2e_ptr.data(), n2e.data(), findrm.data(), colm.data(), nullptr, n_nodes);
  std::vector<int> len((size_t)n_nodes, 0);
#pragma omp parallel
  {
    std::vector<StripEntry> row;
#pragma omp for schedule(dynamic, 1024)
    for (int r = 0; r < n_nodes; r++) {
      if (generic) build_strip_row_generic(loc, nd0.data(), n2e_ptr.data(), n2e.data(), findrm.data(), colm.data(), r, row);
      else build_strip_row(loc, nd0.data(), n2e_ptr.data(), n2e.data(), findrm.data(), colm.data(), r, row);
      len[r] = (int)row.size();
    }
  }
  long long total = 0;
  row_ptr[0] = 0;
  for (int r = 0; r < n_nodes; r++) row_ptr[r + 1] = (total += len[r]);
  if (entries && total <= capacity) {
#pragma omp parallel
    {
      std::vector<StripEntry> row;
#pragma omp for schedule(dynamic, 1024)
      for (int r = 0; r < n_nodes; r++) {
        if (generic) build_strip_row_generic(loc, nd0.data(), n2e_ptr.data(), n2e.data(), findrm.data(), colm.data(), r, row);
        else build_strip_row(loc, nd0.data(), n2e_ptr.data(), n2e.data(), findrm.data(), colm.data(), r, row);
        for (size_t k = 0; k < row.size(); k++) {
          entries[2 * (row_ptr[r] + (long long)k)] = row[k].node + 1;
          entries[2 * (row_ptr[r] + (long long)k) + 1] = row[k].meta;
        }
      }
    }
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

// ---- diagnostics ABI: wall-clock seconds of the host phases of a handle's set-up (no GPU needed) --------
// times[0..5]: connectivity conversion, node->element adjacency, sparsity, Morton order, row blocks, strips
extern "C" int cgasm_plan_host_timing(int dim, int n_nodes, int n_elements, const int* ndglno, const double* X,
                                      double* times, double* entries_per_pair) {
  using namespace cgasm;
  const int loc = dim + 1;
  if ((dim != 2 && dim != 3) || n_nodes <= 0 || n_elements <= 0 || !ndglno || !X || !times)
    CG_FAIL(CGASM_EARG, "cgasm_plan_host_timing: bad argument");
  auto now = [] { return omp_get_wtime(); };
  Handle h;
  h.dim = dim;
  h.loc = loc;
  h.n_nodes = n_nodes;
  h.n_elements = n_elements;
  double t0 = now();
  h.h_nd0.resize((size_t)4 * n_elements);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < n_elements; e++)
    for (int i = 0; i < 4; i++) h.h_nd0[(size_t)4 * e + i] = i < loc ? ndglno[(size_t)loc * e + i] - 1 : -1;
  h.h_X.assign(X, X + (size_t)dim * n_nodes);
  times[0] = now() - t0;
  t0 = now();
  build_node_to_element(n_nodes, n_elements, loc, h.h_nd0.data(), h.n2e_ptr, h.n2e);
  times[1] = now() - t0;
  t0 = now();
  build_sparsity(n_nodes, n_elements, loc, h.h_nd0.data(), h.n2e_ptr, h.n2e, h.h_findrm, h.h_colm);
  times[2] = now() - t0;
  t0 = now();
  std::vector<int> order, rows;
  MortonFrame F;
  morton_order(&h, order, F);
  times[3] = now() - t0;
  t0 = now();
  const int nb = form_row_blocks(&h, order, F, 128, rows);
  order_block_rows(&h, F, rows, nb);
  times[4] = now() - t0;
  t0 = now();
  h.h_nd0.shrink_to_fit();
  StagedPlanHost sp;
  build_staged_plan_host(&h, rows, nb, 64, std::vector<int>(), sp);
  const long long total = sp.total_real;
  times[5] = now() - t0;
  if (entries_per_pair) *entries_per_pair = h.n2e.empty() ? 0.0 : (double)total / (double)h.n2e.size();
  return CGASM_OK;
}

// ---- diagnostics ABI: shape of the staged STRIP plan of a mesh, built on the host (no GPU needed) ----------------
// stats[0] strip entries per (row, element) pair; [1] largest number of distinct nodes a block touches;
// [2] shared-memory wavefronts per quarter-warp LDS.128 of the staged records (1 = bank-conflict free: the 8 lanes
//     read 8 different 16-byte bank groups, lanes on the same node broadcast);
// [3] entries a warp walks / entries its rows need (1 = every lane busy until the warp's last step);
// [4] element computations a warp executes / computations its rows need (a step computes if any lane does).
extern "C" int cgasm_plan_host_stats(int dim, int n_nodes, int n_elements, const int* ndglno, const double* X, double* stats) {
  using namespace cgasm;
  const int loc = dim + 1;
  if ((dim != 2 && dim != 3) || n_nodes <= 0 || n_elements <= 0 || !ndglno || !X || !stats)
    CG_FAIL(CGASM_EARG, "cgasm_plan_host_stats: bad argument");
  Handle h;
  h.dim = dim;
  h.loc = loc;
  h.n_nodes = n_nodes;
  h.n_elements = n_elements;
  h.h_nd0.resize((size_t)4 * n_elements);
  for (int e = 0; e < n_elements; e++)
    for (int i = 0; i < 4; i++) h.h_nd0[(size_t)4 * e + i] = i < loc ? ndglno[(size_t)loc * e + i] - 1 : -1;
  h.h_X.assign(X, X + (size_t)dim * n_nodes);
  build_node_to_element(n_nodes, n_elements, loc, h.h_nd0.data(), h.n2e_ptr, h.n2e);
  build_sparsity(n_nodes, n_elements, loc, h.h_nd0.data(), h.n2e_ptr, h.n2e, h.h_findrm, h.h_colm);
  std::vector<int> order, rows;
  MortonFrame F;
  morton_order(&h, order, F);
  const int nb = form_row_blocks(&h, order, F, kBR, rows);
  order_block_rows(&h, F, rows, nb);
  StagedPlanHost sp;
  build_staged_plan_host(&h, rows, nb, 64, std::vector<int>(), sp);
  stats[0] = h.n2e.empty() ? 0.0 : (double)sp.total_real / (double)h.n2e.size();
  stats[1] = sp.blk_nodes_max;
  double wf = 0, groups = 0, walked = 0, needed = 0, comp_exec = 0, comp_need = 0;
  for (int b = 0; b < nb; b++) {
    const int task = b / sp.task_blocks;
    const long long base = sp.ptr[b] - sp.ptr[(size_t)task * sp.task_blocks];
    const int ldeg = (int)((sp.ptr[(size_t)b + 1] - sp.ptr[b]) / kBR);
    const unsigned* ent = sp.ent[task].data() + base;
    for (int w = 0; w < kBR / 32; w++) {
      int steps = 0;
      for (int t = 32 * w; t < 32 * w + 32; t++) steps = std::max(steps, (sp.row_meta[4 * ((size_t)b * kBR + t) + 2] >> 24) & 0xff);
      const int wdeg = std::min(ldeg, steps * dim);
      for (int t = 32 * w; t < 32 * w + 32; t++) {
        const int mine = ((sp.row_meta[4 * ((size_t)b * kBR + t) + 2] >> 24) & 0xff) * dim;
        needed += mine;
        walked += wdeg;
      }
      for (int k = 0; k < wdeg; k++) {
        int anyc = 0;
        for (int t = 32 * w; t < 32 * w + 32; t++) {
          const unsigned e = ent[(size_t)k * kBR + t];
          comp_need += e & kStagedCompute;
          anyc |= (int)(e & kStagedCompute);
        }
        comp_exec += 32.0 * anyc;
        for (int q = 0; q < 4; q++) {
          int idx[8], cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          for (int l = 0; l < 8; l++) idx[l] = (int)((ent[(size_t)k * kBR + 32 * w + 8 * q + l] >> 4) & 0xfff);
          for (int l = 0; l < 8; l++) {
            bool dup = false;
            for (int m = 0; m < l; m++) dup |= idx[m] == idx[l];
            if (!dup) cnt[idx[l] & 7]++;
          }
          int mx = 0;
          for (int c = 0; c < 8; c++) mx = std::max(mx, cnt[c]);
          wf += mx;
          groups += 1;
        }
      }
    }
  }
  stats[2] = groups > 0 ? wf / groups : 0.0;
  stats[3] = needed > 0 ? walked / needed : 0.0;
  stats[4] = comp_need > 0 ? comp_exec / comp_need : 0.0;
  return CGASM_OK;
}
