// host_mesh.cpp -- once-per-mesh host preprocessing behind the C ABI: the first-order
// node-node sparsity, the greedy CG1 element colouring and node->element adjacency.
//
// Same RESULTS as the reference routines, different algorithms:
//   sparsity  femtools/Sparsity_Patterns.F90:299-428 inserts into per-row sorted linked lists
//             (O(row length) per insert). Here: counting-sort the (node, element) incidences
//             into a node->element CSR, then per row gather the incident elements' nodes into
//             a small scratch, sort + unique. Rows are independent => OpenMP over rows.
//   colouring femtools/Colouring.F90:159-199 greedy in element order, lowest colour unused by
//             lower-numbered neighbours; neighbours = elements sharing a node
//             (make_sparsity_transpose, Sparsity_Patterns.F90:87-148). Inherently sequential;
//             here with a 64-bit colour mask per element instead of a Judy set.
#include "cgasm_internal.h"

#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cstring>
#include <numeric>

#include <omp.h>

namespace cgasm {

// node -> element adjacency (CSR, 0-based), elements ascending inside a node. Parallel counting sort without
// atomics: every thread owns a contiguous range of NODES and streams the whole connectivity, keeping the
// incidences of its own nodes (sequential reads, private writes; elements arrive in ascending order). The
// redundant reads are cheap next to the cache misses of a shared-counter scatter on a randomly numbered mesh.
void build_node_to_element(int n_nodes, int n_elements, int loc, const int* nd0 /*0-based, stride 4*/,
                           I64Vec& ptr, IVec& adj) {
  (void)loc;  // unused connectivity slots hold -1 and fall outside every range
  std::vector<int> cnt((size_t)n_nodes, 0);
  // node range of every chunk of kChunk elements: a thread skips the chunks that cannot hold its nodes (on a
  // mesh numbered with some locality that is nearly all of them)
  constexpr int kChunk = 512;
  const int nchunks = (n_elements + kChunk - 1) / kChunk;
  std::vector<int> cmin((size_t)nchunks), cmax((size_t)nchunks);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nchunks; c++) {
    int lo = n_nodes, hi = -1;
    const size_t k1 = (size_t)4 * std::min(n_elements, (c + 1) * kChunk);
    for (size_t k = (size_t)4 * c * kChunk; k < k1; k++) {
      const int v = nd0[k];
      if (v < 0) continue;
      lo = std::min(lo, v);
      hi = std::max(hi, v);
    }
    cmin[c] = lo;
    cmax[c] = hi;
  }
  auto scan = [&](int n0, int n1, auto&& visit) {
    const unsigned span = (unsigned)(n1 - n0);
    for (int c = 0; c < nchunks; c++) {
      if (cmax[c] < n0 || cmin[c] >= n1) continue;
      const size_t k1 = (size_t)4 * std::min(n_elements, (c + 1) * kChunk);
      for (size_t k = (size_t)4 * c * kChunk; k < k1; k++) {
        const unsigned v = (unsigned)(nd0[k] - n0);
        if (v < span) visit((size_t)n0 + v, (int)(k >> 2));
      }
    }
  };
#pragma omp parallel
  {
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
    const int n0 = (int)((int64_t)n_nodes * t / nt), n1 = (int)((int64_t)n_nodes * (t + 1) / nt);
    scan(n0, n1, [&](size_t node, int) { cnt[node]++; });
  }
  ptr.resize((size_t)n_nodes + 1);
  ptr[0] = 0;
  for (int i = 0; i < n_nodes; i++) ptr[(size_t)i + 1] = ptr[i] + cnt[i];
  adj.resize((size_t)ptr[n_nodes]);
#pragma omp parallel
  {
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
    const int n0 = (int)((int64_t)n_nodes * t / nt), n1 = (int)((int64_t)n_nodes * (t + 1) / nt);
    for (int i = n0; i < n1; i++) cnt[i] = 0;
    scan(n0, n1, [&](size_t node, int e) { adj[(size_t)ptr[node] + (size_t)cnt[node]++] = e; });
  }
}

// sorted distinct nodes of the elements around r -> scratch[0 .. count); returns the count. The ~4m gathered ids
// hold only ~m/2 + 3 distinct ones, so they are deduplicated through a 64-slot table before the sort.
static inline int row_columns(int r, int loc, const int* nd0, const I64Vec& n2e_ptr, const IVec& n2e,
                              std::vector<int>& scratch) {
  scratch.clear();
  const int64_t k0 = n2e_ptr[r], k1 = n2e_ptr[(size_t)r + 1];
  if (k1 == k0) return 0;
  int table[64];
  bool small = true;
  for (int q = 0; q < 64; q++) table[q] = -1;
  scratch.push_back(r);
  table[((uint32_t)r * 2654435761u) >> 26] = r;
  for (int64_t k = k0; k < k1 && small; k++) {
    const int* nd = nd0 + (size_t)4 * n2e[(size_t)k];
    for (int j = 0; j < loc; j++) {
      const int v = nd[j];
      unsigned p = ((uint32_t)v * 2654435761u) >> 26;
      while (table[p] != -1 && table[p] != v) p = (p + 1) & 63;
      if (table[p] == -1) {
        if (scratch.size() >= 40) {  // too many distinct nodes for the table: plain sort + unique below
          small = false;
          break;
        }
        table[p] = v;
        scratch.push_back(v);
      }
    }
  }
  if (small) {
    std::sort(scratch.begin(), scratch.end());
    return (int)scratch.size();
  }
  scratch.clear();
  for (int64_t k = k0; k < k1; k++) {
    const int* nd = nd0 + (size_t)4 * n2e[(size_t)k];
    for (int j = 0; j < loc; j++) scratch.push_back(nd[j]);
  }
  std::sort(scratch.begin(), scratch.end());
  return (int)(std::unique(scratch.begin(), scratch.end()) - scratch.begin());
}

// findrm (n+1), colm (nnz), 0-based; rows sorted ascending, unique. One pass: every thread owns a contiguous
// range of rows and appends their columns to its own buffer; the buffers are then copied into place.
// Returns nnz; if it does not fit 32 bits nothing is filled (the caller refuses the mesh).
int64_t build_sparsity(int n_nodes, int n_elements, int loc, const int* nd0,
                       const I64Vec& n2e_ptr, const IVec& n2e,
                       IVec& findrm, IVec& colm) {
  (void)n_elements;
  const int maxt = omp_get_max_threads();
  std::vector<std::vector<int>> chunk((size_t)maxt);
  std::vector<int> first((size_t)maxt + 1, n_nodes);
  std::vector<int> rowlen((size_t)n_nodes);
#pragma omp parallel
  {
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
    const int r0 = (int)((int64_t)n_nodes * t / nt), r1 = (int)((int64_t)n_nodes * (t + 1) / nt);
    first[t] = r0;
    std::vector<int>& out = chunk[t];
    out.reserve((size_t)(r1 - r0) * 16);
    std::vector<int> scratch;
    for (int r = r0; r < r1; r++) {
      const int len = row_columns(r, loc, nd0, n2e_ptr, n2e, scratch);
      rowlen[r] = len;
      out.insert(out.end(), scratch.begin(), scratch.begin() + len);
    }
  }
  int64_t acc = 0;
  for (int r = 0; r < n_nodes; r++) acc += rowlen[r];
  if (acc >= (int64_t)1 << 31) return acc;
  findrm.resize((size_t)n_nodes + 1);
  acc = 0;
  for (int r = 0; r < n_nodes; r++) {
    findrm[r] = (int)acc;
    acc += rowlen[r];
  }
  findrm[n_nodes] = (int)acc;
  colm.resize((size_t)acc);
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < maxt; t++)
    if (!chunk[t].empty()) std::copy(chunk[t].begin(), chunk[t].end(), colm.begin() + findrm[first[t]]);
  return acc;
}

int64_t count_nnz(int n_nodes, int loc, const int* nd0, const I64Vec& n2e_ptr,
                  const IVec& n2e) {
  int64_t total = 0;
#pragma omp parallel reduction(+ : total)
  {
    std::vector<int> scratch;
#pragma omp for schedule(static)
    for (int r = 0; r < n_nodes; r++) total += row_columns(r, loc, nd0, n2e_ptr, n2e, scratch);
  }
  return total;
}

// Greedy colouring in element order. colour_of is 0-based; returns the number of colours, or
// -1 if more than 64*kWords colours would be needed.
int greedy_colouring(int n_elements, int loc, const int* nd0, const I64Vec& n2e_ptr,
                     const IVec& n2e, std::vector<int>& colour_of) {
  constexpr int kWords = 4;  // up to 256 colours
  colour_of.assign((size_t)n_elements, -1);
  int ncol = 0;
  for (int e = 0; e < n_elements; e++) {
    uint64_t used[kWords] = {0, 0, 0, 0};
    const int* nd = nd0 + (size_t)4 * e;
    for (int i = 0; i < loc; i++) {
      const int node = nd[i];
      for (int64_t k = n2e_ptr[node]; k < n2e_ptr[node + 1]; k++) {
        const int e2 = n2e[(size_t)k];
        if (e2 >= e) break;  // adjacency is ascending: only lower-numbered neighbours count
        const int c = colour_of[e2];
        used[c >> 6] |= (uint64_t)1 << (c & 63);
      }
    }
    int c = -1;
    for (int w = 0; w < kWords && c < 0; w++)
      if (~used[w]) c = w * 64 + __builtin_ctzll(~used[w]);
    if (c < 0) return -1;
    colour_of[e] = c;
    if (c + 1 > ncol) ncol = c + 1;
  }
  return ncol;
}

// colour_sets: offsets + element ids ascending inside each colour (0-based).
void colour_sets(int n_elements, int ncol, const std::vector<int>& colour_of,
                 std::vector<int>& colour_ptr, std::vector<int>& colour_elements) {
  colour_ptr.assign((size_t)ncol + 1, 0);
  for (int e = 0; e < n_elements; e++) colour_ptr[(size_t)colour_of[e] + 1]++;
  for (int c = 0; c < ncol; c++) colour_ptr[c + 1] += colour_ptr[c];
  colour_elements.resize((size_t)n_elements);
  std::vector<int> fill(colour_ptr.begin(), colour_ptr.end() - 1);
  for (int e = 0; e < n_elements; e++) colour_elements[(size_t)fill[colour_of[e]]++] = e;
}

// order = the stable ascending order of key (ties keep ascending index): parallel LSD radix sort over the
// bits actually used, 11 bits per pass (std::stable_sort of 17 M keys was ~2 s of the S3 set-up).
static void radix_order(const std::vector<uint64_t>& key, std::vector<int>& order) {
  const size_t n = key.size();
  order.resize(n);
  uint64_t kmax = 0;
#pragma omp parallel for schedule(static) reduction(max : kmax)
  for (size_t i = 0; i < n; i++) kmax = std::max(kmax, key[i]);
  int bits = 0;
  while (bits < 64 && (kmax >> bits)) bits++;
  constexpr int kDigit = 11, kBuckets = 1 << kDigit;
  std::vector<int> a(n), b(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) a[i] = (int)i;
  const int maxt = omp_get_max_threads();
  std::vector<size_t> hist((size_t)maxt * kBuckets);
  for (int shift = 0; shift < std::max(bits, 1); shift += kDigit) {
    std::fill(hist.begin(), hist.end(), 0);
    int nt_used = 1;
#pragma omp parallel
    {
      const int t = omp_get_thread_num(), nt = omp_get_num_threads();
      const size_t i0 = n * t / nt, i1 = n * (t + 1) / nt;
      size_t* ht = hist.data() + (size_t)t * kBuckets;
      for (size_t i = i0; i < i1; i++) ht[(key[a[i]] >> shift) & (kBuckets - 1)]++;
#pragma omp barrier
#pragma omp single
      {
        nt_used = nt;
        size_t run = 0;
        for (int d = 0; d < kBuckets; d++)
          for (int q = 0; q < nt; q++) {
            const size_t c = hist[(size_t)q * kBuckets + d];
            hist[(size_t)q * kBuckets + d] = run;
            run += c;
          }
      }
      for (size_t i = i0; i < i1; i++) b[ht[(key[a[i]] >> shift) & (kBuckets - 1)]++] = a[i];
    }
    (void)nt_used;
    a.swap(b);
  }
  order.swap(a);
}

// Locality order of the nodes (used to group CSR rows into tiles / gather blocks).
void morton_order(const Handle* h, std::vector<int>& order, MortonFrame& F) {
  const int n = h->n_nodes, dim = h->dim;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  {
    double l0 = 1e300, l1 = 1e300, l2 = 1e300, h0 = -1e300, h1 = -1e300, h2 = -1e300;
    const double* X = h->h_X.data();
#pragma omp parallel for schedule(static) reduction(min : l0, l1, l2) reduction(max : h0, h1, h2)
    for (int i = 0; i < n; i++) {
      const double* x = X + (size_t)dim * i;
      l0 = std::min(l0, x[0]);
      h0 = std::max(h0, x[0]);
      l1 = std::min(l1, x[1]);
      h1 = std::max(h1, x[1]);
      if (dim == 3) {
        l2 = std::min(l2, x[2]);
        h2 = std::max(h2, x[2]);
      }
    }
    lo[0] = l0; lo[1] = l1; lo[2] = l2;
    hi[0] = h0; hi[1] = h1; hi[2] = h2;
  }
  // Quantise to a lattice about as fine as the mesh itself, per axis: the spacing ratio between axes
  // comes from the mean |edge component| over all element edges, the absolute scale from
  // prod_a(ext_a / h_a + 1) = n (points per axis), and cells_a = round(ext_a / h_a): exact for a
  // lattice also when the spacing differs between axes, e.g. slab partitions of a box.
  // Jittered lattice nodes then snap to their own lattice point, so fixed-count cuts of the Morton
  // sequence are aligned bricks (measured tile redundancy 1.31 at 1024 rows vs 1.51 with a fine
  // quantisation).
  double mean_edge[3] = {0, 0, 0};
  {
    const int loc = h->loc, ne = h->n_elements;
    // a sample is enough, but not a strided one: structured meshes repeat their element shapes with a
    // short period (6 Kuhn tets per cube), and a stride sharing a factor with it sees one shape only
    const long long nsample = std::min<long long>(ne, 4000000);
    double acc0 = 0, acc1 = 0, acc2 = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc0, acc1, acc2)
    for (long long s = 0; s < nsample; s++) {
      const int e = nsample == ne ? (int)s : (int)(((unsigned long long)s * 2654435761ull + 12345ull) % (unsigned long long)ne);
      const int* nd = h->h_nd0.data() + (size_t)4 * e;
      for (int i = 0; i < loc; i++)
        for (int j = i + 1; j < loc; j++) {
          const double* xi = &h->h_X[(size_t)dim * nd[i]];
          const double* xj = &h->h_X[(size_t)dim * nd[j]];
          acc0 += std::fabs(xi[0] - xj[0]);
          acc1 += std::fabs(xi[1] - xj[1]);
          if (dim == 3) acc2 += std::fabs(xi[2] - xj[2]);
        }
    }
    mean_edge[0] = acc0;
    mean_edge[1] = acc1;
    mean_edge[2] = acc2;
  }
  // h_a = c * mean_edge_a with c solving prod_a(ext_a / h_a + 1) = n (points per axis = cells + 1)
  double lo_c = 1e-300, hi_c = 1e300;
  {
    auto points = [&](double c) {
      double p = 1.0;
      for (int a = 0; a < dim; a++)
        if (hi[a] > lo[a] && mean_edge[a] > 0.0) p *= (hi[a] - lo[a]) / (c * mean_edge[a]) + 1.0;
      return p;
    };
    // bracket: points(c) decreases with c
    double c0 = 1.0;
    while (points(c0) < (double)n && c0 > 1e-280) c0 *= 0.5;
    lo_c = c0;
    hi_c = c0;
    while (points(hi_c) > (double)n && hi_c < 1e280) hi_c *= 2.0;
    for (int it = 0; it < 200; it++) {
      const double mid = 0.5 * (lo_c + hi_c);
      if (points(mid) > (double)n) lo_c = mid;
      else hi_c = mid;
    }
  }
  const double cscale = 0.5 * (lo_c + hi_c);
  const double maxcells = dim == 3 ? 2097151.0 : 4294967295.0;
  F.dim = dim;
  for (int a = 0; a < dim; a++) {
    F.lo[a] = lo[a];
    F.scale[a] = 0.0;
    if (hi[a] > lo[a] && mean_edge[a] > 0.0) {
      const double ha = cscale * mean_edge[a];
      const double cells = std::min(maxcells, std::max(1.0, std::round((hi[a] - lo[a]) / ha)));
      F.scale[a] = cells / (hi[a] - lo[a]);
    }
  }
  std::vector<uint64_t> key((size_t)n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) key[i] = F.key_round(&h->h_X[(size_t)dim * i]);
  radix_order(key, order);
}


namespace {
// set of node ids with O(1) clear (generation stamps), open addressing; grows on demand
struct NodeSet {
  std::vector<int> key;
  std::vector<unsigned> gen;
  unsigned cur = 1;
  size_t mask = 0, count = 0;
  explicit NodeSet(size_t cap = 4096) { reset(cap); }
  void reset(size_t cap) {
    key.assign(cap, 0);
    gen.assign(cap, 0);
    mask = cap - 1;
    cur = 1;
    count = 0;
  }
  void clear() {
    count = 0;
    if (++cur == 0) {
      std::fill(gen.begin(), gen.end(), 0u);
      cur = 1;
    }
  }
  static size_t hash(int v) { return (size_t)((uint32_t)v * 2654435761u) >> 7; }
  bool insert(int node) {  // true if it was not in the set
    if (2 * (count + 1) > mask + 1) grow();
    size_t p = hash(node) & mask;
    while (gen[p] == cur) {
      if (key[p] == node) return false;
      p = (p + 1) & mask;
    }
    key[p] = node;
    gen[p] = cur;
    count++;
    return true;
  }
  void grow() {
    std::vector<int> live;
    live.reserve(count);
    for (size_t p = 0; p <= mask; p++)
      if (gen[p] == cur) live.push_back(key[p]);
    reset(2 * (mask + 1));
    for (int v : live) insert(v);
  }
};
}  // namespace

int form_row_blocks(const Handle* h, const std::vector<int>& order, const MortonFrame& F, int block_rows,
                    std::vector<int>& rows) {
  const int n = h->n_nodes;
  int nblocks = 0;
  // Row blocks = runs of the Morton sequence. Cutting every block_rows nodes drifts off the lattice bricks as
  // soon as one brick is partly filled (domain boundary), and a block that straddles two bricks touches
  // up to twice as many distinct nodes. So: group the sequence by brick (block_rows lattice points: the low
  // log2(block_rows) key bits), merge consecutive groups while they fit AND while the distinct nodes the
  // block touches (the union of its CSR rows: what the staged kernels keep in shared memory) stay
  // near what a full brick needs, split over-full groups; keep that only if the padding stays below 15 %.
  // The merge runs in parallel over fixed chunks of kChunk consecutive groups (a block never spans two
  // chunks; the chunking does not depend on the thread count, so neither does the result).
  int shift = 0;
  while ((1 << shift) < block_rows) shift++;
  std::vector<uint64_t> brick((size_t)n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) brick[i] = F.key_round(&h->h_X[(size_t)h->dim * order[i]]) >> shift;
  std::vector<int> gstart;  // start of every brick group in `order`
  for (int i = 0; i < n; i++)
    if (i == 0 || brick[i] != brick[i - 1] || i - gstart.back() == block_rows) gstart.push_back(i);
  gstart.push_back(n);
  const int ng = (int)gstart.size() - 1;
  const int* findrm = h->h_findrm.data();
  const int* colm = h->h_colm.data();
  // distinct columns of rows order[i0..i1) that are not in the set yet (they are added)
  auto touch = [&](NodeSet& S, int i0, int i1) {
    int fresh = 0;
    for (int i = i0; i < i1; i++) {
      const int r = order[i];
      for (int q = findrm[r]; q < findrm[r + 1]; q++) fresh += S.insert(colm[q]);
    }
    return fresh;
  };
  // what a full brick touches (median over a sample of full groups)
  int cap = 1 << 30;
  {
    std::vector<int> sizes;
    NodeSet S;
    const int step = std::max(1, ng / 2000);
    for (int g = 0; g < ng; g += step)
      if (gstart[g + 1] - gstart[g] == block_rows) {
        S.clear();
        sizes.push_back(touch(S, gstart[g], gstart[g + 1]));
      }
    if (!sizes.empty()) {
      std::nth_element(sizes.begin(), sizes.begin() + sizes.size() / 2, sizes.end());
      cap = (int)(1.15 * sizes[sizes.size() / 2]);
    }
  }
  constexpr int kChunk = 2048;
  const int nchunks = (ng + kChunk - 1) / kChunk;
  std::vector<std::vector<int>> cuts((size_t)std::max(nchunks, 1));  // block starts (positions in `order`) per chunk
#pragma omp parallel
  {
    NodeSet S;
#pragma omp for schedule(dynamic, 1)
    for (int ch = 0; ch < nchunks; ch++) {
      std::vector<int>& cut = cuts[ch];
      int count = 0, touched = 0;
      S.clear();
      for (int g = ch * kChunk; g < std::min(ng, (ch + 1) * kChunk); g++) {
        const int len = gstart[g + 1] - gstart[g];
        bool start_block = count == 0;
        int fresh = 0;
        if (!start_block) {
          fresh = touch(S, gstart[g], gstart[g + 1]);  // tentatively part of the current block
          if (count + len > block_rows || touched + fresh > cap) start_block = true;
        }
        if (start_block) {
          S.clear();
          fresh = touch(S, gstart[g], gstart[g + 1]);
          cut.push_back(gstart[g]);
          count = 0;
          touched = 0;
        }
        touched += fresh;
        count += len;
      }
    }
  }
  std::vector<int> cut;
  for (int ch = 0; ch < nchunks; ch++) cut.insert(cut.end(), cuts[ch].begin(), cuts[ch].end());
  cut.push_back(n);
  const int nb_aligned = (int)cut.size() - 1, nb_plain = (n + block_rows - 1) / block_rows;
  if ((double)nb_aligned <= 1.15 * nb_plain && !getenv("CGASM_GATHER_PLAIN_BLOCKS")) {
    nblocks = nb_aligned;
    rows.assign((size_t)nb_aligned * block_rows, -1);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb_aligned; b++)
      std::copy(order.begin() + cut[b], order.begin() + cut[b + 1], rows.begin() + (size_t)b * block_rows);
  } else {
    nblocks = nb_plain;
    rows.assign((size_t)nb_plain * block_rows, -1);
    std::copy(order.begin(), order.end(), rows.begin());
  }
  if (getenv("CGASM_DEBUG"))
    fprintf(stderr, "[cgasm] row blocks: %d (plain %d, brick-aligned %d, node cap %d)\n", nblocks, nb_plain, nb_aligned, cap);
  return nblocks;
}

// Geometric keys + the order of the rows inside every block.
//   geokey[node] = the node's lattice point, lexicographic (z, y, x): orders the rows of a block, the block's staged node
//   list and the labels of the strip builder independently of the caller's numbering.
//   Rows of a block: by descending number of incident elements, then by key, then by id. Rows with equally long element
//   lists share a warp, so a warp of the row-owner kernels stops after ITS longest row instead of the block's
//   (unstructured meshes: node degrees 8-58 made a warp execute 1.9x the element computations its lanes needed); where
//   the degree is uniform (the interior of a structured mesh) this is the lexicographic order of the brick -- the plain
//   order by global id on a lexicographically numbered mesh -- whatever the numbering is.
void order_block_rows(Handle* h, const MortonFrame& F, std::vector<int>& rows, int nblocks) {
  const int n = h->n_nodes, dim = h->dim;
  h->geokey.assign((size_t)n, 0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    int64_t q[3] = {0, 0, 0};
    for (int a = 0; a < dim; a++) q[a] = (int64_t)std::llround((h->h_X[(size_t)dim * i + a] - F.lo[a]) * F.scale[a]) & 0x1fffff;
    h->geokey[i] = q[2] << 42 | q[1] << 21 | q[0];
  }
  const int64_t* np = h->n2e_ptr.data();
  const int64_t* gk = h->geokey.data();
  const int br = nblocks > 0 ? (int)(rows.size() / (size_t)nblocks) : 0;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; b++) {
    int* rb = rows.data() + (size_t)b * br;
    std::sort(rb, std::find(rb, rb + br, -1), [np, gk](int x, int y) {
      const int64_t dx = np[x + 1] - np[x], dy = np[y + 1] - np[y];
      if (dx != dy) return dx > dy;
      return gk[x] != gk[y] ? gk[x] < gk[y] : x < y;
    });
  }
}

}  // namespace cgasm
