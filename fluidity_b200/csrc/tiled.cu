// tiled.cu -- CGASM_SCATTER_TILED: node-tile owner-computes assembly.
//
// The reference scatters every element contribution straight into the global CSR
// (femtools/Sparse_Tools.F90:2680-2703, Sparse_Tools_Petsc.F90:848-879) and, under OpenMP,
// serialises conflicting elements with a mesh-wide colouring (femtools/Colouring.F90).
// On a B200 neither maps well (measured, profiles/): global FP64 reds are bound by the SM's
// REDG issue rate (~1.3 cycles/lane, ~2e11 reds/s for the whole chip) and mesh-wide colours
// destroy locality. This variant re-uses the reference's own PARALLEL decomposition idea one
// level down -- node ownership + redundant assembly of halo elements
// (every MPI rank assembles all its elements and keeps only owned rows):
//
//   * nodes are ordered along a Morton curve of their coordinates and cut into tiles of
//     <= max_rows rows whose CSR rows fit in shared memory;
//   * one CTA owns one tile: it visits EVERY element touching an owned node (elements on a
//     tile surface are visited by each tile they touch: ~1.3-1.4x redundant element math,
//     zero inter-CTA communication), accumulates the owned rows in shared memory and writes
//     each CSR value / rhs entry exactly once with coalesced stores -- no atomics, no
//     pre-zeroing of the outputs, bitwise reproducible from run to run;
//   * inside the CTA, conflicting read-modify-writes are serialised the way Colouring.F90
//     does it, but per tile: elements are greedily coloured (balanced, conflicts only through
//     OWNED nodes) and the colours become __syncthreads-separated phases.
//
// The per-element record (node ids, tile-local row of each node, slot of every (i,j) inside
// that row) is precomputed once per mesh+sparsity, so the hot loop does no searching at all
// (the reference bisects every entry, Sparse_Tools.F90:2438-2497).
#include "cgasm_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace cgasm {

struct TileClassPlan {
  int nb = 0, nvec = 0;     // accumulated matrix blocks / vector components this plan is sized for
  int ntiles = 0;
  int max_tile_entries = 0, max_tile_rows = 0, max_phase = 0;  // max_phase: clusters in the largest phase
  int cluster = 1;  // elements per cluster (records are padded to whole clusters per tile)
  size_t smem_bytes = 0;
  long long n_tile_elements = 0;
  // device arrays
  int* d_tile_row_ptr = nullptr;    // [ntiles+1] into rows
  int* d_rows = nullptr;            // owned global node per tile row (ascending inside a tile)
  int* d_rowoff = nullptr;          // offset of the row inside the tile's matrix accumulator
  int* d_tile_entries = nullptr;    // [ntiles] CSR entries owned by the tile
  int* d_tile_phase_off = nullptr;  // [ntiles+1] into phase_ptr
  int* d_phase_ptr = nullptr;       // per tile: nphase+1 offsets in CLUSTER units (records = cluster*K + j)
  int* d_tile_run_ptr = nullptr;    // [ntiles+1] into runs
  int2* d_runs = nullptr;           // {first tile-local row, nrows}: consecutive global node ids
  int4* d_el_nodes = nullptr;       // per tile-element: global node ids (0-based)
  uint2* d_el_rows = nullptr;       // 4 x uint16 tile-local row (0xFFFF = not owned)
  uint4* d_el_slots = nullptr;      // 16 x uint8: slot of (i,j) inside row i  (i*4+j)
};

struct TilePlan {
  TileClassPlan cls[2];  // [0]: one matrix block, [1]: dim matrix blocks
  bool built[2] = {false, false};
};

static void free_class(TileClassPlan& p) {
  void* ptrs[] = {p.d_tile_row_ptr, p.d_rows, p.d_rowoff, p.d_tile_entries, p.d_tile_phase_off,
                  p.d_phase_ptr, p.d_tile_run_ptr, p.d_runs, p.d_el_nodes, p.d_el_rows, p.d_el_slots};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  p = TileClassPlan();
}

void tiles_free(Handle* h) {
  if (!h->tiles) return;
  for (auto& c : h->tiles->cls) free_class(c);
  delete h->tiles;
  h->tiles = nullptr;
}

template <class T>
static int upload(T** d, const std::vector<T>& v) {
  CG_CUDA(cudaMalloc(d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CG_CUDA(cg_upload(*d, v.data(), sizeof(T) * v.size()));
  return CGASM_OK;
}

// Fills el_rows-derived slots on the device: slot of column node j inside CSR row of node i.
__global__ void tile_slots_kernel(long long n, const int4* __restrict__ el_nodes, int loc,
                                  const int* __restrict__ findrm, const int* __restrict__ colm,
                                  uint4* __restrict__ el_slots) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int4 nd = el_nodes[k];
  const int nodes[4] = {nd.x, nd.y, nd.z, nd.w};
  unsigned packed[4] = {0, 0, 0, 0};
  for (int i = 0; i < loc; i++) {
    const int s = findrm[nodes[i]], e = findrm[nodes[i] + 1];
    for (int q = s; q < e; q++) {
      const int c = colm[q];
      for (int j = 0; j < loc; j++)
        if (c == nodes[j]) packed[i] |= (unsigned)((q - s) & 0xff) << (8 * j);
    }
  }
  el_slots[k] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

static int build_class(Handle* h, TileClassPlan& P, int nb, int nvec, const std::vector<int>& order,
                       const MortonFrame& F) {
  const int loc = h->loc, n_nodes = h->n_nodes;
  const int* nd0 = h->h_nd0.data();
  const IVec& fr = h->h_findrm;
  int max_rows = 1024;
  size_t smem_cap = 200 * 1024;
  if (const char* s = getenv("CGASM_TILE_ROWS")) max_rows = std::max(32, atoi(s));
  if (const char* s = getenv("CGASM_TILE_SMEM_KB")) smem_cap = (size_t)std::max(16, atoi(s)) * 1024;
  max_rows = std::min(max_rows, 65534);
  P.nb = nb;
  P.nvec = nvec;

  // ---- cut the Morton sequence into tiles ---------------------------------------------------
  std::vector<int> tile_row_ptr{0}, rows;
  rows.reserve((size_t)n_nodes);
  {
    size_t entries = 0;
    int nrows = 0;
    auto bytes = [&](size_t ent, int nr) {
      return sizeof(double) * ((size_t)nb * ent + (size_t)nvec * nr) + sizeof(int) * (size_t)(nr + 1);
    };
    for (int k = 0; k < n_nodes; k++) {
      const int node = order[k];
      const int len = fr[node + 1] - fr[node];
      if (len > 255) CG_FAIL(CGASM_EUNSUPPORTED, "CSR row longer than 255 entries: tiled scatter slot index is 8 bit");
      if (nrows && (nrows == max_rows || bytes(entries + len, nrows + 1) > smem_cap)) {
        tile_row_ptr.push_back((int)rows.size());
        entries = 0;
        nrows = 0;
      }
      if (bytes(len, 1) > smem_cap) CG_FAIL(CGASM_EUNSUPPORTED, "a single CSR row does not fit in shared memory");
      rows.push_back(node);
      entries += len;
      nrows++;
    }
    tile_row_ptr.push_back((int)rows.size());
  }
  const int ntiles = (int)tile_row_ptr.size() - 1;
  P.ntiles = ntiles;

  std::vector<int> rowoff((size_t)n_nodes), tile_entries((size_t)ntiles), tile_of_node((size_t)n_nodes),
      lrow_of_node((size_t)n_nodes);
  std::vector<int> tile_run_ptr((size_t)ntiles + 1, 0);
  std::vector<std::vector<int2>> runs_t((size_t)ntiles);
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < ntiles; t++) {
    int* b = rows.data() + tile_row_ptr[t];
    int* e = rows.data() + tile_row_ptr[t + 1];
    std::sort(b, e);  // ascending global id => contiguous CSR ranges on output
    int off = 0;
    for (int* p = b; p < e; p++) {
      const int l = (int)(p - b);
      rowoff[tile_row_ptr[t] + l] = off;
      off += fr[*p + 1] - fr[*p];
      tile_of_node[*p] = t;
      lrow_of_node[*p] = l;
      if (l && *p == *(p - 1) + 1) runs_t[t].back().y++;
      else runs_t[t].push_back(make_int2(l, 1));
    }
    tile_entries[t] = off;
  }
  std::vector<int2> runs;
  for (int t = 0; t < ntiles; t++) {
    tile_run_ptr[t] = (int)runs.size();
    runs.insert(runs.end(), runs_t[t].begin(), runs_t[t].end());
    runs_t[t].clear();
    runs_t[t].shrink_to_fit();
  }
  tile_run_ptr[ntiles] = (int)runs.size();
  P.max_tile_entries = *std::max_element(tile_entries.begin(), tile_entries.end());
  for (int t = 0; t < ntiles; t++) P.max_tile_rows = std::max(P.max_tile_rows, tile_row_ptr[t + 1] - tile_row_ptr[t]);
  P.smem_bytes = sizeof(double) * ((size_t)nb * P.max_tile_entries + (size_t)nvec * P.max_tile_rows) +
                 sizeof(int) * (size_t)(P.max_tile_rows + 1);
  P.smem_bytes = (P.smem_bytes + 15) & ~(size_t)15;

  // ---- per tile: element set -> clusters of K spatially adjacent elements -> balanced greedy
  // colouring of the CLUSTERS (two clusters conflict iff they share an OWNED node); a colour is a
  // __syncthreads-separated phase in which one thread walks one cluster element by element.
  // Fewer, fatter phases than colouring single elements (a node has ~24 incident tets but only
  // ~8 incident 6-tet clusters), and same-thread RMWs inside a cluster need no ordering at all.
  int K = 1;  // measured: sequential clusters lengthen the per-thread critical path (latency-bound)
  if (const char* s = getenv("CGASM_TILE_CLUSTER")) K = std::min(64, std::max(1, atoi(s)));
  P.cluster = K;
  const int dim = h->dim;
  std::vector<long long> tile_cl_count((size_t)ntiles + 1, 0);  // clusters per tile (prefix later)
  std::vector<std::vector<int>> tile_elems((size_t)ntiles);     // element ids, cluster-major, phase order, -1 = pad
  std::vector<std::vector<int>> tile_phase((size_t)ntiles);     // nphase+1 offsets in clusters
#pragma omp parallel
  {
    std::vector<int> els, colour, cnt, start, perm;
    std::vector<uint64_t> mask, ekey;
#pragma omp for schedule(dynamic, 16)
    for (int t = 0; t < ntiles; t++) {
      const int r0 = tile_row_ptr[t], r1 = tile_row_ptr[t + 1], nr = r1 - r0;
      els.clear();
      for (int r = r0; r < r1; r++) {
        const int node = rows[r];
        for (int64_t k = h->n2e_ptr[node]; k < h->n2e_ptr[node + 1]; k++) els.push_back(h->n2e[(size_t)k]);
      }
      std::sort(els.begin(), els.end());
      els.erase(std::unique(els.begin(), els.end()), els.end());
      const int ne = (int)els.size();
      // spatial order: Morton key of the containing lattice cell of the centroid
      ekey.resize((size_t)ne);
      perm.resize((size_t)ne);
      for (int k = 0; k < ne; k++) {
        const int* nd = nd0 + (size_t)4 * els[k];
        double c[3] = {0, 0, 0};
        for (int i = 0; i < loc; i++)
          for (int a = 0; a < dim; a++) c[a] += h->h_X[(size_t)dim * nd[i] + a];
        for (int a = 0; a < dim; a++) c[a] /= loc;
        ekey[k] = F.key_floor(c);
        perm[k] = k;
      }
      std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return ekey[a] < ekey[b]; });
      const int ncl = (ne + K - 1) / K;
      mask.assign((size_t)nr, 0);
      colour.assign((size_t)ncl, 0);
      cnt.assign(64, 0);
      int palette = 8, ncol = 0;
      bool ok = true;
      for (int c = 0; c < ncl && ok; c++) {
        uint64_t used = 0;
        for (int k = c * K; k < std::min(ne, (c + 1) * K); k++) {
          const int* nd = nd0 + (size_t)4 * els[perm[k]];
          for (int i = 0; i < loc; i++)
            if (tile_of_node[nd[i]] == t) used |= mask[lrow_of_node[nd[i]]];
        }
        int best = -1;
        for (int q = 0; q < palette; q++)
          if (!(used >> q & 1) && (best < 0 || cnt[q] < cnt[best])) best = q;
        if (best < 0) {
          for (int q = palette; q < 64; q++)
            if (!(used >> q & 1)) {
              best = q;
              break;
            }
          if (best < 0) {
            ok = false;
            break;
          }
          palette = best + 1;
        }
        colour[c] = best;
        cnt[best]++;
        ncol = std::max(ncol, best + 1);
        for (int k = c * K; k < std::min(ne, (c + 1) * K); k++) {
          const int* nd = nd0 + (size_t)4 * els[perm[k]];
          for (int i = 0; i < loc; i++)
            if (tile_of_node[nd[i]] == t) mask[lrow_of_node[nd[i]]] |= (uint64_t)1 << best;
        }
      }
      if (!ok) {
        tile_phase[t].clear();  // flagged below
        continue;
      }
      start.assign((size_t)ncol + 1, 0);
      for (int q = 0; q < ncol; q++) start[q + 1] = start[q] + cnt[q];
      tile_phase[t] = start;
      tile_elems[t].assign((size_t)ncl * K, -1);
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (int c = 0; c < ncl; c++) {
        const int dst = fill[colour[c]]++;
        for (int k = c * K; k < std::min(ne, (c + 1) * K); k++) tile_elems[t][(size_t)dst * K + (k - c * K)] = els[perm[k]];
      }
      tile_cl_count[t + 1] = ncl;
    }
  }
  for (int t = 0; t < ntiles; t++) {
    if (tile_phase[t].empty()) CG_FAIL(CGASM_EUNSUPPORTED, "a tile needs more than 64 colours");
    tile_cl_count[t + 1] += tile_cl_count[t];
  }
  const long long ntel = tile_cl_count[ntiles] * K;
  if (ntel >= ((long long)1 << 31)) CG_FAIL(CGASM_EUNSUPPORTED, "tile element records exceed 2^31");
  P.n_tile_elements = ntel;

  std::vector<int> tile_phase_off((size_t)ntiles + 1, 0), phase_ptr;
  for (int t = 0; t < ntiles; t++) {
    tile_phase_off[t] = (int)phase_ptr.size();
    for (int v : tile_phase[t]) phase_ptr.push_back((int)(tile_cl_count[t] + v));
    for (size_t c = 0; c + 1 < tile_phase[t].size(); c++)
      P.max_phase = std::max(P.max_phase, tile_phase[t][c + 1] - tile_phase[t][c]);
  }
  tile_phase_off[ntiles] = (int)phase_ptr.size();

  std::vector<int4> el_nodes((size_t)ntel);
  std::vector<uint2> el_rows((size_t)ntel);
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < ntiles; t++) {
    const long long base = tile_cl_count[t] * K;
    int last_valid = -1;
    for (size_t k = 0; k < tile_elems[t].size(); k++) {
      const int e = tile_elems[t][k];
      unsigned short lr[4] = {0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF};
      // padding record: a valid element's nodes (so the math stays finite) but no owned row
      const int* nd = nd0 + (size_t)4 * (e >= 0 ? e : (last_valid >= 0 ? last_valid : tile_elems[t][0]));
      if (e >= 0) {
        last_valid = e;
        for (int i = 0; i < loc; i++)
          if (tile_of_node[nd[i]] == t) lr[i] = (unsigned short)lrow_of_node[nd[i]];
      }
      // unused 4th node of a triangle: repeat node 0 so gathers stay in range
      el_nodes[(size_t)(base + k)] = make_int4(nd[0], nd[1], nd[2], loc == 4 ? nd[3] : nd[0]);
      el_rows[(size_t)(base + k)] = make_uint2((unsigned)lr[0] | (unsigned)lr[1] << 16, (unsigned)lr[2] | (unsigned)lr[3] << 16);
    }
  }

  int st;
  if ((st = upload(&P.d_tile_row_ptr, tile_row_ptr)) || (st = upload(&P.d_rows, rows)) ||
      (st = upload(&P.d_rowoff, rowoff)) || (st = upload(&P.d_tile_entries, tile_entries)) ||
      (st = upload(&P.d_tile_phase_off, tile_phase_off)) || (st = upload(&P.d_phase_ptr, phase_ptr)) ||
      (st = upload(&P.d_tile_run_ptr, tile_run_ptr)) || (st = upload(&P.d_runs, runs)) ||
      (st = upload(&P.d_el_nodes, el_nodes)) || (st = upload(&P.d_el_rows, el_rows)))
    return st;
  CG_CUDA(cudaMalloc(&P.d_el_slots, sizeof(uint4) * (size_t)std::max<long long>(ntel, 1)));
  {
    const int block = 256;
    const long long grid = (ntel + block - 1) / block;
    tile_slots_kernel<<<(unsigned)grid, block, 0, h->stream>>>(ntel, P.d_el_nodes, loc, h->d_findrm, h->d_colm, P.d_el_slots);
    h->launches++;
    CG_CUDA(cudaStreamSynchronize(h->stream));
  }
  return CGASM_OK;
}

int tiles_build(Handle* h) {
  if (!h->have_X) CG_FAIL(CGASM_ESTATE, "tiled scatter orders nodes by their coordinates: set coordinates first");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "tiled scatter needs the sparsity first");
  tiles_free(h);
  h->tiles = new TilePlan();
  std::vector<int> order;
  MortonFrame F;
  morton_order(h, order, F);
  int st = build_class(h, h->tiles->cls[0], 1, 2 * h->dim, order, F);
  if (st) {
    tiles_free(h);
    return st;
  }
  h->tiles->built[0] = true;
  return CGASM_OK;
}

static int ensure_class(Handle* h, int c) {
  if (h->tiles->built[c]) return CGASM_OK;
  std::vector<int> order;
  MortonFrame F;
  morton_order(h, order, F);
  int st = build_class(h, h->tiles->cls[c], c == 0 ? 1 : h->dim, 2 * h->dim, order, F);
  if (st) return st;
  h->tiles->built[c] = true;
  return CGASM_OK;
}

// ---- device side ---------------------------------------------------------------------------------
struct TileArgs {
  const int* __restrict__ tile_row_ptr;
  const int* __restrict__ rows;
  const int* __restrict__ rowoff;
  const int* __restrict__ tile_entries;
  const int* __restrict__ tile_phase_off;
  const int* __restrict__ phase_ptr;
  const int* __restrict__ tile_run_ptr;
  const int2* __restrict__ runs;
  const int4* __restrict__ el_nodes;
  const uint2* __restrict__ el_rows;
  const uint4* __restrict__ el_slots;
  const int* __restrict__ findrm;
  size_t nnz;
  int cluster;
};

__device__ __forceinline__ unsigned lrow_of(const uint2& r, int i) {
  const unsigned w = i < 2 ? r.x : r.y;
  return (i & 1) ? (w >> 16) : (w & 0xffffu);
}
__device__ __forceinline__ unsigned slot_of(const uint4& s, int i, int j) {
  const unsigned w = i == 0 ? s.x : (i == 1 ? s.y : (i == 2 ? s.z : s.w));
  return (w >> (8 * j)) & 0xffu;
}

// Shared-memory layout of one tile: [NB][entries] matrix accumulators, [rows][NVEC] vector
// accumulators, [rows+1] row offsets.
template <int NB, int NVEC>
struct TileSmem {
  double* mat;
  double* vec;
  int* off;
  __device__ TileSmem(unsigned char* base, int max_entries, int max_rows) {
    mat = reinterpret_cast<double*>(base);
    vec = mat + (size_t)NB * max_entries;
    off = reinterpret_cast<int*>(vec + (size_t)NVEC * max_rows);
  }
};

// ABS: 0 = no absorption (the dim diagonal blocks are identical: ONE accumulator, written dim
// times), 1 = lumped absorption (blocks differ on the diagonal), 2 = full absorption matrix.
// MLD: masslump differs per component (pressure-corrected lumped absorption).
template <int DIM, int ABS, bool MLD>
__global__ void __launch_bounds__(256, 1)
tiled_momentum_kernel(const MomentumArgs A, const TileArgs T, int max_entries, int max_rows,
                      double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  constexpr int LOC = DIM + 1;
  constexpr bool LABS = ABS == 2;
  constexpr bool PERD = ABS >= 1;
  constexpr int NB = PERD ? DIM : 1;
  constexpr int MLC = MLD ? DIM : 1;
  constexpr int NVEC = DIM + MLC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<NB, NVEC> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < NB * max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < NVEC * nrows; k += nthr) S.vec[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) S.off[k] = T.rowoff[r0 + k];
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();

  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], ce = T.phase_ptr[p0 + ph + 1];
    for (int cl = cb + tid; cl < ce; cl += nthr)
#pragma unroll 1
    for (int k = cl * T.cluster, kend = k + T.cluster; k < kend; k++) {
      const int4 nd = __ldg(T.el_nodes + k);
      const uint2 lr = __ldg(T.el_rows + k);
      const uint4 sl = __ldg(T.el_slots + k);
      MomentumLocal<DIM, LABS> R;
      Geom<DIM> G;
      momentum_element<DIM, LABS>(A, nd, R, G);
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        const unsigned r = lrow_of(lr, i);
        if (r != 0xffffu) {
          const int base = S.off[r];
#pragma unroll
          for (int j = 0; j < LOC; j++) {
            const int q = base + (int)slot_of(sl, i, j);
#pragma unroll
            for (int b = 0; b < NB; b++) {
              double v = R.L[i][j];
              if constexpr (LABS) v += R.Labs[b][i][j];
              if (i == j) v += R.diag[PERD ? b : 0][i];
              S.mat[(size_t)b * max_entries + q] += v;
            }
          }
#pragma unroll
          for (int d = 0; d < DIM; d++) S.vec[r * NVEC + d] += R.rhs[d][i];
#pragma unroll
          for (int d = 0; d < MLC; d++) S.vec[r * NVEC + DIM + d] += R.ml[d][i];
        }
      }
    }
    __syncthreads();
  }

  // ---- write every owned value exactly once, run by run (consecutive global rows) --------------
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int ru0 = T.tile_run_ptr[t], ru1 = T.tile_run_ptr[t + 1];
  for (int ru = ru0 + warp; ru < ru1; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double* s = S.mat + (size_t)(PERD ? d : 0) * max_entries + src;
      double* o = big_m + (size_t)d * T.nnz + dst;
      for (int k = lane; k < n; k += 32) o[k] = s[k];
    }
    // rhs(dim, node), masslump(dim, node): run.y consecutive nodes
    for (int k = lane; k < run.y * DIM; k += 32) {
      const int rr = k / DIM, d = k - rr * DIM;
      rhs[(size_t)DIM * g0 + k] = S.vec[(run.x + rr) * NVEC + d];
      if (masslump) masslump[(size_t)DIM * g0 + k] = S.vec[(run.x + rr) * NVEC + DIM + (MLD ? d : 0)];
    }
  }
}

// ct_m: dim blocks grad_p_u_mat (Momentum_CG.F90:1401,1469); first assembly only, so it is a
// separate pass over the same plan machinery instead of widening the hot kernel.
template <int DIM>
__global__ void __launch_bounds__(256, 1)
tiled_ct_kernel(const MomentumArgs A, const TileArgs T, int max_entries, int max_rows, double* __restrict__ ct_m) {
  constexpr int LOC = DIM + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<DIM, 0> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < DIM * max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) S.off[k] = T.rowoff[r0 + k];
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();
  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], ce = T.phase_ptr[p0 + ph + 1];
    for (int cl = cb + tid; cl < ce; cl += nthr)
#pragma unroll 1
    for (int k = cl * T.cluster, kend = k + T.cluster; k < kend; k++) {
      const int4 nd = __ldg(T.el_nodes + k);
      const uint2 lr = __ldg(T.el_rows + k);
      const uint4 sl = __ldg(T.el_slots + k);
      double X[LOC][DIM];
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        double unused;
        unpack<DIM>(ld256(A.rec.r0 + node_of(nd, i)), X[i], unused);
      }
      Geom<DIM> G;
      geometry<DIM>(X, G);
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        const unsigned r = lrow_of(lr, i);
        if (r != 0xffffu) {
          const int base = S.off[r];
#pragma unroll
          for (int j = 0; j < LOC; j++)
#pragma unroll
            for (int d = 0; d < DIM; d++)
              S.mat[(size_t)d * max_entries + base + (int)slot_of(sl, i, j)] += grad_p_u<DIM>(A.tab, G, d, i, j, A.o.integrate_continuity_by_parts != 0);
        }
      }
    }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int ru = T.tile_run_ptr[t] + warp; ru < T.tile_run_ptr[t + 1]; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
#pragma unroll
    for (int d = 0; d < DIM; d++)
      for (int k = lane; k < n; k += 32) ct_m[(size_t)d * T.nnz + dst + k] = S.mat[(size_t)d * max_entries + src + k];
  }
}

template <int DIM>
__global__ void __launch_bounds__(256, 1)
tiled_advdiff_kernel(const AdvDiffArgs A, const TileArgs T, int max_entries, int max_rows,
                     double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int LOC = DIM + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<1, 1> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) {
    S.vec[k] = 0.0;
    S.off[k] = T.rowoff[r0 + k];
  }
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();
  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], ce = T.phase_ptr[p0 + ph + 1];
    for (int cl = cb + tid; cl < ce; cl += nthr)
#pragma unroll 1
    for (int k = cl * T.cluster, kend = k + T.cluster; k < kend; k++) {
      const int4 nd = __ldg(T.el_nodes + k);
      const uint2 lr = __ldg(T.el_rows + k);
      const uint4 sl = __ldg(T.el_slots + k);
      AdvDiffLocal<DIM> R;
      advdiff_element<DIM>(A, nd, R);
#pragma unroll
      for (int i = 0; i < LOC; i++) {
        const unsigned r = lrow_of(lr, i);
        if (r != 0xffffu) {
          const int base = S.off[r];
#pragma unroll
          for (int j = 0; j < LOC; j++) S.mat[base + (int)slot_of(sl, i, j)] += R.A[i][j];
          S.vec[r] += R.rhs[i];
        }
      }
    }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int ru = T.tile_run_ptr[t] + warp; ru < T.tile_run_ptr[t + 1]; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
    for (int k = lane; k < n; k += 32) matrix[dst + k] = S.mat[src + k];
    for (int k = lane; k < run.y; k += 32) rhs[g0 + k] = S.vec[run.x + k];
  }
}

// ---- fast-path kernels: fused row-by-row accumulate (element_math.cuh momentum_fast) ---------
template <int DIM, bool PERD, bool MLD>
struct MomTileSink {
  static constexpr int MLC = MLD ? DIM : 1;
  static constexpr int NVEC = DIM + MLC;
  double* mat_;
  double* vec_;
  const int* off_;
  int max_entries;
  uint2 lr;
  uint4 sl;
  int base, r;
  __device__ __forceinline__ bool owned(int i) {
    r = (int)lrow_of(lr, i);
    if (r == 0xffff) return false;
    base = off_[r];
    return true;
  }
  __device__ __forceinline__ void mat(int i, int j, int d, double v) {
    mat_[(size_t)d * max_entries + base + (int)slot_of(sl, i, j)] += v;
  }
  __device__ __forceinline__ void vec(int, int d, double v) { vec_[r * NVEC + d] += v; }
  __device__ __forceinline__ void ml(int, int d, double v) {
    if (MLD || d == 0) vec_[r * NVEC + DIM + (MLD ? d : 0)] += v;
  }
  __device__ __forceinline__ void row_end(int) {}
};

template <int DIM, bool PERD, bool MLD>
__global__ void __launch_bounds__(384, 1)
tiled_momentum_fast_kernel(const MomentumArgs A, const TileArgs T, int max_entries, int max_rows,
                           double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  constexpr int NB = PERD ? DIM : 1;
  constexpr int MLC = MLD ? DIM : 1;
  constexpr int NVEC = DIM + MLC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<NB, NVEC> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < NB * max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < NVEC * nrows; k += nthr) S.vec[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) S.off[k] = T.rowoff[r0 + k];
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();
  MomTileSink<DIM, PERD, MLD> sink;
  sink.mat_ = S.mat;
  sink.vec_ = S.vec;
  sink.off_ = S.off;
  sink.max_entries = max_entries;
  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], ce = T.phase_ptr[p0 + ph + 1];
    for (int cl = cb + tid; cl < ce; cl += nthr)
#pragma unroll 1
      for (int k = cl * T.cluster, kend = k + T.cluster; k < kend; k++) {
        const int4 nd = __ldg(T.el_nodes + k);
        sink.lr = __ldg(T.el_rows + k);
        sink.sl = __ldg(T.el_slots + k);
        momentum_fast<DIM, PERD>(A, nd, sink);
      }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int ru = T.tile_run_ptr[t] + warp; ru < T.tile_run_ptr[t + 1]; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double* s = S.mat + (size_t)(PERD ? d : 0) * max_entries + src;
      double* o = big_m + (size_t)d * T.nnz + dst;
      for (int k = lane; k < n; k += 32) o[k] = s[k];
    }
    for (int k = lane; k < run.y * DIM; k += 32) {
      const int rr = k / DIM, d = k - rr * DIM;
      rhs[(size_t)DIM * g0 + k] = S.vec[(run.x + rr) * NVEC + d];
      if (masslump) masslump[(size_t)DIM * g0 + k] = S.vec[(run.x + rr) * NVEC + DIM + (MLD ? d : 0)];
    }
  }
}

template <int DIM>
struct AdvTileSink {
  double* mat_;
  double* vec_;
  const int* off_;
  uint2 lr;
  uint4 sl;
  int base, r;
  __device__ __forceinline__ bool owned(int i) {
    r = (int)lrow_of(lr, i);
    if (r == 0xffff) return false;
    base = off_[r];
    return true;
  }
  __device__ __forceinline__ void mat(int i, int j, double v) { mat_[base + (int)slot_of(sl, i, j)] += v; }
  __device__ __forceinline__ void vec(int, double v) { vec_[r] += v; }
  __device__ __forceinline__ void row_end(int) {}
};

template <int DIM>
__global__ void __launch_bounds__(384, 1)
tiled_advdiff_fast_kernel(const AdvDiffArgs A, const TileArgs T, int max_entries, int max_rows,
                          double* __restrict__ matrix, double* __restrict__ rhs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<1, 1> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) {
    S.vec[k] = 0.0;
    S.off[k] = T.rowoff[r0 + k];
  }
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();
  AdvTileSink<DIM> sink;
  sink.mat_ = S.mat;
  sink.vec_ = S.vec;
  sink.off_ = S.off;
  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], ce = T.phase_ptr[p0 + ph + 1];
    for (int cl = cb + tid; cl < ce; cl += nthr)
#pragma unroll 1
      for (int k = cl * T.cluster, kend = k + T.cluster; k < kend; k++) {
        const int4 nd = __ldg(T.el_nodes + k);
        sink.lr = __ldg(T.el_rows + k);
        sink.sl = __ldg(T.el_slots + k);
        advdiff_fast<DIM>(A, nd, sink);
      }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int ru = T.tile_run_ptr[t] + warp; ru < T.tile_run_ptr[t + 1]; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
    for (int k = lane; k < n; k += 32) matrix[dst + k] = S.mat[src + k];
    for (int k = lane; k < run.y; k += 32) rhs[g0 + k] = S.vec[run.x + k];
  }
}

// ---- quad kernels: one lane = one local row of one element (element_math.cuh momentum_row0) ----
__device__ __forceinline__ int sel4(const int4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ unsigned sel4u(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

template <int DIM, bool PERD, bool MLD>
struct MomQuadSink {
  static constexpr int LOC = DIM + 1;
  static constexpr int MLC = MLD ? DIM : 1;
  static constexpr int NVEC = DIM + MLC;
  double* mat_;   // accumulator of block 0 at this row's base
  double* vec_;   // this row's vector accumulators
  int max_entries;
  unsigned slots;  // 4 x 8 bit slot of column j inside this row
  int i;
  __device__ __forceinline__ void mat(int jj, int d, double v) {
    int j = i + jj;
    if (j >= LOC) j -= LOC;
    mat_[(size_t)d * max_entries + ((slots >> (8 * j)) & 0xffu)] += v;
  }
  __device__ __forceinline__ void vec(int d, double v) { vec_[d] += v; }
  __device__ __forceinline__ void ml(int d, double v) {
    if (MLD || d == 0) vec_[DIM + (MLD ? d : 0)] += v;
  }
};

template <int DIM, bool PERD, bool MLD>
__global__ void __launch_bounds__(768, 1)
tiled_momentum_quad_kernel(const MomentumArgs A, const TileArgs T, int max_entries, int max_rows,
                           double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  constexpr int LOC = DIM + 1;
  constexpr int NB = PERD ? DIM : 1;
  constexpr int MLC = MLD ? DIM : 1;
  constexpr int NVEC = DIM + MLC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<NB, NVEC> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < NB * max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < NVEC * nrows; k += nthr) S.vec[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) S.off[k] = T.rowoff[r0 + k];
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();
  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], nwork = (T.phase_ptr[p0 + ph + 1] - cb) * 4;
    for (int w = tid; w < nwork; w += nthr) {
      const int k = cb + (w >> 2), i = w & 3;
      if (DIM == 2 && i == 3) continue;
      const uint2 lr = __ldg(T.el_rows + k);
      const unsigned r = ((i < 2 ? lr.x : lr.y) >> (16 * (i & 1))) & 0xffffu;
      if (r == 0xffffu) continue;  // row of a node another tile owns
      const int4 nd = __ldg(T.el_nodes + k);
      const uint4 sl = __ldg(T.el_slots + k);
      int n[4];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        int j = i + jj;
        if (j >= LOC) j -= LOC;
        n[jj] = sel4(nd, jj < LOC ? j : 0);
      }
      MomQuadSink<DIM, PERD, MLD> sink;
      sink.mat_ = S.mat + S.off[r];
      sink.vec_ = S.vec + r * NVEC;
      sink.max_entries = max_entries;
      sink.slots = sel4u(sl, i);
      sink.i = i;
      OwnNode<DIM> own;
      own.load(A.rec, n[0]);
      momentum_row0<DIM, PERD>(A, n, own, sink, MomRuntimeFlags{A.o, A.viscosity.stride});
    }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int ru = T.tile_run_ptr[t] + warp; ru < T.tile_run_ptr[t + 1]; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const double* s = S.mat + (size_t)(PERD ? d : 0) * max_entries + src;
      double* o = big_m + (size_t)d * T.nnz + dst;
      for (int k = lane; k < n; k += 32) o[k] = s[k];
    }
    for (int k = lane; k < run.y * DIM; k += 32) {
      const int rr = k / DIM, d = k - rr * DIM;
      rhs[(size_t)DIM * g0 + k] = S.vec[(run.x + rr) * NVEC + d];
      if (masslump) masslump[(size_t)DIM * g0 + k] = S.vec[(run.x + rr) * NVEC + DIM + (MLD ? d : 0)];
    }
  }
}

template <int DIM>
struct AdvQuadSink {
  static constexpr int LOC = DIM + 1;
  double* mat_;
  double* vec_;
  unsigned slots;
  int i;
  __device__ __forceinline__ void mat(int jj, double v) {
    int j = i + jj;
    if (j >= LOC) j -= LOC;
    mat_[(slots >> (8 * j)) & 0xffu] += v;
  }
  __device__ __forceinline__ void vec(double v) { *vec_ += v; }
};

template <int DIM>
__global__ void __launch_bounds__(768, 1)
tiled_advdiff_quad_kernel(const AdvDiffArgs A, const TileArgs T, int max_entries, int max_rows,
                          double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int LOC = DIM + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<1, 1> S(smem_raw, max_entries, max_rows);
  const int t = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int r0 = T.tile_row_ptr[t], nrows = T.tile_row_ptr[t + 1] - r0;
  const int entries = T.tile_entries[t];
  for (int k = tid; k < max_entries; k += nthr) S.mat[k] = 0.0;
  for (int k = tid; k < nrows; k += nthr) {
    S.vec[k] = 0.0;
    S.off[k] = T.rowoff[r0 + k];
  }
  if (tid == 0) S.off[nrows] = entries;
  __syncthreads();
  const int p0 = T.tile_phase_off[t], nphase = T.tile_phase_off[t + 1] - p0 - 1;
  for (int ph = 0; ph < nphase; ph++) {
    const int cb = T.phase_ptr[p0 + ph], nwork = (T.phase_ptr[p0 + ph + 1] - cb) * 4;
    for (int w = tid; w < nwork; w += nthr) {
      const int k = cb + (w >> 2), i = w & 3;
      if (DIM == 2 && i == 3) continue;
      const uint2 lr = __ldg(T.el_rows + k);
      const unsigned r = ((i < 2 ? lr.x : lr.y) >> (16 * (i & 1))) & 0xffffu;
      if (r == 0xffffu) continue;
      const int4 nd = __ldg(T.el_nodes + k);
      const uint4 sl = __ldg(T.el_slots + k);
      int n[4];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        int j = i + jj;
        if (j >= LOC) j -= LOC;
        n[jj] = sel4(nd, jj < LOC ? j : 0);
      }
      AdvQuadSink<DIM> sink;
      sink.mat_ = S.mat + S.off[r];
      sink.vec_ = S.vec + r;
      sink.slots = sel4u(sl, i);
      sink.i = i;
      OwnNode<DIM> own;
      own.load_tracer(A.rec, n[0]);
      advdiff_row0<DIM>(A, n, own, sink, AdvRuntimeFlags{A.o, A.diffusivity.stride});
    }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  for (int ru = T.tile_run_ptr[t] + warp; ru < T.tile_run_ptr[t + 1]; ru += nwarps) {
    const int2 run = T.runs[ru];
    const int g0 = T.rows[r0 + run.x];
    const int src = S.off[run.x];
    const int n = S.off[run.x + run.y] - src;
    const size_t dst = (size_t)T.findrm[g0];
    for (int k = lane; k < n; k += 32) matrix[dst + k] = S.mat[src + k];
    for (int k = lane; k < run.y; k += 32) rhs[g0 + k] = S.vec[run.x + k];
  }
}

static TileArgs tile_args(const Handle* h, const TileClassPlan& P) {
  TileArgs T;
  T.tile_row_ptr = P.d_tile_row_ptr;
  T.rows = P.d_rows;
  T.rowoff = P.d_rowoff;
  T.tile_entries = P.d_tile_entries;
  T.tile_phase_off = P.d_tile_phase_off;
  T.phase_ptr = P.d_phase_ptr;
  T.tile_run_ptr = P.d_tile_run_ptr;
  T.runs = P.d_runs;
  T.el_nodes = P.d_el_nodes;
  T.el_rows = P.d_el_rows;
  T.el_slots = P.d_el_slots;
  T.findrm = h->d_findrm;
  T.nnz = (size_t)h->nnz;
  T.cluster = P.cluster;
  return T;
}

static int block_threads(const TileClassPlan& P, int cap) {
  int thr;
  if (const char* s = getenv("CGASM_TILE_THREADS")) thr = atoi(s);
  else thr = std::max(64, ((P.max_phase + 31) / 32) * 32);
  return std::min(cap, std::max(32, (thr / 32) * 32));
}

// quad kernels: 4 work items per element; pick the thread count that splits the largest phase into
// equal rounds
static int quad_threads(const TileClassPlan& P, int cap) {
  if (const char* s = getenv("CGASM_TILE_THREADS")) return std::min(cap, std::max(32, (atoi(s) / 32) * 32));
  const int work = std::max(1, P.max_phase * 4);
  const int rounds = (work + cap - 1) / cap;
  const int thr = (((work + rounds - 1) / rounds) + 31) / 32 * 32;
  return std::min(cap, std::max(64, thr));
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  CG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CGASM_OK;
}

template <int DIM>
static int tiles_momentum_dim(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  const int abs_mode = !A.o.have_absorption ? 0 : (A.o.lump_absorption ? 1 : 2);
  const bool mld = abs_mode == 1 && A.o.pressure_corrected_absorption;
  int st = ensure_class(h, abs_mode ? 1 : 0);
  if (st) return st;
  const TileClassPlan& P = h->tiles->cls[abs_mode ? 1 : 0];
  const TileArgs T = tile_args(h, P);
  double* ml = want_ml ? h->d_masslump : nullptr;
  const bool fast_ok = abs_mode != 2 && momentum_fast_ok(A.o, A.gravity.stride, A.absorption.stride) &&
                       !getenv("CGASM_TILE_GENERIC");
  if (fast_ok && A.tab.sym && P.cluster == 1 && !getenv("CGASM_TILE_NOQUAD")) {
    const int thr = quad_threads(P, 768);
#define LAUNCH_QUAD(PERD_, MLD_)                                                                        \
  do {                                                                                                  \
    if ((st = set_smem(tiled_momentum_quad_kernel<DIM, PERD_, MLD_>, P.smem_bytes))) return st;         \
    tiled_momentum_quad_kernel<DIM, PERD_, MLD_><<<P.ntiles, thr, P.smem_bytes, h->stream>>>(           \
        A, T, P.max_tile_entries, P.max_tile_rows, h->d_big_m, h->d_mom_rhs, ml);                       \
  } while (0)
    if (abs_mode == 1 && mld) LAUNCH_QUAD(true, true);
    else if (abs_mode == 1) LAUNCH_QUAD(true, false);
    else LAUNCH_QUAD(false, false);
#undef LAUNCH_QUAD
    h->launches++;
  } else if (fast_ok) {
    const int thr = block_threads(P, 384);
#define LAUNCH_FAST(PERD_, MLD_)                                                                        \
  do {                                                                                                  \
    if ((st = set_smem(tiled_momentum_fast_kernel<DIM, PERD_, MLD_>, P.smem_bytes))) return st;         \
    tiled_momentum_fast_kernel<DIM, PERD_, MLD_><<<P.ntiles, thr, P.smem_bytes, h->stream>>>(           \
        A, T, P.max_tile_entries, P.max_tile_rows, h->d_big_m, h->d_mom_rhs, ml);                       \
  } while (0)
    if (abs_mode == 1 && mld) LAUNCH_FAST(true, true);
    else if (abs_mode == 1) LAUNCH_FAST(true, false);
    else LAUNCH_FAST(false, false);
#undef LAUNCH_FAST
    h->launches++;
  } else {
  const int thr = block_threads(P, 256);
#define LAUNCH_MOM(ABS_, MLD_)                                                                        \
  do {                                                                                                \
    if ((st = set_smem(tiled_momentum_kernel<DIM, ABS_, MLD_>, P.smem_bytes))) return st;             \
    tiled_momentum_kernel<DIM, ABS_, MLD_><<<P.ntiles, thr, P.smem_bytes, h->stream>>>(               \
        A, T, P.max_tile_entries, P.max_tile_rows, h->d_big_m, h->d_mom_rhs, ml);                     \
  } while (0)
  if (abs_mode == 2) LAUNCH_MOM(2, false);
  else if (abs_mode == 1 && mld) LAUNCH_MOM(1, true);
  else if (abs_mode == 1) LAUNCH_MOM(1, false);
  else LAUNCH_MOM(0, false);
#undef LAUNCH_MOM
  h->launches++;
  }
  if (want_ct) {
    if ((st = ensure_class(h, 1))) return st;
    const TileClassPlan& P1 = h->tiles->cls[1];
    const TileArgs T1 = tile_args(h, P1);
    if ((st = set_smem(tiled_ct_kernel<DIM>, P1.smem_bytes))) return st;
    tiled_ct_kernel<DIM><<<P1.ntiles, block_threads(P1, 256), P1.smem_bytes, h->stream>>>(
        A, T1, P1.max_tile_entries, P1.max_tile_rows, h->d_ct_m);
    h->launches++;
  }
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int tiles_momentum(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  if (!h->tiles) CG_FAIL(CGASM_ESTATE, "tile plan missing");
  return h->dim == 3 ? tiles_momentum_dim<3>(h, A, want_ml, want_ct) : tiles_momentum_dim<2>(h, A, want_ml, want_ct);
}

int tiles_advdiff(Handle* h, const AdvDiffArgs& A) {
  if (!h->tiles) CG_FAIL(CGASM_ESTATE, "tile plan missing");
  int st = ensure_class(h, 0);
  if (st) return st;
  const TileClassPlan& P = h->tiles->cls[0];
  const TileArgs T = tile_args(h, P);
  const bool fast = advdiff_fast_ok(A.o) && !getenv("CGASM_TILE_GENERIC");
  const bool quad = fast && A.tab.sym && P.cluster == 1 && !getenv("CGASM_TILE_NOQUAD");
  const int thr = quad ? quad_threads(P, 768) : block_threads(P, fast ? 384 : 256);
#define LAUNCH_ADV(KERNEL)                                                                            \
  do {                                                                                                \
    if ((st = set_smem(KERNEL, P.smem_bytes))) return st;                                             \
    KERNEL<<<P.ntiles, thr, P.smem_bytes, h->stream>>>(A, T, P.max_tile_entries, P.max_tile_rows,     \
                                                       h->d_adv_matrix, h->d_adv_rhs);                \
  } while (0)
  if (h->dim == 3) {
    if (quad) LAUNCH_ADV(tiled_advdiff_quad_kernel<3>);
    else if (fast) LAUNCH_ADV(tiled_advdiff_fast_kernel<3>);
    else LAUNCH_ADV(tiled_advdiff_kernel<3>);
  } else {
    if (quad) LAUNCH_ADV(tiled_advdiff_quad_kernel<2>);
    else if (fast) LAUNCH_ADV(tiled_advdiff_fast_kernel<2>);
    else LAUNCH_ADV(tiled_advdiff_kernel<2>);
  }
#undef LAUNCH_ADV
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

}  // namespace cgasm
