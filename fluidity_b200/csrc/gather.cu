// gather.cu -- CGASM_SCATTER_GATHER: "local rows + row gather", two barrier-free passes.
//
// The reference adds every element contribution into the global matrix the moment it is
// computed (addto: femtools/Sparse_Tools.F90:2680-2703, Sparse_Tools_Petsc.F90:848-879,
// Fields_Manipulation.F90:255-379), which on a GPU means either atomics (bound by the SM's red
// issue rate, profiles/), colours (femtools/Colouring.F90: dozens of dependent phases) or
// shared-memory tiles (latency-bound lock-step phases, tiled.cu). Here the addto is split:
//
//   pass A (one thread per element, nothing shared): the element routine runs exactly once per
//     element and streams its loc ROW RECORDS -- row i of the local matrix block(s), rhs(:,i),
//     lumped mass -- to a staging buffer with 256-bit stores;
//   pass B (one thread per CSR row): walks the precomputed list of (element, local row) pairs
//     incident to its node, reads each 64-byte record once, adds the loc matrix values into its
//     own slots (thread-private shared-memory column, conflict-free by layout) and the vector
//     part in registers, then the warp writes the finished row -- every output entry is written
//     exactly once, no atomics, no pre-zeroing, summation order fixed by the plan (bitwise
//     reproducible), no redundant element math.
//
// The pair list plays the role of csr_sparsity_pos (Sparse_Tools.F90:2411-2517): positions are
// found once per mesh instead of once per entry per assembly.
#include "cgasm_internal.h"
#include "gather_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace cgasm {

void gather_free(Handle* h) {
  GatherPlan* p = h->gather;
  if (!p) return;
  if (p->d_rows) cudaFree(p->d_rows);
  if (p->d_block_ptr) cudaFree(p->d_block_ptr);
  if (p->d_pairs) cudaFree(p->d_pairs);
  if (p->d_pair_nodes) cudaFree(p->d_pair_nodes);
  if (p->d_stage) cudaFree(p->d_stage);
  if (p->d_walk_ptr) cudaFree(p->d_walk_ptr);
  if (p->d_walk) cudaFree(p->d_walk);
  if (p->d_own_slot) cudaFree(p->d_own_slot);
  strip_free(p);
  delete p;
  h->gather = nullptr;
  set_permutation(h, std::vector<int>());  // the mirrors follow the row blocks
}

// One thread per row slot: writes the row's pair list (element order = ascending element id, the
// order the serial reference visits them) with the slot of every element node inside the row.
__global__ void gather_pairs_kernel(int nblocks, int loc, const int* __restrict__ rows,
                                    const long long* __restrict__ block_ptr,
                                    const long long* __restrict__ n2e_ptr, const int* __restrict__ n2e,
                                    const int4* __restrict__ ndglno, const int* __restrict__ findrm,
                                    const int* __restrict__ colm, uint2* __restrict__ pairs,
                                    int4* __restrict__ pair_nodes) {
  const int b = blockIdx.x, t = threadIdx.x;
  if (b >= nblocks) return;
  const int r = rows[b * kBR + t];
  const long long base = block_ptr[b];
  const int deg_block = (int)((block_ptr[b + 1] - base) / kBR);
  int deg = 0;
  if (r >= 0) {
    const long long k0 = n2e_ptr[r];
    deg = (int)(n2e_ptr[r + 1] - k0);
    const int s = findrm[r], e_ = findrm[r + 1];
    for (int k = 0; k < deg; k++) {
      const int e = n2e[k0 + k];
      const int4 nd = ndglno[e];
      const int nodes[4] = {nd.x, nd.y, nd.z, nd.w};
      unsigned slots = 0, irow = 0;
      for (int j = 0; j < loc; j++) {
        if (nodes[j] == r) irow = j;
        for (int q = s; q < e_; q++)
          if (colm[q] == nodes[j]) slots |= (unsigned)((q - s) & 0xff) << (8 * j);
      }
      pairs[base + (long long)k * kBR + t] = make_uint2((unsigned)e * 4u + irow, slots);
      int rn[4];
      unsigned rs = 0;
      for (int jj = 0; jj < 4; jj++) {
        int j = (int)irow + jj;
        if (j >= loc) j -= loc;
        if (jj >= loc) j = (int)irow;
        rn[jj] = nodes[j];
        rs |= ((slots >> (8 * j)) & 0xffu) << (8 * jj);
      }
      pair_nodes[base + (long long)k * kBR + t] = make_int4(rn[1], rn[2], loc == 4 ? rn[3] : rn[0], (int)rs);
    }
  }
  for (int k = deg; k < deg_block; k++) {
    pairs[base + (long long)k * kBR + t] = make_uint2(0xFFFFFFFFu, 0u);
    pair_nodes[base + (long long)k * kBR + t] = make_int4(-1, -1, -1, 0);
  }
}

// ---- walk plan -------------------------------------------------------------------------------------
// For one row (node r) with incident elements E_r: two elements are face-adjacent around r when they
// share two of their other nodes. A greedy walk (next = unvisited face-neighbour with the fewest
// unvisited neighbours, else any unvisited element sharing most nodes with the held ones) orders
// E_r; each step emits one entry per node of the next element that is not already held.
static void build_walk_row(const Handle* h, int r, std::vector<int2>& out) {
  const int loc = h->loc, no = loc - 1;  // other nodes per element
  const int64_t k0 = h->n2e_ptr[r];
  const int m = (int)(h->n2e_ptr[r + 1] - k0);
  out.clear();
  if (m == 0) return;
  // other nodes of each incident element
  std::vector<int> oth((size_t)m * 3, -1);
  for (int k = 0; k < m; k++) {
    const int* nd = h->h_nd0.data() + (size_t)4 * h->n2e[(size_t)(k0 + k)];
    int q = 0;
    for (int i = 0; i < loc; i++)
      if (nd[i] != r) oth[(size_t)k * 3 + q++] = nd[i];
  }
  auto shared = [&](int a, int b) {
    int c = 0;
    for (int i = 0; i < no; i++)
      for (int j = 0; j < no; j++) c += oth[(size_t)a * 3 + i] == oth[(size_t)b * 3 + j];
    return c;
  };
  // face adjacency (shares no-1 other nodes); m is ~24 (3-D) / ~6 (2-D): quadratic is fine
  std::vector<std::vector<int>> adj((size_t)m);
  for (int a = 0; a < m; a++)
    for (int b = a + 1; b < m; b++)
      if (shared(a, b) == no - 1) {
        adj[a].push_back(b);
        adj[b].push_back(a);
      }
  std::vector<char> visited((size_t)m, 0);
  auto unvisited_deg = [&](int e) {
    int c = 0;
    for (int x : adj[e]) c += !visited[x];
    return c;
  };
  int held[4] = {r, -1, -1, -1};
  const int s0 = h->h_findrm[r], s1 = h->h_findrm[r + 1];
  auto slot_of = [&](int node) {
    const int* b = h->h_colm.data() + s0;
    const int* e = h->h_colm.data() + s1;
    return (int)(std::lower_bound(b, e, node) - b);
  };
  auto emit = [&](int e) {
    // which positions keep a node of e, which nodes of e are missing
    bool keep[4] = {true, false, false, false};
    int missing[3], nmiss = 0;
    for (int i = 0; i < no; i++) {
      const int node = oth[(size_t)e * 3 + i];
      bool found = false;
      for (int p = 1; p <= no; p++)
        if (held[p] == node) {
          keep[p] = true;
          found = true;
        }
      if (!found) missing[nmiss++] = node;
    }
    int p = 1;
    for (int q = 0; q < nmiss; q++) {
      while (keep[p]) p++;
      held[p] = missing[q];
      keep[p] = true;
      const int compute = (q == nmiss - 1) ? 4 : 0;
      out.push_back(make_int2(missing[q], p | compute | (slot_of(missing[q]) << 8)));
    }
  };
  // start from an element with the fewest face neighbours (an end of the fan on boundaries)
  int cur = 0;
  for (int e = 1; e < m; e++)
    if (adj[e].size() < adj[cur].size()) cur = e;
  for (int step = 0; step < m; step++) {
    visited[cur] = 1;
    emit(cur);
    int next = -1, best = 1 << 30;
    for (int x : adj[cur])
      if (!visited[x]) {
        const int d = unvisited_deg(x);
        if (d < best) {
          best = d;
          next = x;
        }
      }
    if (next < 0) {  // dead end: jump to the unvisited element sharing most held nodes
      int bs = -1;
      for (int e = 0; e < m; e++)
        if (!visited[e]) {
          int c = 0;
          for (int i = 0; i < no; i++)
            for (int p = 1; p <= no; p++) c += held[p] == oth[(size_t)e * 3 + i];
          if (c > bs) {
            bs = c;
            next = e;
          }
        }
    }
    if (next < 0) break;
    cur = next;
  }
}

static int build_walk_plan(Handle* h, GatherPlan* P, const std::vector<int>& rows) {
  const int nb = P->nblocks;
  std::vector<long long> walk_ptr((size_t)nb + 1, 0);
  std::vector<std::vector<int2>> rowplans(rows.size());
  std::vector<unsigned char> own_slot(rows.size(), 0);
  std::vector<int> block_deg((size_t)nb, 0);
#pragma omp parallel
  {
    std::vector<int2> tmp;
#pragma omp for schedule(dynamic, 8)
    for (int b = 0; b < nb; b++) {
      int deg = 0;
      for (int t = 0; t < kBR; t++) {
        const size_t q = (size_t)b * kBR + t;
        const int r = rows[q];
        if (r < 0) continue;
        build_walk_row(h, r, tmp);
        rowplans[q] = tmp;
        deg = std::max(deg, (int)tmp.size());
        const int* cb = h->h_colm.data() + h->h_findrm[r];
        const int* ce = h->h_colm.data() + h->h_findrm[r + 1];
        own_slot[q] = (unsigned char)(std::lower_bound(cb, ce, r) - cb);
      }
      block_deg[b] = deg;
    }
  }
  long long total_real = 0;
  for (int b = 0; b < nb; b++) walk_ptr[b + 1] = walk_ptr[b] + (long long)block_deg[b] * kBR;
  P->n_walk = walk_ptr[nb];
  std::vector<int2> walk((size_t)std::max<long long>(P->n_walk, 1), make_int2(-1, 0));
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total_real)
  for (int b = 0; b < nb; b++)
    for (int t = 0; t < kBR; t++) {
      const auto& rp = rowplans[(size_t)b * kBR + t];
      for (size_t k = 0; k < rp.size(); k++) walk[(size_t)(walk_ptr[b] + (long long)k * kBR + t)] = rp[k];
      total_real += (long long)rp.size();
    }
  P->walk_entries_per_pair = h->n2e.empty() ? 0.0 : (double)total_real / (double)h->n2e.size();
  if (getenv("CGASM_DEBUG"))
    fprintf(stderr, "[cgasm] walk plan: %.3f entries per (row, element) pair, %lld padded entries\n",
            P->walk_entries_per_pair, P->n_walk);
  CG_CUDA(cudaMalloc(&P->d_walk_ptr, sizeof(long long) * walk_ptr.size()));
  CG_CUDA(cg_upload(P->d_walk_ptr, walk_ptr.data(), sizeof(long long) * walk_ptr.size()));
  CG_CUDA(cudaMalloc(&P->d_walk, sizeof(int2) * walk.size()));
  CG_CUDA(cg_upload(P->d_walk, walk.data(), sizeof(int2) * walk.size()));
  if (!P->d_own_slot) {  // the STRIP plan may have made it already
    CG_CUDA(cudaMalloc(&P->d_own_slot, own_slot.size()));
    CG_CUDA(cg_upload(P->d_own_slot, own_slot.data(), own_slot.size()));
  }
  return CGASM_OK;
}

// Row blocks only: what every row-owner variant needs (rows of each block, block degrees, longest row).
int gather_build_rows(Handle* h) {
  if (!h->have_X) CG_FAIL(CGASM_ESTATE, "gather scatter orders rows by node coordinates: set coordinates first");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "gather scatter needs the sparsity first");
  if ((long long)h->n_elements >= (1ll << 30)) CG_FAIL(CGASM_EUNSUPPORTED, "gather scatter packs element*4+row in 32 bits");
  gather_free(h);
  // built into a local plan and installed on the handle only when everything succeeded: a half-built plan
  // (e.g. rows longer than the slot index allows) must not make the next cgasm_set_scatter skip the rebuild
  GatherPlan* P = new GatherPlan();
  static long long plan_serial = 0;
  P->serial = ++plan_serial;
  std::vector<int> order;
  MortonFrame F;
  morton_order(h, order, F);
  std::vector<int> rows;
  P->nblocks = form_row_blocks(h, order, F, kBR, rows);
  order_block_rows(h, F, rows, P->nblocks);
  std::vector<long long> block_ptr((size_t)P->nblocks + 1, 0);
  int maxlen = 0;
#pragma omp parallel for schedule(static) reduction(max : maxlen)
  for (int b = 0; b < P->nblocks; b++) {
    int deg = 0;
    for (int t = 0; t < kBR; t++) {
      const int r = rows[(size_t)b * kBR + t];
      if (r < 0) continue;
      deg = std::max(deg, (int)(h->n2e_ptr[r + 1] - h->n2e_ptr[r]));
      maxlen = std::max(maxlen, h->h_findrm[r + 1] - h->h_findrm[r]);
    }
    block_ptr[b + 1] = (long long)deg * kBR;
  }
  for (int b = 0; b < P->nblocks; b++) block_ptr[b + 1] += block_ptr[b];
  P->maxlen = maxlen;
  P->n_entries = block_ptr[P->nblocks];
  auto fail = [&](int code) {
    h->gather = P;  // gather_free releases whatever was allocated
    gather_free(h);
    return code;
  };
  if (maxlen > 255) {
    set_error("CSR row longer than 255 entries: gather slot index is 8 bit");
    return fail(CGASM_EUNSUPPORTED);
  }
  if (cudaMalloc(&P->d_rows, sizeof(int) * rows.size()) != cudaSuccess ||
      cg_upload(P->d_rows, rows.data(), sizeof(int) * rows.size()) != cudaSuccess ||
      cudaMalloc(&P->d_block_ptr, sizeof(long long) * block_ptr.size()) != cudaSuccess ||
      cg_upload(P->d_block_ptr, block_ptr.data(), sizeof(long long) * block_ptr.size()) != cudaSuccess) {
    set_error(std::string("gather_build_rows: ") + cudaGetErrorString(cudaGetLastError()));
    return fail(CGASM_ECUDA);
  }
  P->h_rows.swap(rows);
  h->gather = P;
  return CGASM_OK;
}

// Pair lists + walk plan of the GATHER kernels. The STRIP variant builds them only when an option set
// outside its own kernels is first assembled (they are 13 GB and ~40 % of the plan time at S3).
int gather_build_pairs(Handle* h) {
  GatherPlan* P = h->gather;
  if (!P) CG_FAIL(CGASM_ESTATE, "gather plan missing");
  if (P->d_pairs) return CGASM_OK;
  const int n = h->n_nodes;
  CG_CUDA(cudaMalloc(&P->d_pairs, sizeof(uint2) * (size_t)std::max<long long>(P->n_entries, 1)));
  CG_CUDA(cudaMalloc(&P->d_pair_nodes, sizeof(int4) * (size_t)std::max<long long>(P->n_entries, 1)));
  // node->element adjacency goes to the device only for the duration of the plan build
  long long* d_n2e_ptr = nullptr;
  int* d_n2e = nullptr;
  CG_CUDA(cudaMalloc(&d_n2e_ptr, sizeof(long long) * ((size_t)n + 1)));
  CG_CUDA(cudaMalloc(&d_n2e, sizeof(int) * std::max<size_t>(h->n2e.size(), 1)));
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long");
  cudaError_t e1 = cg_upload(d_n2e_ptr, h->n2e_ptr.data(), sizeof(long long) * ((size_t)n + 1));
  cudaError_t e2 = cg_upload(d_n2e, h->n2e.data(), sizeof(int) * h->n2e.size());
  if (e1 == cudaSuccess && e2 == cudaSuccess) {
    gather_pairs_kernel<<<P->nblocks, kBR, 0, h->stream>>>(P->nblocks, h->loc, P->d_rows, P->d_block_ptr, d_n2e_ptr,
                                                          d_n2e, h->d_ndglno, h->d_findrm, h->d_colm, P->d_pairs,
                                                          P->d_pair_nodes);
    h->launches++;
    e1 = cudaStreamSynchronize(h->stream);
  }
  cudaFree(d_n2e_ptr);
  cudaFree(d_n2e);
  CG_CUDA(e1);
  CG_CUDA(e2);
  if (!getenv("CGASM_GATHER_NOWALK")) return build_walk_plan(h, P, P->h_rows);
  return CGASM_OK;
}

int gather_build(Handle* h) {
  int st = gather_build_rows(h);
  if (st) return st;
  return gather_build_pairs(h);
}

static int ensure_stage(GatherPlan* P, size_t doubles) {
  if (P->stage_doubles >= doubles) return CGASM_OK;
  if (P->d_stage) cudaFree(P->d_stage);
  P->d_stage = nullptr;
  P->stage_doubles = 0;
  CG_CUDA(cudaMalloc(&P->d_stage, sizeof(double) * doubles));
  P->stage_doubles = doubles;
  return CGASM_OK;
}

// ---- record layout -------------------------------------------------------------------------------
// record of (element e, local row i) at stage + (e*LOC + i)*RS doubles:
//   [b*LOC + j]  matrix value (i,j) of block b, b < NB     [NB*LOC + c]  vector component c < NV
template <int LOC, int NB, int NV>
struct Rec {
  static constexpr int N = NB * LOC + NV;
  static constexpr int RS = (N + 3) / 4 * 4;  // whole 32-byte sectors
};

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int RS>
__device__ __forceinline__ void store_rec(double* dst, const double (&v)[RS]) {
#pragma unroll
  for (int q = 0; q < RS; q += 4) st256(dst + q, v[q], v[q + 1], v[q + 2], v[q + 3]);
}

// ---- pass A ----------------------------------------------------------------------------------------
// momentum, generic options: NB = 1 (no absorption) or DIM; NV = DIM + MLC
#ifndef CGASM_SU_MINB
#define CGASM_SU_MINB 1
#endif
template <int DIM, bool LABS, int NB, int MLC, int STAB>
__global__ void __launch_bounds__(128, STAB != 0 ? CGASM_SU_MINB : 1)
gather_momentum_stage_kernel(const MomentumArgs A, double* __restrict__ stage) {
  constexpr int LOC = DIM + 1;
  using R_ = Rec<LOC, NB, DIM + MLC>;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_elements) return;
  const int4 nd = __ldg(A.ndglno + e);
  MomentumLocal<DIM, LABS> R;
  Geom<DIM> G;
  momentum_element<DIM, LABS, STAB>(A, nd, R, G);
#pragma unroll
  for (int i = 0; i < LOC; i++) {
    double v[R_::RS];
#pragma unroll
    for (int q = 0; q < R_::RS; q++) v[q] = 0.0;
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
      for (int j = 0; j < LOC; j++) {
        double x = R.L[i][j];
        if constexpr (LABS) x += R.Labs[b][i][j];
        if (i == j) x += R.diag[NB > 1 ? b : 0][i];
        v[b * LOC + j] = x;
      }
#pragma unroll
    for (int d = 0; d < DIM; d++) v[NB * LOC + d] = R.rhs[d][i];
#pragma unroll
    for (int d = 0; d < MLC; d++) v[NB * LOC + DIM + d] = R.ml[d][i];
    store_rec<R_::RS>(stage + ((size_t)e * LOC + i) * R_::RS, v);
  }
}

template <int DIM, bool PERD, int MLC>
struct MomStageSink {
  static constexpr int LOC = DIM + 1;
  static constexpr int NB = PERD ? DIM : 1;
  using R_ = Rec<LOC, NB, DIM + MLC>;
  double* dst;  // first record of the element
  double v[R_::RS];
  __device__ __forceinline__ bool owned(int) {
#pragma unroll
    for (int q = 0; q < R_::RS; q++) v[q] = 0.0;
    return true;
  }
  __device__ __forceinline__ void mat(int, int j, int d, double x) { v[d * LOC + j] = x; }
  __device__ __forceinline__ void vec(int, int d, double x) { v[NB * LOC + d] = x; }
  __device__ __forceinline__ void ml(int, int d, double x) {
    if (d < MLC) v[NB * LOC + DIM + d] = x;
  }
  __device__ __forceinline__ void row_end(int i) { store_rec<R_::RS>(dst + (size_t)i * R_::RS, v); }
};

template <int DIM, bool PERD, int MLC>
__global__ void __launch_bounds__(128)
gather_momentum_stage_fast_kernel(const MomentumArgs A, double* __restrict__ stage) {
  constexpr int LOC = DIM + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_elements) return;
  const int4 nd = __ldg(A.ndglno + e);
  MomStageSink<DIM, PERD, MLC> sink;
  sink.dst = stage + (size_t)e * LOC * MomStageSink<DIM, PERD, MLC>::R_::RS;
  momentum_fast<DIM, PERD>(A, nd, sink);
}

template <int DIM>
__global__ void __launch_bounds__(128) gather_ct_stage_kernel(const MomentumArgs A, double* __restrict__ stage) {
  constexpr int LOC = DIM + 1;
  using R_ = Rec<LOC, DIM, 0>;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_elements) return;
  const int4 nd = __ldg(A.ndglno + e);
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
    double v[R_::RS];
#pragma unroll
    for (int q = 0; q < R_::RS; q++) v[q] = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
      for (int j = 0; j < LOC; j++) v[d * LOC + j] = grad_p_u<DIM>(A.tab, G, d, i, j, A.o.integrate_continuity_by_parts != 0);
    store_rec<R_::RS>(stage + ((size_t)e * LOC + i) * R_::RS, v);
  }
}

template <int DIM, int STAB>
__global__ void __launch_bounds__(128, STAB != 0 ? CGASM_SU_MINB : 1) gather_advdiff_stage_kernel(const AdvDiffArgs A, double* __restrict__ stage) {
  constexpr int LOC = DIM + 1;
  using R_ = Rec<LOC, 1, 1>;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_elements) return;
  AdvDiffLocal<DIM> R;
  advdiff_element<DIM, STAB>(A, __ldg(A.ndglno + e), R);
#pragma unroll
  for (int i = 0; i < LOC; i++) {
    double v[R_::RS];
#pragma unroll
    for (int q = 0; q < R_::RS; q++) v[q] = 0.0;
#pragma unroll
    for (int j = 0; j < LOC; j++) v[j] = R.A[i][j];
    v[LOC] = R.rhs[i];
    store_rec<R_::RS>(stage + ((size_t)e * LOC + i) * R_::RS, v);
  }
}

template <int DIM>
struct AdvStageSink {
  static constexpr int LOC = DIM + 1;
  using R_ = Rec<LOC, 1, 1>;
  double* dst;
  double v[R_::RS];
  __device__ __forceinline__ bool owned(int) {
#pragma unroll
    for (int q = 0; q < R_::RS; q++) v[q] = 0.0;
    return true;
  }
  __device__ __forceinline__ void mat(int, int j, double x) { v[j] = x; }
  __device__ __forceinline__ void vec(int, double x) { v[LOC] = x; }
  __device__ __forceinline__ void row_end(int i) { store_rec<R_::RS>(dst + (size_t)i * R_::RS, v); }
};

template <int DIM>
__global__ void __launch_bounds__(128)
gather_advdiff_stage_fast_kernel(const AdvDiffArgs A, double* __restrict__ stage) {
  constexpr int LOC = DIM + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_elements) return;
  AdvStageSink<DIM> sink;
  sink.dst = stage + (size_t)e * LOC * AdvStageSink<DIM>::R_::RS;
  advdiff_fast<DIM>(A, __ldg(A.ndglno + e), sink);
}

// ---- pass B ----------------------------------------------------------------------------------------
// KIND 0: momentum (out0 = big_m [DIM blocks], out1 = rhs, out2 = masslump or null)
// KIND 1: tracer   (out0 = matrix, out1 = rhs)        KIND 2: ct_m (out0 = ct_m [DIM blocks])
template <int DIM, int NB, int NV, int KIND>
__global__ void __launch_bounds__(kBR)
gather_rows_kernel(const int* __restrict__ rows, const long long* __restrict__ block_ptr,
                   const uint2* __restrict__ pairs, const double* __restrict__ stage,
                   const int* __restrict__ findrm, size_t nnz, int maxlen, double* __restrict__ out0,
                   double* __restrict__ out1, double* __restrict__ out2) {
  constexpr int LOC = DIM + 1;
  using R_ = Rec<LOC, NB, NV>;
  extern __shared__ double acc[];  // [NB][maxlen][kBR]: column t is private to thread t
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = rows[b * kBR + t];
  const long long base = block_ptr[b];
  const int deg = (int)((block_ptr[b + 1] - base) / kBR);
  const int len = r >= 0 ? findrm[r + 1] - findrm[r] : 0;
  for (int q = 0; q < NB * maxlen; q++) acc[q * kBR + t] = 0.0;
  double vec[NV > 0 ? NV : 1];
#pragma unroll
  for (int c = 0; c < (NV > 0 ? NV : 1); c++) vec[c] = 0.0;
  const uint2* p = pairs + base + t;
#pragma unroll 2
  for (int k = 0; k < deg; k++) {
    const uint2 ent = __ldg(p + (long long)k * kBR);
    if (ent.x == 0xFFFFFFFFu) continue;
    const double* rec = stage + ((size_t)(ent.x >> 2) * LOC + (ent.x & 3u)) * R_::RS;
    double v[R_::RS];
#pragma unroll
    for (int q = 0; q < R_::RS; q += 4) {
      const double4 x = ld256(reinterpret_cast<const double4*>(rec + q));
      v[q] = x.x;
      v[q + 1] = x.y;
      v[q + 2] = x.z;
      v[q + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < LOC; j++) {
      const int s = (ent.y >> (8 * j)) & 0xff;
#pragma unroll
      for (int bb = 0; bb < NB; bb++) acc[(bb * maxlen + s) * kBR + t] += v[bb * LOC + j];
    }
#pragma unroll
    for (int c = 0; c < NV; c++) vec[c] += v[NB * LOC + c];
  }
  // vector outputs: one thread = one node
  if (r >= 0) {
    if (KIND == 0) {
#pragma unroll
      for (int d = 0; d < DIM; d++) out1[(size_t)DIM * r + d] = vec[d];
      if (out2) {
#pragma unroll
        for (int d = 0; d < DIM; d++) out2[(size_t)DIM * r + d] = vec[DIM + (NV == 2 * DIM ? d : 0)];
      }
    } else if (KIND == 1) {
      out1[r] = vec[0];
    }
  }
  __syncthreads();
  // matrix rows: a warp writes the rows of its 32 threads one after another, lanes = entries
  const int warp = t >> 5, lane = t & 31;
  for (int rr = 0; rr < 32; rr++) {
    const int tt = warp * 32 + rr;
    const int row = rows[b * kBR + tt];
    if (row < 0) continue;
    const int s0 = findrm[row], n = findrm[row + 1] - s0;
    for (int s = lane; s < n; s += 32) {
      if (KIND == 0) {
#pragma unroll
        for (int d = 0; d < DIM; d++) out0[(size_t)d * nnz + s0 + s] = acc[((NB > 1 ? d : 0) * maxlen + s) * kBR + tt];
      } else if (KIND == 1) {
        out0[(size_t)s0 + s] = acc[s * kBR + tt];
      } else {
#pragma unroll
        for (int d = 0; d < DIM; d++) out0[(size_t)d * nnz + s0 + s] = acc[(d * maxlen + s) * kBR + tt];
      }
    }
  }
  (void)len;
}

// ---- single pass: the row thread computes its own row of every incident element ---------------------
// For the fast option set (momentum_fast_ok / advdiff_fast_ok) with a node-symmetric quadrature rule
// the row of local node i costs ~120 FP64 operations in closed form (element_math.cuh
// momentum_row0), so recomputing it per (row, element) pair is cheaper than staging 64 bytes per
// pair through HBM: no staging buffer, one kernel, every output written once.
template <int DIM, bool PERD, bool MLD>
struct MomDirectSink {
  static constexpr int LOC = DIM + 1;
  static constexpr int NV = DIM + (MLD ? DIM : 1);
  double* acc;  // this thread's column: acc[(b*maxlen + s)*kBR]
  int maxlen;
  unsigned slots;
  int i;
  double vec_[NV];
  __device__ __forceinline__ void mat(int jj, int d, double v) {
    acc[(d * maxlen + (int)((slots >> (8 * jj)) & 0xffu)) * kBR] += v;
  }
  __device__ __forceinline__ void vec(int d, double v) { vec_[d] += v; }
  __device__ __forceinline__ void ml(int d, double v) {
    if (MLD || d == 0) vec_[DIM + (MLD ? d : 0)] += v;
  }
};

template <int DIM>
struct AdvDirectSink {
  static constexpr int LOC = DIM + 1;
  double* acc;
  unsigned slots;
  int i;
  double rhs;
  __device__ __forceinline__ void mat(int jj, double v) { acc[(int)((slots >> (8 * jj)) & 0xffu) * kBR] += v; }
  __device__ __forceinline__ void vec(double v) { rhs += v; }
};

__device__ __forceinline__ int rot_node(const int4& nd, int j) { return j == 0 ? nd.x : (j == 1 ? nd.y : (j == 2 ? nd.z : nd.w)); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Pair stream of one row thread: rotated node ids + slots of the current pair, with the next
// pair already in flight. Plan data is read once, so it bypasses L1 (L1::no_allocate) and leaves
// the cache to the node records, which neighbouring rows re-read ~24 times.
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
struct PairStream {
  const int4* pn;
  int deg, k;
  int4 nd0, nd1;
  __device__ __forceinline__ void load_next(int kk) {
    if (kk < deg) nd1 = ldg_stream(pn + (long long)kk * kBR);
    else nd1 = make_int4(-1, -1, -1, 0);
  }
  __device__ __forceinline__ void init(const int4* pn_, int deg_) {
    pn = pn_;
    deg = deg_;
    k = 0;
    load_next(0);
    nd0 = nd1;
    load_next(1);
  }
  __device__ __forceinline__ void advance() {
    nd0 = nd1;
    k++;
    load_next(k + 1);
  }
  __device__ __forceinline__ bool valid() const { return nd0.x >= 0; }
};

template <int NREC>
__device__ __forceinline__ void prefetch_nodes(const NodeRecs& rec, const int4& nd, bool valid) {
  if (!valid) return;
  const int nn[3] = {nd.x, nd.y, nd.z};
#pragma unroll
  for (int q = 0; q < 3; q++) {
    prefetch_l1(rec.r0 + nn[q]);
    prefetch_l1(rec.r1 + nn[q]);
    if (NREC > 2) prefetch_l1(rec.r2 + nn[q]);
  }
}

template <int DIM, bool PERD, bool MLD, bool COMMON, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
gather_momentum_direct_kernel(const MomentumArgs A, const int* __restrict__ rows, const long long* __restrict__ block_ptr,
                              const int4* __restrict__ pair_nodes, const int* __restrict__ findrm, size_t nnz, int maxlen,
                              int prefetch, double* __restrict__ big_m, double* __restrict__ rhs,
                              double* __restrict__ masslump) {
  constexpr int LOC = DIM + 1;
  constexpr int NB = PERD ? DIM : 1;
  extern __shared__ double acc[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = rows[b * kBR + t];
  const long long base = block_ptr[b];
  const int deg = (int)((block_ptr[b + 1] - base) / kBR);
  PairStream ps;
  ps.init(pair_nodes + base + t, deg);
  for (int q = 0; q < NB * maxlen; q++) acc[q * kBR + t] = 0.0;
  MomDirectSink<DIM, PERD, MLD> sink;
  sink.acc = acc + t;
  sink.maxlen = maxlen;
#pragma unroll
  for (int c = 0; c < MomDirectSink<DIM, PERD, MLD>::NV; c++) sink.vec_[c] = 0.0;
  OwnNode<DIM> own;
  own.load(A.rec, r >= 0 ? r : 0);
  for (; ps.k < deg; ps.advance()) {
    if (prefetch) prefetch_nodes<3>(A.rec, ps.nd1, ps.nd1.x >= 0);
    if (!ps.valid()) continue;
    const int n[4] = {r, ps.nd0.x, ps.nd0.y, ps.nd0.z};
    sink.slots = (unsigned)ps.nd0.w;
    sink.i = 0;
    if constexpr (COMMON) momentum_row0<DIM, PERD>(A, n, own, sink, MomCommonFlags());
    else momentum_row0<DIM, PERD>(A, n, own, sink, MomRuntimeFlags{A.o, A.viscosity.stride});
  }
  if (r >= 0) {
#pragma unroll
    for (int d = 0; d < DIM; d++) rhs[(size_t)DIM * r + d] = sink.vec_[d];
    if (masslump) {
#pragma unroll
      for (int d = 0; d < DIM; d++) masslump[(size_t)DIM * r + d] = sink.vec_[DIM + (MLD ? d : 0)];
    }
  }
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int rr = 0; rr < 32; rr++) {
    const int tt = warp * 32 + rr;
    const int row = rows[b * kBR + tt];
    if (row < 0) continue;
    const int s0 = findrm[row], n = findrm[row + 1] - s0;
    for (int s = lane; s < n; s += 32) {
#pragma unroll
      for (int d = 0; d < DIM; d++) big_m[(size_t)d * nnz + s0 + s] = acc[((PERD ? d : 0) * maxlen + s) * kBR + tt];
    }
  }
}

template <int DIM, bool COMMON, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
gather_advdiff_direct_kernel(const AdvDiffArgs A, const int* __restrict__ rows, const long long* __restrict__ block_ptr,
                             const int4* __restrict__ pair_nodes, const int* __restrict__ findrm, int maxlen, int prefetch,
                             double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int LOC = DIM + 1;
  extern __shared__ double acc[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = rows[b * kBR + t];
  const long long base = block_ptr[b];
  const int deg = (int)((block_ptr[b + 1] - base) / kBR);
  PairStream ps;
  ps.init(pair_nodes + base + t, deg);
  for (int q = 0; q < maxlen; q++) acc[q * kBR + t] = 0.0;
  AdvDirectSink<DIM> sink;
  sink.acc = acc + t;
  sink.rhs = 0.0;
  OwnNode<DIM> own;
  own.load_tracer(A.rec, r >= 0 ? r : 0);
  for (; ps.k < deg; ps.advance()) {
    if (prefetch) prefetch_nodes<2>(A.rec, ps.nd1, ps.nd1.x >= 0);
    if (!ps.valid()) continue;
    const int n[4] = {r, ps.nd0.x, ps.nd0.y, ps.nd0.z};
    sink.slots = (unsigned)ps.nd0.w;
    sink.i = 0;
    if constexpr (COMMON) advdiff_row0<DIM>(A, n, own, sink, AdvCommonFlags());
    else advdiff_row0<DIM>(A, n, own, sink, AdvRuntimeFlags{A.o, A.diffusivity.stride});
  }
  if (r >= 0) rhs[r] = sink.rhs;
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int rr = 0; rr < 32; rr++) {
    const int tt = warp * 32 + rr;
    const int row = rows[b * kBR + tt];
    if (row < 0) continue;
    const int s0 = findrm[row], n = findrm[row + 1] - s0;
    for (int s = lane; s < n; s += 32) matrix[(size_t)s0 + s] = acc[s * kBR + tt];
  }
}

// ---- single pass, walk order: one new node per entry ---------------------------------------------------
__device__ __forceinline__ int2 ldg_stream(const int2* p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

template <int DIM, bool COMMON, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
gather_momentum_walk_kernel(const MomentumArgs A, const int* __restrict__ rows, const long long* __restrict__ walk_ptr,
                            const int2* __restrict__ walk, const unsigned char* __restrict__ own_slot,
                            const int* __restrict__ findrm, size_t nnz, int maxlen, int prefetch,
                            double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ masslump) {
  constexpr int LOC = DIM + 1;
  extern __shared__ double acc[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = rows[b * kBR + t];
  const long long base = walk_ptr[b];
  const int deg = (int)((walk_ptr[b + 1] - base) / kBR);
  const int2* p = walk + base + t;
  int2 e0 = deg > 0 ? ldg_stream(p) : make_int2(-1, 0);
  int2 e1 = deg > 1 ? ldg_stream(p + kBR) : make_int2(-1, 0);
  for (int q = 0; q < maxlen; q++) acc[q * kBR + t] = 0.0;
  MomDirectSink<DIM, false, false> sink;
  sink.acc = acc + t;
  sink.maxlen = maxlen;
  sink.i = 0;
#pragma unroll
  for (int c = 0; c < MomDirectSink<DIM, false, false>::NV; c++) sink.vec_[c] = 0.0;
  // node data in registers: position 0 = the row's own node, 1..LOC-1 = the walk's current element
  double X[LOC][DIM], nu[LOC][DIM], oldu[LOC][DIM], rho[LOC], bb[LOC];
  int n[4] = {r >= 0 ? r : 0, 0, 0, 0};
  unsigned slots = own_slot[b * kBR + t];
  {
    double unused;
    unpack<DIM>(ld256(A.rec.r0 + n[0]), X[0], unused);
    unpack<DIM>(ld256(A.rec.r1 + n[0]), nu[0], rho[0]);
    unpack<DIM>(ld256(A.rec.r2 + n[0]), oldu[0], bb[0]);
#pragma unroll
    for (int q = 1; q < LOC; q++) {
#pragma unroll
      for (int a = 0; a < DIM; a++) X[q][a] = nu[q][a] = oldu[q][a] = 0.0;
      rho[q] = bb[q] = 0.0;
    }
  }
  for (int k = 0; k < deg; k++) {
    const int2 ent = e0;
    e0 = e1;
    e1 = (k + 2 < deg) ? ldg_stream(p + (long long)(k + 2) * kBR) : make_int2(-1, 0);
    if (prefetch && e0.x >= 0) {
      prefetch_l1(A.rec.r0 + e0.x);
      prefetch_l1(A.rec.r1 + e0.x);
      prefetch_l1(A.rec.r2 + e0.x);
    }
    if (ent.x < 0) continue;
    const int pos = ent.y & 3;
#pragma unroll
    for (int q = 1; q < LOC; q++)
      if (pos == q) {
        double unused;
        unpack<DIM>(ld256(A.rec.r0 + ent.x), X[q], unused);
        unpack<DIM>(ld256(A.rec.r1 + ent.x), nu[q], rho[q]);
        unpack<DIM>(ld256(A.rec.r2 + ent.x), oldu[q], bb[q]);
        n[q] = ent.x;
        slots = (slots & ~(0xffu << (8 * q))) | (((unsigned)ent.y >> 8 & 0xffu) << (8 * q));
      }
    if (ent.y & 4) {
      sink.slots = slots;
      if constexpr (COMMON) momentum_row0_data<DIM, false>(A, n, X, nu, rho, oldu, bb, sink, MomCommonFlags());
      else momentum_row0_data<DIM, false>(A, n, X, nu, rho, oldu, bb, sink, MomRuntimeFlags{A.o, A.viscosity.stride});
    }
  }
  if (r >= 0) {
#pragma unroll
    for (int d = 0; d < DIM; d++) rhs[(size_t)DIM * r + d] = sink.vec_[d];
    if (masslump) {
#pragma unroll
      for (int d = 0; d < DIM; d++) masslump[(size_t)DIM * r + d] = sink.vec_[DIM];
    }
  }
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int rr = 0; rr < 32; rr++) {
    const int tt = warp * 32 + rr;
    const int row = rows[b * kBR + tt];
    if (row < 0) continue;
    const int s0 = findrm[row], nn = findrm[row + 1] - s0;
    for (int s = lane; s < nn; s += 32) {
#pragma unroll
      for (int d = 0; d < DIM; d++) big_m[(size_t)d * nnz + s0 + s] = acc[s * kBR + tt];
    }
  }
}

template <int DIM, bool COMMON, int MINB>
__global__ void __launch_bounds__(kBR, MINB)
gather_advdiff_walk_kernel(const AdvDiffArgs A, const int* __restrict__ rows, const long long* __restrict__ walk_ptr,
                           const int2* __restrict__ walk, const unsigned char* __restrict__ own_slot,
                           const int* __restrict__ findrm, int maxlen, int prefetch, double* __restrict__ matrix,
                           double* __restrict__ rhs) {
  constexpr int LOC = DIM + 1;
  extern __shared__ double acc[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int r = rows[b * kBR + t];
  const long long base = walk_ptr[b];
  const int deg = (int)((walk_ptr[b + 1] - base) / kBR);
  const int2* p = walk + base + t;
  int2 e0 = deg > 0 ? ldg_stream(p) : make_int2(-1, 0);
  int2 e1 = deg > 1 ? ldg_stream(p + kBR) : make_int2(-1, 0);
  for (int q = 0; q < maxlen; q++) acc[q * kBR + t] = 0.0;
  AdvDirectSink<DIM> sink;
  sink.acc = acc + t;
  sink.rhs = 0.0;
  sink.i = 0;
  double X[LOC][DIM], u[LOC][DIM], T[LOC];
  int n[4] = {r >= 0 ? r : 0, 0, 0, 0};
  unsigned slots = own_slot[b * kBR + t];
  {
    double unused;
    unpack<DIM>(ld256(A.rec.r0 + n[0]), X[0], T[0]);
    unpack<DIM>(ld256(A.rec.r1 + n[0]), u[0], unused);
#pragma unroll
    for (int q = 1; q < LOC; q++) {
#pragma unroll
      for (int a = 0; a < DIM; a++) X[q][a] = u[q][a] = 0.0;
      T[q] = 0.0;
    }
  }
  for (int k = 0; k < deg; k++) {
    const int2 ent = e0;
    e0 = e1;
    e1 = (k + 2 < deg) ? ldg_stream(p + (long long)(k + 2) * kBR) : make_int2(-1, 0);
    if (prefetch && e0.x >= 0) {
      prefetch_l1(A.rec.r0 + e0.x);
      prefetch_l1(A.rec.r1 + e0.x);
    }
    if (ent.x < 0) continue;
    const int pos = ent.y & 3;
#pragma unroll
    for (int q = 1; q < LOC; q++)
      if (pos == q) {
        double unused;
        unpack<DIM>(ld256(A.rec.r0 + ent.x), X[q], T[q]);
        unpack<DIM>(ld256(A.rec.r1 + ent.x), u[q], unused);
        n[q] = ent.x;
        slots = (slots & ~(0xffu << (8 * q))) | (((unsigned)ent.y >> 8 & 0xffu) << (8 * q));
      }
    if (ent.y & 4) {
      sink.slots = slots;
      if constexpr (COMMON) advdiff_row0_data<DIM>(A, n, X, T, u, sink, AdvCommonFlags());
      else advdiff_row0_data<DIM>(A, n, X, T, u, sink, AdvRuntimeFlags{A.o, A.diffusivity.stride});
    }
  }
  if (r >= 0) rhs[r] = sink.rhs;
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int rr = 0; rr < 32; rr++) {
    const int tt = warp * 32 + rr;
    const int row = rows[b * kBR + tt];
    if (row < 0) continue;
    const int s0 = findrm[row], nn = findrm[row + 1] - s0;
    for (int s = lane; s < nn; s += 32) matrix[(size_t)s0 + s] = acc[s * kBR + tt];
  }
}

template <class K>
static int set_dyn_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CGASM_OK;
}

template <int DIM, int NB, int NV, int KIND>
static int launch_rows(Handle* h, double* out0, double* out1, double* out2) {
  GatherPlan* P = h->gather;
  const size_t smem = sizeof(double) * (size_t)NB * P->maxlen * kBR;
  if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "gather scatter: CSR rows too long for the shared-memory accumulator");
  int st = set_dyn_smem(gather_rows_kernel<DIM, NB, NV, KIND>, smem);
  if (st) return st;
  gather_rows_kernel<DIM, NB, NV, KIND><<<P->nblocks, kBR, smem, h->stream>>>(
      P->d_rows, P->d_block_ptr, P->d_pairs, P->d_stage, h->d_findrm, (size_t)h->nnz, P->maxlen, out0, out1, out2);
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

template <int DIM>
static int gather_momentum_dim(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  constexpr int LOC = DIM + 1;
  GatherPlan* P = h->gather;
  const int ne = h->n_elements, grid = (ne + 127) / 128;
  const int abs_mode = !A.o.have_absorption ? 0 : (A.o.lump_absorption ? 1 : 2);
  const bool mld = abs_mode == 1 && A.o.pressure_corrected_absorption;
  const bool fast = abs_mode != 2 && momentum_fast_ok(A.o, A.gravity.stride, A.absorption.stride) &&
                    !getenv("CGASM_GATHER_GENERIC");
  double* ml = want_ml ? h->d_masslump : nullptr;
  int st;
  // the STRIP kernels cover more option sets than the GATHER fast path (full absorption, sources, the reference
  // profile through the additive pass): ask them first; ct_m is a separate pass over the pair lists either way
  const bool strip = h->scatter == CGASM_SCATTER_STRIP && strip_momentum_ok(h, A, want_ml);
  if (strip && (st = strip_momentum(h, A))) return st;
  if (!strip) h->mom_path = (fast && A.tab.sym && !getenv("CGASM_GATHER_STAGED")) ? CGASM_PATH_GATHER_ROWS : CGASM_PATH_GATHER_STAGED;
  const bool direct = !strip && fast && A.tab.sym && !getenv("CGASM_GATHER_STAGED");
#define STAGE_SIZE(NB_, NV_) ((size_t)ne * LOC * Rec<LOC, NB_, NV_>::RS)
  const int stab = A.o.stabilisation_scheme;
#define STAGE_GENERIC(LABS_, NB_, MLC_)                                                                             \
  do {                                                                                                              \
    if (stab == 1) gather_momentum_stage_kernel<DIM, LABS_, NB_, MLC_, 1><<<grid, 128, 0, h->stream>>>(A, P->d_stage);      \
    else if (stab == 2) gather_momentum_stage_kernel<DIM, LABS_, NB_, MLC_, 2><<<grid, 128, 0, h->stream>>>(A, P->d_stage); \
    else gather_momentum_stage_kernel<DIM, LABS_, NB_, MLC_, 0><<<grid, 128, 0, h->stream>>>(A, P->d_stage);                \
  } while (0)
  if (direct) {
    const int nb = abs_mode ? DIM : 1;
    const size_t smem = sizeof(double) * (size_t)nb * P->maxlen * kBR;
    if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "gather scatter: CSR rows too long for the shared-memory accumulator");
    const int prefetch = getenv("CGASM_GATHER_PREFETCH") ? atoi(getenv("CGASM_GATHER_PREFETCH")) : 0;
    const int minb = getenv("CGASM_GATHER_MINB") ? atoi(getenv("CGASM_GATHER_MINB")) : 4;
#define LAUNCH_DIRECT(PERD_, MLD_, COMMON_, MINB_)                                                             \
  do {                                                                                                         \
    if ((st = set_dyn_smem(gather_momentum_direct_kernel<DIM, PERD_, MLD_, COMMON_, MINB_>, smem))) return st; \
    gather_momentum_direct_kernel<DIM, PERD_, MLD_, COMMON_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(     \
        A, P->d_rows, P->d_block_ptr, P->d_pair_nodes, h->d_findrm, (size_t)h->nnz,                           \
        P->maxlen, prefetch, h->d_big_m,                                                                       \
        h->d_mom_rhs, ml);                                                                                     \
  } while (0)
    const bool use_walk = P->d_walk && abs_mode == 0 && !getenv("CGASM_GATHER_DIRECT");
    if (use_walk) {
      const int wminb = getenv("CGASM_WALK_MINB") ? atoi(getenv("CGASM_WALK_MINB")) : 3;
      const int wprefetch = getenv("CGASM_WALK_PREFETCH") ? atoi(getenv("CGASM_WALK_PREFETCH")) : 0;
#define LAUNCH_WALK(COMMON_, MINB_)                                                                            \
  do {                                                                                                         \
    if ((st = set_dyn_smem(gather_momentum_walk_kernel<DIM, COMMON_, MINB_>, smem))) return st;                \
    gather_momentum_walk_kernel<DIM, COMMON_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(                    \
        A, P->d_rows, P->d_walk_ptr, P->d_walk, P->d_own_slot, h->d_findrm, (size_t)h->nnz, P->maxlen,         \
        wprefetch, h->d_big_m, h->d_mom_rhs, ml);                                                              \
  } while (0)
      if (momentum_common_ok(A.o, A.viscosity.stride) && want_ml) {
        if (wminb >= 4) LAUNCH_WALK(true, 4);
        else if (wminb == 3) LAUNCH_WALK(true, 3);
        else LAUNCH_WALK(true, 2);
      } else LAUNCH_WALK(false, 2);
#undef LAUNCH_WALK
    } else if (abs_mode == 0 && momentum_common_ok(A.o, A.viscosity.stride) && want_ml) {
      if (minb >= 6) LAUNCH_DIRECT(false, false, true, 6);
      else if (minb == 5) LAUNCH_DIRECT(false, false, true, 5);
      else LAUNCH_DIRECT(false, false, true, 4);
    } else if (abs_mode == 0) LAUNCH_DIRECT(false, false, false, 4);
    else if (!mld) LAUNCH_DIRECT(true, false, false, 4);
    else LAUNCH_DIRECT(true, true, false, 4);
#undef LAUNCH_DIRECT
    h->launches++;
  } else if (strip) {
    // done above
  } else if (abs_mode == 0) {
    if ((st = ensure_stage(P, STAGE_SIZE(1, DIM + 1)))) return st;
    if (fast) gather_momentum_stage_fast_kernel<DIM, false, 1><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
    else STAGE_GENERIC(false, 1, 1);
    h->launches++;
    if ((st = launch_rows<DIM, 1, DIM + 1, 0>(h, h->d_big_m, h->d_mom_rhs, ml))) return st;
  } else if (!mld) {
    if ((st = ensure_stage(P, STAGE_SIZE(DIM, DIM + 1)))) return st;
    if (fast) gather_momentum_stage_fast_kernel<DIM, true, 1><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
    else if (abs_mode == 2) STAGE_GENERIC(true, DIM, 1);
    else STAGE_GENERIC(false, DIM, 1);
    h->launches++;
    if ((st = launch_rows<DIM, DIM, DIM + 1, 0>(h, h->d_big_m, h->d_mom_rhs, ml))) return st;
  } else {
    if ((st = ensure_stage(P, STAGE_SIZE(DIM, 2 * DIM)))) return st;
    if (fast) gather_momentum_stage_fast_kernel<DIM, true, DIM><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
    else STAGE_GENERIC(false, DIM, DIM);
    h->launches++;
    if ((st = launch_rows<DIM, DIM, 2 * DIM, 0>(h, h->d_big_m, h->d_mom_rhs, ml))) return st;
  }
  if (want_ct) {
    if ((st = ensure_stage(P, STAGE_SIZE(DIM, 0)))) return st;
    gather_ct_stage_kernel<DIM><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
    h->launches++;
    if ((st = launch_rows<DIM, DIM, 0, 2>(h, h->d_ct_m, nullptr, nullptr))) return st;
  }
#undef STAGE_SIZE
#undef STAGE_GENERIC
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int gather_momentum(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  if (!h->gather) CG_FAIL(CGASM_ESTATE, "gather plan missing");
  const bool strip = h->scatter == CGASM_SCATTER_STRIP && strip_momentum_ok(h, A, want_ml);
  if (!strip || want_ct) {
    int st = gather_build_pairs(h);  // no-op when they exist
    if (st) return st;
  }
  if (!strip) {  // only the STRIP kernels overlap a pending halo exchange with their halo-independent blocks
    int st = halo_join(h);
    if (st) return st;
  }
  return h->dim == 3 ? gather_momentum_dim<3>(h, A, want_ml, want_ct) : gather_momentum_dim<2>(h, A, want_ml, want_ct);
}

template <int DIM>
static int gather_advdiff_dim(Handle* h, const AdvDiffArgs& A) {
  constexpr int LOC = DIM + 1;
  GatherPlan* P = h->gather;
  const int ne = h->n_elements, grid = (ne + 127) / 128;
  int st;
  // the STRIP kernels cover more option sets than the GATHER fast path (absorption, sources): ask them first
  if (h->scatter == CGASM_SCATTER_STRIP && strip_advdiff_ok(h, A)) {
    if ((st = strip_advdiff(h, A))) return st;
    CG_CUDA(cudaGetLastError());
    return CGASM_OK;
  }
  h->adv_path = CGASM_PATH_GATHER_STAGED;
  if (advdiff_fast_ok(A.o) && A.tab.sym && !getenv("CGASM_GATHER_GENERIC") && !getenv("CGASM_GATHER_STAGED")) {
    h->adv_path = CGASM_PATH_GATHER_ROWS;
    const size_t smem = sizeof(double) * (size_t)P->maxlen * kBR;
    if (smem > 200 * 1024) CG_FAIL(CGASM_EUNSUPPORTED, "gather scatter: CSR rows too long for the shared-memory accumulator");
    const int prefetch = getenv("CGASM_GATHER_PREFETCH") ? atoi(getenv("CGASM_GATHER_PREFETCH")) : 0;
    const int minb = getenv("CGASM_GATHER_MINB") ? atoi(getenv("CGASM_GATHER_MINB")) : 6;
#define LAUNCH_ADIRECT(COMMON_, MINB_)                                                                   \
  do {                                                                                                   \
    if ((st = set_dyn_smem(gather_advdiff_direct_kernel<DIM, COMMON_, MINB_>, smem))) return st;         \
    gather_advdiff_direct_kernel<DIM, COMMON_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(             \
        A, P->d_rows, P->d_block_ptr, P->d_pair_nodes, h->d_findrm, P->maxlen,                          \
        prefetch, h->d_adv_matrix,                                                                       \
        h->d_adv_rhs);                                                                                   \
  } while (0)
    if (P->d_walk && !getenv("CGASM_GATHER_DIRECT")) {
      const int wminb = getenv("CGASM_WALK_MINB") ? atoi(getenv("CGASM_WALK_MINB")) : 4;
      const int wprefetch = getenv("CGASM_WALK_PREFETCH") ? atoi(getenv("CGASM_WALK_PREFETCH")) : 0;
#define LAUNCH_AWALK(COMMON_, MINB_)                                                                     \
  do {                                                                                                   \
    if ((st = set_dyn_smem(gather_advdiff_walk_kernel<DIM, COMMON_, MINB_>, smem))) return st;           \
    gather_advdiff_walk_kernel<DIM, COMMON_, MINB_><<<P->nblocks, kBR, smem, h->stream>>>(               \
        A, P->d_rows, P->d_walk_ptr, P->d_walk, P->d_own_slot, h->d_findrm, P->maxlen, wprefetch,        \
        h->d_adv_matrix, h->d_adv_rhs);                                                                  \
  } while (0)
      if (advdiff_common_ok(A.o, A.diffusivity.stride)) {
        if (wminb >= 5) LAUNCH_AWALK(true, 5);
        else if (wminb == 4) LAUNCH_AWALK(true, 4);
        else LAUNCH_AWALK(true, 3);
      } else LAUNCH_AWALK(false, 3);
#undef LAUNCH_AWALK
    } else if (advdiff_common_ok(A.o, A.diffusivity.stride)) {
      if (minb >= 6) LAUNCH_ADIRECT(true, 6);
      else if (minb == 5) LAUNCH_ADIRECT(true, 5);
      else LAUNCH_ADIRECT(true, 4);
    } else LAUNCH_ADIRECT(false, 4);
#undef LAUNCH_ADIRECT
    h->launches++;
    CG_CUDA(cudaGetLastError());
    return CGASM_OK;
  }
  if ((st = ensure_stage(P, (size_t)ne * LOC * Rec<LOC, 1, 1>::RS))) return st;
  if (advdiff_fast_ok(A.o) && !getenv("CGASM_GATHER_GENERIC"))
    gather_advdiff_stage_fast_kernel<DIM><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
  else if (A.o.stabilisation_scheme == 1)
    gather_advdiff_stage_kernel<DIM, 1><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
  else if (A.o.stabilisation_scheme == 2)
    gather_advdiff_stage_kernel<DIM, 2><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
  else
    gather_advdiff_stage_kernel<DIM, 0><<<grid, 128, 0, h->stream>>>(A, P->d_stage);
  h->launches++;
  return launch_rows<DIM, 1, 1, 1>(h, h->d_adv_matrix, h->d_adv_rhs, nullptr);
}

int gather_advdiff(Handle* h, const AdvDiffArgs& A) {
  if (!h->gather) CG_FAIL(CGASM_ESTATE, "gather plan missing");
  if (!(h->scatter == CGASM_SCATTER_STRIP && strip_advdiff_ok(h, A))) {
    int st = gather_build_pairs(h);
    if (st) return st;
    if ((st = halo_join(h))) return st;
  }
  return h->dim == 3 ? gather_advdiff_dim<3>(h, A) : gather_advdiff_dim<2>(h, A);
}

}  // namespace cgasm
