// cmc.cu -- lumped-mass pressure matrix cmc_m = C_P^T M_L^-1 C next to the momentum loop (SURVEY.md 8(f) #3;
// assemble_masslumped_cmc, assemble/Assemble_CMC.F90:119-135 -> mult_div_vector_div_T,
// femtools/Sparse_Matrices_Fields.F90:590-671). Its inputs are results of cgasm_momentum_dev that are already
// resident: ct_m (assemble_ct_matrix_here) and the lumped mass; P1-P1, ctp_m = ct_m (single phase, not
// compressible). The second-order sparsity is the reference's make_sparsity_mult
// (femtools/Sparsity_Patterns.F90:150-210), adopted from the caller or built here bit-identically.
//
// Integer/byte-bound sparse work, no tensor-core shape. Two kernels, same bits:
//   expand (default)  half-warp per row i, the second-order row accumulated in shared memory by expanding
//                     (k in row i) x (j in row k) with precomputed slots and transposed positions;
//   merge (fallback: asymmetric first-order pattern, rows longer than the accumulator, CGASM_CMC_MERGE=1)
//                     warp per row i, every lane merges the sorted rows i and j like the reference's loop.
// No atomics, every entry written once, summation in the reference's order.
#include <omp.h>

#include <cstdlib>

#include "cgasm_internal.h"
#include "cmc_math.h"
#include "cmc_plan.h"

namespace cgasm {


void cmc_free(Handle* h) {
  CmcPlan* P = h->cmc;
  if (!P) return;
  cudaFree(P->d_findrm2);
  cudaFree(P->d_colm2);
  cudaFree(P->d_val);
  cudaFree(P->d_ct);
  cudaFree(P->d_inv);
  cudaFree(P->d_tpos);
  cudaFree(P->d_pptr);
  cudaFree(P->d_slots);
  cudaFree(P->d_ctT);
  cudaFree(P->d_kt);
  cudaFree(P->d_ktT);
  cudaFree(P->d_pml);
  cudaFree(P->d_pinv);
  cudaFree(P->d_kmk);
  delete P;
  h->cmc = nullptr;
}

// Row i of the second-order pattern = ascending union of the first-order rows of the nodes in row i.
template <class V>
static void second_order_sparsity(int n, const V& findrm, const V& colm,
                                  std::vector<int>& findrm2, std::vector<int>& colm2, long long* nnz2) {
  std::vector<long long> ptr((size_t)n + 1, 0);
#pragma omp parallel
  {
    std::vector<int> mark((size_t)n, -1);
#pragma omp for schedule(static)
    for (int i = 0; i < n; i++) {
      long long c = 0;
      for (int a = findrm[i]; a < findrm[i + 1]; a++) {
        const int k = colm[a];
        for (int b = findrm[k]; b < findrm[k + 1]; b++)
          if (mark[colm[b]] != i) {
            mark[colm[b]] = i;
            c++;
          }
      }
      ptr[(size_t)i + 1] = c;
    }
  }
  for (int i = 0; i < n; i++) ptr[(size_t)i + 1] += ptr[i];
  *nnz2 = ptr[n];
  if (*nnz2 > 2147483647LL) return;  // the reference's integer findrm cannot hold it either
  findrm2.resize((size_t)n + 1);
  for (int i = 0; i <= n; i++) findrm2[i] = (int)ptr[i];
  colm2.resize((size_t)*nnz2);
#pragma omp parallel
  {
    std::vector<int> mark((size_t)n, -1);
#pragma omp for schedule(static)
    for (int i = 0; i < n; i++) {
      int* out = colm2.data() + findrm2[i];
      int c = 0;
      for (int a = findrm[i]; a < findrm[i + 1]; a++) {
        const int k = colm[a];
        for (int b = findrm[k]; b < findrm[k + 1]; b++)
          if (mark[colm[b]] != i) {
            mark[colm[b]] = i;
            out[c++] = colm[b];
          }
      }
      std::sort(out, out + c);
    }
  }
}

// Expansion plan. Returns false (no plan: the merge kernel runs) if the first-order pattern is not structurally
// symmetric, a second-order row is not the union the expansion needs, or a row is too long for the accumulator.
constexpr int kExpandMaxRow2 = 384;  // 16 half-warps x 384 doubles = 48 KB of shared memory per block
template <class V>
static bool build_expand_plan(int n, const V& findrm, const V& colm,
                              const std::vector<int>& findrm2, const std::vector<int>& colm2, CmcPlan* P) {
  P->have_expand = false;
  const size_t nnz = colm.size();
  P->h_tpos.assign(nnz, -1);
  bool ok = true;
  int n2max = 0;
#pragma omp parallel for schedule(static) reduction(&& : ok) reduction(max : n2max)
  for (int k = 0; k < n; k++) {
    n2max = std::max(n2max, findrm2[k + 1] - findrm2[k]);
    for (int p = findrm[k]; p < findrm[k + 1]; p++) {
      const int j = colm[p];
      const int* b = colm.data() + findrm[j];
      const int* e = colm.data() + findrm[j + 1];
      const int* it = std::lower_bound(b, e, k);
      if (it == e || *it != k) ok = false;
      else P->h_tpos[p] = (int)(it - colm.data());
    }
  }
  if (!ok || n2max > kExpandMaxRow2) return false;
  P->n2max = n2max;
  P->slot_bytes = n2max <= 256 ? 1 : 2;
  P->h_pptr.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; i++) {
    long long c = 0;
    for (int a = findrm[i]; a < findrm[i + 1]; a++) c += findrm[colm[a] + 1] - findrm[colm[a]];
    P->h_pptr[(size_t)i + 1] = P->h_pptr[i] + c;
  }
  P->h_slots.assign((size_t)P->h_pptr[n] * P->slot_bytes, 0);
#pragma omp parallel
  {
    std::vector<int> slot_of((size_t)n, -1);
#pragma omp for schedule(static) reduction(&& : ok)
    for (int i = 0; i < n; i++) {
      for (int s = findrm2[i]; s < findrm2[i + 1]; s++) slot_of[colm2[s]] = s - findrm2[i];
      long long w = P->h_pptr[i];
      for (int a = findrm[i]; a < findrm[i + 1]; a++) {
        const int k = colm[a];
        for (int p = findrm[k]; p < findrm[k + 1]; p++, w++) {
          const int s = slot_of[colm[p]];
          if (s < 0) ok = false;  // the adopted second-order pattern lacks an entry of S.S
          else if (P->slot_bytes == 1) P->h_slots[(size_t)w] = (unsigned char)s;
          else reinterpret_cast<unsigned short*>(P->h_slots.data())[(size_t)w] = (unsigned short)s;
        }
      }
      for (int s = findrm2[i]; s < findrm2[i + 1]; s++) slot_of[colm2[s]] = -1;
    }
  }
  P->have_expand = ok;
  return ok;
}

static int upload_pattern(Handle* h) {
  CmcPlan* P = h->cmc;
  cudaFree(P->d_findrm2);
  cudaFree(P->d_colm2);
  cudaFree(P->d_val);
  P->d_findrm2 = P->d_colm2 = nullptr;
  P->d_val = nullptr;
  P->valid = false;
  cudaFree(P->d_kmk);
  P->d_kmk = nullptr;
  P->kmk_valid = false;
  CG_CUDA(cudaMalloc(&P->d_findrm2, sizeof(int) * P->h_findrm2.size()));
  CG_CUDA(cudaMalloc(&P->d_colm2, sizeof(int) * std::max<size_t>(P->h_colm2.size(), 1)));
  CG_CUDA(cudaMalloc(&P->d_val, sizeof(double) * std::max<size_t>(P->h_colm2.size(), 1)));
  CG_CUDA(cudaMemcpyAsync(P->d_findrm2, P->h_findrm2.data(), sizeof(int) * P->h_findrm2.size(), cudaMemcpyHostToDevice, h->stream));
  CG_CUDA(cudaMemcpyAsync(P->d_colm2, P->h_colm2.data(), sizeof(int) * P->h_colm2.size(), cudaMemcpyHostToDevice, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(P->d_tpos);
  cudaFree(P->d_pptr);
  cudaFree(P->d_slots);
  cudaFree(P->d_ctT);
  P->d_ctT = nullptr;
  P->d_tpos = nullptr;
  P->d_pptr = nullptr;
  P->d_slots = nullptr;
  if (build_expand_plan(h->n_nodes, h->h_findrm, h->h_colm, P->h_findrm2, P->h_colm2, P)) {
    CG_CUDA(cudaMalloc(&P->d_tpos, sizeof(int) * std::max<size_t>(P->h_tpos.size(), 1)));
    CG_CUDA(cudaMalloc(&P->d_pptr, sizeof(long long) * P->h_pptr.size()));
    CG_CUDA(cudaMalloc(&P->d_slots, std::max<size_t>(P->h_slots.size(), 1)));
    CG_CUDA(cudaMemcpyAsync(P->d_tpos, P->h_tpos.data(), sizeof(int) * P->h_tpos.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaMemcpyAsync(P->d_pptr, P->h_pptr.data(), sizeof(long long) * P->h_pptr.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaMemcpyAsync(P->d_slots, P->h_slots.data(), P->h_slots.size(), cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaStreamSynchronize(h->stream));
  }
  // the host copies of the plan are only needed by the diagnostics entry point
  std::vector<unsigned char>().swap(P->h_slots);
  std::vector<int>().swap(P->h_tpos);
  std::vector<long long>().swap(P->h_pptr);
  return CGASM_OK;
}

constexpr int kCmcWarps = 8;  // rows per block

template <int DIM>
__global__ void __launch_bounds__(kCmcWarps * 32)
cmc_kernel(int n_rows, const int* __restrict__ findrm, const int* __restrict__ colm, const double* __restrict__ ct, size_t nnz,
           const double* __restrict__ inv, const int* __restrict__ findrm2, const int* __restrict__ colm2,
           double* __restrict__ out) {
  const int i = blockIdx.x * kCmcWarps + (threadIdx.x >> 5);
  if (i >= n_rows) return;
  const int lane = threadIdx.x & 31;
  for (int e = findrm2[i] + lane; e < findrm2[i + 1]; e += 32)
    out[e] = cmc_entry<DIM>(findrm, colm, ct, ct, nnz, inv, i, colm2[e]);
}

// Expansion kernel: one HALF-warp per row i (first-order rows of P1 meshes have ~15-27 entries: 16 lanes keep them
// busy), 16 rows per block. The second-order row is accumulated in shared memory; for each column k of row i, in
// ascending order, the lanes take the entries (k, j) of row k -- distinct j, so distinct accumulator slots, plain
// read-modify-write -- and a __syncwarp separates the k steps. Slots and transposed positions come from the plan:
// no searching, no merging, the plan and row k are read coalesced. Same bits as the merge kernel.

// ctT[d][p] = ct[d][tpos[p]]: one gathered pass over the first-order entries, so that the expansion reads the
// factors C(j,k), j in row k, contiguously (every row k is read by all ~15-27 rows i that contain k)
template <int DIM>
__global__ void transpose_ct_kernel(size_t nnz, const int* __restrict__ tpos, const double* __restrict__ ct,
                                    double* __restrict__ ctT) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int t = tpos[p];
  for (int d = 0; d < DIM; d++) ctT[d * nnz + p] = ct[d * nnz + t];
}

template <int DIM, class SLOT>
__global__ void __launch_bounds__(kExpandRows * 16)
cmc_expand_kernel(int n_rows, const int* __restrict__ findrm, const int* __restrict__ colm, const double* __restrict__ ct,
                  const double* __restrict__ ctT, size_t nnz, const double* __restrict__ inv, const long long* __restrict__ pptr,
                  const SLOT* __restrict__ slots, const int* __restrict__ findrm2, int n2max, double* __restrict__ out) {
  extern __shared__ double cmc_acc[];
  const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
  const int i = blockIdx.x * kExpandRows + hw;
  double* acc = cmc_acc + (size_t)hw * n2max;
  const bool live = i < n_rows;
  const int r0 = live ? findrm[i] : 0;
  const int n1 = live ? findrm[i + 1] - r0 : 0;
  const int o0 = live ? findrm2[i] : 0;
  const int n2 = live ? findrm2[i + 1] - o0 : 0;
  for (int s = hl; s < n2; s += 16) acc[s] = 0.0;
  const int n1_both = max(n1, __shfl_xor_sync(0xffffffffu, n1, 16));  // the two halves of a warp step together
  __syncwarp();
  long long pp = live ? pptr[i] : 0;
  for (int a = 0; a < n1_both; a++) {
    if (a < n1) {
      const int k = colm[r0 + a];
      double Ad[DIM], Wd[DIM];
      for (int d = 0; d < DIM; d++) {
        Ad[d] = ct[d * nnz + r0 + a];
        Wd[d] = inv[(size_t)DIM * k + d];
      }
      const int kb = findrm[k], kn = findrm[k + 1] - kb;
      for (int q = hl; q < kn; q += 16) {
        const int s = (int)slots[pp + q];
        acc[s] = cmc_accumulate<DIM>(acc[s], Ad, Wd, ctT, nnz, kb + q);
      }
      pp += kn;
    }
    __syncwarp();
  }
  for (int s = hl; s < n2; s += 16) out[o0 + s] = acc[s];
}

// invert(inverse_masslump) (assemble/Momentum_CG.F90:873): 1/x per entry
__global__ void invert_kernel(size_t n, const double* __restrict__ x, double* __restrict__ y) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) y[k] = 1.0 / x[k];
}

// K diag(w) K^T for a scalar matrix on the first-order sparsity (kmk.cu): the DIM = 1 instances of the kernels above
int cmc_scalar_product(Handle* h, const double* d_a, double* d_aT, const double* d_w, double* d_out) {
  CmcPlan* P = h->cmc;
  const size_t nnz = (size_t)h->nnz;
  const bool expand = P->d_slots && !getenv("CGASM_CMC_MERGE");
  if (expand) {
    const int blocks = (h->n_nodes + kExpandRows - 1) / kExpandRows;
    const size_t smem = sizeof(double) * (size_t)kExpandRows * P->n2max;
    transpose_ct_kernel<1><<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(nnz, P->d_tpos, d_a, d_aT);
    h->launches++;
    if (P->slot_bytes == 1)
      cmc_expand_kernel<1, unsigned char><<<blocks, kExpandRows * 16, smem, h->stream>>>(
          h->n_nodes, h->d_findrm, h->d_colm, d_a, d_aT, nnz, d_w, P->d_pptr, P->d_slots, P->d_findrm2, P->n2max, d_out);
    else
      cmc_expand_kernel<1, unsigned short><<<blocks, kExpandRows * 16, smem, h->stream>>>(
          h->n_nodes, h->d_findrm, h->d_colm, d_a, d_aT, nnz, d_w, P->d_pptr, reinterpret_cast<const unsigned short*>(P->d_slots),
          P->d_findrm2, P->n2max, d_out);
  } else {
    const int blocks = (h->n_nodes + kCmcWarps - 1) / kCmcWarps;
    cmc_kernel<1><<<blocks, kCmcWarps * 32, 0, h->stream>>>(h->n_nodes, h->d_findrm, h->d_colm, d_a, nnz, d_w, P->d_findrm2,
                                                            P->d_colm2, d_out);
  }
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

}  // namespace cgasm

using namespace cgasm;

#define GET_HANDLE(h, id)                                         \
  Handle* h = get_handle(id);                                     \
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");         \
  CG_CUDA(cudaSetDevice(h->device))

extern "C" {

int cgasm_cmc_build_sparsity(int id, long long* nnz2) {
  GET_HANDLE(h, id);
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no first-order sparsity: call cgasm_build_sparsity or cgasm_set_sparsity");
  if (!h->cmc) h->cmc = new CmcPlan();
  CmcPlan* P = h->cmc;
  second_order_sparsity(h->n_nodes, h->h_findrm, h->h_colm, P->h_findrm2, P->h_colm2, &P->nnz2);
  if (nnz2) *nnz2 = P->nnz2;
  if (P->nnz2 > 2147483647LL) CG_FAIL(CGASM_EUNSUPPORTED, "second-order sparsity exceeds 32-bit indices: partition the mesh");
  return upload_pattern(h);
}

int cgasm_cmc_sparsity_host(int n_nodes, const int* findrm, const int* colm, int* findrm2, int* colm2, long long capacity,
                            long long* needed) {
  if (n_nodes < 1 || !findrm || !colm || !needed) CG_FAIL(CGASM_EARG, "null argument");
  std::vector<int> f0((size_t)n_nodes + 1), c0((size_t)(findrm[n_nodes] - 1)), f2, c2;
  for (int i = 0; i <= n_nodes; i++) f0[i] = findrm[i] - 1;
  for (size_t k = 0; k < c0.size(); k++) c0[k] = colm[k] - 1;
  second_order_sparsity(n_nodes, f0, c0, f2, c2, needed);
  if (*needed > 2147483647LL) CG_FAIL(CGASM_EUNSUPPORTED, "second-order sparsity exceeds 32-bit indices");
  if (findrm2)
    for (int i = 0; i <= n_nodes; i++) findrm2[i] = f2[i] + 1;
  if (colm2 && *needed <= capacity)
    for (size_t k = 0; k < c2.size(); k++) colm2[k] = c2[k] + 1;
  return CGASM_OK;
}

int cgasm_cmc_expand_plan_host(int n_nodes, const int* findrm, const int* colm, const int* findrm2, const int* colm2,
                               int* tpos, long long* pptr, unsigned short* slots, long long capacity, long long* needed,
                               int* n2max) {
  if (n_nodes < 1 || !findrm || !colm || !findrm2 || !colm2 || !needed) CG_FAIL(CGASM_EARG, "null argument");
  auto zero_based = [](const int* a, size_t n) {
    std::vector<int> v(n);
    for (size_t k = 0; k < n; k++) v[k] = a[k] - 1;
    return v;
  };
  const std::vector<int> f0 = zero_based(findrm, (size_t)n_nodes + 1), c0 = zero_based(colm, (size_t)(findrm[n_nodes] - 1));
  const std::vector<int> f2 = zero_based(findrm2, (size_t)n_nodes + 1), c2 = zero_based(colm2, (size_t)(findrm2[n_nodes] - 1));
  CmcPlan P;
  if (!build_expand_plan(n_nodes, f0, c0, f2, c2, &P)) {
    *needed = -1;  // no expansion plan for these patterns: the merge kernel would run
    return CGASM_OK;
  }
  *needed = P.h_pptr[n_nodes];
  if (n2max) *n2max = P.n2max;
  if (tpos)
    for (size_t k = 0; k < P.h_tpos.size(); k++) tpos[k] = P.h_tpos[k];
  if (pptr)
    for (size_t k = 0; k < P.h_pptr.size(); k++) pptr[k] = P.h_pptr[k];
  if (slots && *needed <= capacity)
    for (long long k = 0; k < *needed; k++)
      slots[k] = P.slot_bytes == 1 ? (unsigned short)P.h_slots[(size_t)k]
                                   : reinterpret_cast<const unsigned short*>(P.h_slots.data())[(size_t)k];
  return CGASM_OK;
}

int cgasm_cmc_get_sparsity(int id, int* findrm2, int* colm2) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || P->h_findrm2.empty()) CG_FAIL(CGASM_ESTATE, "no second-order sparsity");
  if (findrm2)
    for (size_t i = 0; i < P->h_findrm2.size(); i++) findrm2[i] = P->h_findrm2[i] + 1;
  if (colm2)
    for (size_t k = 0; k < P->h_colm2.size(); k++) colm2[k] = P->h_colm2[k] + 1;
  return CGASM_OK;
}

int cgasm_cmc_set_sparsity(int id, int rows, int nnz2, const int* findrm2, const int* colm2) {
  GET_HANDLE(h, id);
  if (!findrm2 || !colm2 || rows != h->n_nodes || nnz2 < 0) CG_FAIL(CGASM_EARG, "bad second-order sparsity");
  if (findrm2[0] != 1 || findrm2[rows] != nnz2 + 1) CG_FAIL(CGASM_EARG, "findrm does not span colm");
  for (int i = 0; i < rows; i++) {
    if (findrm2[i + 1] < findrm2[i]) CG_FAIL(CGASM_EARG, "findrm is not monotone");
    for (int k = findrm2[i] - 1; k < findrm2[i + 1] - 1; k++)
      if (colm2[k] < 1 || colm2[k] > rows || (k > findrm2[i] - 1 && colm2[k] <= colm2[k - 1]))
        CG_FAIL(CGASM_EARG, "rows of the second-order sparsity must be sorted, unique and in range");
  }
  if (!h->cmc) h->cmc = new CmcPlan();
  CmcPlan* P = h->cmc;
  P->nnz2 = nnz2;
  P->h_findrm2.resize((size_t)rows + 1);
  P->h_colm2.resize((size_t)nnz2);
  for (int i = 0; i <= rows; i++) P->h_findrm2[i] = findrm2[i] - 1;
  for (int k = 0; k < nnz2; k++) P->h_colm2[k] = colm2[k] - 1;
  return upload_pattern(h);
}

int cgasm_cmc_dev(int id, const double* ct_m, const double* inverse_masslump) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || !P->d_findrm2) CG_FAIL(CGASM_ESTATE, "no second-order sparsity: call cgasm_cmc_build_sparsity or cgasm_cmc_set_sparsity");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no first-order sparsity");
  const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes, dim = (size_t)h->dim;
  const double* ct = nullptr;
  if (ct_m) {
    if (!P->d_ct) CG_CUDA(cudaMalloc(&P->d_ct, sizeof(double) * dim * std::max<size_t>(nnz, 1)));
    CG_CUDA(cudaMemcpyAsync(P->d_ct, ct_m, sizeof(double) * dim * nnz, cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaStreamSynchronize(h->stream));
    ct = P->d_ct;
  } else {
    if (!h->mom_valid || !h->mom_has_ct)
      CG_FAIL(CGASM_ESTATE, "no resident ct_m: run cgasm_momentum_dev with assemble_ct_matrix_here, or pass ct_m");
    ct = h->d_ct_m;
  }
  if (!P->d_inv) CG_CUDA(cudaMalloc(&P->d_inv, sizeof(double) * dim * nn));
  if (inverse_masslump) {
    CG_CUDA(cudaMemcpyAsync(P->d_inv, inverse_masslump, sizeof(double) * dim * nn, cudaMemcpyHostToDevice, h->stream));
    CG_CUDA(cudaStreamSynchronize(h->stream));
  } else {
    if (!h->mom_valid || !h->mom_has_masslump)
      CG_FAIL(CGASM_ESTATE, "no resident lumped mass: run cgasm_momentum_dev with assemble_inverse_masslump, or pass inverse_masslump");
    const size_t cnt = dim * nn;
    invert_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(cnt, h->d_masslump, P->d_inv);
    h->launches++;
  }
  CG_CUDA(cudaEventRecord(h->ev0, h->stream));
  const bool expand = P->d_slots && !getenv("CGASM_CMC_MERGE");
  if (expand) {
    const int blocks = (h->n_nodes + kExpandRows - 1) / kExpandRows;
    const size_t smem = sizeof(double) * (size_t)kExpandRows * P->n2max;
    if (!P->d_ctT) CG_CUDA(cudaMalloc(&P->d_ctT, sizeof(double) * dim * std::max<size_t>(nnz, 1)));
    const unsigned tb = (unsigned)((nnz + 255) / 256);
    if (h->dim == 3) transpose_ct_kernel<3><<<tb, 256, 0, h->stream>>>(nnz, P->d_tpos, ct, P->d_ctT);
    else transpose_ct_kernel<2><<<tb, 256, 0, h->stream>>>(nnz, P->d_tpos, ct, P->d_ctT);
    h->launches++;
#define EXPAND(DIM_, SLOT_)                                                                                              \
  cmc_expand_kernel<DIM_, SLOT_><<<blocks, kExpandRows * 16, smem, h->stream>>>(                                          \
      h->n_nodes, h->d_findrm, h->d_colm, ct, P->d_ctT, nnz, P->d_inv, P->d_pptr, reinterpret_cast<const SLOT_*>(P->d_slots), \
      P->d_findrm2, P->n2max, P->d_val)
    if (h->dim == 3) {
      if (P->slot_bytes == 1) EXPAND(3, unsigned char);
      else EXPAND(3, unsigned short);
    } else {
      if (P->slot_bytes == 1) EXPAND(2, unsigned char);
      else EXPAND(2, unsigned short);
    }
#undef EXPAND
  } else {
    const int blocks = (h->n_nodes + kCmcWarps - 1) / kCmcWarps;
    if (h->dim == 3)
      cmc_kernel<3><<<blocks, kCmcWarps * 32, 0, h->stream>>>(h->n_nodes, h->d_findrm, h->d_colm, ct, nnz, P->d_inv, P->d_findrm2,
                                                              P->d_colm2, P->d_val);
    else
      cmc_kernel<2><<<blocks, kCmcWarps * 32, 0, h->stream>>>(h->n_nodes, h->d_findrm, h->d_colm, ct, nnz, P->d_inv, P->d_findrm2,
                                                              P->d_colm2, P->d_val);
  }
  h->launches++;
  CG_CUDA(cudaEventRecord(h->ev1, h->stream));
  CG_CUDA(cudaGetLastError());
  P->valid = true;
  return CGASM_OK;
}

int cgasm_cmc_fetch(int id, double* cmc_val) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || !P->valid) CG_FAIL(CGASM_ESTATE, "no cmc result to fetch");
  if (!cmc_val) CG_FAIL(CGASM_EARG, "null output");
  CG_CUDA(cudaMemcpyAsync(cmc_val, P->d_val, sizeof(double) * (size_t)P->nnz2, cudaMemcpyDeviceToHost, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

int cgasm_cmc_result_dev(int id, double** cmc_val_dev) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || !P->valid) CG_FAIL(CGASM_ESTATE, "no cmc result");
  if (!cmc_val_dev) CG_FAIL(CGASM_EARG, "null output");
  *cmc_val_dev = P->d_val;
  return CGASM_OK;
}

}  // extern "C"
