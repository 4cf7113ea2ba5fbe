// coo.cu -- device hand-off of the assembled matrices to PETSc without the host round trip.
//
// The reference inserts every element block into a petsc_csr_matrix with MatSetValues through the row / column
// numberings gnn2unn (femtools/Sparse_Tools_Petsc.F90:848-879, femtools/Petsc_Tools.F90:141-306; rows of nodes the
// process does not own are masked with -1, Sparse_Tools_Petsc.F90:220-227, and PETSc skips negative indices). With the
// matrices resident on the device the same insertion is MatSetPreallocationCOO(A, n, i, j) once per sparsity and
// MatSetValuesCOO(A, v, INSERT_VALUES) per assembly, both of which take device pointers in PETSc's CUDA back ends:
//   pattern   (i, j) = (gnn2unn_row(row node, d), gnn2unn_col(column node, d)) for every entry of every diagonal block,
//             in the order [block d][CSR entry k] -- built once by a kernel from findrm / colm and the caller's numberings;
//   values    uncompacted: the assembly's own device array big_m[d][k] (zero copy: v IS the result buffer);
//             compacted (masked entries removed): one gather kernel per assembly through a stored source index.
// PETSc itself is not linked here (absent from the image): the entry points return device pointers, tests read them
// back and compare with formats.blocks_to_petsc (the matrix Sparse_Tools_Petsc.F90 would have built).
#include "cgasm_internal.h"

#include <cub/device/device_scan.cuh>

namespace cgasm {

struct CooPlan {
  long long ncoo = 0;       // entries handed to PETSc
  long long nall = 0;       // nblocks * nnz
  int nblocks = 0;
  bool compact = false;
  int* d_i = nullptr;
  int* d_j = nullptr;
  long long* d_src = nullptr;  // compact: position of entry q in the [block][nnz] value array
  double* d_v = nullptr;       // compact: gathered values
};

void coo_free(Handle* h) {
  for (CooPlan*& p : h->coo) {
    if (!p) continue;
    cudaFree(p->d_i);
    cudaFree(p->d_j);
    cudaFree(p->d_src);
    cudaFree(p->d_v);
    delete p;
    p = nullptr;
  }
}

// one thread per CSR row: (i, j) of its entries in every block; keep[q] = 1 if neither index is masked
__global__ void coo_pattern_kernel(int n_nodes, int nblocks, size_t nnz, const int* __restrict__ findrm,
                                   const int* __restrict__ colm, const int* __restrict__ rown, const int* __restrict__ coln,
                                   int* __restrict__ ci, int* __restrict__ cj, int* __restrict__ keep) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int k0 = findrm[r], k1 = findrm[r + 1];
  for (int d = 0; d < nblocks; d++) {
    const int gi = rown[(size_t)d * n_nodes + r];
    for (int k = k0; k < k1; k++) {
      const int gj = coln[(size_t)d * n_nodes + colm[k]];
      const size_t q = (size_t)d * nnz + k;
      ci[q] = gi;
      cj[q] = gj;
      if (keep) keep[q] = (gi >= 0 && gj >= 0) ? 1 : 0;
    }
  }
}

__global__ void coo_compact_kernel(long long nall, const int* __restrict__ keep, const long long* __restrict__ pos,
                                   const int* __restrict__ ci, const int* __restrict__ cj, int* __restrict__ oi,
                                   int* __restrict__ oj, long long* __restrict__ src) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nall || !keep[q]) return;
  const long long o = pos[q];
  oi[o] = ci[q];
  oj[o] = cj[q];
  src[o] = q;
}

__global__ void coo_gather_kernel(long long n, const long long* __restrict__ src, const double* __restrict__ val,
                                  double* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = val[src[q]];
}

__global__ void coo_keep_to_ll_kernel(long long n, const int* __restrict__ keep, long long* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = keep[q];
}

}  // namespace cgasm

using namespace cgasm;

extern "C" {

int cgasm_coo_pattern_dev(int id, int which, const int* row_gnn2unn, const int* col_gnn2unn, int compact,
                          long long* ncoo, int** coo_i_dev, int** coo_j_dev) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  if (which != CGASM_COO_MOMENTUM && which != CGASM_COO_TRACER) CG_FAIL(CGASM_EARG, "which: CGASM_COO_MOMENTUM or CGASM_COO_TRACER");
  if (!row_gnn2unn || !ncoo) CG_FAIL(CGASM_EARG, "null argument");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no sparsity");
  if (!col_gnn2unn) col_gnn2unn = row_gnn2unn;
  const int nb = which == CGASM_COO_MOMENTUM ? h->dim : 1, nn = h->n_nodes;
  const size_t nnz = (size_t)h->nnz;
  const long long nall = (long long)nb * (long long)nnz;
  CooPlan*& slot = h->coo[which];
  if (slot) {
    cudaFree(slot->d_i);
    cudaFree(slot->d_j);
    cudaFree(slot->d_src);
    cudaFree(slot->d_v);
    delete slot;
    slot = nullptr;
  }
  CooPlan* P = new CooPlan();
  P->nblocks = nb;
  P->nall = nall;
  P->compact = compact != 0;
  int *d_rown = nullptr, *d_coln = nullptr, *d_keep = nullptr, *d_ai = nullptr, *d_aj = nullptr;
  long long *d_keepll = nullptr, *d_pos = nullptr;
  void* d_tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_rown);
    cudaFree(d_coln);
    cudaFree(d_keep);
    cudaFree(d_keepll);
    cudaFree(d_pos);
    cudaFree(d_tmp);
  };
  auto fail = [&](cudaError_t e, const char* what) {
    cleanup();
    cudaFree(d_ai);
    cudaFree(d_aj);
    cudaFree(P->d_i);
    cudaFree(P->d_j);
    cudaFree(P->d_src);
    delete P;
    set_error(std::string("cgasm_coo_pattern_dev: ") + what + ": " + cudaGetErrorString(e));
    return CGASM_ECUDA;
  };
  cudaError_t e;
  const size_t nmap = sizeof(int) * (size_t)nb * nn;
  if ((e = cudaMalloc(&d_rown, nmap)) != cudaSuccess || (e = cudaMalloc(&d_coln, nmap)) != cudaSuccess ||
      (e = cudaMalloc(&d_ai, sizeof(int) * (size_t)std::max<long long>(nall, 1))) != cudaSuccess ||
      (e = cudaMalloc(&d_aj, sizeof(int) * (size_t)std::max<long long>(nall, 1))) != cudaSuccess)
    return fail(e, "cudaMalloc");
  if (P->compact && (e = cudaMalloc(&d_keep, sizeof(int) * (size_t)std::max<long long>(nall, 1))) != cudaSuccess)
    return fail(e, "cudaMalloc");
  if ((e = cudaMemcpyAsync(d_rown, row_gnn2unn, nmap, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_coln, col_gnn2unn, nmap, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess)
    return fail(e, "cudaMemcpy");
  coo_pattern_kernel<<<(nn + 127) / 128, 128, 0, h->stream>>>(nn, nb, nnz, h->d_findrm, h->d_colm, d_rown, d_coln, d_ai, d_aj,
                                                                d_keep);
  h->launches++;
  if (!P->compact) {
    P->d_i = d_ai;
    P->d_j = d_aj;
    d_ai = d_aj = nullptr;
    P->ncoo = nall;
  } else {
    if ((e = cudaMalloc(&d_keepll, sizeof(long long) * (size_t)std::max<long long>(nall, 1))) != cudaSuccess ||
        (e = cudaMalloc(&d_pos, sizeof(long long) * (size_t)std::max<long long>(nall, 1))) != cudaSuccess)
      return fail(e, "cudaMalloc");
    const unsigned grid = (unsigned)((nall + 255) / 256);
    if (nall) coo_keep_to_ll_kernel<<<grid, 256, 0, h->stream>>>(nall, d_keep, d_keepll);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_keepll, d_pos, nall, h->stream);
    if ((e = cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1))) != cudaSuccess) return fail(e, "cudaMalloc");
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_keepll, d_pos, nall, h->stream);
    long long last_pos = 0;
    int last_keep = 0;
    if (nall) {
      if ((e = cudaMemcpyAsync(&last_pos, d_pos + nall - 1, sizeof last_pos, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess ||
          (e = cudaMemcpyAsync(&last_keep, d_keep + nall - 1, sizeof last_keep, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
        return fail(e, "cudaMemcpy");
    }
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e, "scan");
    P->ncoo = last_pos + last_keep;
    const size_t nk = (size_t)std::max<long long>(P->ncoo, 1);
    if ((e = cudaMalloc(&P->d_i, sizeof(int) * nk)) != cudaSuccess || (e = cudaMalloc(&P->d_j, sizeof(int) * nk)) != cudaSuccess ||
        (e = cudaMalloc(&P->d_src, sizeof(long long) * nk)) != cudaSuccess)
      return fail(e, "cudaMalloc");
    if (nall) coo_compact_kernel<<<grid, 256, 0, h->stream>>>(nall, d_keep, d_pos, d_ai, d_aj, P->d_i, P->d_j, P->d_src);
    h->launches += 4;
  }
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e, "pattern kernels");
  cleanup();
  cudaFree(d_ai);
  cudaFree(d_aj);
  slot = P;
  *ncoo = P->ncoo;
  if (coo_i_dev) *coo_i_dev = P->d_i;
  if (coo_j_dev) *coo_j_dev = P->d_j;
  return CGASM_OK;
}

int cgasm_coo_values_dev(int id, int which, double** coo_v_dev) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  if (which != CGASM_COO_MOMENTUM && which != CGASM_COO_TRACER) CG_FAIL(CGASM_EARG, "which: CGASM_COO_MOMENTUM or CGASM_COO_TRACER");
  if (!coo_v_dev) CG_FAIL(CGASM_EARG, "null out");
  CooPlan* P = h->coo[which];
  if (!P) CG_FAIL(CGASM_ESTATE, "cgasm_coo_pattern_dev has not been called for this matrix (or the sparsity changed since)");
  const bool valid = which == CGASM_COO_MOMENTUM ? h->mom_valid : h->adv_valid;
  if (!valid) CG_FAIL(CGASM_ESTATE, "no assembled result");
  const double* val = which == CGASM_COO_MOMENTUM ? h->d_big_m : h->d_adv_matrix;
  if (!P->compact) {
    *coo_v_dev = const_cast<double*>(val);  // the result buffer itself, [block][CSR entry]
    return CGASM_OK;
  }
  if (!P->d_v) CG_CUDA(cudaMalloc(&P->d_v, sizeof(double) * (size_t)std::max<long long>(P->ncoo, 1)));
  if (P->ncoo) {
    coo_gather_kernel<<<(unsigned)((P->ncoo + 255) / 256), 256, 0, h->stream>>>(P->ncoo, P->d_src, val, P->d_v);
    h->launches++;
    CG_CUDA(cudaGetLastError());
  }
  *coo_v_dev = P->d_v;
  return CGASM_OK;
}

int cgasm_coo_fetch(int id, int which, long long ncoo, int* coo_i, int* coo_j, double* coo_v) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  if (which != CGASM_COO_MOMENTUM && which != CGASM_COO_TRACER) CG_FAIL(CGASM_EARG, "which: CGASM_COO_MOMENTUM or CGASM_COO_TRACER");
  CooPlan* P = h->coo[which];
  if (!P) CG_FAIL(CGASM_ESTATE, "cgasm_coo_pattern_dev has not been called for this matrix");
  if (ncoo != P->ncoo) CG_FAIL(CGASM_EARG, "ncoo does not match the pattern");
  if (coo_i) CG_CUDA(cudaMemcpyAsync(coo_i, P->d_i, sizeof(int) * (size_t)ncoo, cudaMemcpyDeviceToHost, h->stream));
  if (coo_j) CG_CUDA(cudaMemcpyAsync(coo_j, P->d_j, sizeof(int) * (size_t)ncoo, cudaMemcpyDeviceToHost, h->stream));
  if (coo_v) {
    double* v = nullptr;
    int st = cgasm_coo_values_dev(id, which, &v);
    if (st) return st;
    CG_CUDA(cudaMemcpyAsync(coo_v, v, sizeof(double) * (size_t)ncoo, cudaMemcpyDeviceToHost, h->stream));
  }
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

}  // extern "C"
