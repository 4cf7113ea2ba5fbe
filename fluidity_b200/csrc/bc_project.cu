// bc_project.cu -- the two steps next to the momentum element loop that touch its device-resident results
// (SURVEY.md 8(f) #1 / #3): strong Dirichlet conditions on big_m and the lumped-mass velocity correction.
#include "cgasm_internal.h"
#include "surface_math.h"

namespace cgasm {

// apply_dirichlet_conditions_vector_petsc_csr (femtools/Boundary_Conditions.F90:2198-2218): lift_boundary_conditions
// (femtools/Sparse_Tools_Petsc.F90:1139-1254) = MatZeroRowsColumns(pivot 1, x, b) + fix_scaling, on the dim diagonal
// blocks. One thread per CSR row, every block: rows are independent (a thread writes only its own row's entries and its
// own rhs entries; the boundary values come from the caller's list, not from rhs), so the result is deterministic.
//   listed (r, d):     off-diagonal entries of row r in block d = 0, the diagonal keeps its value, rhs(d, r) = diag * x
//   other rows i:      rhs(d, i) -= A_d(i, r) x(r, d) for listed columns r, then A_d(i, r) = 0
// bcidx[d * n_nodes + node] = position in the value list, -1 = not listed.
template <int DIM>
__global__ void dirichlet_vector_kernel(int n_nodes, size_t nnz, const int* __restrict__ findrm, const int* __restrict__ colm,
                                        const int* __restrict__ bcidx, const double* __restrict__ x,
                                        double* __restrict__ big_m, double* __restrict__ rhs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int k0 = findrm[i], k1 = findrm[i + 1];
  for (int d = 0; d < DIM; d++) {
    const int* idx = bcidx + (size_t)d * n_nodes;
    double* A = big_m + (size_t)d * nnz;
    const int mine = idx[i];
    if (mine >= 0) {
      double diag = 0.0;
      for (int k = k0; k < k1; k++) {
        if (colm[k] == i) diag = A[k];
        else A[k] = 0.0;
      }
      rhs[(size_t)DIM * i + d] = __dmul_rn(diag, x[mine]);
    } else {
      double r = rhs[(size_t)DIM * i + d];
      bool touched = false;
      for (int k = k0; k < k1; k++) {
        const int b = idx[colm[k]];
        if (b < 0) continue;
        r = __dadd_rn(r, -__dmul_rn(A[k], x[b]));
        A[k] = 0.0;
        touched = true;
      }
      if (touched) rhs[(size_t)DIM * i + d] = r;
    }
  }
}

__global__ void dirichlet_index_kernel(int n, int n_nodes, const int* __restrict__ nodes, const int* __restrict__ comps,
                                       int* __restrict__ bcidx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) bcidx[(size_t)comps[j] * n_nodes + nodes[j]] = j;
}

// correct_masslumped_velocity (assemble/Momentum_CG.F90:2544-2575): u(d, j) += inverse_masslump(d, j) * sum_i C_d(i, j) p_i.
// One thread per velocity node j gathers column j of ct_m through the symmetric pattern (entry (i, j) sits in row i at
// the position of column j: bisection), rows i ascending = the order mult_T adds them (Sparse_Tools.F90:3908-3914);
// explicit multiplies and adds, so the sum is the reference's bit for bit.
template <int DIM>
__global__ void correct_velocity_kernel(int n_nodes, size_t nnz, const int* __restrict__ findrm, const int* __restrict__ colm,
                                        const double* __restrict__ ct, const double* __restrict__ inv_ml,
                                        const double* __restrict__ dp, double* __restrict__ u) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_nodes) return;
  double du[DIM];
  for (int d = 0; d < DIM; d++) du[d] = 0.0;
  for (int k = findrm[j]; k < findrm[j + 1]; k++) {
    const int i = colm[k];
    const int pos = csr_pos0(findrm, colm, i, j);
    if (pos < 0) continue;
    const double p = dp[i];
    for (int d = 0; d < DIM; d++) du[d] = __dadd_rn(du[d], __dmul_rn(ct[(size_t)d * nnz + pos], p));
  }
  for (int d = 0; d < DIM; d++)
    u[(size_t)DIM * j + d] = __dadd_rn(u[(size_t)DIM * j + d], __dmul_rn(du[d], inv_ml[(size_t)DIM * j + d]));
}

}  // namespace cgasm

using namespace cgasm;

#define GET_HANDLE(h, id)                                         \
  Handle* h = get_handle(id);                                     \
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");         \
  CG_CUDA(cudaSetDevice(h->device))

extern "C" {

int cgasm_momentum_dirichlet_dev(int id, int n, const int* nodes, const int* comps, const double* values) {
  GET_HANDLE(h, id);
  if (n < 0 || (n > 0 && (!nodes || !comps || !values))) CG_FAIL(CGASM_EARG, "null argument");
  if (!h->mom_valid) CG_FAIL(CGASM_ESTATE, "no momentum result: call cgasm_momentum_dev first");
  if (n == 0) return CGASM_OK;
  const int dim = h->dim, nn = h->n_nodes;
  // a (node, component) listed by two boundary conditions: the later value wins (collect_vector_dirichlet_conditions sets)
  std::vector<int> nd((size_t)n), cp((size_t)n);
  for (int j = 0; j < n; j++) {
    if (nodes[j] < 1 || nodes[j] > nn) CG_FAIL(CGASM_EARG, "Dirichlet node out of range");
    if (comps[j] < 1 || comps[j] > dim) CG_FAIL(CGASM_EARG, "Dirichlet component out of range");
    nd[j] = nodes[j] - 1;
    cp[j] = comps[j] - 1;
  }
  int *d_idx = nullptr, *d_nd = nullptr, *d_cp = nullptr;
  double* d_x = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_idx);
    cudaFree(d_nd);
    cudaFree(d_cp);
    cudaFree(d_x);
  };
  cudaError_t e;
  if ((e = cudaMalloc(&d_idx, sizeof(int) * (size_t)dim * nn)) != cudaSuccess || (e = cudaMalloc(&d_nd, sizeof(int) * n)) != cudaSuccess ||
      (e = cudaMalloc(&d_cp, sizeof(int) * n)) != cudaSuccess || (e = cudaMalloc(&d_x, sizeof(double) * n)) != cudaSuccess ||
      (e = cudaMemsetAsync(d_idx, 0xff, sizeof(int) * (size_t)dim * nn, h->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_nd, nd.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_cp, cp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_x, values, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) {
    cleanup();
    CG_CUDA(e);
  }
  if (h->mom_copy_pending) {
    cudaStreamWaitEvent(h->stream, h->ev_mom_copied, 0);
    h->mom_copy_pending = false;
  }
  // duplicates resolve to the LAST list position: launch the index kernel in list order chunks is not needed -- a plain
  // serial pass over the (short) list on the host decides, then one kernel writes the winners
  {
    std::vector<int> last((size_t)n);
    std::vector<long long> key((size_t)n);
    for (int j = 0; j < n; j++) key[j] = (long long)cp[j] * nn + nd[j];
    std::vector<int> order((size_t)n);
    for (int j = 0; j < n; j++) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
    // keep only the last occurrence of every key: overwrite earlier duplicates' node with the winner's data
    for (int q = 0; q + 1 < n; q++)
      if (key[order[q]] == key[order[q + 1]]) nd[order[q]] = -1;
    int m = 0;
    std::vector<double> x((size_t)n);
    for (int j = 0; j < n; j++)
      if (nd[j] >= 0) {
        nd[m] = nd[j];
        cp[m] = cp[j];
        x[m] = values[j];
        m++;
      }
    if (m != n) {
      cudaMemcpyAsync(d_nd, nd.data(), sizeof(int) * m, cudaMemcpyHostToDevice, h->stream);
      cudaMemcpyAsync(d_cp, cp.data(), sizeof(int) * m, cudaMemcpyHostToDevice, h->stream);
      cudaMemcpyAsync(d_x, x.data(), sizeof(double) * m, cudaMemcpyHostToDevice, h->stream);
      cudaStreamSynchronize(h->stream);  // x goes out of scope
      n = m;
    }
  }
  dirichlet_index_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(n, nn, d_nd, d_cp, d_idx);
  if (dim == 3)
    dirichlet_vector_kernel<3><<<(nn + 127) / 128, 128, 0, h->stream>>>(nn, (size_t)h->nnz, h->d_findrm, h->d_colm, d_idx, d_x,
                                                                        h->d_big_m, h->d_mom_rhs);
  else
    dirichlet_vector_kernel<2><<<(nn + 127) / 128, 128, 0, h->stream>>>(nn, (size_t)h->nnz, h->d_findrm, h->d_colm, d_idx, d_x,
                                                                        h->d_big_m, h->d_mom_rhs);
  h->launches += 2;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cleanup();
  CG_CUDA(e);
  h->mom_identical_blocks = false;  // conditions on some components only make the blocks differ
  return CGASM_OK;
}

int cgasm_correct_masslumped_velocity(int id, const double* ct_m, const double* inverse_masslump, const double* delta_p,
                                      double* u) {
  GET_HANDLE(h, id);
  if (!inverse_masslump || !delta_p || !u) CG_FAIL(CGASM_EARG, "null argument");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no sparsity");
  if (!ct_m && !(h->mom_valid && h->mom_has_ct))
    CG_FAIL(CGASM_ESTATE, "no resident ct_m: run cgasm_momentum_dev with assemble_ct_matrix_here, or pass ct_m");
  const int dim = h->dim, nn = h->n_nodes;
  const size_t nnz = (size_t)h->nnz;
  double *d_ct = nullptr, *d_iml = nullptr, *d_dp = nullptr, *d_u = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_ct);
    cudaFree(d_iml);
    cudaFree(d_dp);
    cudaFree(d_u);
  };
  cudaError_t e = cudaSuccess;
  if (ct_m && ((e = cudaMalloc(&d_ct, sizeof(double) * dim * nnz)) != cudaSuccess ||
               (e = cudaMemcpyAsync(d_ct, ct_m, sizeof(double) * dim * nnz, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess)) {
    cleanup();
    CG_CUDA(e);
  }
  const size_t nv = sizeof(double) * (size_t)dim * nn;
  if ((e = cudaMalloc(&d_iml, nv)) != cudaSuccess || (e = cudaMalloc(&d_u, nv)) != cudaSuccess ||
      (e = cudaMalloc(&d_dp, sizeof(double) * nn)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_iml, inverse_masslump, nv, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_u, u, nv, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_dp, delta_p, sizeof(double) * nn, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) {
    cleanup();
    CG_CUDA(e);
  }
  const double* ct = ct_m ? d_ct : h->d_ct_m;
  if (dim == 3)
    correct_velocity_kernel<3><<<(nn + 127) / 128, 128, 0, h->stream>>>(nn, nnz, h->d_findrm, h->d_colm, ct, d_iml, d_dp, d_u);
  else
    correct_velocity_kernel<2><<<(nn + 127) / 128, 128, 0, h->stream>>>(nn, nnz, h->d_findrm, h->d_colm, ct, d_iml, d_dp, d_u);
  h->launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(u, d_u, nv, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cleanup();
  CG_CUDA(e);
  return CGASM_OK;
}

}  // extern "C"
