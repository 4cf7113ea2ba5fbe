// kmk.cu -- P1-P1 pressure stabilisation matrix next to the pressure matrix:
// assemble_kmk_matrix, assemble/Momentum_CG.F90:2707-2766.
//   kt  = sum_e 0.5 dshape_tensor_dshape(dp_t, h_bar, dp_t, detwei) on the first-order pressure sparsity (:2748-2753),
//         h_bar = edge_length_from_eigenvalue(simplex_tensor(X, ele)) (error_measures/Edge_lengths.F90:68-79,
//         femtools/Metric_tools.F90:852-941,157-164): the element's metric M (e^T M e = 1 on every edge) to the power -1/2;
//   kmk = kt diag(1 / (theta_pg p_masslump)) kt^T (mult_div_invscalar_div_T, femtools/Sparse_Matrices_Fields.F90:673-748)
//         on the second-order sparsity, p_masslump = get_lumped_mass(pressure_mesh) = row sums of the P1 mass matrix.
// O(elements) work off the hot path: one thread per element, FP64 atomics into kt and the lumped mass (the scatter the
// ATOMIC variant uses), then the DIM = 1 instance of the CMC expansion kernel (cmc.cu). The metric's d x d linear
// system (d = 3 or 6) is solved by LU with partial pivoting and its eigen-decomposition taken by cyclic Jacobi
// rotations, like the oracle; the reference calls LAPACK (DGESV, DSPEV) for both.
#include <cstdlib>

#include "cmc_plan.h"
#include "surface_math.h"  // csr_pos0

namespace cgasm {

template <int DIM>
struct KmkMath {
  static constexpr int LOC = DIM + 1, D = DIM * (DIM + 1) / 2;
  // position of (k, l) in the packed unknown vector (Metric_tools.F90:919-933, 0-based)
  __host__ __device__ static int idx(int k, int l) {
    const int a = k < l ? k : l, b = k < l ? l : k;
    if (a == 0) return b;
    return DIM == 3 ? a + b + 1 : a + b;
  }

  __device__ static bool simplex_tensor(const double (&X)[LOC][DIM], double (&M)[DIM][DIM]) {
    double A[D][D], x[D];
    int n = 0;
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int j = i + 1; j < LOC; j++) {
        double diff[DIM];
#pragma unroll
        for (int a = 0; a < DIM; a++) diff[a] = X[j][a] - X[i][a];
#pragma unroll
        for (int k = 0; k < DIM; k++)
#pragma unroll
          for (int l = 0; l < DIM; l++) A[n][idx(k, l)] = diff[k] * diff[l] * (k == l ? 1.0 : 2.0);
        n++;
      }
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = 1.0;
    for (int c = 0; c < D; c++) {
      int piv = c;
      for (int r = c + 1; r < D; r++)
        if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
      if (A[piv][c] == 0.0) return false;
      if (piv != c) {
        for (int q = 0; q < D; q++) {
          const double t = A[c][q];
          A[c][q] = A[piv][q];
          A[piv][q] = t;
        }
        const double t = x[c];
        x[c] = x[piv];
        x[piv] = t;
      }
      for (int r = c + 1; r < D; r++) {
        const double f = A[r][c] / A[c][c];
        for (int q = c + 1; q < D; q++) A[r][q] -= f * A[c][q];
        x[r] -= f * x[c];
      }
    }
    for (int c = D - 1; c >= 0; c--) {
      double t = x[c];
      for (int q = c + 1; q < D; q++) t -= A[c][q] * x[q];
      x[c] = t / A[c][c];
    }
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
      for (int j = 0; j < DIM; j++) M[i][j] = x[idx(i, j)];
    return true;
  }

  // H = V diag(1 / sqrt|lambda|) V^T
  __device__ static void edge_lengths(const double (&M)[DIM][DIM], double (&H)[DIM][DIM]) {
    double A[DIM][DIM], V[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
      for (int j = 0; j < DIM; j++) {
        A[i][j] = M[i][j];
        V[i][j] = i == j ? 1.0 : 0.0;
      }
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0.0, diag = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) {
          if (i != j) off += A[i][j] * A[i][j];
          else diag += A[i][i] * A[i][i];
        }
      if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
      for (int p = 0; p < DIM; p++)
#pragma unroll
        for (int q = p + 1; q < DIM; q++) {
          const double apq = A[p][q];
          if (apq == 0.0) continue;
          const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
          for (int k = 0; k < DIM; k++) {
            const double akp = A[k][p], akq = A[k][q];
            A[k][p] = c * akp - sn * akq;
            A[k][q] = sn * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < DIM; k++) {
            const double apk = A[p][k], aqk = A[q][k];
            A[p][k] = c * apk - sn * aqk;
            A[q][k] = sn * apk + c * aqk;
          }
#pragma unroll
          for (int k = 0; k < DIM; k++) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - sn * vkq;
            V[k][q] = sn * vkp + c * vkq;
          }
        }
    }
    double ev[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) ev[i] = 1.0 / sqrt(fabs(A[i][i]));
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
      for (int j = 0; j < DIM; j++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; k++) s += V[i][k] * ev[k] * V[j][k];
        H[i][j] = s;
      }
  }
};

// one thread per element: kt += 0.5 |J| Wsum gradN_i . H gradN_j, p_masslump_i += |J| W1
template <int DIM>
__global__ void __launch_bounds__(128)
kt_element_kernel(int n_elements, const int4* __restrict__ ndglno, const double* __restrict__ X, const int* __restrict__ findrm,
                  const int* __restrict__ colm, double wsum, double w1, double* __restrict__ kt, double* __restrict__ pml,
                  int* __restrict__ bad) {
  constexpr int LOC = DIM + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elements) return;
  const int4 nd4 = ndglno[e];
  const int nodes[4] = {nd4.x, nd4.y, nd4.z, nd4.w};
  double P[LOC][DIM];
#pragma unroll
  for (int i = 0; i < LOC; i++)
#pragma unroll
    for (int a = 0; a < DIM; a++) P[i][a] = X[(size_t)DIM * nodes[i] + a];
  // gradients of the P1 basis: rows of the inverse of the edge matrix (transform_to_physical, Transform_elements.F90:807-887)
  double E[DIM][DIM], G[LOC][DIM], det;
#pragma unroll
  for (int k = 0; k < DIM; k++)
#pragma unroll
    for (int a = 0; a < DIM; a++) E[k][a] = P[k + 1][a] - P[0][a];
  if constexpr (DIM == 2) {
    det = E[0][0] * E[1][1] - E[0][1] * E[1][0];
    G[1][0] = E[1][1] / det;
    G[1][1] = -E[1][0] / det;
    G[2][0] = -E[0][1] / det;
    G[2][1] = E[0][0] / det;
  } else {
    double c[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double(&p)[3] = E[(k + 1) % 3];
      const double(&q)[3] = E[(k + 2) % 3];
      c[k][0] = p[1] * q[2] - p[2] * q[1];
      c[k][1] = p[2] * q[0] - p[0] * q[2];
      c[k][2] = p[0] * q[1] - p[1] * q[0];
    }
    det = E[0][0] * c[0][0] + E[0][1] * c[0][1] + E[0][2] * c[0][2];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int a = 0; a < 3; a++) G[k + 1][a] = c[k][a] / det;
  }
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    double s = 0.0;
#pragma unroll
    for (int k = 1; k < LOC; k++) s += G[k][a];
    G[0][a] = -s;
  }
  double M[DIM][DIM], H[DIM][DIM];
  if (!KmkMath<DIM>::simplex_tensor(P, M)) {
    atomicAdd(bad, 1);
    return;
  }
  KmkMath<DIM>::edge_lengths(M, H);
  const double scale = 0.5 * fabs(det) * wsum;
  double HG[LOC][DIM];
#pragma unroll
  for (int j = 0; j < LOC; j++)
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < DIM; b++) s += H[a][b] * G[j][b];
      HG[j][a] = s;
    }
#pragma unroll
  for (int i = 0; i < LOC; i++) {
#pragma unroll
    for (int j = 0; j < LOC; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) s += G[i][a] * HG[j][a];
      const int pos = csr_pos0(findrm, colm, nodes[i], nodes[j]);
      if (pos >= 0) atomicAdd(kt + pos, scale * s);
    }
    atomicAdd(pml + nodes[i], fabs(det) * w1);
  }
}

__global__ void kmk_weight_kernel(int n, double theta_pg, const double* __restrict__ pml, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = 1.0 / (theta_pg * pml[i]);
}

}  // namespace cgasm

using namespace cgasm;

#define GET_HANDLE(h, id)                                         \
  Handle* h = get_handle(id);                                     \
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");         \
  CG_CUDA(cudaSetDevice(h->device))

extern "C" {

int cgasm_kmk_dev(int id, double theta_pg) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || !P->d_findrm2) CG_FAIL(CGASM_ESTATE, "no second-order sparsity: call cgasm_cmc_build_sparsity or cgasm_cmc_set_sparsity");
  if (!h->have_sparsity || !h->have_X) CG_FAIL(CGASM_ESTATE, "kmk needs coordinates and the first-order sparsity");
  if (!(theta_pg > 0.0)) CG_FAIL(CGASM_EARG, "theta_pg must be positive");
  const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes, nnz2 = (size_t)P->nnz2;
  if (!P->d_kt) CG_CUDA(cudaMalloc(&P->d_kt, sizeof(double) * std::max<size_t>(nnz, 1)));
  if (!P->d_ktT) CG_CUDA(cudaMalloc(&P->d_ktT, sizeof(double) * std::max<size_t>(nnz, 1)));
  if (!P->d_pml) CG_CUDA(cudaMalloc(&P->d_pml, sizeof(double) * (nn + 1)));  // + the degenerate-element counter
  if (!P->d_pinv) CG_CUDA(cudaMalloc(&P->d_pinv, sizeof(double) * nn));
  if (!P->d_kmk) CG_CUDA(cudaMalloc(&P->d_kmk, sizeof(double) * std::max<size_t>(nnz2, 1)));
  P->kmk_valid = false;
  CG_CUDA(cudaMemsetAsync(P->d_kt, 0, sizeof(double) * nnz, h->stream));
  CG_CUDA(cudaMemsetAsync(P->d_pml, 0, sizeof(double) * (nn + 1), h->stream));
  int* d_bad = reinterpret_cast<int*>(P->d_pml + nn);
  CG_CUDA(cudaEventRecord(h->ev0, h->stream));
  const unsigned eb = (unsigned)((h->n_elements + 127) / 128);
  const Tables& t = h->tab;
  if (h->dim == 3)
    kt_element_kernel<3><<<eb, 128, 0, h->stream>>>(h->n_elements, h->d_ndglno, h->d_X, h->d_findrm, h->d_colm, t.Wsum, t.W1,
                                                    P->d_kt, P->d_pml, d_bad);
  else
    kt_element_kernel<2><<<eb, 128, 0, h->stream>>>(h->n_elements, h->d_ndglno, h->d_X, h->d_findrm, h->d_colm, t.Wsum, t.W1,
                                                    P->d_kt, P->d_pml, d_bad);
  h->launches++;
  kmk_weight_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(h->n_nodes, theta_pg, P->d_pml, P->d_pinv);
  h->launches++;
  const int st = cmc_scalar_product(h, P->d_kt, P->d_ktT, P->d_pinv, P->d_kmk);
  if (st) return st;
  CG_CUDA(cudaEventRecord(h->ev1, h->stream));
  int bad = 0;
  CG_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  if (bad) CG_FAIL(CGASM_EARG, "kmk: degenerate element (singular metric system)");
  P->kmk_valid = true;
  return CGASM_OK;
}

int cgasm_kmk_fetch(int id, double* kmk, double* kt, double* p_masslump) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || !P->kmk_valid) CG_FAIL(CGASM_ESTATE, "no kmk result to fetch");
  if (kmk) CG_CUDA(cudaMemcpyAsync(kmk, P->d_kmk, sizeof(double) * (size_t)P->nnz2, cudaMemcpyDeviceToHost, h->stream));
  if (kt) CG_CUDA(cudaMemcpyAsync(kt, P->d_kt, sizeof(double) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
  if (p_masslump) CG_CUDA(cudaMemcpyAsync(p_masslump, P->d_pml, sizeof(double) * (size_t)h->n_nodes, cudaMemcpyDeviceToHost, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

int cgasm_kmk_result_dev(int id, double** kmk_dev) {
  GET_HANDLE(h, id);
  CmcPlan* P = h->cmc;
  if (!P || !P->kmk_valid) CG_FAIL(CGASM_ESTATE, "no kmk result");
  if (!kmk_dev) CG_FAIL(CGASM_EARG, "null output");
  *kmk_dev = P->d_kmk;
  return CGASM_OK;
}

}  // extern "C"
