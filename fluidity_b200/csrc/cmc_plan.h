// cmc_plan.h -- the second-order sparsity and expansion plan shared by cmc.cu (C M_L^-1 C^T) and kmk.cu (K M_L^-1 K^T).
#pragma once
#include <vector>

#include "cgasm_internal.h"

namespace cgasm {

struct CmcPlan {
  long long nnz2 = 0;
  std::vector<int> h_findrm2, h_colm2;  // 0-based
  int* d_findrm2 = nullptr;
  int* d_colm2 = nullptr;
  double* d_val = nullptr;
  double* d_ct = nullptr;   // uploaded ct_m when the caller passes one
  double* d_inv = nullptr;  // inverse lumped mass (dim, n_nodes)
  bool valid = false;
  // expansion plan (expand kernel): transposed positions of the first-order entries, and for every row i the
  // second-order slot of each (k in row i, j in row k) pair, rows of k back to back
  bool have_expand = false;
  int n2max = 0;
  int slot_bytes = 0;  // 1 or 2
  std::vector<int> h_tpos;
  std::vector<long long> h_pptr;
  std::vector<unsigned char> h_slots;
  int* d_tpos = nullptr;
  long long* d_pptr = nullptr;
  unsigned char* d_slots = nullptr;
  double* d_ctT = nullptr;  // ct_m with every entry moved to its transposed position: row k holds C(j,k) for j in row k
  // P1-P1 stabilisation (kmk.cu): pressure diffusion matrix kt (nnz), its transposed copy, lumped pressure mass and
  // 1 / (theta_pg * mass) (n_nodes each), kmk on the second-order sparsity (nnz2)
  double* d_kt = nullptr;
  double* d_ktT = nullptr;
  double* d_pml = nullptr;
  double* d_pinv = nullptr;
  double* d_kmk = nullptr;
  bool kmk_valid = false;
};

constexpr int kExpandRows = 16;  // rows (half-warps) per block of the expansion kernels

// cmc.cu: product = A diag(w) A^T on the second-order sparsity for a scalar first-order matrix A (nnz values) with its
// transposed copy made here; the expansion kernel where the plan exists, the merge kernel otherwise. out: nnz2 values.
int cmc_scalar_product(Handle* h, const double* d_a, double* d_aT, const double* d_w, double* d_out);

}  // namespace cgasm
