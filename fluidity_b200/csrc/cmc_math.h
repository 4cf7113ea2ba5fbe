// cmc_math.h -- one entry of the lumped-mass pressure matrix C M_L^-1 C^T, host/device
// like surface_math.h: the device kernel in cmc.cu and the CPU harness of tests/test_cmc.py run this function.
#pragma once
#include "surface_math.h"  // CG_HD

namespace cgasm {

// mult_div_vector_div_T (femtools/Sparse_Matrices_Fields.F90:590-671), one (i, j) entry:
//   sum_k sum_d A_d(i,k) B_d(j,k) v(d,k) over the columns k the sorted rows i and j share, walked left to
// right, components innermost -- the reference's order, with explicit multiply/add (no contraction into FMA)
// so that the result does not depend on the compiler. findrm/colm 0-based; ct: [dim][nnz]; v(dim, node).
template <int DIM>
CG_HD double cmc_entry(const int* findrm, const int* colm, const double* ct1, const double* ct2, size_t nnz,
                       const double* v, int i, int j) {
  int k1 = findrm[i], k2 = findrm[j];
  const int e1 = findrm[i + 1], e2 = findrm[j + 1];
  double acc = 0.0;
  while (k1 < e1 && k2 < e2) {
    const int a = colm[k1], b = colm[k2];
    if (a < b) {
      k1++;
    } else if (a == b) {
      for (int d = 0; d < DIM; d++) {
#if defined(__CUDA_ARCH__)
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(ct1[d * nnz + k1], ct2[d * nnz + k2]), v[(size_t)DIM * a + d]));
#else
        const volatile double p = ct1[d * nnz + k1] * ct2[d * nnz + k2];
        const volatile double q = p * v[(size_t)DIM * a + d];
        acc = acc + q;
#endif
      }
      k1++;
      k2++;
    } else {
      k2++;
    }
  }
  return acc;
}

// The same sum by EXPANSION (cmc.cu, expand kernel): for a fixed row i the columns k of row i are visited in
// ascending order and every entry (k, j) of row k adds A_d(i,k) B_d(j,k) v(d,k) to the accumulator of (i, j);
// B_d(j,k) is read from the transposed copy of ct (ctT[d][p] = ct[d][tpos[p]]) at the position p of entry (k, j).
// For a fixed (i, j) the terms arrive in ascending k,
// components innermost: the order of the merge above, hence the same bits. One call = one (k, j) entry.
template <int DIM>
CG_HD double cmc_accumulate(double acc, const double (&Ad)[DIM], const double (&Wd)[DIM], const double* ctT, size_t nnz,
                            int p) {
  for (int d = 0; d < DIM; d++) {
#if defined(__CUDA_ARCH__)
    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(Ad[d], ctT[d * nnz + p]), Wd[d]));
#else
    const volatile double m = Ad[d] * ctT[d * nnz + p];
    const volatile double q = m * Wd[d];
    acc = acc + q;
#endif
  }
  return acc;
}

}  // namespace cgasm
