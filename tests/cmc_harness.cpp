// CPU harness around fluidity_b200/csrc/cmc_math.h (the entry function the device kernel runs). Test code only.
#include <cstddef>
#include "../fluidity_b200/csrc/cmc_math.h"
extern "C" void harness_cmc(int dim, int n, const int* findrm, const int* colm, const double* ct, long long nnz, const double* v,
                            const int* findrm2, const int* colm2, double* out) {
  for (int i = 0; i < n; i++)
    for (int e = findrm2[i]; e < findrm2[i + 1]; e++)
      out[e] = dim == 3 ? cgasm::cmc_entry<3>(findrm, colm, ct, ct, (size_t)nnz, v, i, colm2[e])
                        : cgasm::cmc_entry<2>(findrm, colm, ct, ct, (size_t)nnz, v, i, colm2[e]);
}

// Emulation of cmc_expand_kernel (fluidity_b200/csrc/cmc.cu) with the library's own plan: half-warps of 16 lanes,
// the k steps in order, the lanes of one step in ANY order (they touch distinct slots) -- here descending, to
// show that the result does not depend on it. Returns the number of slot collisions inside a step (must be 0).
extern "C" long long harness_cmc_expand(int dim, int n, const int* findrm, const int* colm, const double* ct, long long nnz,
                                        const double* v, const int* tpos, const long long* pptr, const unsigned short* slots,
                                        const int* findrm2, int n2max, double* out) {
  long long collisions = 0;
  // transpose_ct_kernel: ctT[d][p] = ct[d][tpos[p]]
  double* ctT = new double[(size_t)dim * nnz];
  for (int d = 0; d < dim; d++)
    for (long long p = 0; p < nnz; p++) ctT[d * nnz + p] = ct[d * nnz + tpos[p]];
  double* acc = new double[n2max];
  int* stamp = new int[n2max];
  for (int i = 0; i < n; i++) {
    const int r0 = findrm[i], n1 = findrm[i + 1] - r0, o0 = findrm2[i], n2 = findrm2[i + 1] - o0;
    for (int s = 0; s < n2; s++) acc[s] = 0.0;
    long long pp = pptr[i];
    for (int a = 0; a < n1; a++) {
      const int k = colm[r0 + a];
      const int kb = findrm[k], kn = findrm[k + 1] - kb;
      for (int s = 0; s < n2; s++) stamp[s] = 0;
      for (int hl = 15; hl >= 0; hl--)
        for (int q = hl; q < kn; q += 16) {
          const int s = slots[pp + q];
          if (s >= n2 || stamp[s]++) collisions++;
          if (dim == 3) {
            double Ad[3], Wd[3];
            for (int d = 0; d < 3; d++) { Ad[d] = ct[d * nnz + r0 + a]; Wd[d] = v[(size_t)3 * k + d]; }
            acc[s] = cgasm::cmc_accumulate<3>(acc[s], Ad, Wd, ctT, (size_t)nnz, kb + q);
          } else {
            double Ad[2], Wd[2];
            for (int d = 0; d < 2; d++) { Ad[d] = ct[d * nnz + r0 + a]; Wd[d] = v[(size_t)2 * k + d]; }
            acc[s] = cgasm::cmc_accumulate<2>(acc[s], Ad, Wd, ctT, (size_t)nnz, kb + q);
          }
        }
      pp += kn;
    }
    if (pp != pptr[i + 1]) collisions += 1000000;
    for (int s = 0; s < n2; s++) out[o0 + s] = acc[s];
  }
  delete[] acc;
  delete[] ctT;
  delete[] stamp;
  return collisions;
}
