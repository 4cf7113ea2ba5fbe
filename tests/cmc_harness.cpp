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
