/* Minimal C client of libcgasm.so: the call sequence a Fortran shim makes, from plain C (no CUDA, no C++, no torch
 * in the signatures). Two triangles on the unit square, tracer assembly with the default CG options, the matrix
 * printed row by row.
 *
 *   gcc -Iinclude examples/c_client.c -Lfluidity_b200 -lcgasm -Wl,-rpath,$PWD/fluidity_b200 -o /tmp/c_client && /tmp/c_client
 *
 * Without a B200 the library refuses at cgasm_create (CGASM_ENODEVICE = 7): there is no CPU path. */
#include <stdio.h>
#include <string.h>

#include "cgasm.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int st_ = (call);                                                            \
    if (st_ != CGASM_OK) {                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #call, st_, cgasm_last_error());         \
      return st_;                                                                \
    }                                                                            \
  } while (0)

int main(void) {
  /* mesh%ndglno (1-based, element-major), Coordinate%val(dim, nodes) */
  const int ndglno[6] = {1, 2, 3, 2, 4, 3};
  const double X[8] = {0, 0, 1, 0, 0, 1, 1, 1};
  /* element_type tables of the P1 triangle at quadrature degree 3 (femtools/Quadrature.F90:951-970):
   * n(loc, ngi), dn(loc, ngi, dim), weight(ngi), column-major */
  const double a = 0.2, b = 0.6, t = 0.333333333333333333333333333333333;
  const double l[4][3] = {{t, t, t}, {a, a, b}, {a, b, a}, {b, a, a}};
  const double weight[4] = {-0.28125, 0.260416666666666666666666666666666, 0.260416666666666666666666666666666,
                            0.260416666666666666666666666666666};
  double n[3 * 4], dn[3 * 4 * 2];
  for (int g = 0; g < 4; g++)
    for (int i = 0; i < 3; i++) {
      n[i + 3 * g] = l[g][i];
      for (int k = 0; k < 2; k++) dn[i + 3 * (g + 4 * k)] = i < 2 ? (i == k ? 1.0 : 0.0) : -1.0;
    }
  int id = 0, nnz = 0;
  CHECK(cgasm_create(&id, -1, 2, 3, 4, 4, 2, ndglno, n, dn, weight));
  CHECK(cgasm_set_coordinates(id, X));
  CHECK(cgasm_build_sparsity(id, &nnz));
  int findrm[5], colm[16], centrm[4];
  CHECK(cgasm_get_sparsity(id, findrm, colm, centrm));
  const double T[4] = {0.0, 1.0, 0.5, 0.25};
  const double nu[8] = {1, 0, 1, 0, 1, 0, 1, 0};
  const double kappa[4] = {1e-3, 0, 0, 1e-3};
  CHECK(cgasm_set_field(id, CGASM_F_T, 0, CGASM_FIELD_NORMAL, T, 4));
  CHECK(cgasm_set_field(id, CGASM_F_NU, 1, CGASM_FIELD_NORMAL, nu, 4));
  CHECK(cgasm_set_field(id, CGASM_F_T_DIFFUSIVITY, 2, CGASM_FIELD_CONSTANT, kappa, 1));
  cgasm_advdiff_opts o;
  memset(&o, 0, sizeof o);
  o.dt = 0.01;
  o.theta = 0.5;
  o.have_mass = o.have_advection = o.have_diffusivity = 1;
  o.diffusivity_shape = CGASM_TENSOR_ISOTROPIC;
  double matrix[16], rhs[4];
  CHECK(cgasm_advdiff(id, &o, matrix, rhs));
  for (int i = 0; i < 4; i++) {
    printf("row %d:", i + 1);
    for (int k = findrm[i] - 1; k < findrm[i + 1] - 1; k++) printf("  (%d) % .6e", colm[k], matrix[k]);
    printf("   | rhs % .6e\n", rhs[i]);
  }
  CHECK(cgasm_destroy(id));
  return 0;
}
