// CPU harness around fluidity_b200/csrc/surface_math.h (the per-face arithmetic the device kernels run):
// compiled with g++ by tests/test_surface_math.py and compared with the oracle. Test code only.
#include "../fluidity_b200/csrc/surface_math.h"

using namespace cgasm;

template <int DIM>
static void fill(SurfTables& t, int sngi, const double* n, const double* dn, const double* w) {
  t.sloc = DIM;
  t.sngi = sngi;
  for (int k = 0; k < DIM * sngi; k++) t.n[k] = n[k];
  for (int k = 0; k < DIM * sngi * (DIM - 1); k++) t.dn[k] = dn[k];
  for (int g = 0; g < sngi; g++) t.w[g] = w[g];
}

template <int DIM>
static void adv(int sngi, const double* n, const double* dn, const double* w, const cgasm_advdiff_opts* o, int bc_type,
                const double* Xf_, const double* Xc_, const double* Tf_, const double* Uf_, const double* bc_,
                const double* bc2_, double* A_, double* r_) {
  SurfTables t;
  fill<DIM>(t, sngi, n, dn, w);
  double Xf[DIM][DIM], Xc[DIM], Tf[DIM], Uf[DIM][DIM], bc[DIM], bc2[DIM], A[DIM][DIM], r[DIM];
  for (int i = 0; i < DIM; i++) {
    Xc[i] = Xc_[i];
    Tf[i] = Tf_[i];
    bc[i] = bc_[i];
    bc2[i] = bc2_[i];
    for (int a = 0; a < DIM; a++) {
      Xf[i][a] = Xf_[i * DIM + a];
      Uf[i][a] = Uf_[i * DIM + a];
    }
  }
  advdiff_face<DIM>(t, *o, bc_type, Xf, Xc, Tf, Uf, bc, bc2, A, r);
  for (int i = 0; i < DIM; i++) {
    r_[i] = r[i];
    for (int j = 0; j < DIM; j++) A_[i * DIM + j] = A[i][j];
  }
}

template <int DIM>
static int mom(int sngi, const double* n, const double* dn, const double* w, const cgasm_momentum_opts* o, const int* bt_,
               int ptype, const double* Xf_, const double* Xc_, const double* Uf_, const double* Of_, const double* rho_,
               const double* bc_, double* B_, double* r_, const double* Gf_ = nullptr, double* ml_ = nullptr) {
  SurfTables t;
  fill<DIM>(t, sngi, n, dn, w);
  int bt[DIM];
  double Xf[DIM][DIM], Xc[DIM], Uf[DIM][DIM], Of[DIM][DIM], rho[DIM], bc[DIM][DIM], B[DIM][DIM][DIM], r[DIM][DIM];
  for (int i = 0; i < DIM; i++) {
    bt[i] = bt_[i];
    Xc[i] = Xc_[i];
    rho[i] = rho_[i];
    for (int a = 0; a < DIM; a++) {
      Xf[i][a] = Xf_[i * DIM + a];
      Uf[i][a] = Uf_[i * DIM + a];
      Of[i][a] = Of_[i * DIM + a];
      bc[i][a] = bc_[i * DIM + a];
    }
  }
  if (momentum_face_skipped<DIM>(bt, ptype)) return 1;
  double Gf[DIM][DIM], ml[DIM][DIM];
  for (int i = 0; i < DIM; i++)
    for (int a = 0; a < DIM; a++) Gf[i][a] = Gf_ ? Gf_[i * DIM + a] : 0.0;
  momentum_face<DIM>(t, *o, bt, Xf, Xc, Uf, Of, rho, bc, Gf, B, r, ml);
  for (int d = 0; d < DIM; d++)
    for (int i = 0; i < DIM; i++) {
      if (ml_) ml_[d * DIM + i] = ml[d][i];
      r_[d * DIM + i] = r[d][i];
      for (int j = 0; j < DIM; j++) B_[(d * DIM + i) * DIM + j] = B[d][i][j];
    }
  return 0;
}

extern "C" {
void harness_advdiff_face(int dim, int sngi, const double* n, const double* dn, const double* w, const cgasm_advdiff_opts* o,
                          int bc_type, const double* Xf, const double* Xc, const double* Tf, const double* Uf,
                          const double* bc, const double* bc2, double* A, double* r) {
  if (dim == 3) adv<3>(sngi, n, dn, w, o, bc_type, Xf, Xc, Tf, Uf, bc, bc2, A, r);
  else adv<2>(sngi, n, dn, w, o, bc_type, Xf, Xc, Tf, Uf, bc, bc2, A, r);
}
int harness_momentum_face(int dim, int sngi, const double* n, const double* dn, const double* w, const cgasm_momentum_opts* o,
                          const int* bt, int ptype, const double* Xf, const double* Xc, const double* Uf, const double* Of,
                          const double* rho, const double* bc, double* B, double* r) {
  return dim == 3 ? mom<3>(sngi, n, dn, w, o, bt, ptype, Xf, Xc, Uf, Of, rho, bc, B, r)
                  : mom<2>(sngi, n, dn, w, o, bt, ptype, Xf, Xc, Uf, Of, rho, bc, B, r);
}
// with the gravity direction at the face nodes (free-surface stabilisation) and the masslump contribution
int harness_momentum_face_fs(int dim, int sngi, const double* n, const double* dn, const double* w, const cgasm_momentum_opts* o,
                             const int* bt, int ptype, const double* Xf, const double* Xc, const double* Uf, const double* Of,
                             const double* rho, const double* bc, const double* Gf, double* B, double* r, double* ml) {
  return dim == 3 ? mom<3>(sngi, n, dn, w, o, bt, ptype, Xf, Xc, Uf, Of, rho, bc, B, r, Gf, ml)
                  : mom<2>(sngi, n, dn, w, o, bt, ptype, Xf, Xc, Uf, Of, rho, bc, B, r, Gf, ml);
}
int harness_csr_pos0(const int* findrm, const int* colm, int i, int j) { return csr_pos0(findrm, colm, i, j); }
}
