// scatter_kernels.cu -- element-centric assembly kernels: one thread = one element, local
// matrix in registers, then the addto scatter of the reference
//   tracer   csr_vaddto                 femtools/Sparse_Tools.F90:2680-2703
//   big_m    petsc_csr addto (per block) femtools/Sparse_Tools_Petsc.F90:848-879
//   ct_m     block_csr_blocks_addto     femtools/Sparse_Tools.F90:2764-2812
//   rhs      vector/scalar field addto  femtools/Fields_Manipulation.F90:255-379
// in three flavours (north-star: compare them):
//   MODE_ATOMIC   red.global.add.f64 per entry
//   MODE_PLAIN    plain load-add-store; only legal when the launch covers ONE colour of the
//                 femtools/Colouring.F90 colouring (no two elements share a node)
//   MODE_WARPAGG  match.any on the destination, lanes with the same slot reduce through
//                 shuffles and the lowest lane issues one red
// Positions in colm are found by scanning the (sorted, <= ~30 entry) row once per local row
// instead of loc bisections (csr_sparsity_pos, Sparse_Tools.F90:2438-2497).
#include "cgasm_internal.h"

namespace cgasm {

enum { MODE_ATOMIC = 0, MODE_PLAIN = 1, MODE_WARPAGG = 2 };

template <int MODE>
__device__ __forceinline__ void add_to(double* p, double v) {
  if constexpr (MODE == MODE_ATOMIC) {
    atomicAdd(p, v);
  } else if constexpr (MODE == MODE_PLAIN) {
    *p += v;
  } else {
    const unsigned peers = __match_any_sync(__activemask(), (unsigned long long)p);
    double sum = 0.0;
    for (unsigned m = peers; m; m &= m - 1) sum += __shfl_sync(peers, v, __ffs(m) - 1);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(p, sum);
  }
}

template <int LOC>
__device__ __forceinline__ void row_positions(const int* __restrict__ findrm,
                                              const int* __restrict__ colm, int row,
                                              const int (&cols)[LOC], int (&pos)[LOC]) {
  const int s = __ldg(findrm + row), e = __ldg(findrm + row + 1);
#pragma unroll
  for (int j = 0; j < LOC; j++) pos[j] = s;
  for (int k = s; k < e; k++) {
    const int c = __ldg(colm + k);
#pragma unroll
    for (int j = 0; j < LOC; j++)
      if (c == cols[j]) pos[j] = k;
  }
}

template <int DIM, int MODE, bool LABS, int STAB>
__global__ void __launch_bounds__(128)
momentum_scatter_kernel(const MomentumArgs A, const int* __restrict__ elist, int count,
                        const int* __restrict__ findrm, const int* __restrict__ colm, size_t nnz,
                        double* __restrict__ big_m, double* __restrict__ rhs,
                        double* __restrict__ masslump, double* __restrict__ ct_m) {
  constexpr int LOC = DIM + 1;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count) return;
  const int e = elist ? __ldg(elist + tid) : tid;
  const int4 nd = __ldg(A.ndglno + e);
  MomentumLocal<DIM, LABS> R;
  Geom<DIM> G;
  momentum_element<DIM, LABS, STAB>(A, nd, R, G);

  int cols[LOC];
#pragma unroll
  for (int j = 0; j < LOC; j++) cols[j] = node_of(nd, j);
#pragma unroll
  for (int i = 0; i < LOC; i++) {
    int pos[LOC];
    row_positions<LOC>(findrm, colm, cols[i], cols, pos);
#pragma unroll
    for (int j = 0; j < LOC; j++) {
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        double v = R.L[i][j];
        if constexpr (LABS) v += R.Labs[d][i][j];
        if (i == j) v += R.diag[d][i];
        add_to<MODE>(big_m + (size_t)d * nnz + (size_t)pos[j], v);
        if (ct_m) add_to<MODE>(ct_m + (size_t)d * nnz + (size_t)pos[j], grad_p_u<DIM>(A.tab, G, d, i, j, A.o.integrate_continuity_by_parts != 0));
      }
    }
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      add_to<MODE>(rhs + (size_t)DIM * cols[i] + d, R.rhs[d][i]);
      if (masslump) add_to<MODE>(masslump + (size_t)DIM * cols[i] + d, R.ml[d][i]);
    }
  }
}

template <int DIM, int MODE, int STAB>
__global__ void __launch_bounds__(128)
advdiff_scatter_kernel(const AdvDiffArgs P, const int* __restrict__ elist, int count,
                       const int* __restrict__ findrm, const int* __restrict__ colm,
                       double* __restrict__ matrix, double* __restrict__ rhs) {
  constexpr int LOC = DIM + 1;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count) return;
  const int e = elist ? __ldg(elist + tid) : tid;
  const int4 nd = __ldg(P.ndglno + e);
  AdvDiffLocal<DIM> R;
  advdiff_element<DIM, STAB>(P, nd, R);
  int cols[LOC];
#pragma unroll
  for (int j = 0; j < LOC; j++) cols[j] = node_of(nd, j);
#pragma unroll
  for (int i = 0; i < LOC; i++) {
    int pos[LOC];
    row_positions<LOC>(findrm, colm, cols[i], cols, pos);
#pragma unroll
    for (int j = 0; j < LOC; j++) add_to<MODE>(matrix + pos[j], R.A[i][j]);
    add_to<MODE>(rhs + cols[i], R.rhs[i]);
  }
}

// Single element -> dense local arrays (element-matrix parity checks).
template <int DIM, int STAB>
__global__ void momentum_one_kernel(const MomentumArgs A, int e, double* __restrict__ T,
                                    double* __restrict__ rhs, double* __restrict__ ml,
                                    double* __restrict__ gp) {
  constexpr int LOC = DIM + 1;
  if (threadIdx.x || blockIdx.x) return;
  const int4 nd = A.ndglno[e];
  MomentumLocal<DIM, true> R;
  Geom<DIM> G;
#pragma unroll
  for (int d = 0; d < DIM; d++)
#pragma unroll
    for (int i = 0; i < LOC; i++)
#pragma unroll
      for (int j = 0; j < LOC; j++) R.Labs[d][i][j] = 0.0;
  momentum_element<DIM, true, STAB>(A, nd, R, G);
  for (int a = 0; a < DIM * DIM * LOC * LOC; a++) T[a] = 0.0;
  for (int d = 0; d < DIM; d++)
    for (int i = 0; i < LOC; i++) {
      for (int j = 0; j < LOC; j++) {
        double v = R.L[i][j] + R.Labs[d][i][j];
        if (i == j) v += R.diag[d][i];
        T[d + DIM * (d + DIM * (i + LOC * j))] = v;
        gp[d + DIM * (i + LOC * j)] = grad_p_u<DIM>(A.tab, G, d, i, j, A.o.integrate_continuity_by_parts != 0);
      }
      rhs[d + DIM * i] = R.rhs[d][i];
      ml[d + DIM * i] = R.ml[d][i];
    }
}

template <int DIM, int STAB>
__global__ void advdiff_one_kernel(const AdvDiffArgs P, int e, double* __restrict__ Aout,
                                   double* __restrict__ rhs) {
  constexpr int LOC = DIM + 1;
  if (threadIdx.x || blockIdx.x) return;
  AdvDiffLocal<DIM> R;
  advdiff_element<DIM, STAB>(P, P.ndglno[e], R);
  for (int i = 0; i < LOC; i++) {
    for (int j = 0; j < LOC; j++) Aout[i + LOC * j] = R.A[i][j];
    rhs[i] = R.rhs[i];
  }
}

// ---- launchers -----------------------------------------------------------------------------
template <int DIM, int MODE>
static void launch_momentum_mode(Handle* h, const MomentumArgs& A, const int* elist, int count,
                                 double* ml, double* ct) {
  if (count <= 0) return;
  const int block = 128, grid = (count + block - 1) / block;
  const bool labs = A.o.have_absorption && !A.o.lump_absorption;
  const int stab = A.o.stabilisation_scheme;
#define LAUNCH(LABS_, STAB_)                                                                       \
  momentum_scatter_kernel<DIM, MODE, LABS_, STAB_><<<grid, block, 0, h->stream>>>(                 \
      A, elist, count, h->d_findrm, h->d_colm, (size_t)h->nnz, h->d_big_m, h->d_mom_rhs, ml, ct)
  if constexpr (MODE == MODE_ATOMIC) {
    if (stab == CGASM_STAB_STREAMLINE_UPWIND) {
      if (labs) LAUNCH(true, 1);
      else LAUNCH(false, 1);
      h->launches++;
      return;
    }
    if (stab == CGASM_STAB_SUPG) {
      if (labs) LAUNCH(true, 2);
      else LAUNCH(false, 2);
      h->launches++;
      return;
    }
  }
  if (labs) LAUNCH(true, 0);
  else LAUNCH(false, 0);
#undef LAUNCH
  h->launches++;
}

template <int DIM, int MODE>
static void launch_advdiff_mode(Handle* h, const AdvDiffArgs& P, const int* elist, int count) {
  if (count <= 0) return;
  const int block = 128, grid = (count + block - 1) / block;
  const int stab = P.o.stabilisation_scheme;
#define LAUNCH(STAB_)                                                              \
  advdiff_scatter_kernel<DIM, MODE, STAB_><<<grid, block, 0, h->stream>>>(         \
      P, elist, count, h->d_findrm, h->d_colm, h->d_adv_matrix, h->d_adv_rhs)
  if constexpr (MODE == MODE_ATOMIC) {
    if (stab == CGASM_STAB_STREAMLINE_UPWIND) {
      LAUNCH(1);
      h->launches++;
      return;
    }
    if (stab == CGASM_STAB_SUPG) {
      LAUNCH(2);
      h->launches++;
      return;
    }
  }
  LAUNCH(0);
#undef LAUNCH
  h->launches++;
}

template <int DIM>
static int scatter_momentum_dim(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  double* ml = want_ml ? h->d_masslump : nullptr;
  double* ct = want_ct ? h->d_ct_m : nullptr;
  switch (h->scatter) {
    case CGASM_SCATTER_ATOMIC:
      launch_momentum_mode<DIM, MODE_ATOMIC>(h, A, nullptr, h->n_elements, ml, ct);
      break;
    case CGASM_SCATTER_WARPAGG:
      launch_momentum_mode<DIM, MODE_WARPAGG>(h, A, nullptr, h->n_elements, ml, ct);
      break;
    case CGASM_SCATTER_COLOURED:
      // colour loop of Momentum_CG.F90:726-752: one launch per colour, stream order is the
      // barrier the OpenMP version gets from the end of each !$OMP DO
      for (int c = 0; c < h->ncolours; c++)
        launch_momentum_mode<DIM, MODE_PLAIN>(h, A, h->d_colour_elements + h->h_colour_ptr[c],
                                              h->h_colour_ptr[c + 1] - h->h_colour_ptr[c], ml, ct);
      break;
    default:
      CG_FAIL(CGASM_EARG, "unknown scatter variant");
  }
  return CGASM_OK;
}

template <int DIM>
static int scatter_advdiff_dim(Handle* h, const AdvDiffArgs& P) {
  switch (h->scatter) {
    case CGASM_SCATTER_ATOMIC:
      launch_advdiff_mode<DIM, MODE_ATOMIC>(h, P, nullptr, h->n_elements);
      break;
    case CGASM_SCATTER_WARPAGG:
      launch_advdiff_mode<DIM, MODE_WARPAGG>(h, P, nullptr, h->n_elements);
      break;
    case CGASM_SCATTER_COLOURED:
      for (int c = 0; c < h->ncolours; c++)
        launch_advdiff_mode<DIM, MODE_PLAIN>(h, P, h->d_colour_elements + h->h_colour_ptr[c],
                                             h->h_colour_ptr[c + 1] - h->h_colour_ptr[c]);
      break;
    default:
      CG_FAIL(CGASM_EARG, "unknown scatter variant");
  }
  return CGASM_OK;
}

int scatter_momentum(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  return h->dim == 3 ? scatter_momentum_dim<3>(h, A, want_ml, want_ct)
                     : scatter_momentum_dim<2>(h, A, want_ml, want_ct);
}
int scatter_advdiff(Handle* h, const AdvDiffArgs& P) {
  return h->dim == 3 ? scatter_advdiff_dim<3>(h, P) : scatter_advdiff_dim<2>(h, P);
}

void one_momentum(Handle* h, const MomentumArgs& A, int e, double* T, double* rhs, double* ml,
                  double* gp) {
  const int stab = A.o.stabilisation_scheme;
#define ONE(DIM_)                                                                        \
  do {                                                                                   \
    if (stab == 1) momentum_one_kernel<DIM_, 1><<<1, 32, 0, h->stream>>>(A, e, T, rhs, ml, gp);      \
    else if (stab == 2) momentum_one_kernel<DIM_, 2><<<1, 32, 0, h->stream>>>(A, e, T, rhs, ml, gp); \
    else momentum_one_kernel<DIM_, 0><<<1, 32, 0, h->stream>>>(A, e, T, rhs, ml, gp);                \
  } while (0)
  if (h->dim == 3) ONE(3);
  else ONE(2);
#undef ONE
  h->launches++;
}
void one_advdiff(Handle* h, const AdvDiffArgs& P, int e, double* Aout, double* rhs) {
  const int stab = P.o.stabilisation_scheme;
#define ONE(DIM_)                                                                     \
  do {                                                                                \
    if (stab == 1) advdiff_one_kernel<DIM_, 1><<<1, 32, 0, h->stream>>>(P, e, Aout, rhs);      \
    else if (stab == 2) advdiff_one_kernel<DIM_, 2><<<1, 32, 0, h->stream>>>(P, e, Aout, rhs); \
    else advdiff_one_kernel<DIM_, 0><<<1, 32, 0, h->stream>>>(P, e, Aout, rhs);                \
  } while (0)
  if (h->dim == 3) ONE(3);
  else ONE(2);
#undef ONE
  h->launches++;
}

}  // namespace cgasm
