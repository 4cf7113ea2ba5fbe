// surface.cu -- surface-element loops and strong Dirichlet conditions on the device:
// the step right after the two element loops, so that the whole result of
// assemble_advection_diffusion_cg / construct_momentum_cg can stay device-resident.
//
// Work is O(boundary faces) = O(N^(2/3)): one thread per face, the per-face arithmetic of surface_math.h,
// entries added to the CSR values of the preceding element loop with FP64 atomics at positions found by
// bisection of the (sorted) row -- the reference's csr_sparsity_pos. A boundary node is shared by ~dim faces,
// so contention is negligible; summation order differs from the serial reference within rounding.
#include <cstring>

#include "cgasm_internal.h"
#include "surface_math.h"

namespace cgasm {

struct SurfacePlan {
  int n_faces = 0;
  SurfTables tab{};
  int* d_sndgln = nullptr;    // (sloc, n_faces) 0-based
  int* d_face_ele = nullptr;  // 0-based
  // per-call boundary-condition inputs (grown on demand)
  int* d_itype = nullptr;
  size_t itype_cap = 0;
  int* d_ptype = nullptr;
  size_t ptype_cap = 0;
  double* d_bc = nullptr;
  size_t bc_cap = 0;
  double* d_bc2 = nullptr;
  size_t bc2_cap = 0;
  int* d_nodes = nullptr;  // Dirichlet node list
  size_t nodes_cap = 0;
  double* d_vals = nullptr;
  size_t vals_cap = 0;
};

void surface_free(Handle* h) {
  SurfacePlan* S = h->surface;
  if (!S) return;
  cudaFree(S->d_sndgln);
  cudaFree(S->d_face_ele);
  cudaFree(S->d_itype);
  cudaFree(S->d_ptype);
  cudaFree(S->d_bc);
  cudaFree(S->d_bc2);
  cudaFree(S->d_nodes);
  cudaFree(S->d_vals);
  delete S;
  h->surface = nullptr;
}

template <class T>
static int grow(T** p, size_t* cap, size_t count) {
  if (count <= *cap && *p) return CGASM_OK;
  if (*p) CG_CUDA(cudaFree(*p));
  *p = nullptr;
  CG_CUDA(cudaMalloc(p, sizeof(T) * std::max<size_t>(count, 1)));
  *cap = count;
  return CGASM_OK;
}

// Upload on the handle's stream and wait: the host arrays are the caller's and may change after the call
// returns; these inputs are O(boundary) small.
template <class T>
static int upload(Handle* h, T* dst, const T* src, size_t count) {
  if (!count) return CGASM_OK;
  CG_CUDA(cudaMemcpyAsync(dst, src, sizeof(T) * count, cudaMemcpyHostToDevice, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

struct RawField {  // the caller's array on the device: val(ncomp, nodes), stride 0 = CONSTANT field (node 1)
  const double* val;
  int stride;
};
static RawField raw_of(const Handle* h, int slot, int comps) {
  const DeviceField& f = h->fields[slot];
  return RawField{f.d, f.field_type == CGASM_FIELD_CONSTANT ? 0 : comps};
}

struct FaceMesh {
  int n_faces;
  const int* sndgln;
  const int* face_ele;
  const int4* ndglno;
  const double* X;
  const int* findrm;
  const int* colm;
};

// positions of the face nodes and the centroid of the owning element
template <int DIM>
__device__ __forceinline__ void load_face(const FaceMesh& m, int f, int (&nodes)[DIM], double (&Xf)[DIM][DIM],
                                          double (&Xc)[DIM]) {
  for (int i = 0; i < DIM; i++) {
    nodes[i] = m.sndgln[(size_t)DIM * f + i];
    for (int a = 0; a < DIM; a++) Xf[i][a] = m.X[(size_t)DIM * nodes[i] + a];
  }
  const int4 e = m.ndglno[m.face_ele[f]];
  const int en[4] = {e.x, e.y, e.z, e.w};
  for (int a = 0; a < DIM; a++) {
    double s = 0.0;
    for (int i = 0; i < DIM + 1; i++) s += m.X[(size_t)DIM * en[i] + a];
    Xc[a] = s / (DIM + 1);
  }
}

template <int DIM>
__global__ void advdiff_surface_kernel(const SurfTables t, const FaceMesh m, const cgasm_advdiff_opts o, const RawField T,
                                       const RawField U, const int* __restrict__ bc_type, const double* __restrict__ t_bc,
                                       const double* __restrict__ t_bc_2, double* __restrict__ matrix,
                                       double* __restrict__ rhs) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= m.n_faces) return;
  const int type = bc_type[f];
  if (type == CGASM_TBC_INTERNAL) return;  // Advection_Diffusion_CG.F90:629
  int nodes[DIM];
  double Xf[DIM][DIM], Xc[DIM], Tf[DIM], Uf[DIM][DIM], bc[DIM], bc2[DIM], A[DIM][DIM], r[DIM];
  load_face<DIM>(m, f, nodes, Xf, Xc);
  for (int i = 0; i < DIM; i++) {
    Tf[i] = T.val[(size_t)T.stride * nodes[i]];
    for (int a = 0; a < DIM; a++) Uf[i][a] = U.val ? U.val[(size_t)U.stride * nodes[i] + a] : 0.0;
    bc[i] = t_bc ? t_bc[(size_t)DIM * f + i] : 0.0;
    bc2[i] = t_bc_2 ? t_bc_2[(size_t)DIM * f + i] : 0.0;
  }
  advdiff_face<DIM>(t, o, type, Xf, Xc, Tf, Uf, bc, bc2, A, r);
  for (int i = 0; i < DIM; i++) {
    for (int j = 0; j < DIM; j++) {
      if (A[i][j] == 0.0) continue;  // csr_addto skips exact zeros (Sparse_Tools.F90:2640)
      const int pos = csr_pos0(m.findrm, m.colm, nodes[i], nodes[j]);
      if (pos >= 0) atomicAdd(matrix + pos, A[i][j]);
    }
    if (r[i] != 0.0) atomicAdd(rhs + nodes[i], r[i]);
  }
}

template <int DIM>
__global__ void momentum_surface_kernel(const SurfTables t, const FaceMesh m, const cgasm_momentum_opts o, const RawField U,
                                        const RawField O, const RawField R, const RawField G, const int* __restrict__ vtype,
                                        const double* __restrict__ vbc, const int* __restrict__ ptype, size_t nnz,
                                        double* __restrict__ big_m, double* __restrict__ rhs, double* __restrict__ ct_m,
                                        double* __restrict__ masslump) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= m.n_faces) return;
  int bt[DIM];
  for (int d = 0; d < DIM; d++) bt[d] = vtype[(size_t)DIM * f + d];
  if (momentum_face_skipped<DIM>(bt, ptype ? ptype[f] : 0)) return;
  int nodes[DIM];
  double Xf[DIM][DIM], Xc[DIM], Uf[DIM][DIM], Of[DIM][DIM], rho[DIM], bc[DIM][DIM], Gf[DIM][DIM], B[DIM][DIM][DIM], r[DIM][DIM],
      ml[DIM][DIM];
  load_face<DIM>(m, f, nodes, Xf, Xc);
  for (int i = 0; i < DIM; i++) {
    rho[i] = R.val[(size_t)R.stride * nodes[i]];
    for (int a = 0; a < DIM; a++) {
      Uf[i][a] = U.val[(size_t)U.stride * nodes[i] + a];
      Of[i][a] = O.val[(size_t)O.stride * nodes[i] + a];
      Gf[i][a] = G.val ? G.val[(size_t)G.stride * nodes[i] + a] : 0.0;
      bc[i][a] = vbc ? vbc[((size_t)f * DIM + i) * DIM + a] : 0.0;  // velocity_bc(dim, sloc, n_faces)
    }
  }
  momentum_face<DIM>(t, o, bt, Xf, Xc, Uf, Of, rho, bc, Gf, B, r, ml);
  for (int i = 0; i < DIM; i++) {
    for (int j = 0; j < DIM; j++) {
      const int pos = csr_pos0(m.findrm, m.colm, nodes[i], nodes[j]);
      if (pos < 0) continue;
      for (int d = 0; d < DIM; d++)
        if (B[d][i][j] != 0.0) atomicAdd(big_m + (size_t)d * nnz + pos, B[d][i][j]);
    }
    for (int d = 0; d < DIM; d++) {
      if (r[d][i] != 0.0) atomicAdd(rhs + (size_t)DIM * nodes[i] + d, r[d][i]);
      if (masslump && ml[d][i] != 0.0) atomicAdd(masslump + (size_t)DIM * nodes[i] + d, ml[d][i]);
    }
  }
  // continuity by parts: the boundary blocks of ct_m (:1073-1088; weak-Dirichlet ct_rhs and pressure conditions are
  // outside the device path: cgasm_momentum_surface_dev refuses them)
  if (ct_m && bt[0] != CGASM_VBC_NO_NORMAL_FLOW && bt[0] != CGASM_VBC_FREE_SURFACE) {
    double CB[DIM][DIM][DIM];
    momentum_face_ct<DIM>(t, Xf, Xc, CB);
    for (int i = 0; i < DIM; i++)
      for (int j = 0; j < DIM; j++) {
        const int pos = csr_pos0(m.findrm, m.colm, nodes[i], nodes[j]);
        if (pos < 0) continue;
        for (int d = 0; d < DIM; d++) atomicAdd(ct_m + (size_t)d * nnz + pos, CB[d][i][j]);
      }
  }
}

// apply_dirichlet_conditions_scalar (femtools/Boundary_Conditions.F90:2008-2021)
__global__ void dirichlet_scalar_kernel(int n, const int* __restrict__ nodes, const double* __restrict__ values,
                                        const RawField T, int have_dt, double dt, double* __restrict__ rhs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int node = nodes[j];
  rhs[node] = have_dt ? (values[j] - T.val[(size_t)T.stride * node]) / dt : values[j];
}

static FaceMesh face_mesh(const Handle* h) {
  const SurfacePlan* S = h->surface;
  return FaceMesh{S->n_faces, S->d_sndgln, S->d_face_ele, h->d_ndglno, h->d_X, h->d_findrm, h->d_colm};
}

}  // namespace cgasm

using namespace cgasm;

#define GET_HANDLE(h, id)                                         \
  Handle* h = get_handle(id);                                     \
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");         \
  CG_CUDA(cudaSetDevice(h->device))

extern "C" {

int cgasm_set_surface(int id, int n_faces, int sloc, int sngi, const int* sndgln, const int* face_ele,
                      const double* n_f, const double* dn_f, const double* weight_f) {
  GET_HANDLE(h, id);
  if (n_faces < 0 || (n_faces > 0 && (!sndgln || !face_ele)) || !n_f || !dn_f || !weight_f) CG_FAIL(CGASM_EARG, "null argument");
  const int dim = h->dim;
  if (sloc != dim) CG_FAIL(CGASM_EUNSUPPORTED, "only P1 simplex faces (sloc = dim) are on the device path");
  if (sngi < 1 || sngi > kMaxSngi) CG_FAIL(CGASM_EUNSUPPORTED, "face quadrature with more than 4 points");
  for (int k = 0; k < dim - 1; k++)
    for (int g = 0; g < sngi; g++)
      for (int i = 0; i < sloc; i++) {
        const double want = (i < dim - 1) ? (i == k ? 1.0 : 0.0) : -1.0;
        if (std::fabs(dn_f[i + sloc * (g + sngi * k)] - want) > 1e-14)
          CG_FAIL(CGASM_EUNSUPPORTED, "dn_f is not the P1 Lagrange simplex derivative table");
      }
  std::vector<int> sn((size_t)sloc * n_faces), fe((size_t)n_faces);
  for (int f = 0; f < n_faces; f++) {
    if (face_ele[f] < 1 || face_ele[f] > h->n_elements) CG_FAIL(CGASM_EARG, "face_ele out of range");
    fe[f] = face_ele[f] - 1;
    const int* en = &h->h_nd0[(size_t)4 * fe[f]];
    for (int i = 0; i < sloc; i++) {
      const int node = sndgln[(size_t)sloc * f + i];
      if (node < 1 || node > h->n_nodes) CG_FAIL(CGASM_EARG, "sndgln out of range");
      bool found = false;
      for (int q = 0; q < h->loc; q++) found = found || en[q] == node - 1;
      if (!found) CG_FAIL(CGASM_EARG, "a face node is not a node of the face's element");
      sn[(size_t)sloc * f + i] = node - 1;
    }
  }
  surface_free(h);
  SurfacePlan* S = new SurfacePlan();
  h->surface = S;
  S->n_faces = n_faces;
  S->tab.sloc = sloc;
  S->tab.sngi = sngi;
  std::memset(S->tab.n, 0, sizeof S->tab.n);
  std::memset(S->tab.dn, 0, sizeof S->tab.dn);
  std::memset(S->tab.w, 0, sizeof S->tab.w);
  for (int k = 0; k < sloc * sngi; k++) S->tab.n[k] = n_f[k];
  for (int k = 0; k < sloc * sngi * (dim - 1); k++) S->tab.dn[k] = dn_f[k];
  for (int g = 0; g < sngi; g++) S->tab.w[g] = weight_f[g];
  CG_CUDA(cudaMalloc(&S->d_sndgln, sizeof(int) * std::max<size_t>(sn.size(), 1)));
  CG_CUDA(cudaMalloc(&S->d_face_ele, sizeof(int) * std::max<size_t>(fe.size(), 1)));
  int st;
  if ((st = upload(h, S->d_sndgln, sn.data(), sn.size()))) return st;
  if ((st = upload(h, S->d_face_ele, fe.data(), fe.size()))) return st;
  return CGASM_OK;
}

int cgasm_advdiff_surface_dev(int id, const cgasm_advdiff_opts* opts, const int* bc_type, const double* t_bc,
                              const double* t_bc_2) {
  GET_HANDLE(h, id);
  if (!opts) CG_FAIL(CGASM_EARG, "null opts");
  SurfacePlan* S = h->surface;
  if (!S || !S->d_sndgln) CG_FAIL(CGASM_ESTATE, "cgasm_set_surface has not been called");
  if (!h->adv_valid) CG_FAIL(CGASM_ESTATE, "no tracer result to add to: call cgasm_advdiff_dev first");
  if (opts->move_mesh || opts->multiphase || opts->equation_type_not_advdiff)
    CG_FAIL(CGASM_EUNSUPPORTED, "tracer option outside the device path; keep the Fortran loop");
  const bool by_parts = opts->have_advection && opts->integrate_advection_by_parts;
  if (!(by_parts || opts->have_diffusivity) || S->n_faces == 0) return CGASM_OK;  // :611-614
  if (!bc_type) CG_FAIL(CGASM_EARG, "null bc_type");
  bool need_bc = false, need_bc2 = false;
  for (int f = 0; f < S->n_faces; f++) {
    const int t = bc_type[f];
    if (t < CGASM_TBC_NONE || t > CGASM_TBC_ROBIN) CG_FAIL(CGASM_EARG, "bad tracer boundary-condition type");
    if (t == CGASM_TBC_WEAKDIRICHLET && opts->have_diffusivity)
      CG_FAIL(CGASM_EUNSUPPORTED, "weak Dirichlet boundary conditions with diffusivity are not supported by CG advection-diffusion");
    need_bc = need_bc || (t == CGASM_TBC_WEAKDIRICHLET && by_parts) ||
              (opts->have_diffusivity && (t == CGASM_TBC_NEUMANN || t == CGASM_TBC_ROBIN));
    need_bc2 = need_bc2 || (opts->have_diffusivity && t == CGASM_TBC_ROBIN);
  }
  if ((need_bc && !t_bc) || (need_bc2 && !t_bc_2)) CG_FAIL(CGASM_EARG, "a face needs boundary values that were not given");
  if (!h->fields[CGASM_F_T].set) CG_FAIL(CGASM_ESTATE, "field slot not set: T");
  if (by_parts && !h->fields[CGASM_F_NU].set) CG_FAIL(CGASM_ESTATE, "field slot not set: NU");
  const size_t nf = (size_t)S->n_faces, nv = nf * (size_t)h->dim;
  int st;
  if ((st = grow(&S->d_itype, &S->itype_cap, nf))) return st;
  if ((st = upload(h, S->d_itype, bc_type, nf))) return st;
  if (need_bc) {
    if ((st = grow(&S->d_bc, &S->bc_cap, nv))) return st;
    if ((st = upload(h, S->d_bc, t_bc, nv))) return st;
  }
  if (need_bc2) {
    if ((st = grow(&S->d_bc2, &S->bc2_cap, nv))) return st;
    if ((st = upload(h, S->d_bc2, t_bc_2, nv))) return st;
  }
  if (h->adv_copy_pending) {  // an asynchronous fetch may still be reading the result we add to
    CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_adv_copied, 0));
    h->adv_copy_pending = false;
  }
  const RawField T = raw_of(h, CGASM_F_T, 1);
  RawField U = raw_of(h, CGASM_F_NU, h->dim);
  if (!by_parts) U.val = nullptr;
  const int threads = 128, blocks = (S->n_faces + threads - 1) / threads;
  const FaceMesh m = face_mesh(h);
  if (h->dim == 3)
    advdiff_surface_kernel<3><<<blocks, threads, 0, h->stream>>>(S->tab, m, *opts, T, U, S->d_itype, need_bc ? S->d_bc : nullptr,
                                                                 need_bc2 ? S->d_bc2 : nullptr, h->d_adv_matrix, h->d_adv_rhs);
  else
    advdiff_surface_kernel<2><<<blocks, threads, 0, h->stream>>>(S->tab, m, *opts, T, U, S->d_itype, need_bc ? S->d_bc : nullptr,
                                                                 need_bc2 ? S->d_bc2 : nullptr, h->d_adv_matrix, h->d_adv_rhs);
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int cgasm_advdiff_dirichlet_dev(int id, int n, const int* nodes, const double* values, int have_dt, double dt) {
  GET_HANDLE(h, id);
  if (n < 0 || (n > 0 && (!nodes || !values))) CG_FAIL(CGASM_EARG, "null argument");
  if (have_dt && dt == 0.0) CG_FAIL(CGASM_EARG, "dt = 0");
  if (!h->adv_valid) CG_FAIL(CGASM_ESTATE, "no tracer result: call cgasm_advdiff_dev first");
  if (!h->fields[CGASM_F_T].set) CG_FAIL(CGASM_ESTATE, "field slot not set: T");
  if (n == 0) return CGASM_OK;
  std::vector<int> nd((size_t)n);
  for (int j = 0; j < n; j++) {
    if (nodes[j] < 1 || nodes[j] > h->n_nodes) CG_FAIL(CGASM_EARG, "Dirichlet node out of range");
    nd[j] = nodes[j] - 1;
  }
  if (!h->surface) h->surface = new SurfacePlan();
  SurfacePlan* S = h->surface;
  int st;
  if ((st = grow(&S->d_nodes, &S->nodes_cap, (size_t)n))) return st;
  if ((st = grow(&S->d_vals, &S->vals_cap, (size_t)n))) return st;
  if ((st = upload(h, S->d_nodes, nd.data(), (size_t)n))) return st;
  if ((st = upload(h, S->d_vals, values, (size_t)n))) return st;
  if (h->adv_copy_pending) {
    CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_adv_copied, 0));
    h->adv_copy_pending = false;
  }
  const int threads = 128, blocks = (n + threads - 1) / threads;
  dirichlet_scalar_kernel<<<blocks, threads, 0, h->stream>>>(n, S->d_nodes, S->d_vals, raw_of(h, CGASM_F_T, 1), have_dt, dt,
                                                             h->d_adv_rhs);
  h->launches++;
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int cgasm_momentum_surface_dev(int id, const cgasm_momentum_opts* opts, const int* velocity_bc_type,
                               const double* velocity_bc, const int* pressure_bc_type) {
  GET_HANDLE(h, id);
  if (!opts) CG_FAIL(CGASM_EARG, "null opts");
  SurfacePlan* S = h->surface;
  if (!S || !S->d_sndgln) CG_FAIL(CGASM_ESTATE, "cgasm_set_surface has not been called");
  if (!h->mom_valid) CG_FAIL(CGASM_ESTATE, "no momentum result to add to: call cgasm_momentum_dev first");
  if (opts->have_les || opts->multiphase || opts->on_sphere || opts->move_mesh)
    CG_FAIL(CGASM_EUNSUPPORTED, "momentum option outside the device path; keep the Fortran loop");
  // integrate_continuity_by_parts: the boundary blocks go into the ct_m of the element loop; the ct_rhs of weak-Dirichlet
  // velocities and the pressure-condition terms (:1084-1098) need data this call does not take
  const bool ct_bdy = opts->integrate_continuity_by_parts && opts->assemble_ct_matrix_here;
  if (ct_bdy && !h->mom_has_ct) CG_FAIL(CGASM_ESTATE, "no ct_m to add the boundary blocks to: run cgasm_momentum_dev with assemble_ct_matrix_here");
  if (opts->integrate_continuity_by_parts && pressure_bc_type)
    for (int f = 0; f < S->n_faces; f++)
      if (pressure_bc_type[f] > 0)
        CG_FAIL(CGASM_EUNSUPPORTED, "pressure boundary conditions with integrate_continuity_by_parts are outside the device path");
  if (S->n_faces == 0) return CGASM_OK;
  if (!velocity_bc_type) CG_FAIL(CGASM_EARG, "null velocity_bc_type");
  const int dim = h->dim;
  bool need_bc = false, adds_matrix = false, fs_faces = false;
  const bool by_parts = opts->integrate_advection_by_parts && !opts->exclude_advection;
  for (size_t k = 0; k < (size_t)S->n_faces * dim; k++) {
    const int t = velocity_bc_type[k];
    if (t < CGASM_VBC_NONE || t > CGASM_VBC_FLUX) CG_FAIL(CGASM_EARG, "bad velocity boundary-condition type");
    need_bc = need_bc || t == CGASM_VBC_FLUX || (t == CGASM_VBC_WEAKDIRICHLET && by_parts);
    adds_matrix = adds_matrix || (by_parts && t != CGASM_VBC_WEAKDIRICHLET);
    fs_faces = fs_faces || (t == CGASM_VBC_FREE_SURFACE && opts->have_surface_fs_stabilisation && k % dim == 0);
  }
  // free-surface stabilisation (:1108-1176): needs the gravity direction; the reference exits on a consistent mass with
  // pressure-corrected absorption (:1161-1163) and adds to masslump only when it assembles one (:1168-1172)
  if (fs_faces) {
    if (!h->fields[CGASM_F_GRAVITY].set) CG_FAIL(CGASM_ESTATE, "free-surface stabilisation needs the gravity direction field");
    if (!opts->lump_mass && opts->pressure_corrected_absorption)
      CG_FAIL(CGASM_EUNSUPPORTED, "free-surface stabilisation requires a lumped mass or absorption outside the pressure correction");
    if (opts->lump_mass && opts->pressure_corrected_absorption && !(opts->assemble_inverse_masslump && h->mom_has_masslump))
      CG_FAIL(CGASM_ESTATE, "free-surface stabilisation with pressure-corrected absorption adds to masslump: assemble it");
    adds_matrix = true;
  }
  if (need_bc && !velocity_bc) CG_FAIL(CGASM_EARG, "a face needs boundary values that were not given");
  const size_t nf = (size_t)S->n_faces;
  int st;
  if ((st = grow(&S->d_itype, &S->itype_cap, nf * dim))) return st;
  if ((st = upload(h, S->d_itype, velocity_bc_type, nf * dim))) return st;
  if (pressure_bc_type) {
    if ((st = grow(&S->d_ptype, &S->ptype_cap, nf))) return st;
    if ((st = upload(h, S->d_ptype, pressure_bc_type, nf))) return st;
  }
  if (need_bc) {
    if ((st = grow(&S->d_bc, &S->bc_cap, nf * dim * dim))) return st;
    if ((st = upload(h, S->d_bc, velocity_bc, nf * dim * dim))) return st;
  }
  if (h->mom_copy_pending) {
    CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_mom_copied, 0));
    h->mom_copy_pending = false;
  }
  const RawField U = raw_of(h, CGASM_F_NU, dim), O = raw_of(h, CGASM_F_OLDU, dim), R = raw_of(h, CGASM_F_DENSITY, 1);
  const RawField G = fs_faces ? raw_of(h, CGASM_F_GRAVITY, dim) : RawField{nullptr, 0};
  double* ml_out = (fs_faces && opts->lump_mass && opts->pressure_corrected_absorption) ? h->d_masslump : nullptr;
  const int threads = 128, blocks = (S->n_faces + threads - 1) / threads;
  const FaceMesh m = face_mesh(h);
  if (dim == 3)
    momentum_surface_kernel<3><<<blocks, threads, 0, h->stream>>>(S->tab, m, *opts, U, O, R, G, S->d_itype, need_bc ? S->d_bc : nullptr,
                                                                  pressure_bc_type ? S->d_ptype : nullptr, (size_t)h->nnz,
                                                                  h->d_big_m, h->d_mom_rhs, ct_bdy ? h->d_ct_m : nullptr, ml_out);
  else
    momentum_surface_kernel<2><<<blocks, threads, 0, h->stream>>>(S->tab, m, *opts, U, O, R, G, S->d_itype, need_bc ? S->d_bc : nullptr,
                                                                  pressure_bc_type ? S->d_ptype : nullptr, (size_t)h->nnz,
                                                                  h->d_big_m, h->d_mom_rhs, ct_bdy ? h->d_ct_m : nullptr, ml_out);
  h->launches++;
  CG_CUDA(cudaGetLastError());
  // weak Dirichlet on some components only makes the diagonal blocks differ
  if (adds_matrix) {
    bool uniform = true;
    for (size_t f = 0; f < nf && uniform; f++)
      for (int d = 1; d < dim; d++)
        uniform = uniform && ((velocity_bc_type[f * dim + d] == CGASM_VBC_WEAKDIRICHLET) ==
                              (velocity_bc_type[f * dim] == CGASM_VBC_WEAKDIRICHLET));
    if (!uniform) h->mom_identical_blocks = false;
  }
  if (fs_faces) h->mom_identical_blocks = false;  // the stabilisation acts along the gravity direction only
  return CGASM_OK;
}

}  // extern "C"
