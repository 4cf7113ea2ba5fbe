// cgasm_api.cu -- the C ABI of include/cgasm.h: handle table, device residency of mesh /
// coordinates / fields, option guards, and dispatch to the assembly kernels.
//
// The guard logic mirrors what the Fortran shim evaluates before leaving the reference's
// element loop: anything outside the device path returns
// CGASM_EUNSUPPORTED so the caller keeps the Fortran loop; there is no CPU path in here.
#include "cgasm_internal.h"
#include "gather_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace cgasm {

static thread_local std::string g_last_error;
static std::mutex g_mutex;
static std::vector<Handle*> g_handles;  // id = index + 1

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
  g_last_error = buf;
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CGASM_ENODEVICE : CGASM_ECUDA;
}

Handle* get_handle(int id) {
  std::lock_guard<std::mutex> lk(g_mutex);
  if (id < 1 || id > (int)g_handles.size()) return nullptr;
  return g_handles[id - 1];
}

int scatter_momentum(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct);
int scatter_advdiff(Handle* h, const AdvDiffArgs& P);
void one_momentum(Handle* h, const MomentumArgs& A, int e, double* T, double* rhs, double* ml, double* gp);
void one_advdiff(Handle* h, const AdvDiffArgs& P, int e, double* Aout, double* rhs);

static void free_dev(void* p) {
  if (p) cudaFree(p);
}

// rec[node].lane[lane0 + c] = src[node*stride + c]   (stride 0: CONSTANT field, broadcast)
// (recw = doubles per record: 4 for the 32-byte node records, 2 for the tracer absorption / source pairs)
__global__ void pack_lanes_kernel(double* __restrict__ rec, int recw, int lane0, int ncomp,
                                  const double* __restrict__ src, int stride, const int* __restrict__ nodes, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int node = nodes ? nodes[k] : k;
  double* r = rec + (size_t)recw * node;
  for (int c = 0; c < ncomp; c++) r[lane0 + c] = src[(size_t)stride * node + c];
}

// Which packed-record lanes mirror field `slot` (-1 = coordinates): up to two targets {record array, doubles per
// record, first lane}; ncomp components each. Returns the number of targets (0: not a packed field).
int record_targets(Handle* h, int slot, double* rec[2], int recw[2], int lane0[2], int* ncomp) {
  const int dim = h->dim;
  rec[0] = rec[1] = nullptr;
  recw[0] = recw[1] = 4;
  lane0[0] = lane0[1] = 0;
  auto r4 = [](double4* p) { return reinterpret_cast<double*>(p); };
  switch (slot) {
    case -1:  // X: rec0 and rec3
      rec[0] = r4(h->d_rec0); rec[1] = r4(h->d_rec3); *ncomp = dim; return 2;
    case CGASM_F_T: rec[0] = r4(h->d_rec0); lane0[0] = 3; *ncomp = 1; return 1;
    case CGASM_F_NU: rec[0] = r4(h->d_rec1); *ncomp = dim; return 1;
    case CGASM_F_DENSITY: rec[0] = r4(h->d_rec1); lane0[0] = 3; *ncomp = 1; return 1;
    case CGASM_F_OLDU: rec[0] = r4(h->d_rec2); *ncomp = dim; return 1;
    case CGASM_F_BUOYANCY:  // rec2 and rec3 = { X, buoyancy }
      rec[0] = r4(h->d_rec2); rec[1] = r4(h->d_rec3); lane0[0] = lane0[1] = 3; *ncomp = 1; return 2;
    case CGASM_F_T_ABSORPTION:
    case CGASM_F_T_SOURCE:
      // { absorption, source } pairs of the tracer STRIP kernel, made on first use and zeroed (a lane that is
      // never set is multiplied by a zero coefficient: it must hold a finite number)
      if (!h->d_rec4) {
        if (cudaMalloc(&h->d_rec4, sizeof(double2) * (size_t)h->n_nodes) != cudaSuccess) return -1;
        if (cudaMemsetAsync(h->d_rec4, 0, sizeof(double2) * (size_t)h->n_nodes, h->stream) != cudaSuccess) return -1;
      }
      rec[0] = reinterpret_cast<double*>(h->d_rec4); recw[0] = 2; lane0[0] = slot == CGASM_F_T_ABSORPTION ? 0 : 1;
      *ncomp = 1;
      return 1;
    case CGASM_F_ABSORPTION:
    case CGASM_F_HB_DENSITY:
    case CGASM_F_SOURCE:
      if (ensure_extra_records(h) != CGASM_OK) return -1;
      if (slot == CGASM_F_SOURCE) {
        rec[0] = r4(h->d_rec6); *ncomp = dim;
      } else if (slot == CGASM_F_ABSORPTION) {
        rec[0] = r4(h->d_rec5); *ncomp = dim;
      } else {
        rec[0] = r4(h->d_rec5); lane0[0] = 3; *ncomp = 1;
      }
      return 1;
    default: return 0;
  }
}

// { absorption, hb_density } and { source, - } records of the additive STRIP pass: zeroed when made (unset lanes
// are multiplied by zero coefficients and must hold finite numbers)
int ensure_extra_records(Handle* h) {
  for (double4** p : {&h->d_rec5, &h->d_rec6})
    if (!*p) {
      CG_CUDA(cudaMalloc(p, sizeof(double4) * (size_t)h->n_nodes));
      CG_CUDA(cudaMemsetAsync(*p, 0, sizeof(double4) * (size_t)h->n_nodes, h->stream));
    }
  return CGASM_OK;
}

void* rec_array(const Handle* h, int k) {
  switch (k) {
    case 0: return h->d_rec0;
    case 1: return h->d_rec1;
    case 2: return h->d_rec2;
    case 3: return h->d_rec3;
    case 4: return h->d_rec4;
    case 5: return h->d_rec5;
    default: return h->d_rec6;
  }
}
int rec_width(int k) { return k == 4 ? 2 : 4; }  // doubles per record

const void* staged_rec(const Handle* h, int k) { return (h->d_perm && h->d_prec[k]) ? h->d_prec[k] : rec_array(h, k); }

// dst[perm[node]] = src[node], whole records
__global__ void permute_records_kernel(double* __restrict__ dst, const double* __restrict__ src, int recw,
                                       const int* __restrict__ perm, const int* __restrict__ nodes, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int node = nodes ? nodes[k] : k;
  const double* a = src + (size_t)recw * node;
  double* b = dst + (size_t)recw * perm[node];
  if (recw == 4) *reinterpret_cast<double4*>(b) = *reinterpret_cast<const double4*>(a);
  else *reinterpret_cast<double2*>(b) = *reinterpret_cast<const double2*>(a);
}

int refresh_permuted(Handle* h, unsigned recmask, const int* d_nodes, int n, cudaStream_t stream) {
  if (!h->d_perm) return CGASM_OK;
  const int count = d_nodes ? n : h->n_nodes;
  if (count <= 0) return CGASM_OK;
  for (int k = 0; k < 7; k++) {
    if (!(recmask & (1u << k)) || !rec_array(h, k)) continue;
    if (!h->d_prec[k]) {
      CG_CUDA(cudaMalloc(&h->d_prec[k], sizeof(double) * rec_width(k) * (size_t)h->n_nodes));
      // a mirror made late starts as a full copy
      permute_records_kernel<<<(h->n_nodes + 255) / 256, 256, 0, stream>>>(
          (double*)h->d_prec[k], (const double*)rec_array(h, k), rec_width(k), h->d_perm, nullptr, h->n_nodes);
      h->launches++;
      continue;
    }
    permute_records_kernel<<<(count + 255) / 256, 256, 0, stream>>>((double*)h->d_prec[k], (const double*)rec_array(h, k),
                                                                    rec_width(k), h->d_perm, d_nodes, count);
    h->launches++;
  }
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

int set_permutation(Handle* h, const std::vector<int>& perm) {
  free_dev(h->d_perm);
  h->d_perm = nullptr;
  for (void*& p : h->d_prec) {
    free_dev(p);
    p = nullptr;
  }
  if (perm.empty()) return CGASM_OK;
  CG_CUDA(cudaMalloc(&h->d_perm, sizeof(int) * perm.size()));
  CG_CUDA(cudaMemcpyAsync(h->d_perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return refresh_permuted(h, 0x7f, nullptr, 0, h->stream);
}

// which record arrays mirror a field slot (bit k = record k)
static unsigned slot_recmask(int slot) {
  switch (slot) {
    case -1: return 1u << 0 | 1u << 3;
    case CGASM_F_T: return 1u << 0;
    case CGASM_F_NU:
    case CGASM_F_DENSITY: return 1u << 1;
    case CGASM_F_OLDU: return 1u << 2;
    case CGASM_F_BUOYANCY: return 1u << 2 | 1u << 3;
    case CGASM_F_T_ABSORPTION:
    case CGASM_F_T_SOURCE: return 1u << 4;
    case CGASM_F_ABSORPTION:
    case CGASM_F_HB_DENSITY: return 1u << 5;
    case CGASM_F_SOURCE: return 1u << 6;
    default: return 0;
  }
}
unsigned slot_record_mask(int slot) { return slot_recmask(slot); }

int repack_slot(Handle* h, int slot, const int* d_nodes, int n) {
  double* rec[2];
  int recw[2], lane0[2], ncomp = 0;
  const int nt = record_targets(h, slot, rec, recw, lane0, &ncomp);
  if (nt < 0) CG_FAIL(CGASM_ECUDA, "cannot allocate the absorption / source records");
  if (nt == 0) return CGASM_OK;  // not a packed field
  const double* src;
  int stride;
  if (slot < 0) {
    src = h->d_X;
    stride = h->dim;
  } else {
    const DeviceField& f = h->fields[slot];
    src = f.d;
    stride = f.field_type == CGASM_FIELD_CONSTANT ? 0 : ncomp;
  }
  const int count = d_nodes ? n : h->n_nodes;
  if (count <= 0) return CGASM_OK;
  for (int m = 0; m < nt; m++) {
    pack_lanes_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(rec[m], recw[m], lane0[m], ncomp, src, stride, d_nodes, count);
    h->launches++;
  }
  CG_CUDA(cudaGetLastError());
  return refresh_permuted(h, slot_recmask(slot), d_nodes, n, h->stream);
}

static void destroy_handle(Handle* h) {
  cudaSetDevice(h->device);
  tiles_free(h);
  gather_free(h);
  halo_free(h);
  surface_free(h);
  cmc_free(h);
  coo_free(h);
  free_dev(h->d_ndglno);
  free_dev(h->d_X);
  free_dev(h->d_rec0);
  free_dev(h->d_rec1);
  free_dev(h->d_rec2);
  free_dev(h->d_rec3);
  free_dev(h->d_rec4);
  free_dev(h->d_rec5);
  free_dev(h->d_rec6);
  free_dev(h->d_perm);
  for (void* p : h->d_prec) free_dev(p);
  free_dev(h->d_findrm);
  free_dev(h->d_colm);
  free_dev(h->d_colour_elements);
  for (auto& f : h->fields) free_dev(f.d);
  free_dev(h->d_big_m);
  free_dev(h->d_mom_rhs);
  free_dev(h->d_masslump);
  free_dev(h->d_mass);
  free_dev(h->d_mass_rhs);
  free_dev(h->d_ct_m);
  free_dev(h->d_adv_matrix);
  free_dev(h->d_adv_rhs);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->evc0) cudaEventDestroy(h->evc0);
  if (h->evc1) cudaEventDestroy(h->evc1);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_mom_copied) cudaEventDestroy(h->ev_mom_copied);
  if (h->ev_adv_copied) cudaEventDestroy(h->ev_adv_copied);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int upload_sparsity(Handle* h) {
  free_dev(h->d_findrm);
  free_dev(h->d_colm);
  h->d_findrm = h->d_colm = nullptr;
  CG_CUDA(cudaMalloc(&h->d_findrm, sizeof(int) * ((size_t)h->n_nodes + 1)));
  CG_CUDA(cudaMalloc(&h->d_colm, sizeof(int) * (size_t)std::max(h->nnz, 1)));
  CG_CUDA(cudaMemcpyAsync(h->d_findrm, h->h_findrm.data(), sizeof(int) * ((size_t)h->n_nodes + 1),
                          cudaMemcpyHostToDevice, h->stream));
  CG_CUDA(cudaMemcpyAsync(h->d_colm, h->h_colm.data(), sizeof(int) * (size_t)h->nnz,
                          cudaMemcpyHostToDevice, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  // result buffers depend on nnz
  free_dev(h->d_big_m);
  free_dev(h->d_ct_m);
  free_dev(h->d_adv_matrix);
  free_dev(h->d_mass);
  h->d_big_m = h->d_ct_m = h->d_adv_matrix = h->d_mass = nullptr;
  h->mom_valid = h->adv_valid = false;
  h->have_sparsity = true;
  // every plan derived from the first-order pattern is stale now (the CMC plan holds slots / transposed positions
  // sized for the old nnz: cgasm_cmc_dev answers CGASM_ESTATE until the second-order pattern is set again)
  tiles_free(h);
  gather_free(h);
  cmc_free(h);
  coo_free(h);
  return CGASM_OK;
}

static FieldView view_of(const Handle* h, int slot, int comps) {
  const DeviceField& f = h->fields[slot];
  FieldView v;
  v.val = f.d;
  v.stride = (f.field_type == CGASM_FIELD_CONSTANT) ? 0 : comps;
  return v;
}

static int need_field(const Handle* h, int slot, const char* name) {
  if (!h->fields[slot].set) {
    set_error(std::string("field slot not set: ") + name);
    return CGASM_ESTATE;
  }
  return CGASM_OK;
}

static int check_momentum_opts(const cgasm_momentum_opts* o) {
  if (o->have_les || o->multiphase || o->on_sphere || o->move_mesh || o->have_coriolis ||
      o->have_geostrophic_pressure || o->have_surfacetension || o->have_vertical_stabilization ||
      o->have_swe_bottom_drag || o->have_wd_abs || o->have_temperature_dependent_viscosity ||
      o->stress_form || o->partial_stress_form || o->radial_gravity || o->vel_lump_on_submesh ||
      o->cmc_lump_on_submesh || o->abs_lump_on_submesh)
    CG_FAIL(CGASM_EUNSUPPORTED, "momentum option outside the device path; keep the Fortran loop");
  if (o->stabilisation_scheme < CGASM_STAB_NONE || o->stabilisation_scheme > CGASM_STAB_SUPG)
    CG_FAIL(CGASM_EARG, "bad stabilisation_scheme");
  if (o->stabilisation_scheme != CGASM_STAB_NONE &&
      (o->nu_bar_scheme < CGASM_NU_BAR_OPTIMAL || o->nu_bar_scheme > CGASM_NU_BAR_UNITY || o->nu_bar_scale < 0.0))
    CG_FAIL(CGASM_EARG, "bad nu_bar scheme / scale");
  if (o->have_viscosity && (o->viscosity_shape < 0 || o->viscosity_shape > CGASM_TENSOR_FULL))
    CG_FAIL(CGASM_EARG, "bad viscosity_shape");
  return CGASM_OK;
}

static int check_advdiff_opts(const cgasm_advdiff_opts* o) {
  if (o->move_mesh || o->multiphase || o->equation_type_not_advdiff)
    CG_FAIL(CGASM_EUNSUPPORTED, "tracer option outside the device path; keep the Fortran loop");
  if (o->stabilisation_scheme < CGASM_STAB_NONE || o->stabilisation_scheme > CGASM_STAB_SUPG)
    CG_FAIL(CGASM_EARG, "bad stabilisation_scheme");
  if (o->stabilisation_scheme != CGASM_STAB_NONE &&
      (o->nu_bar_scheme < CGASM_NU_BAR_OPTIMAL || o->nu_bar_scheme > CGASM_NU_BAR_UNITY || o->nu_bar_scale < 0.0))
    CG_FAIL(CGASM_EARG, "bad nu_bar scheme / scale");
  if (o->have_diffusivity && o->diffusivity_shape != CGASM_TENSOR_ISOTROPIC &&
      o->diffusivity_shape != CGASM_TENSOR_FULL)
    CG_FAIL(CGASM_EARG, "bad diffusivity_shape");
  return CGASM_OK;
}

static int make_momentum_args(Handle* h, const cgasm_momentum_opts* o, MomentumArgs& A) {
  int st;
  if ((st = check_momentum_opts(o))) return st;
  if (!h->have_X) CG_FAIL(CGASM_ESTATE, "cgasm_set_coordinates has not been called");
  if ((st = need_field(h, CGASM_F_OLDU, "OLDU"))) return st;
  if ((st = need_field(h, CGASM_F_DENSITY, "DENSITY"))) return st;
  if ((st = need_field(h, CGASM_F_NU, "NU"))) return st;
  if (o->have_viscosity && (st = need_field(h, CGASM_F_VISCOSITY, "VISCOSITY"))) return st;
  if (o->have_gravity) {
    if ((st = need_field(h, CGASM_F_BUOYANCY, "BUOYANCY"))) return st;
    if ((st = need_field(h, CGASM_F_GRAVITY, "GRAVITY"))) return st;
    if (o->subtract_out_reference_profile && (st = need_field(h, CGASM_F_HB_DENSITY, "HB_DENSITY")))
      return st;
  }
  if (o->have_absorption && (st = need_field(h, CGASM_F_ABSORPTION, "ABSORPTION"))) return st;
  if (o->have_source && (st = need_field(h, CGASM_F_SOURCE, "SOURCE"))) return st;
  const int dim = h->dim;
  A.tab = h->tab;
  A.o = *o;
  A.ndglno = h->d_ndglno;
  A.rec.r0 = h->d_rec0;
  A.rec.r1 = h->d_rec1;
  A.rec.r2 = h->d_rec2;
  A.viscosity = view_of(h, CGASM_F_VISCOSITY, dim * dim);
  A.hb_density = view_of(h, CGASM_F_HB_DENSITY, 1);
  A.gravity = view_of(h, CGASM_F_GRAVITY, dim);
  A.absorption = view_of(h, CGASM_F_ABSORPTION, dim);
  A.source = view_of(h, CGASM_F_SOURCE, dim);
  A.n_elements = h->n_elements;
  return CGASM_OK;
}

static int make_advdiff_args(Handle* h, const cgasm_advdiff_opts* o, AdvDiffArgs& P) {
  int st;
  if ((st = check_advdiff_opts(o))) return st;
  if (!h->have_X) CG_FAIL(CGASM_ESTATE, "cgasm_set_coordinates has not been called");
  if ((st = need_field(h, CGASM_F_T, "T"))) return st;
  if (o->have_advection && (st = need_field(h, CGASM_F_NU, "NU"))) return st;
  if (o->have_diffusivity && (st = need_field(h, CGASM_F_T_DIFFUSIVITY, "T_DIFFUSIVITY"))) return st;
  if (o->have_source && (st = need_field(h, CGASM_F_T_SOURCE, "T_SOURCE"))) return st;
  if (o->have_absorption && (st = need_field(h, CGASM_F_T_ABSORPTION, "T_ABSORPTION"))) return st;
  const int dim = h->dim;
  P.tab = h->tab;
  P.o = *o;
  P.ndglno = h->d_ndglno;
  P.rec.r0 = h->d_rec0;
  P.rec.r1 = h->d_rec1;
  P.rec.r2 = h->d_rec2;
  P.source = view_of(h, CGASM_F_T_SOURCE, 1);
  P.absorption = view_of(h, CGASM_F_T_ABSORPTION, 1);
  P.diffusivity = view_of(h, CGASM_F_T_DIFFUSIVITY, dim * dim);
  P.n_elements = h->n_elements;
  return CGASM_OK;
}

static int ensure(double** p, size_t count) {
  if (*p) return CGASM_OK;
  CG_CUDA(cudaMalloc(p, sizeof(double) * std::max<size_t>(count, 1)));
  return CGASM_OK;
}

}  // namespace cgasm

using namespace cgasm;

#define GET_HANDLE(h, id)                                         \
  Handle* h = get_handle(id);                                     \
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");         \
  CG_CUDA(cudaSetDevice(h->device))

extern "C" {

const char* cgasm_last_error(void) { return g_last_error.c_str(); }

int cgasm_create(int* id, int device, int dim, int loc, int ngi, int n_nodes, int n_elements,
                 const int* ndglno, const double* n, const double* dn, const double* weight) {
  if (!id || !ndglno || !n || !dn || !weight) CG_FAIL(CGASM_EARG, "null argument");
  if (dim != 2 && dim != 3) CG_FAIL(CGASM_EUNSUPPORTED, "only dim 2 and 3 are on the device path");
  if (loc != dim + 1) CG_FAIL(CGASM_EUNSUPPORTED, "only P1 simplices (loc = dim+1) are on the device path");
  if (ngi != (dim == 3 ? 5 : 4))
    CG_FAIL(CGASM_EUNSUPPORTED, "only the degree-3 simplex quadrature (ngi 4/5) is on the device path");
  if (n_nodes < 1 || n_elements < 1) CG_FAIL(CGASM_EARG, "empty mesh");
  // P1 check on the derivative table: dn(i,g,k) = delta_ik, dn(loc,g,k) = -1 for all g
  for (int k = 0; k < dim; k++)
    for (int g = 0; g < ngi; g++)
      for (int i = 0; i < loc; i++) {
        const double want = (i < dim) ? (i == k ? 1.0 : 0.0) : -1.0;
        if (std::fabs(dn[i + loc * (g + ngi * k)] - want) > 1e-14)
          CG_FAIL(CGASM_EUNSUPPORTED, "dn is not the P1 Lagrange simplex derivative table");
      }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) CG_FAIL(CGASM_ENODEVICE, "no CUDA device: libcgasm has no CPU path");
  if (device < 0) CG_CUDA(cudaGetDevice(&device));
  if (device >= ndev) CG_FAIL(CGASM_EARG, "device index out of range");
  CG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) CG_FAIL(CGASM_ENODEVICE, "libcgasm is built for sm_100a only");

  Handle* h = new Handle();
  h->device = device;
  h->dim = dim;
  h->loc = loc;
  h->ngi = ngi;
  h->n_nodes = n_nodes;
  h->n_elements = n_elements;
  for (int i = 0; i < loc; i++)
    for (int g = 0; g < ngi; g++) h->tab.N[i * ngi + g] = n[i + loc * g];
  for (int g = 0; g < ngi; g++) h->tab.w[g] = weight[g];
  {
    // moments of the rule and the node-permutation symmetry the quad kernels rely on
    Tables& T = h->tab;
    auto Nf = [&](int i, int g) { return T.N[i * ngi + g]; };
    double P[4][4], Q[4][4][4], W1v[4];
    T.Wsum = 0.0;
    for (int g = 0; g < ngi; g++) T.Wsum += T.w[g];
    for (int i = 0; i < loc; i++) {
      W1v[i] = 0.0;
      for (int g = 0; g < ngi; g++) W1v[i] += Nf(i, g) * T.w[g];
      for (int k = 0; k < loc; k++) {
        P[i][k] = 0.0;
        for (int g = 0; g < ngi; g++) P[i][k] += Nf(i, g) * Nf(k, g) * T.w[g];
        for (int l = 0; l < loc; l++) {
          Q[i][k][l] = 0.0;
          for (int g = 0; g < ngi; g++) Q[i][k][l] += Nf(i, g) * Nf(k, g) * Nf(l, g) * T.w[g];
        }
      }
    }
    T.Pd = P[0][0];
    T.Po = P[0][1];
    T.Qaaa = Q[0][0][0];
    T.Qaab = Q[0][0][1];
    T.Qabc = loc > 2 ? Q[0][1][2] : 0.0;
    T.W1 = W1v[0];
    bool sym = true;
    const double tol = 1e-15;
    for (int i = 0; i < loc; i++) {
      sym = sym && std::fabs(W1v[i] - T.W1) < tol;
      for (int k = 0; k < loc; k++) {
        sym = sym && std::fabs(P[i][k] - (i == k ? T.Pd : T.Po)) < tol;
        for (int l = 0; l < loc; l++) {
          const int eq = (i == k) + (k == l) + (i == l);
          const double want = eq == 3 ? T.Qaaa : (eq == 1 ? T.Qaab : T.Qabc);
          sym = sym && std::fabs(Q[i][k][l] - want) < tol;
        }
      }
    }
    T.sym = sym ? 1 : 0;
  }
  h->h_nd0.resize((size_t)4 * n_elements);
  {
    int bad = 0;
    int* nd0 = h->h_nd0.data();
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int e = 0; e < n_elements; e++)
      for (int i = 0; i < 4; i++) {
        int v = -1;
        if (i < loc) {
          v = ndglno[(size_t)loc * e + i] - 1;
          if (v < 0 || v >= n_nodes) bad++;
        }
        nd0[(size_t)4 * e + i] = v;
      }
    if (bad) {
      delete h;
      CG_FAIL(CGASM_EARG, "ndglno entry out of range (expects 1-based node numbers)");
    }
  }
  auto fail = [&](int code) {
    destroy_handle(h);
    return code;
  };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
      cudaEventCreate(&h->evc0) != cudaSuccess || cudaEventCreate(&h->evc1) != cudaSuccess ||
      cudaMalloc(&h->d_ndglno, sizeof(int4) * (size_t)n_elements) != cudaSuccess ||
      cudaMalloc(&h->d_X, sizeof(double) * (size_t)dim * n_nodes) != cudaSuccess ||
      cudaMalloc(&h->d_rec0, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMalloc(&h->d_rec1, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMalloc(&h->d_rec2, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMalloc(&h->d_rec3, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMemset(h->d_rec3, 0, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMemset(h->d_rec0, 0, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMemset(h->d_rec1, 0, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cudaMemset(h->d_rec2, 0, sizeof(double4) * (size_t)n_nodes) != cudaSuccess ||
      cg_upload(h->d_ndglno, h->h_nd0.data(), sizeof(int4) * (size_t)n_elements) != cudaSuccess) {
    set_error(std::string("cgasm_create: ") + cudaGetErrorString(cudaGetLastError()));
    return fail(CGASM_ECUDA);
  }
  build_node_to_element(n_nodes, n_elements, loc, h->h_nd0.data(), h->n2e_ptr, h->n2e);
  std::lock_guard<std::mutex> lk(g_mutex);
  g_handles.push_back(h);
  *id = (int)g_handles.size();
  return CGASM_OK;
}

int cgasm_destroy(int id) {
  Handle* h;
  {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (id < 1 || id > (int)g_handles.size() || !g_handles[id - 1])
      CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
    h = g_handles[id - 1];
    g_handles[id - 1] = nullptr;
  }
  destroy_handle(h);
  return CGASM_OK;
}

int cgasm_set_coordinates(int id, const double* X) {
  GET_HANDLE(h, id);
  if (!X) CG_FAIL(CGASM_EARG, "null X");
  const size_t cnt = (size_t)h->dim * h->n_nodes;
  h->h_X.assign(X, X + cnt);
  CG_CUDA(cudaMemcpyAsync(h->d_X, X, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->stream));
  int st = repack_slot(h, -1, nullptr, 0);
  if (st) return st;
  CG_CUDA(cudaStreamSynchronize(h->stream));
  h->have_X = true;
  return CGASM_OK;
}

int cgasm_build_sparsity(int id, int* nnz) {
  GET_HANDLE(h, id);
  IVec findrm, colm;
  const int64_t total = build_sparsity(h->n_nodes, h->n_elements, h->loc, h->h_nd0.data(), h->n2e_ptr, h->n2e, findrm, colm);
  if (total >= (int64_t)1 << 31) CG_FAIL(CGASM_EUNSUPPORTED, "nnz does not fit the reference's 32-bit integers");
  h->h_findrm.swap(findrm);
  h->h_colm.swap(colm);
  h->nnz = (int)total;
  if (nnz) *nnz = h->nnz;
  return upload_sparsity(h);
}

int cgasm_get_sparsity(int id, int* findrm, int* colm, int* centrm) {
  GET_HANDLE(h, id);
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no sparsity yet");
  if (findrm)
    for (int r = 0; r <= h->n_nodes; r++) findrm[r] = h->h_findrm[r] + 1;
  if (colm)
    for (int k = 0; k < h->nnz; k++) colm[k] = h->h_colm[k] + 1;
  if (centrm)
    for (int r = 0; r < h->n_nodes; r++) {
      centrm[r] = 0;  // lists2csr_sparsity: 0 when the diagonal is missing
      for (int k = h->h_findrm[r]; k < h->h_findrm[r + 1]; k++)
        if (h->h_colm[k] == r) {
          centrm[r] = k + 1;
          break;
        }
    }
  return CGASM_OK;
}

int cgasm_set_sparsity(int id, int rows, int nnz, const int* findrm, const int* colm) {
  GET_HANDLE(h, id);
  if (!findrm || !colm) CG_FAIL(CGASM_EARG, "null argument");
  if (rows != h->n_nodes) CG_FAIL(CGASM_EARG, "sparsity rows != n_nodes");
  if (nnz < 0 || findrm[0] != 1 || findrm[rows] != nnz + 1) CG_FAIL(CGASM_EARG, "findrm is not a 1-based CSR row pointer");
  // Validate everything against the caller's arrays into temporaries first; the handle keeps its previous
  // pattern (host copy, nnz, device copy, plans) unless every check passes.
  for (int r = 0; r < rows; r++)
    if (findrm[r] < 1 || findrm[r + 1] < findrm[r] || findrm[r + 1] > nnz + 1) CG_FAIL(CGASM_EARG, "findrm not monotone inside [1, nnz+1]");
  IVec new_findrm((size_t)rows + 1), new_colm((size_t)nnz);
  int bad_range = 0, bad_order = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_range, bad_order)
  for (int r = 0; r < rows; r++) {
    new_findrm[r] = findrm[r] - 1;
    for (int k = findrm[r] - 1; k < findrm[r + 1] - 1; k++) {
      const int c = colm[k] - 1;
      if (c < 0 || c >= h->n_nodes) bad_range++;
      if (k > findrm[r] - 1 && colm[k] <= colm[k - 1]) bad_order++;
      new_colm[k] = c;
    }
  }
  new_findrm[rows] = nnz;
  if (bad_range) CG_FAIL(CGASM_EARG, "colm entry out of range");
  if (bad_order) CG_FAIL(CGASM_EARG, "rows must be sorted ascending (sorted_rows)");
  // every element pair must be present (Sparse_Tools.F90:2644 would FLAbort otherwise)
  int missing = 0;
#pragma omp parallel for schedule(static) reduction(+ : missing)
  for (int e = 0; e < h->n_elements; e++)
    for (int i = 0; i < h->loc; i++) {
      const int r = h->h_nd0[(size_t)4 * e + i];
      for (int j = 0; j < h->loc; j++) {
        const int c = h->h_nd0[(size_t)4 * e + j];
        const int* b = new_colm.data() + new_findrm[r];
        const int* en = new_colm.data() + new_findrm[r + 1];
        if (!std::binary_search(b, en, c)) missing++;
      }
    }
  if (missing) CG_FAIL(CGASM_EARG, "sparsity misses an element node pair");
  h->h_findrm.swap(new_findrm);
  h->h_colm.swap(new_colm);
  h->nnz = nnz;
  return upload_sparsity(h);
}

int cgasm_build_colouring(int id, int* ncolours) {
  GET_HANDLE(h, id);
  std::vector<int> colour_of;
  const int nc = greedy_colouring(h->n_elements, h->loc, h->h_nd0.data(), h->n2e_ptr, h->n2e, colour_of);
  if (nc < 0) CG_FAIL(CGASM_EUNSUPPORTED, "mesh needs more than 256 colours");
  h->ncolours = nc;
  colour_sets(h->n_elements, nc, colour_of, h->h_colour_ptr, h->h_colour_elements);
  free_dev(h->d_colour_elements);
  h->d_colour_elements = nullptr;
  CG_CUDA(cudaMalloc(&h->d_colour_elements, sizeof(int) * (size_t)h->n_elements));
  CG_CUDA(cg_upload(h->d_colour_elements, h->h_colour_elements.data(), sizeof(int) * (size_t)h->n_elements));
  if (ncolours) *ncolours = nc;
  return CGASM_OK;
}

int cgasm_get_colouring(int id, int* colour_ptr, int* colour_elements) {
  GET_HANDLE(h, id);
  if (!h->ncolours) CG_FAIL(CGASM_ESTATE, "no colouring yet");
  if (colour_ptr)
    for (int c = 0; c <= h->ncolours; c++) colour_ptr[c] = h->h_colour_ptr[c] + 1;
  if (colour_elements)
    for (int e = 0; e < h->n_elements; e++) colour_elements[e] = h->h_colour_elements[e] + 1;
  return CGASM_OK;
}

int cgasm_set_colouring(int id, int ncolours, const int* colour_ptr, const int* colour_elements) {
  GET_HANDLE(h, id);
  if (ncolours < 1 || !colour_ptr || !colour_elements) CG_FAIL(CGASM_EARG, "bad colouring");
  if (colour_ptr[0] != 1 || colour_ptr[ncolours] != h->n_elements + 1)
    CG_FAIL(CGASM_EARG, "colour_ptr must cover every element exactly once (1-based)");
  // validity: inside a colour no two elements may share a node, or the plain-store scatter races
  std::vector<int> stamp((size_t)h->n_nodes, -1);
  std::vector<char> seen((size_t)h->n_elements, 0);
  for (int c = 0; c < ncolours; c++)
    for (int k = colour_ptr[c] - 1; k < colour_ptr[c + 1] - 1; k++) {
      const int e = colour_elements[k] - 1;
      if (e < 0 || e >= h->n_elements || seen[e]) CG_FAIL(CGASM_EARG, "colour_elements is not a permutation");
      seen[e] = 1;
      for (int i = 0; i < h->loc; i++) {
        const int node = h->h_nd0[(size_t)4 * e + i];
        if (stamp[node] == c) CG_FAIL(CGASM_EARG, "invalid colouring: two elements of a colour share a node");
        stamp[node] = c;
      }
    }
  h->ncolours = ncolours;
  h->h_colour_ptr.resize((size_t)ncolours + 1);
  h->h_colour_elements.resize((size_t)h->n_elements);
  for (int c = 0; c <= ncolours; c++) h->h_colour_ptr[c] = colour_ptr[c] - 1;
  for (int e = 0; e < h->n_elements; e++) h->h_colour_elements[e] = colour_elements[e] - 1;
  free_dev(h->d_colour_elements);
  h->d_colour_elements = nullptr;
  CG_CUDA(cudaMalloc(&h->d_colour_elements, sizeof(int) * (size_t)h->n_elements));
  CG_CUDA(cg_upload(h->d_colour_elements, h->h_colour_elements.data(), sizeof(int) * (size_t)h->n_elements));
  return CGASM_OK;
}

int cgasm_set_scatter(int id, int variant) {
  GET_HANDLE(h, id);
  if (variant < CGASM_SCATTER_ATOMIC || variant > CGASM_SCATTER_STRIP) CG_FAIL(CGASM_EARG, "unknown scatter variant");
  if (variant == CGASM_SCATTER_COLOURED && !h->ncolours) {
    int st = cgasm_build_colouring(id, nullptr);
    if (st) return st;
  }
  if (variant == CGASM_SCATTER_TILED) {
    if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "tiled scatter needs the sparsity first");
    if (!h->tiles) {
      int st = tiles_build(h);
      if (st) return st;
    }
  }
  if (variant == CGASM_SCATTER_GATHER || variant == CGASM_SCATTER_STRIP) {
    if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "gather scatter needs the sparsity first");
    if (!h->gather) {
      int st = gather_build_rows(h);
      if (st) return st;
    }
    // GATHER needs its pair lists now; STRIP builds them only if an option set falls back to them
    int st = variant == CGASM_SCATTER_GATHER ? gather_build_pairs(h) : strip_build(h);
    if (st) return st;
  }
  h->scatter = variant;
  return CGASM_OK;
}

int cgasm_set_field(int id, int slot, int rank, int field_type, const double* val, int n_val_nodes) {
  GET_HANDLE(h, id);
  if (slot < 0 || slot >= CGASM_F_NSLOTS || !val) CG_FAIL(CGASM_EARG, "bad slot or null val");
  if (int js = halo_join(h)) return js;  // a pending halo exchange writes the fields
  static const int kRank[CGASM_F_NSLOTS] = {1, 1, 0, 2, 0, 0, 1, 1, 1, 0, 2, 0, 0};
  if (rank != kRank[slot]) CG_FAIL(CGASM_EARG, "field rank does not match the slot");
  if (field_type == CGASM_FIELD_CONSTANT) {
    if (n_val_nodes != 1) CG_FAIL(CGASM_EARG, "a CONSTANT field has one node");
  } else if (field_type == CGASM_FIELD_NORMAL) {
    // a field on another mesh (P2 density, P0 viscosity ...) is an option set outside the device path, not a caller bug:
    // the shim keeps the Fortran loop
    if (n_val_nodes != h->n_nodes) CG_FAIL(CGASM_EUNSUPPORTED, "a NORMAL field must live on the velocity mesh nodes");
  } else {
    CG_FAIL(CGASM_EUNSUPPORTED, "only NORMAL and CONSTANT fields are on the device path");
  }
  size_t comps = 1;
  for (int r = 0; r < rank; r++) comps *= (size_t)h->dim;
  const size_t count = comps * (size_t)n_val_nodes;
  DeviceField& f = h->fields[slot];
  if (f.d && f.count != count) {
    cudaFree(f.d);
    f.d = nullptr;
  }
  if (!f.d) CG_CUDA(cudaMalloc(&f.d, sizeof(double) * count));
  CG_CUDA(cudaMemcpyAsync(f.d, val, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
  f.rank = rank;
  f.field_type = field_type;
  f.n_val_nodes = n_val_nodes;
  f.count = count;
  f.set = true;
  if (field_type == CGASM_FIELD_CONSTANT)
    for (size_t c = 0; c < count && c < 9; c++) f.h_const[c] = val[c];
  int st = repack_slot(h, slot, nullptr, 0);
  if (st) return st;
  if (!h->async) CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

int cgasm_get_field(int id, int slot, double* val, int n_val_nodes) {
  GET_HANDLE(h, id);
  if (slot < 0 || slot >= CGASM_F_NSLOTS || !val) CG_FAIL(CGASM_EARG, "bad slot or null val");
  const DeviceField& f = h->fields[slot];
  if (!f.set) CG_FAIL(CGASM_ESTATE, "field slot not set");
  if (n_val_nodes != f.n_val_nodes) CG_FAIL(CGASM_EARG, "n_val_nodes mismatch");
  if (int js = halo_join(h)) return js;
  CG_CUDA(cudaMemcpyAsync(val, f.d, sizeof(double) * f.count, cudaMemcpyDeviceToHost, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

// runs the momentum element loop of the handle's scatter variant into the handle's result buffers
static int run_momentum(Handle* h, const MomentumArgs& A, bool want_ml, bool want_ct) {
  const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes, dim = (size_t)h->dim;
  int st;
  if (h->scatter != CGASM_SCATTER_GATHER && h->scatter != CGASM_SCATTER_STRIP && (st = halo_join(h))) return st;
  if (h->scatter == CGASM_SCATTER_TILED) {
    st = tiles_momentum(h, A, want_ml, want_ct);
    h->mom_path = CGASM_PATH_TILED;
  } else if (h->scatter == CGASM_SCATTER_GATHER || h->scatter == CGASM_SCATTER_STRIP) {
    st = gather_momentum(h, A, want_ml, want_ct);
  } else {
    // zero(big_m), zero(rhs) ... (Momentum_Equation.F90:593-606) then accumulate
    CG_CUDA(cudaMemsetAsync(h->d_big_m, 0, sizeof(double) * dim * nnz, h->stream));
    CG_CUDA(cudaMemsetAsync(h->d_mom_rhs, 0, sizeof(double) * dim * nn, h->stream));
    if (want_ml) CG_CUDA(cudaMemsetAsync(h->d_masslump, 0, sizeof(double) * dim * nn, h->stream));
    if (want_ct) CG_CUDA(cudaMemsetAsync(h->d_ct_m, 0, sizeof(double) * dim * nnz, h->stream));
    st = scatter_momentum(h, A, want_ml, want_ct);
    h->mom_path = CGASM_PATH_ELEMENT;
  }
  return st;
}

int cgasm_momentum_dev(int id, const cgasm_momentum_opts* opts) {
  GET_HANDLE(h, id);
  if (!opts) CG_FAIL(CGASM_EARG, "null opts");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no sparsity: call cgasm_build_sparsity or cgasm_set_sparsity");
  MomentumArgs A;
  int st = make_momentum_args(h, opts, A);
  if (st) return st;
  const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes, dim = (size_t)h->dim;
  const bool want_ml = opts->assemble_inverse_masslump != 0;
  const bool want_ct = opts->assemble_ct_matrix_here != 0;
  if ((st = ensure(&h->d_big_m, dim * nnz))) return st;
  if ((st = ensure(&h->d_mom_rhs, dim * nn))) return st;
  if (want_ml && (st = ensure(&h->d_masslump, dim * nn))) return st;
  if (want_ct && (st = ensure(&h->d_ct_m, dim * nnz))) return st;
  if (opts->stabilisation_scheme != CGASM_STAB_NONE && h->scatter != CGASM_SCATTER_ATOMIC &&
      h->scatter != CGASM_SCATTER_GATHER && h->scatter != CGASM_SCATTER_STRIP)
    CG_FAIL(CGASM_EUNSUPPORTED, "SU/SUPG stabilisation runs on the ATOMIC and GATHER scatter variants only");
  if (opts->assemble_mass_matrix && opts->stabilisation_scheme == CGASM_STAB_SUPG)
    CG_FAIL(CGASM_EUNSUPPORTED, "assemble_mass_matrix with a SUPG test function is outside the device path");
  if (h->mom_copy_pending) {  // an asynchronous fetch may still be reading the previous result
    CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_mom_copied, 0));
    h->mom_copy_pending = false;
  }
  CG_CUDA(cudaEventRecord(h->ev0, h->stream));
  if ((st = run_momentum(h, A, want_ml, want_ct))) return st;
  h->mom_has_mass = false;
  if (opts->assemble_mass_matrix) {
    // The `mass` matrix (Momentum_CG.F90:1567-1571, :2073-2078) is the big_m of ANOTHER option set of the same loop:
    // consistent mass, no advection / viscosity / gravity / sources, and the full absorption matrix when the
    // absorption is pressure-corrected -- T = mass_mat + dt theta absorption_mat. Run that set into the mass buffers.
    if ((st = ensure(&h->d_mass, dim * nnz)) || (st = ensure(&h->d_mass_rhs, dim * nn))) return st;
    cgasm_momentum_opts mo = *opts;
    mo.assemble_mass_matrix = 0;
    mo.lump_mass = 0;
    mo.exclude_mass = 0;
    mo.exclude_advection = 1;
    mo.have_viscosity = 0;
    mo.have_gravity = 0;
    mo.have_source = 0;
    mo.have_absorption = (opts->have_absorption && opts->pressure_corrected_absorption) ? 1 : 0;
    mo.lump_absorption = 0;
    mo.pressure_corrected_absorption = 0;
    mo.stabilisation_scheme = CGASM_STAB_NONE;
    mo.assemble_inverse_masslump = 0;
    mo.assemble_ct_matrix_here = 0;
    MomentumArgs Am;
    if ((st = make_momentum_args(h, &mo, Am))) return st;
    const int path = h->mom_path;
    std::swap(h->d_big_m, h->d_mass);
    std::swap(h->d_mom_rhs, h->d_mass_rhs);
    st = run_momentum(h, Am, false, false);
    std::swap(h->d_big_m, h->d_mass);
    std::swap(h->d_mom_rhs, h->d_mass_rhs);
    h->mom_path = path;
    if (st) return st;
    h->mom_has_mass = true;
  }
  CG_CUDA(cudaEventRecord(h->ev1, h->stream));
  CG_CUDA(cudaGetLastError());
  h->last_combined = false;
  h->mom_has_masslump = want_ml;
  h->mom_has_ct = want_ct;
  h->mom_identical_blocks = !opts->have_absorption;
  h->mom_valid = true;
  return CGASM_OK;
}

int cgasm_momentum_mass_fetch(int id, double* mass) {
  GET_HANDLE(h, id);
  if (!mass) CG_FAIL(CGASM_EARG, "null mass");
  if (!h->mom_valid || !h->mom_has_mass) CG_FAIL(CGASM_ESTATE, "the mass matrix was not assembled (assemble_mass_matrix = 0)");
  CG_CUDA(cudaMemcpyAsync(mass, h->d_mass, sizeof(double) * (size_t)h->dim * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
  CG_CUDA(cudaStreamSynchronize(h->stream));
  return CGASM_OK;
}

int cgasm_momentum_mass_dev(int id, double** mass_dev) {
  GET_HANDLE(h, id);
  if (!mass_dev) CG_FAIL(CGASM_EARG, "null out");
  if (!h->mom_valid || !h->mom_has_mass) CG_FAIL(CGASM_ESTATE, "the mass matrix was not assembled (assemble_mass_matrix = 0)");
  *mass_dev = h->d_mass;
  return CGASM_OK;
}

int cgasm_advdiff_dev(int id, const cgasm_advdiff_opts* opts) {
  GET_HANDLE(h, id);
  if (!opts) CG_FAIL(CGASM_EARG, "null opts");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no sparsity: call cgasm_build_sparsity or cgasm_set_sparsity");
  AdvDiffArgs P;
  int st = make_advdiff_args(h, opts, P);
  if (st) return st;
  const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes;
  if ((st = ensure(&h->d_adv_matrix, nnz))) return st;
  if ((st = ensure(&h->d_adv_rhs, nn))) return st;
  if (opts->stabilisation_scheme != CGASM_STAB_NONE && h->scatter != CGASM_SCATTER_ATOMIC &&
      h->scatter != CGASM_SCATTER_GATHER && h->scatter != CGASM_SCATTER_STRIP)
    CG_FAIL(CGASM_EUNSUPPORTED, "SU/SUPG stabilisation runs on the ATOMIC and GATHER scatter variants only");
  if (h->adv_copy_pending) {
    CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_adv_copied, 0));
    h->adv_copy_pending = false;
  }
  if (h->scatter != CGASM_SCATTER_GATHER && h->scatter != CGASM_SCATTER_STRIP && (st = halo_join(h))) return st;
  CG_CUDA(cudaEventRecord(h->ev0, h->stream));
  if (h->scatter == CGASM_SCATTER_TILED) {
    st = tiles_advdiff(h, P);
    h->adv_path = CGASM_PATH_TILED;
  } else if (h->scatter == CGASM_SCATTER_GATHER || h->scatter == CGASM_SCATTER_STRIP) {
    st = gather_advdiff(h, P);
  } else {
    CG_CUDA(cudaMemsetAsync(h->d_adv_matrix, 0, sizeof(double) * nnz, h->stream));
    CG_CUDA(cudaMemsetAsync(h->d_adv_rhs, 0, sizeof(double) * nn, h->stream));
    st = scatter_advdiff(h, P);
    h->adv_path = CGASM_PATH_ELEMENT;
  }
  if (st) return st;
  CG_CUDA(cudaEventRecord(h->ev1, h->stream));
  CG_CUDA(cudaGetLastError());
  h->last_combined = false;
  h->adv_valid = true;
  return CGASM_OK;
}

// Both element loops of a time step in one call. When both option sets are the common STRIP sets (lumped or excluded
// mass, plain advection, constant isotropic viscosity / diffusivity, constant gravity direction, no absorption / sources)
// one kernel assembles both systems, sharing the strip, the staged node records and the element geometry
// (strip_fused.cu); otherwise the two loops run one after the other. Results: exactly those of the two calls.
int cgasm_momentum_advdiff_dev(int id, const cgasm_momentum_opts* mopts, const cgasm_advdiff_opts* aopts) {
  GET_HANDLE(h, id);
  if (!mopts || !aopts) CG_FAIL(CGASM_EARG, "null opts");
  if (!h->have_sparsity) CG_FAIL(CGASM_ESTATE, "no sparsity: call cgasm_build_sparsity or cgasm_set_sparsity");
  MomentumArgs M;
  AdvDiffArgs A;
  int st = make_momentum_args(h, mopts, M);
  if (st) return st;
  if ((st = make_advdiff_args(h, aopts, A))) return st;
  CG_CUDA(cudaEventRecord(h->evc0, h->stream));
  if (aopts->have_advection && strip_fused_ok(h, M, A)) {
    const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes, dim = (size_t)h->dim;
    const bool want_ml = mopts->assemble_inverse_masslump != 0;
    if ((st = ensure(&h->d_big_m, dim * nnz)) || (st = ensure(&h->d_mom_rhs, dim * nn)) ||
        (want_ml && (st = ensure(&h->d_masslump, dim * nn))) || (st = ensure(&h->d_adv_matrix, nnz)) ||
        (st = ensure(&h->d_adv_rhs, nn)))
      return st;
    if (h->mom_copy_pending) {
      CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_mom_copied, 0));
      h->mom_copy_pending = false;
    }
    if (h->adv_copy_pending) {
      CG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_adv_copied, 0));
      h->adv_copy_pending = false;
    }
    if ((st = strip_fused(h, M, A))) return st;
    h->mom_has_masslump = want_ml;
    h->mom_has_ct = false;
    h->mom_identical_blocks = true;
    h->mom_valid = h->adv_valid = true;
  } else {
    if ((st = cgasm_momentum_dev(id, mopts))) return st;
    if ((st = cgasm_advdiff_dev(id, aopts))) return st;
  }
  CG_CUDA(cudaEventRecord(h->evc1, h->stream));
  CG_CUDA(cudaGetLastError());
  h->last_combined = true;
  return CGASM_OK;
}

// Device -> host copies of a result. Synchronous flavour: on the handle's stream, returns when the data
// is in the caller's buffers. Asynchronous flavour (cgasm_set_async): the copies are queued on the copy
// stream behind an event that marks the result complete, the call returns at once and the compute
// stream is free for the next element loop; cgasm_synchronize waits for everything.
static int fetch_begin(Handle* h, cudaStream_t* s) {
  if (!h->async) {
    *s = h->stream;
    return CGASM_OK;
  }
  CG_CUDA(cudaEventRecord(h->ev_ready, h->stream));
  CG_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
  *s = h->copy_stream;
  return CGASM_OK;
}
static int fetch_end(Handle* h, bool momentum) {
  if (!h->async) {
    CG_CUDA(cudaStreamSynchronize(h->stream));
    return CGASM_OK;
  }
  if (momentum) {
    CG_CUDA(cudaEventRecord(h->ev_mom_copied, h->copy_stream));
    h->mom_copy_pending = true;
  } else {
    CG_CUDA(cudaEventRecord(h->ev_adv_copied, h->copy_stream));
    h->adv_copy_pending = true;
  }
  return CGASM_OK;
}

int cgasm_momentum_fetch(int id, double* big_m, double* rhs, double* masslump, double* ct_m) {
  GET_HANDLE(h, id);
  if (!h->mom_valid) CG_FAIL(CGASM_ESTATE, "no momentum result to fetch");
  const size_t nnz = (size_t)h->nnz, nn = (size_t)h->n_nodes, dim = (size_t)h->dim;
  if (masslump && !h->mom_has_masslump) CG_FAIL(CGASM_ESTATE, "masslump was not assembled (assemble_inverse_masslump = 0)");
  if (ct_m && !h->mom_has_ct) CG_FAIL(CGASM_ESTATE, "ct_m was not assembled (assemble_ct_matrix_here = 0)");
  cudaStream_t cs;
  int st = fetch_begin(h, &cs);
  if (st) return st;
  if (big_m) CG_CUDA(cudaMemcpyAsync(big_m, h->d_big_m, sizeof(double) * dim * nnz, cudaMemcpyDeviceToHost, cs));
  if (rhs) CG_CUDA(cudaMemcpyAsync(rhs, h->d_mom_rhs, sizeof(double) * dim * nn, cudaMemcpyDeviceToHost, cs));
  if (masslump) CG_CUDA(cudaMemcpyAsync(masslump, h->d_masslump, sizeof(double) * dim * nn, cudaMemcpyDeviceToHost, cs));
  if (ct_m) CG_CUDA(cudaMemcpyAsync(ct_m, h->d_ct_m, sizeof(double) * dim * nnz, cudaMemcpyDeviceToHost, cs));
  return fetch_end(h, true);
}

int cgasm_momentum_identical_blocks(int id, int* identical) {
  GET_HANDLE(h, id);
  if (!identical) CG_FAIL(CGASM_EARG, "null out");
  if (!h->mom_valid) CG_FAIL(CGASM_ESTATE, "no momentum result");
  *identical = h->mom_identical_blocks ? 1 : 0;
  return CGASM_OK;
}

int cgasm_momentum_fetch_blocks(int id, int first_block, int nblocks, double* big_m) {
  GET_HANDLE(h, id);
  if (!h->mom_valid) CG_FAIL(CGASM_ESTATE, "no momentum result to fetch");
  if (!big_m || first_block < 0 || nblocks < 1 || first_block + nblocks > h->dim) CG_FAIL(CGASM_EARG, "bad block range");
  const size_t nnz = (size_t)h->nnz;
  cudaStream_t cs;
  int st = fetch_begin(h, &cs);
  if (st) return st;
  CG_CUDA(cudaMemcpyAsync(big_m, h->d_big_m + (size_t)first_block * nnz, sizeof(double) * nnz * nblocks,
                          cudaMemcpyDeviceToHost, cs));
  return fetch_end(h, true);
}

int cgasm_advdiff_fetch(int id, double* matrix_val, double* rhs) {
  GET_HANDLE(h, id);
  if (!h->adv_valid) CG_FAIL(CGASM_ESTATE, "no tracer result to fetch");
  cudaStream_t cs;
  int st = fetch_begin(h, &cs);
  if (st) return st;
  if (matrix_val)
    CG_CUDA(cudaMemcpyAsync(matrix_val, h->d_adv_matrix, sizeof(double) * (size_t)h->nnz, cudaMemcpyDeviceToHost, cs));
  if (rhs) CG_CUDA(cudaMemcpyAsync(rhs, h->d_adv_rhs, sizeof(double) * (size_t)h->n_nodes, cudaMemcpyDeviceToHost, cs));
  return fetch_end(h, false);
}

int cgasm_set_async(int id, int on) {
  GET_HANDLE(h, id);
  if (on && !h->copy_stream) {
    CG_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CG_CUDA(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    CG_CUDA(cudaEventCreateWithFlags(&h->ev_mom_copied, cudaEventDisableTiming));
    CG_CUDA(cudaEventCreateWithFlags(&h->ev_adv_copied, cudaEventDisableTiming));
  }
  if (!on && h->async) {
    CG_CUDA(cudaStreamSynchronize(h->stream));
    CG_CUDA(cudaStreamSynchronize(h->copy_stream));
    h->mom_copy_pending = h->adv_copy_pending = false;
  }
  h->async = on != 0;
  return CGASM_OK;
}

int cgasm_momentum(int id, const cgasm_momentum_opts* opts, double* big_m, double* rhs,
                   double* masslump, double* ct_m) {
  int st = cgasm_momentum_dev(id, opts);
  if (st) return st;
  return cgasm_momentum_fetch(id, big_m, rhs, opts->assemble_inverse_masslump ? masslump : nullptr,
                              opts->assemble_ct_matrix_here ? ct_m : nullptr);
}

int cgasm_advdiff(int id, const cgasm_advdiff_opts* opts, double* matrix_val, double* rhs) {
  int st = cgasm_advdiff_dev(id, opts);
  if (st) return st;
  return cgasm_advdiff_fetch(id, matrix_val, rhs);
}

int cgasm_momentum_result_dev(int id, double** big_m_dev, double** rhs_dev, double** masslump_dev,
                              double** ct_m_dev) {
  GET_HANDLE(h, id);
  if (!h->mom_valid) CG_FAIL(CGASM_ESTATE, "no momentum result");
  if (big_m_dev) *big_m_dev = h->d_big_m;
  if (rhs_dev) *rhs_dev = h->d_mom_rhs;
  if (masslump_dev) *masslump_dev = h->mom_has_masslump ? h->d_masslump : nullptr;
  if (ct_m_dev) *ct_m_dev = h->mom_has_ct ? h->d_ct_m : nullptr;
  return CGASM_OK;
}

int cgasm_advdiff_result_dev(int id, double** matrix_dev, double** rhs_dev) {
  GET_HANDLE(h, id);
  if (!h->adv_valid) CG_FAIL(CGASM_ESTATE, "no tracer result");
  if (matrix_dev) *matrix_dev = h->d_adv_matrix;
  if (rhs_dev) *rhs_dev = h->d_adv_rhs;
  return CGASM_OK;
}

int cgasm_momentum_element(int id, const cgasm_momentum_opts* opts, int ele, double* big_m_tensor_addto,
                           double* rhs_addto, double* mass_lump, double* grad_p_u_mat) {
  GET_HANDLE(h, id);
  if (!opts) CG_FAIL(CGASM_EARG, "null opts");
  if (ele < 1 || ele > h->n_elements) CG_FAIL(CGASM_EARG, "element number out of range");
  MomentumArgs A;
  int st = make_momentum_args(h, opts, A);
  if (st) return st;
  const int dim = h->dim, loc = h->loc;
  const size_t nT = (size_t)dim * dim * loc * loc, nr = (size_t)dim * loc, ng = (size_t)dim * loc * loc;
  double* buf = nullptr;
  CG_CUDA(cudaMalloc(&buf, sizeof(double) * (nT + 2 * nr + ng)));
  one_momentum(h, A, ele - 1, buf, buf + nT, buf + nT + nr, buf + nT + 2 * nr);
  std::vector<double> hb(nT + 2 * nr + ng);
  cudaError_t e = cudaMemcpyAsync(hb.data(), buf, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  CG_CUDA(e);
  if (big_m_tensor_addto) memcpy(big_m_tensor_addto, hb.data(), sizeof(double) * nT);
  if (rhs_addto) memcpy(rhs_addto, hb.data() + nT, sizeof(double) * nr);
  if (mass_lump) memcpy(mass_lump, hb.data() + nT + nr, sizeof(double) * nr);
  if (grad_p_u_mat) memcpy(grad_p_u_mat, hb.data() + nT + 2 * nr, sizeof(double) * ng);
  return CGASM_OK;
}

int cgasm_advdiff_element(int id, const cgasm_advdiff_opts* opts, int ele, double* matrix_addto,
                          double* rhs_addto) {
  GET_HANDLE(h, id);
  if (!opts) CG_FAIL(CGASM_EARG, "null opts");
  if (ele < 1 || ele > h->n_elements) CG_FAIL(CGASM_EARG, "element number out of range");
  AdvDiffArgs P;
  int st = make_advdiff_args(h, opts, P);
  if (st) return st;
  const int loc = h->loc;
  double* buf = nullptr;
  CG_CUDA(cudaMalloc(&buf, sizeof(double) * (size_t)(loc * loc + loc)));
  one_advdiff(h, P, ele - 1, buf, buf + loc * loc);
  std::vector<double> hb((size_t)loc * loc + loc);
  cudaError_t e = cudaMemcpyAsync(hb.data(), buf, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  CG_CUDA(e);
  if (matrix_addto) memcpy(matrix_addto, hb.data(), sizeof(double) * (size_t)loc * loc);
  if (rhs_addto) memcpy(rhs_addto, hb.data() + loc * loc, sizeof(double) * (size_t)loc);
  return CGASM_OK;
}

int cgasm_synchronize(int id) {
  GET_HANDLE(h, id);
  if (int js = halo_join(h)) return js;
  CG_CUDA(cudaStreamSynchronize(h->stream));
  if (h->copy_stream) {
    CG_CUDA(cudaStreamSynchronize(h->copy_stream));
    h->mom_copy_pending = h->adv_copy_pending = false;
  }
  return CGASM_OK;
}

int cgasm_stream(int id, void** stream) {
  GET_HANDLE(h, id);
  if (!stream) CG_FAIL(CGASM_EARG, "null stream out");
  *stream = (void*)h->stream;
  return CGASM_OK;
}

int cgasm_launch_count(int id, long long* launches) {
  GET_HANDLE(h, id);
  if (!launches) CG_FAIL(CGASM_EARG, "null out");
  *launches = h->launches;
  return CGASM_OK;
}

int cgasm_last_path(int id, int* momentum_path, int* advdiff_path) {
  GET_HANDLE(h, id);
  if (momentum_path) *momentum_path = h->mom_path;
  if (advdiff_path) *advdiff_path = h->adv_path;
  return CGASM_OK;
}

int cgasm_plan_stats(int id, double* stats) {
  GET_HANDLE(h, id);
  if (!stats) CG_FAIL(CGASM_EARG, "null out");
  const GatherPlan* P = h->gather;
  if (!P) CG_FAIL(CGASM_ESTATE, "no row-block plan: call cgasm_set_scatter(GATHER | STRIP) first");
  stats[0] = P->nblocks;
  stats[1] = kBR;
  stats[2] = P->maxlen;
  stats[3] = P->strip_entries_per_pair;
  stats[4] = P->staged_ok ? 1.0 : 0.0;
  stats[5] = P->blk_nodes_max;
  stats[6] = P->nl;
  stats[7] = P->staged_ok ? (double)(sizeof(double) * (size_t)P->maxlen * kAS + (size_t)P->nl * 88 + kBR * 16) : 0.0;
  return CGASM_OK;
}

int cgasm_last_kernel_ms(int id, float* ms) {
  GET_HANDLE(h, id);
  if (!ms) CG_FAIL(CGASM_EARG, "null out");
  cudaEvent_t e0 = h->last_combined ? h->evc0 : h->ev0, e1 = h->last_combined ? h->evc1 : h->ev1;
  CG_CUDA(cudaEventSynchronize(e1));
  CG_CUDA(cudaEventElapsedTime(ms, e0, e1));
  return CGASM_OK;
}

}  // extern "C"
