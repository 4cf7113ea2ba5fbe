// cgasm_internal.h -- handle layout and helpers shared by the translation units of
// libcgasm.so. Not part of the ABI (that is include/cgasm.h).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cgasm.h"
#include "element_math.cuh"

namespace cgasm {

// std::vector whose resize() leaves new elements uninitialised: the big host arrays (connectivity, adjacency,
// sparsity) are filled by parallel loops right after they are sized, and the serial zero-fill of a plain
// std::vector (1.6 GB for the connectivity of a 100 M-tet mesh) was a visible part of the set-up time.
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAlloc<U>;
  };
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
    else ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
using IVec = std::vector<int, DefaultInitAlloc<int>>;
using I64Vec = std::vector<int64_t, DefaultInitAlloc<int64_t>>;

struct DeviceField {
  double* d = nullptr;
  int rank = 0;
  int field_type = CGASM_FIELD_NORMAL;
  int n_val_nodes = 0;
  size_t count = 0;  // doubles
  bool set = false;
  double h_const[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // host copy of a CONSTANT field's single node (<= dim*dim values)
};

struct TilePlan;    // tiled.cu
struct GatherPlan;  // gather.cu
struct HaloPlan;  // halo.cu
struct SurfacePlan;  // surface.cu
struct CmcPlan;  // cmc.cu
struct CooPlan;  // coo.cu

struct Handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  int dim = 0, loc = 0, ngi = 0, n_nodes = 0, n_elements = 0;
  Tables tab{};

  IVec h_nd0;  // 0-based connectivity, stride 4 (4th = -1 on triangles)
  int4* d_ndglno = nullptr;
  double* d_X = nullptr;  // Coordinate%val(dim, n_nodes) as given
  bool have_X = false;
  // packed node records read by the kernels (element_math.cuh NodeRecs)
  double4* d_rec0 = nullptr;
  double4* d_rec1 = nullptr;
  double4* d_rec2 = nullptr;
  double4* d_rec3 = nullptr;  // { X, buoyancy }: with rec1 everything the strip momentum loop reads
  double2* d_rec4 = nullptr;  // { tracer absorption, tracer source }: made when one of them is first set
  double4* d_rec5 = nullptr;  // { momentum absorption, hb_density } and
  double4* d_rec6 = nullptr;  // { momentum source, - }: read by the additive STRIP pass (strip_extra.cu), made on demand
  std::vector<double> h_X;  // kept for locality ordering of the tile plan
  // Node records in the order of the STRIP row blocks (Morton), made when the caller's numbering is scattered: the staging
  // gathers of a row block then read a few contiguous runs instead of ~370 random sectors. d_perm[node] = position;
  // s_rec[k] = what the staged kernels read (the permuted mirror of d_rec<k>, or d_rec<k> itself when d_perm is null).
  int* d_perm = nullptr;
  void* d_prec[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  // node -> element adjacency (host)
  I64Vec n2e_ptr;
  IVec n2e;
  // lexicographic key of every node's lattice point (made with the row blocks, gather.cu): orders the rows of a block,
  // the block's staged node list and the labels of the strip builder independently of the caller's numbering
  std::vector<int64_t> geokey;

  // sparsity, 0-based
  bool have_sparsity = false;
  int nnz = 0;
  IVec h_findrm, h_colm;
  int* d_findrm = nullptr;
  int* d_colm = nullptr;

  // colouring, 0-based
  int ncolours = 0;
  std::vector<int> h_colour_ptr, h_colour_elements;
  int* d_colour_elements = nullptr;

  DeviceField fields[CGASM_F_NSLOTS];

  // results
  double* d_big_m = nullptr;     // [dim][nnz]
  double* d_mom_rhs = nullptr;   // (dim, n_nodes)
  double* d_masslump = nullptr;  // (dim, n_nodes)
  double* d_ct_m = nullptr;      // [dim][nnz]
  double* d_mass = nullptr;      // [dim][nnz] the `mass` matrix (assemble_mass_matrix)
  double* d_mass_rhs = nullptr;  // scratch of the pass that makes it
  double* d_adv_matrix = nullptr;  // [nnz]
  double* d_adv_rhs = nullptr;     // (n_nodes)
  bool mom_has_masslump = false, mom_has_ct = false, mom_valid = false, adv_valid = false;
  bool mom_identical_blocks = false;
  bool mom_has_mass = false;

  int scatter = CGASM_SCATTER_ATOMIC;
  TilePlan* tiles = nullptr;
  GatherPlan* gather = nullptr;
  HaloPlan* halo = nullptr;
  SurfacePlan* surface = nullptr;
  CmcPlan* cmc = nullptr;
  CooPlan* coo[2] = {nullptr, nullptr};  // CGASM_COO_MOMENTUM, CGASM_COO_TRACER

  long long launches = 0;
  int mom_path = 0, adv_path = 0;  // CGASM_PATH_* of the last assembly (cgasm_last_path)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evc0 = nullptr, evc1 = nullptr;  // around cgasm_momentum_advdiff_dev (both loops)
  bool last_combined = false;                   // cgasm_last_kernel_ms reads evc0..evc1 instead of ev0..ev1
  float last_ms = 0.f;
  // asynchronous host flavour (cgasm_set_async): results leave on copy_stream while the next loop runs
  bool async = false;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready = nullptr;       // compute stream -> copy stream: the result is complete
  cudaEvent_t ev_mom_copied = nullptr;  // copy stream -> compute stream: momentum buffers may be overwritten
  cudaEvent_t ev_adv_copied = nullptr;
  bool mom_copy_pending = false, adv_copy_pending = false;
};

// error plumbing
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define CG_CUDA(call)                                                     \
  do {                                                                    \
    cudaError_t _e = (call);                                              \
    if (_e != cudaSuccess) return cgasm::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)
#define CG_FAIL(code, msg)       \
  do {                           \
    cgasm::set_error(msg);       \
    return (code);               \
  } while (0)

// Host -> device copy that has LANDED when it returns. Plain cudaMemcpy does not promise that for pageable sources: it
// returns once the data is staged, the DMA is ordered in the legacy default stream only, and the handles' streams are
// cudaStreamNonBlocking -- a kernel launched there right after the call could read the destination before its tail had
// arrived. Seen at the end of round 2 as a once-in-a-hundred-handles failure of the two-pass GATHER path on the
// cube-parallel fixture (wrong rows at the high node numbers, illegal addresses inside gather_pairs_kernel), never under
// compute-sanitizer or cuda-gdb, which serialise the copies (scripts/stress_surface.py reproduces it in seconds).
inline cudaError_t cg_upload(void* dst, const void* src, size_t bytes) {
  const cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  return e != cudaSuccess ? e : cudaStreamSynchronize(cudaStreamLegacy);
}

Handle* get_handle(int id);

// host_mesh.cpp
void build_node_to_element(int n_nodes, int n_elements, int loc, const int* nd0,
                           I64Vec& ptr, IVec& adj);
int64_t build_sparsity(int n_nodes, int n_elements, int loc, const int* nd0,
                       const I64Vec& n2e_ptr, const IVec& n2e,
                       IVec& findrm, IVec& colm);  // returns nnz (>= 2^31: nothing filled)
int64_t count_nnz(int n_nodes, int loc, const int* nd0, const I64Vec& n2e_ptr,
                  const IVec& n2e);
int greedy_colouring(int n_elements, int loc, const int* nd0, const I64Vec& n2e_ptr,
                     const IVec& n2e, std::vector<int>& colour_of);
void colour_sets(int n_elements, int ncol, const std::vector<int>& colour_of,
                 std::vector<int>& colour_ptr, std::vector<int>& colour_elements);

// ---- Morton order ----------------------------------------------------------------------------
inline uint64_t spread3(uint64_t x) {  // 21 bits -> every third bit
  x &= 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
inline uint64_t spread2(uint64_t x) {  // 32 bits -> every second bit
  x &= 0xffffffffULL;
  x = (x | x << 16) & 0x0000ffff0000ffffULL;
  x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
  x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;
  x = (x | x << 2) & 0x3333333333333333ULL;
  x = (x | x << 1) & 0x5555555555555555ULL;
  return x;
}

struct MortonFrame {
  double lo[3] = {0, 0, 0}, scale[3] = {0, 0, 0};
  int dim = 3;
  // cell of a point on the mesh-spacing lattice (nearest lattice point for nodes)
  inline uint64_t key_round(const double* x) const {
    uint64_t q[3] = {0, 0, 0};
    for (int a = 0; a < dim; a++) q[a] = (uint64_t)std::llround((x[a] - lo[a]) * scale[a]);
    return dim == 3 ? (spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2) : (spread2(q[0]) | spread2(q[1]) << 1);
  }
  // containing cell (floor): element centroids of one lattice cell share a key
  inline uint64_t key_floor(const double* x) const {
    uint64_t q[3] = {0, 0, 0};
    for (int a = 0; a < dim; a++) q[a] = (uint64_t)std::max(0.0, std::floor((x[a] - lo[a]) * scale[a]));
    return dim == 3 ? (spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2) : (spread2(q[0]) | spread2(q[1]) << 1);
  }
};

void morton_order(const Handle* h, std::vector<int>& order, MortonFrame& F);
// Cuts the Morton sequence into blocks of at most block_rows rows (rows: nblocks * block_rows node ids,
// -1 = padding at the end of a block); returns the number of blocks. Needs the sparsity.
int form_row_blocks(const Handle* h, const std::vector<int>& order, const MortonFrame& F, int block_rows,
                    std::vector<int>& rows);

// geometric keys of the nodes (h->geokey) and the order of the rows inside every block (host_mesh.cpp)
void order_block_rows(Handle* h, const MortonFrame& F, std::vector<int>& rows, int nblocks);

// tiled.cu
int tiles_build(Handle* h);
void tiles_free(Handle* h);
int tiles_momentum(Handle* h, const MomentumArgs& args, bool want_ml, bool want_ct);
int tiles_advdiff(Handle* h, const AdvDiffArgs& args);

// gather.cu
int gather_build(Handle* h);        // rows + pairs/walk plans (GATHER)
int gather_build_rows(Handle* h);   // row blocks only
int gather_build_pairs(Handle* h);  // pair lists + walk plan, on demand
void gather_free(Handle* h);
int gather_momentum(Handle* h, const MomentumArgs& args, bool want_ml, bool want_ct);
int gather_advdiff(Handle* h, const AdvDiffArgs& args);

// strip_extra.cu: the additive momentum pass (absorption, sources, reference profile; constant density)
bool strip_extra_needed(const MomentumArgs& args);
bool strip_extra_ok(const Handle* h, const MomentumArgs& args);
int strip_extra(Handle* h, const MomentumArgs& args, bool skip_full_absorption = false);
// strip_absorb.cu: the common STRIP momentum kernel with the full absorption matrix in the same pass
bool strip_absorb_ok(const Handle* h, const MomentumArgs& args);
int strip_absorb_momentum(Handle* h, const MomentumArgs& args);
int ensure_extra_records(Handle* h);  // cgasm_api.cu
// record k (0..6) as the staged kernels read it; refresh of the permuted mirrors after record k changed (all nodes, or the
// listed ones) on `stream`
const void* staged_rec(const Handle* h, int k);
int refresh_permuted(Handle* h, unsigned recmask, const int* d_nodes, int n, cudaStream_t stream);
int set_permutation(Handle* h, const std::vector<int>& perm);  // empty = none
void* rec_array(const Handle* h, int k);
int rec_width(int k);
unsigned slot_record_mask(int slot);  // bit k = record k mirrors the slot (-1 = coordinates)

// strip_fused.cu: both element loops in one kernel (common STRIP option sets)
bool strip_fused_ok(const Handle* h, const MomentumArgs& m, const AdvDiffArgs& a);
int strip_fused(Handle* h, const MomentumArgs& m, const AdvDiffArgs& a);

// strip.cu
int strip_build(Handle* h);
void strip_free(GatherPlan* P);
bool strip_momentum_ok(const Handle* h, const MomentumArgs& args, bool want_ml);
bool strip_advdiff_ok(const Handle* h, const AdvDiffArgs& args);
int strip_momentum(Handle* h, const MomentumArgs& args);
int strip_advdiff(Handle* h, const AdvDiffArgs& args);

// halo.cu
void halo_free(Handle* h);

// surface.cu
void surface_free(Handle* h);

// cmc.cu
void cmc_free(Handle* h);

// coo.cu
void coo_free(Handle* h);

// cgasm_api.cu: refresh the packed record lanes fed by `slot` (-1 = coordinates); nodes == nullptr
// repacks every node, else only the listed ones (device array of 0-based node ids).
int repack_slot(Handle* h, int slot, const int* d_nodes, int n);
// packed-record lanes that mirror a field slot (cgasm_api.cu); returns the number of targets, -1 on allocation failure
int record_targets(Handle* h, int slot, double* rec[2], int recw[2], int lane0[2], int* ncomp);

// halo.cu: overlap of the halo exchange with the assembly
int halo_join(Handle* h);  // compute stream waits for a pending exchange (no-op if none)
// if an exchange is pending and the STRIP block split exists: the two block lists (device) and their lengths; else false
bool halo_split(Handle* h, const int** indep, int* n_indep, const int** dep, int* n_dep);

}  // namespace cgasm
