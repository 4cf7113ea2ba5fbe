// halo.cu -- halo_update of device-resident fields over NCCL point-to-point.
//
// Replaces halo_update -> halo_update_array_real_star (femtools/Halos_Communications.F90
// :497-567, :320-412): there, per neighbour, an MPI indexed-block datatype built from
// halo%sends(p) / halo%receives(p) is sent/received in place on field%val, block_size reals
// per node. Here: one pack kernel gathers every requested field's send nodes into a
// per-neighbour contiguous staging segment, ONE ncclGroup carries one ncclSend + one ncclRecv
// per neighbour for all fields together (NVLink 5 / NVSwitch: latency-bound, so fewer, larger
// messages), and one unpack kernel scatters into the receive nodes. Everything is queued on the
// handle's stream, so the following assembly kernels are ordered after it without host sync.
//
// NCCL is resolved with dlopen at first use so that libcgasm.so has no link-time NCCL
// dependency (a process that already loaded a libnccl.so.2, e.g. through torch, shares it).
#include "cgasm_internal.h"
#include "gather_plan.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <dlfcn.h>
#include <nccl.h>

namespace cgasm {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.ok) return CGASM_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) CG_FAIL(CGASM_ENCCL, std::string("cannot dlopen libnccl: ") + dlerror());
#define SYM(field, name)                                                     \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                        \
  if (!g_nccl.field) CG_FAIL(CGASM_ENCCL, std::string("libnccl lacks ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.ok = true;
  return CGASM_OK;
}

#define CG_NCCL(call)                                                                  \
  do {                                                                                 \
    ncclResult_t _r = (call);                                                          \
    if (_r != ncclSuccess) CG_FAIL(CGASM_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

struct HaloPlan {
  int nprocs = 0, rank = 0;
  ncclComm_t comm = nullptr;
  std::vector<int> nsend, nrecv, send_off, recv_off;  // per process
  int total_send = 0, total_recv = 0;
  // per entry: node (0-based), owning peer's segment offset / count, index inside the segment
  int *d_send_node = nullptr, *d_send_base = nullptr, *d_send_cnt = nullptr, *d_send_kk = nullptr;
  int *d_recv_node = nullptr, *d_recv_base = nullptr, *d_recv_cnt = nullptr, *d_recv_kk = nullptr;
  double* d_send_stage = nullptr;
  double* d_recv_stage = nullptr;
  size_t stage_comps = 0;  // components per node the staging buffers are sized for
  // overlap with the assembly (cgasm_halo_set_overlap): the exchange runs on its own stream; the STRIP row blocks
  // that read no received node are launched before the compute stream waits for it
  bool overlap = false, pending = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_inputs = nullptr, ev_done = nullptr;
  long long split_serial = 0;  // serial of the GatherPlan the block lists were made for
  int* d_blocks_indep = nullptr;
  int* d_blocks_dep = nullptr;
  int n_indep = 0, n_dep = 0;
};

void halo_free(Handle* h) {
  HaloPlan* p = h->halo;
  if (!p) return;
  if (p->comm_stream) cudaStreamSynchronize(p->comm_stream);
  if (p->comm && g_nccl.ok) g_nccl.CommDestroy(p->comm);
  int* ints[] = {p->d_send_node, p->d_send_base, p->d_send_cnt, p->d_send_kk, p->d_recv_node,
                 p->d_recv_base, p->d_recv_cnt,  p->d_recv_kk,  p->d_blocks_indep, p->d_blocks_dep};
  for (int* q : ints)
    if (q) cudaFree(q);
  if (p->d_send_stage) cudaFree(p->d_send_stage);
  if (p->d_recv_stage) cudaFree(p->d_recv_stage);
  if (p->ev_inputs) cudaEventDestroy(p->ev_inputs);
  if (p->ev_done) cudaEventDestroy(p->ev_done);
  if (p->comm_stream) cudaStreamDestroy(p->comm_stream);
  delete p;
  h->halo = nullptr;
}

// All fields of one update in one launch (Halos_Communications.F90:320-412 exchanges one field per call; the
// reference's callers update several fields back to back before an assembly -- here they share one message per
// neighbour and one pack / one unpack kernel).
//   stage[total*base + prefix_i*cnt + kk*comps_i + c] <-> field_i[node*comps_i + c]
// The unpack side also refreshes the packed node records the kernels read (repack of the received nodes only).
constexpr int kHaloMaxFields = CGASM_F_NSLOTS;
struct HaloFields {
  int ns = 0, total = 0;
  int comps[kHaloMaxFields], prefix[kHaloMaxFields];
  double* field[kHaloMaxFields];
  // up to two record arrays mirror a field: rec + recw * node + lane0 + c  (recw = doubles per record)
  double* rec[kHaloMaxFields][2];
  int recw[kHaloMaxFields][2], lane0[kHaloMaxFields][2];
};

template <bool PACK>
__global__ void halo_fields_kernel(const HaloFields F, int n_entries, const int* __restrict__ node,
                                   const int* __restrict__ base, const int* __restrict__ cnt,
                                   const int* __restrict__ kk, double* __restrict__ stage) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_entries * F.total) return;
  const int k = tid / F.total;
  int c = tid - k * F.total, i = 0;
  while (i + 1 < F.ns && c >= F.prefix[i + 1]) i++;
  c -= F.prefix[i];
  const int comps = F.comps[i], nd = node[k];
  const size_t s = (size_t)F.total * base[k] + (size_t)F.prefix[i] * cnt[k] + (size_t)kk[k] * comps + c;
  const size_t f = (size_t)nd * comps + c;
  if (PACK) {
    stage[s] = F.field[i][f];
  } else {
    const double v = stage[s];
    F.field[i][f] = v;
#pragma unroll
    for (int m = 0; m < 2; m++)
      if (F.rec[i][m]) F.rec[i][m][(size_t)F.recw[i][m] * nd + F.lane0[i][m] + c] = v;
  }
}

// dep[b] = 1 if block b's staged node list holds a marked (received) node
__global__ void halo_block_dep_kernel(int nblocks, int nl, const int* __restrict__ blk_nodes,
                                      const unsigned char* __restrict__ mark, unsigned char* __restrict__ dep) {
  const int b = blockIdx.x;
  int any = 0;
  for (int i = threadIdx.x; i < nl; i += blockDim.x) {
    const int nd = blk_nodes[(size_t)b * nl + i];
    if (nd >= 0 && mark[nd]) any = 1;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) dep[b] = (unsigned char)any;
}

// (perm: the block node lists hold record positions when the records are permuted)
__global__ void halo_mark_kernel(int n, const int* __restrict__ node, const int* __restrict__ perm,
                                 unsigned char* __restrict__ mark) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) mark[perm ? perm[node[k]] : node[k]] = 1;
}

static int upload_ints(int** d, const std::vector<int>& v) {
  CG_CUDA(cudaMalloc(d, sizeof(int) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CG_CUDA(cg_upload(*d, v.data(), sizeof(int) * v.size()));
  return CGASM_OK;
}

}  // namespace cgasm

using namespace cgasm;

extern "C" {

int cgasm_nccl_unique_id(void* out128) {
  if (!out128) CG_FAIL(CGASM_EARG, "null out");
  int st = nccl_load();
  if (st) return st;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId uid;
  CG_NCCL(g_nccl.GetUniqueId(&uid));
  memcpy(out128, &uid, sizeof uid);
  return CGASM_OK;
}

int cgasm_halo_create(int id, int nprocs, int rank, const int* nsend, const int* sends,
                      const int* nrecv, const int* recvs, const void* nccl_unique_id) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  if (nprocs < 1 || rank < 0 || rank >= nprocs || !nsend || !nrecv) CG_FAIL(CGASM_EARG, "bad halo arguments");
  if (nsend[rank] != 0 || nrecv[rank] != 0) CG_FAIL(CGASM_EARG, "a process does not send to itself");
  halo_free(h);
  // Built into a local plan; the handle gets it only when every check, upload and the communicator succeeded
  // (a half-built plan would make a later cgasm_halo_update launch kernels on null index arrays).
  HaloPlan* p = new HaloPlan();
  auto fail = [&](int code) {
    h->halo = p;
    halo_free(h);
    return code;
  };
  p->nprocs = nprocs;
  p->rank = rank;
  p->nsend.assign(nsend, nsend + nprocs);
  p->nrecv.assign(nrecv, nrecv + nprocs);
  p->send_off.assign(nprocs + 1, 0);
  p->recv_off.assign(nprocs + 1, 0);
  for (int q = 0; q < nprocs; q++) {
    if (nsend[q] < 0 || nrecv[q] < 0) {
      set_error("negative halo count");
      return fail(CGASM_EARG);
    }
    p->send_off[q + 1] = p->send_off[q] + nsend[q];
    p->recv_off[q + 1] = p->recv_off[q] + nrecv[q];
  }
  p->total_send = p->send_off[nprocs];
  p->total_recv = p->recv_off[nprocs];
  if ((p->total_send && !sends) || (p->total_recv && !recvs)) {
    set_error("null halo node list");
    return fail(CGASM_EARG);
  }
  auto expand = [&](const std::vector<int>& cnt, const std::vector<int>& off, const int* nodes,
                    std::vector<int>& node, std::vector<int>& base, std::vector<int>& c,
                    std::vector<int>& kk) -> int {
    for (int q = 0; q < nprocs; q++)
      for (int k = 0; k < cnt[q]; k++) {
        const int nd = nodes[off[q] + k] - 1;
        if (nd < 0 || nd >= h->n_nodes) CG_FAIL(CGASM_EARG, "halo node out of range (expects 1-based)");
        node.push_back(nd);
        base.push_back(off[q]);
        c.push_back(cnt[q]);
        kk.push_back(k);
      }
    return CGASM_OK;
  };
  std::vector<int> node, base, c, kk;
  int st = expand(p->nsend, p->send_off, sends, node, base, c, kk);
  if (st) return fail(st);
  if ((st = upload_ints(&p->d_send_node, node)) || (st = upload_ints(&p->d_send_base, base)) ||
      (st = upload_ints(&p->d_send_cnt, c)) || (st = upload_ints(&p->d_send_kk, kk)))
    return fail(st);
  node.clear(); base.clear(); c.clear(); kk.clear();
  if ((st = expand(p->nrecv, p->recv_off, recvs, node, base, c, kk))) return fail(st);
  if ((st = upload_ints(&p->d_recv_node, node)) || (st = upload_ints(&p->d_recv_base, base)) ||
      (st = upload_ints(&p->d_recv_cnt, c)) || (st = upload_ints(&p->d_recv_kk, kk)))
    return fail(st);
  if (nprocs > 1) {
    if (!nccl_unique_id) {
      set_error("null nccl_unique_id");
      return fail(CGASM_EARG);
    }
    if ((st = nccl_load())) return fail(st);
    ncclUniqueId uid;
    memcpy(&uid, nccl_unique_id, sizeof uid);
    const ncclResult_t r = g_nccl.CommInitRank(&p->comm, nprocs, uid, rank);
    if (r != ncclSuccess) {
      p->comm = nullptr;
      set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
      return fail(CGASM_ENCCL);
    }
  }
  h->halo = p;
  return CGASM_OK;
}

int cgasm_halo_set_overlap(int id, int on) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  HaloPlan* p = h->halo;
  if (!p) CG_FAIL(CGASM_ESTATE, "cgasm_halo_create has not been called");
  int st = halo_join(h);
  if (st) return st;
  if (on && !p->comm_stream) {
    // the exchange should win SMs over the bulk kernels it overlaps: highest priority
    int lo = 0, hi = 0;
    CG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CG_CUDA(cudaStreamCreateWithPriority(&p->comm_stream, cudaStreamNonBlocking, hi));
    CG_CUDA(cudaEventCreateWithFlags(&p->ev_inputs, cudaEventDisableTiming));
    CG_CUDA(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
  }
  p->overlap = on != 0;
  return CGASM_OK;
}

int cgasm_halo_update(int id, unsigned slot_mask) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  HaloPlan* p = h->halo;
  if (!p) CG_FAIL(CGASM_ESTATE, "cgasm_halo_create has not been called");
  int st = halo_join(h);  // a previous exchange nobody waited for yet
  if (st) return st;
  // fields taking part: set, NORMAL (a CONSTANT field has nothing to exchange)
  HaloFields F;
  for (int s = 0; s < CGASM_F_NSLOTS; s++) {
    if (!(slot_mask & (1u << s))) continue;
    const DeviceField& f = h->fields[s];
    if (!f.set) CG_FAIL(CGASM_ESTATE, "halo_update of a field slot that is not set");
    if (f.field_type != CGASM_FIELD_NORMAL) continue;
    int cpn = 1;
    for (int r = 0; r < f.rank; r++) cpn *= h->dim;
    const int i = F.ns++;
    F.comps[i] = cpn;
    F.prefix[i] = F.total;
    F.field[i] = f.d;
    F.total += cpn;
    int ncomp = 0;
    const int nt = record_targets(h, s, F.rec[i], F.recw[i], F.lane0[i], &ncomp);
    if (nt < 0) CG_FAIL(CGASM_ECUDA, "cannot allocate the tracer absorption / source records");
    for (int m = nt; m < 2; m++) F.rec[i][m] = nullptr;
  }
  if (!F.ns || p->nprocs == 1) return CGASM_OK;
  const int total = F.total;
  if ((size_t)total > p->stage_comps) {
    if (p->d_send_stage) cudaFree(p->d_send_stage);
    if (p->d_recv_stage) cudaFree(p->d_recv_stage);
    p->d_send_stage = p->d_recv_stage = nullptr;
    p->stage_comps = 0;
    CG_CUDA(cudaMalloc(&p->d_send_stage, sizeof(double) * std::max<size_t>((size_t)total * p->total_send, 1)));
    CG_CUDA(cudaMalloc(&p->d_recv_stage, sizeof(double) * std::max<size_t>((size_t)total * p->total_recv, 1)));
    p->stage_comps = (size_t)total;
  }
  cudaStream_t cs = h->stream;
  if (p->overlap) {
    // the fields (and the results of the previous assembly that read them) belong to the compute stream
    CG_CUDA(cudaEventRecord(p->ev_inputs, h->stream));
    CG_CUDA(cudaStreamWaitEvent(p->comm_stream, p->ev_inputs, 0));
    cs = p->comm_stream;
  }
  const int block = 256;
  if (p->total_send) {
    const long long n = (long long)p->total_send * total;
    halo_fields_kernel<true><<<(unsigned)((n + block - 1) / block), block, 0, cs>>>(
        F, p->total_send, p->d_send_node, p->d_send_base, p->d_send_cnt, p->d_send_kk, p->d_send_stage);
    h->launches++;
  }
  CG_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < p->nprocs; q++) {
    if (p->nsend[q])
      CG_NCCL(g_nccl.Send(p->d_send_stage + (size_t)total * p->send_off[q], (size_t)total * p->nsend[q],
                          ncclDouble, q, p->comm, cs));
    if (p->nrecv[q])
      CG_NCCL(g_nccl.Recv(p->d_recv_stage + (size_t)total * p->recv_off[q], (size_t)total * p->nrecv[q],
                          ncclDouble, q, p->comm, cs));
  }
  CG_NCCL(g_nccl.GroupEnd());
  if (p->total_recv) {
    const long long n = (long long)p->total_recv * total;
    halo_fields_kernel<false><<<(unsigned)((n + block - 1) / block), block, 0, cs>>>(
        F, p->total_recv, p->d_recv_node, p->d_recv_base, p->d_recv_cnt, p->d_recv_kk, p->d_recv_stage);
    h->launches++;
  }
  CG_CUDA(cudaGetLastError());
  if (h->d_perm && p->total_recv) {  // the staged kernels read the permuted mirrors of the records just refreshed
    unsigned recmask = 0;
    for (int s = 0; s < CGASM_F_NSLOTS; s++)
      if ((slot_mask & (1u << s)) && h->fields[s].set && h->fields[s].field_type == CGASM_FIELD_NORMAL) recmask |= slot_record_mask(s);
    if ((st = refresh_permuted(h, recmask, p->d_recv_node, p->total_recv, cs))) return st;
  }
  if (p->overlap) {
    CG_CUDA(cudaEventRecord(p->ev_done, cs));
    p->pending = true;
  }
  return CGASM_OK;
}

}  // extern "C"

namespace cgasm {

int halo_join(Handle* h) {
  HaloPlan* p = h->halo;
  if (!p || !p->pending) return CGASM_OK;
  CG_CUDA(cudaStreamWaitEvent(h->stream, p->ev_done, 0));
  p->pending = false;
  return CGASM_OK;
}

// Splits the STRIP row blocks into those that read a received node and those that do not (once per plan).
static int halo_build_split(Handle* h) {
  HaloPlan* p = h->halo;
  GatherPlan* P = h->gather;
  if (p->d_blocks_indep) cudaFree(p->d_blocks_indep);
  if (p->d_blocks_dep) cudaFree(p->d_blocks_dep);
  p->d_blocks_indep = p->d_blocks_dep = nullptr;
  p->n_indep = p->n_dep = 0;
  p->split_serial = P->serial;
  const int nb = P->nblocks;
  unsigned char *d_mark = nullptr, *d_dep = nullptr;
  CG_CUDA(cudaMalloc(&d_mark, (size_t)h->n_nodes));
  CG_CUDA(cudaMalloc(&d_dep, (size_t)std::max(nb, 1)));
  CG_CUDA(cudaMemsetAsync(d_mark, 0, (size_t)h->n_nodes, h->stream));
  if (p->total_recv) halo_mark_kernel<<<(p->total_recv + 255) / 256, 256, 0, h->stream>>>(p->total_recv, p->d_recv_node, h->d_perm, d_mark);
  halo_block_dep_kernel<<<nb, 128, 0, h->stream>>>(nb, P->nl, P->d_blk_nodes, d_mark, d_dep);
  h->launches += 2;
  std::vector<unsigned char> dep((size_t)nb);
  cudaError_t e = cudaMemcpyAsync(dep.data(), d_dep, (size_t)nb, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_mark);
  cudaFree(d_dep);
  CG_CUDA(e);
  std::vector<int> a, b;
  for (int q = 0; q < nb; q++) (dep[q] ? b : a).push_back(q);
  p->n_indep = (int)a.size();
  p->n_dep = (int)b.size();
  int st;
  if ((st = upload_ints(&p->d_blocks_indep, a)) || (st = upload_ints(&p->d_blocks_dep, b))) return st;
  if (getenv("CGASM_DEBUG"))
    fprintf(stderr, "[cgasm] halo overlap: %d of %d row blocks read no received node\n", p->n_indep, nb);
  return CGASM_OK;
}

bool halo_split(Handle* h, const int** indep, int* n_indep, const int** dep, int* n_dep) {
  HaloPlan* p = h->halo;
  GatherPlan* P = h->gather;
  if (!p || !p->pending || !P || !P->staged_ok || !P->d_blk_nodes) return false;
  if (p->split_serial != P->serial && halo_build_split(h) != CGASM_OK) return false;
  if (!p->n_indep) return false;
  *indep = p->d_blocks_indep;
  *n_indep = p->n_indep;
  *dep = p->d_blocks_dep;
  *n_dep = p->n_dep;
  return true;
}

}  // namespace cgasm
