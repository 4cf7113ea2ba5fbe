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
};

void halo_free(Handle* h) {
  HaloPlan* p = h->halo;
  if (!p) return;
  if (p->comm && g_nccl.ok) g_nccl.CommDestroy(p->comm);
  int* ints[] = {p->d_send_node, p->d_send_base, p->d_send_cnt, p->d_send_kk,
                 p->d_recv_node, p->d_recv_base, p->d_recv_cnt, p->d_recv_kk};
  for (int* q : ints)
    if (q) cudaFree(q);
  if (p->d_send_stage) cudaFree(p->d_send_stage);
  if (p->d_recv_stage) cudaFree(p->d_recv_stage);
  delete p;
  h->halo = nullptr;
}

// stage[comps_total*base + comps_prefix*cnt + kk*comps + c] <-> field[node*comps + c]
template <bool PACK>
__global__ void halo_pack_kernel(int n_entries, int comps, int comps_prefix, int comps_total,
                                 const int* __restrict__ node, const int* __restrict__ base,
                                 const int* __restrict__ cnt, const int* __restrict__ kk,
                                 double* __restrict__ field, double* __restrict__ stage) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_entries * comps) return;
  const int k = tid / comps, c = tid - k * comps;
  const size_t s = (size_t)comps_total * base[k] + (size_t)comps_prefix * cnt[k] + (size_t)kk[k] * comps + c;
  const size_t f = (size_t)node[k] * comps + c;
  if (PACK) stage[s] = field[f];
  else field[f] = stage[s];
}

static int upload_ints(int** d, const std::vector<int>& v) {
  CG_CUDA(cudaMalloc(d, sizeof(int) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CG_CUDA(cudaMemcpy(*d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
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
  HaloPlan* p = new HaloPlan();
  h->halo = p;
  p->nprocs = nprocs;
  p->rank = rank;
  p->nsend.assign(nsend, nsend + nprocs);
  p->nrecv.assign(nrecv, nrecv + nprocs);
  p->send_off.assign(nprocs + 1, 0);
  p->recv_off.assign(nprocs + 1, 0);
  for (int q = 0; q < nprocs; q++) {
    if (nsend[q] < 0 || nrecv[q] < 0) CG_FAIL(CGASM_EARG, "negative halo count");
    p->send_off[q + 1] = p->send_off[q] + nsend[q];
    p->recv_off[q + 1] = p->recv_off[q] + nrecv[q];
  }
  p->total_send = p->send_off[nprocs];
  p->total_recv = p->recv_off[nprocs];
  if ((p->total_send && !sends) || (p->total_recv && !recvs)) CG_FAIL(CGASM_EARG, "null halo node list");
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
  if (st) return st;
  if ((st = upload_ints(&p->d_send_node, node)) || (st = upload_ints(&p->d_send_base, base)) ||
      (st = upload_ints(&p->d_send_cnt, c)) || (st = upload_ints(&p->d_send_kk, kk)))
    return st;
  node.clear(); base.clear(); c.clear(); kk.clear();
  if ((st = expand(p->nrecv, p->recv_off, recvs, node, base, c, kk))) return st;
  if ((st = upload_ints(&p->d_recv_node, node)) || (st = upload_ints(&p->d_recv_base, base)) ||
      (st = upload_ints(&p->d_recv_cnt, c)) || (st = upload_ints(&p->d_recv_kk, kk)))
    return st;
  if (nprocs > 1) {
    if (!nccl_unique_id) CG_FAIL(CGASM_EARG, "null nccl_unique_id");
    if ((st = nccl_load())) return st;
    ncclUniqueId uid;
    memcpy(&uid, nccl_unique_id, sizeof uid);
    CG_NCCL(g_nccl.CommInitRank(&p->comm, nprocs, uid, rank));
  }
  return CGASM_OK;
}

int cgasm_halo_update(int id, unsigned slot_mask) {
  Handle* h = get_handle(id);
  if (!h) CG_FAIL(CGASM_EHANDLE, "unknown cgasm handle");
  CG_CUDA(cudaSetDevice(h->device));
  HaloPlan* p = h->halo;
  if (!p) CG_FAIL(CGASM_ESTATE, "cgasm_halo_create has not been called");
  // fields taking part: set, NORMAL (a CONSTANT field has nothing to exchange)
  int slots[CGASM_F_NSLOTS], comps[CGASM_F_NSLOTS], prefix[CGASM_F_NSLOTS], ns = 0, total = 0;
  for (int s = 0; s < CGASM_F_NSLOTS; s++) {
    if (!(slot_mask & (1u << s))) continue;
    const DeviceField& f = h->fields[s];
    if (!f.set) CG_FAIL(CGASM_ESTATE, "halo_update of a field slot that is not set");
    if (f.field_type != CGASM_FIELD_NORMAL) continue;
    int cpn = 1;
    for (int r = 0; r < f.rank; r++) cpn *= h->dim;
    slots[ns] = s;
    comps[ns] = cpn;
    prefix[ns] = total;
    total += cpn;
    ns++;
  }
  if (!ns || p->nprocs == 1) return CGASM_OK;
  if ((size_t)total > p->stage_comps) {
    if (p->d_send_stage) cudaFree(p->d_send_stage);
    if (p->d_recv_stage) cudaFree(p->d_recv_stage);
    p->d_send_stage = p->d_recv_stage = nullptr;
    CG_CUDA(cudaMalloc(&p->d_send_stage, sizeof(double) * std::max<size_t>((size_t)total * p->total_send, 1)));
    CG_CUDA(cudaMalloc(&p->d_recv_stage, sizeof(double) * std::max<size_t>((size_t)total * p->total_recv, 1)));
    p->stage_comps = (size_t)total;
  }
  const int block = 256;
  for (int i = 0; i < ns && p->total_send; i++) {
    const int n = p->total_send * comps[i];
    halo_pack_kernel<true><<<(n + block - 1) / block, block, 0, h->stream>>>(
        p->total_send, comps[i], prefix[i], total, p->d_send_node, p->d_send_base, p->d_send_cnt,
        p->d_send_kk, h->fields[slots[i]].d, p->d_send_stage);
    h->launches++;
  }
  CG_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < p->nprocs; q++) {
    if (p->nsend[q])
      CG_NCCL(g_nccl.Send(p->d_send_stage + (size_t)total * p->send_off[q], (size_t)total * p->nsend[q],
                          ncclDouble, q, p->comm, h->stream));
    if (p->nrecv[q])
      CG_NCCL(g_nccl.Recv(p->d_recv_stage + (size_t)total * p->recv_off[q], (size_t)total * p->nrecv[q],
                          ncclDouble, q, p->comm, h->stream));
  }
  CG_NCCL(g_nccl.GroupEnd());
  for (int i = 0; i < ns && p->total_recv; i++) {
    const int n = p->total_recv * comps[i];
    halo_pack_kernel<false><<<(n + block - 1) / block, block, 0, h->stream>>>(
        p->total_recv, comps[i], prefix[i], total, p->d_recv_node, p->d_recv_base, p->d_recv_cnt,
        p->d_recv_kk, h->fields[slots[i]].d, p->d_recv_stage);
    h->launches++;
  }
  // the kernels read packed node records: refresh the received nodes of the packed fields
  for (int i = 0; i < ns && p->total_recv; i++) {
    int st = repack_slot(h, slots[i], p->d_recv_node, p->total_recv);
    if (st) return st;
  }
  CG_CUDA(cudaGetLastError());
  return CGASM_OK;
}

}  // extern "C"
