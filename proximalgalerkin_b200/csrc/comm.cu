// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch for the two exchanges this path
// has -- the owner->ghost halo of a mixed vector (Vec.ghostUpdate(INSERT, FORWARD),
// src/lvpp/problem.py:56,58,71,73) and the scalar all-reduces of norms / Krylov dot products
// (PETSc VecNorm/VecDot and comm.allreduce at examples/01_obstacle_problem/obstacle_pg.py:50).
// Owned rows are assembled completely from owned + ghost cells, so neither the residual reverse
// scatter (problem.py:66) nor Mat.assemble() (problem.py:77) needs communication.
//
// NCCL is resolved with dlopen at lvpp_comm_init time so that the single-GPU library has no link
// dependency on it (the process usually has torch's libnccl.so.2 loaded already).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "lvpp_internal.cuh"

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    lvpp_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return LVPP_E_COMM;
  }
#define SYM(field, name)                                                    \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                       \
  if (!g_nccl.field) { lvpp_set_error("missing NCCL symbol %s", name); g_nccl.lib = nullptr; return LVPP_E_COMM; }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
}  // namespace

#define NCK(call)                                                                                  \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) {                                                                       \
      lvpp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));      \
      return LVPP_E_COMM;                                                                          \
    }                                                                                              \
  } while (0)

__global__ void k_pack(int64_t n, const int32_t* __restrict__ nodes, const double2* __restrict__ v,
                       double2* __restrict__ buf) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    buf[p] = v[nodes[p]];
}
__global__ void k_unpack(int64_t n, const int32_t* __restrict__ nodes, const double2* __restrict__ buf,
                         double2* __restrict__ v) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    v[nodes[p]] = buf[p];
}

extern "C" int lvpp_comm_unique_id(uint8_t* h_id128) {
  if (!h_id128) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CKR(load_nccl());
  ncclUniqueId id;
  NCK(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(h_id128, &id, 128);
  return LVPP_OK;
}

extern "C" int lvpp_comm_init(lvpp_handle h, const uint8_t* h_id128, int32_t rank, int32_t nranks) {
  if (!h || !h_id128 || nranks < 1 || rank < 0 || rank >= nranks) { lvpp_set_error("bad argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  CKR(load_nccl());
  ncclUniqueId id;
  memcpy(&id, h_id128, 128);
  ncclComm_t comm;
  NCK(g_nccl.CommInitRank(&comm, nranks, id, rank));
  h->nccl_comm = (void*)comm;
  h->rank = rank;
  h->nranks = nranks;
  // global row count
  double* tmp = h->scal->red;
  double rows = 2.0 * (double)h->Vown;
  CK(cudaMemcpyAsync(tmp, &rows, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CKR(lvpp_allreduce_sum(h, tmp, 1));
  CK(cudaMemcpyAsync(&rows, tmp, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->global_rows = (int64_t)(rows + 0.5);
  CKR(lvpp_halo_p2p_setup(h, h->halo));
  return LVPP_OK;
}

static void halo_release(LevelHalo& H) {
  for (void* b : H.peer_base) cudaIpcCloseMemHandle(b);
  H.peer_base.clear();
  if (H.arena) cudaFree(H.arena);
  H.arena = nullptr;
  H.p2p = false;
}

void lvpp_comm_destroy(lvpp_problem* h) {
  halo_release(h->halo);
  for (MgLevel& L : h->levels) halo_release(L.halo);
  if (h->p2p_err) { cudaFreeHost(h->p2p_err); h->p2p_err = nullptr; }
  if (h->nccl_comm && g_nccl.CommDestroy) {
    g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr;
  }
}

int lvpp_allreduce_sum(lvpp_problem* h, double* d_buf, int n) {
  if (h->nranks <= 1) return 0;
  if (!h->nccl_comm) { lvpp_set_error("communicator not initialised"); return LVPP_E_COMM; }
  NCK(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  return 0;
}

// ---- peer-memory halo (NVLink / NVSwitch, no NCCL call on the solve path) ---------------------------
// ONE kernel per exchange.  Phase 1 (push): gather the send lists and store them straight into the neighbours'
// receive buffers over NVLink; the last block to finish raises every neighbour's flag to `seq` (system-scope
// release).  Phase 2 (pull): wait until every neighbour has raised this rank's flags to `seq` (system-scope
// acquire), then scatter the receive buffer (read past L1: the lines were written by a peer) into the ghost
// entries.  The grid never exceeds one block per SM, so every block is resident and the wait of phase 2 cannot
// starve a phase-1 block of this or of the neighbour's kernel.
#define LVPP_MAX_NEIGHBORS 8
struct HaloPush {
  int nb;
  int64_t send_ptr[LVPP_MAX_NEIGHBORS + 1];
  double2* peer_buf[LVPP_MAX_NEIGHBORS];
  unsigned long long* peer_flag[LVPP_MAX_NEIGHBORS];
};

__global__ void __launch_bounds__(256) k_halo_p2p(HaloPush hp, const int32_t* __restrict__ send_nodes,
                                                   const int32_t* __restrict__ recv_nodes, int64_t nrecv, double2* __restrict__ v,
                                                   const double2* __restrict__ rbuf, volatile unsigned long long* flags,
                                                   int* __restrict__ counter, unsigned long long seq, int* __restrict__ err, long long timeout) {
  const int64_t nsend = hp.send_ptr[hp.nb];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nsend; p += (int64_t)gridDim.x * blockDim.x) {
    int b = 0;
    while (p >= hp.send_ptr[b + 1]) ++b;
    hp.peer_buf[b][p - hp.send_ptr[b]] = v[send_nodes[p]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(counter, 1);
    if (done == (int)gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      for (int b = 0; b < hp.nb; ++b) *(volatile unsigned long long*)hp.peer_flag[b] = seq;
    }
    for (int b = 0; b < hp.nb; ++b) {
      const long long t0 = clock64();
      while (flags[b] < seq)
        if (clock64() - t0 > timeout) {  // a peer died or the call sequences diverged (default ~8 s)
          *err = 1;
          break;
        }
    }
    __threadfence_system();
  }
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nrecv; p += (int64_t)gridDim.x * blockDim.x)
    v[recv_nodes[p]] = __ldcg(&rbuf[p]);
}

struct P2pHello {  // what a rank tells each neighbour about its receive arena
  cudaIpcMemHandle_t handle;
  int64_t seg_off;   // doubles: where this neighbour's data starts inside rbuf[parity]
  int64_t slot;      // flag index of this neighbour
  int64_t nrecv;     // total receive count (doubles / 2) of the arena
  int64_t ok;
};

int lvpp_halo_p2p_setup(lvpp_problem* h, LevelHalo& H) {
  H.p2p = false;
  if (h->nranks <= 1) return 0;
  const char* mode = getenv("LVPP_HALO");
  const bool want = !(mode && strcmp(mode, "nccl") == 0);
  if (const char* to = getenv("LVPP_HALO_TIMEOUT_S")) {  // how long k_halo_p2p waits for a neighbour's flag
    const double sec = atof(to);
    if (sec > 0.0) h->halo_timeout_ticks = (long long)(sec * 1.9e9);  // clock64 runs at the SM clock (<= 1.965 GHz)
  }
  const int nb = H.num_neighbors;
  const int64_t nr = H.recv_ptr.back();
  const size_t flag_bytes = 256, buf_bytes = sizeof(double) * 2 * (size_t)(nr > 0 ? nr : 1);
  int ok = want ? 1 : 0;
  if (ok && !h->p2p_err) {
    if (cudaHostAlloc((void**)&h->p2p_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) ok = 0;
    else *h->p2p_err = 0;
  }
  if (ok && nb > LVPP_MAX_NEIGHBORS) ok = 0;
  P2pHello mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaMalloc(&H.arena, flag_bytes + 2 * buf_bytes) != cudaSuccess) { ok = 0; H.arena = nullptr; cudaGetLastError(); }
    else {
      cudaMemsetAsync(H.arena, 0, flag_bytes + 2 * buf_bytes, h->stream);
      if (cudaIpcGetMemHandle(&mine.handle, H.arena) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    }
  }
  // exchange the hello records with every neighbour (through device staging buffers: NCCL moves device memory)
  P2pHello *d_send = nullptr, *d_recv = nullptr;
  CKR(lvpp_dalloc(h, &d_send, (size_t)nb));
  CKR(lvpp_dalloc(h, &d_recv, (size_t)nb));
  std::vector<P2pHello> hs((size_t)nb), hr((size_t)nb);
  for (int b = 0; b < nb; ++b) {
    hs[b] = mine;
    hs[b].seg_off = 2 * H.recv_ptr[b];
    hs[b].slot = b;
    hs[b].nrecv = nr > 0 ? nr : 1;
    hs[b].ok = ok;
  }
  CK(cudaMemcpyAsync(d_send, hs.data(), sizeof(P2pHello) * nb, cudaMemcpyHostToDevice, h->stream));
  NCK(g_nccl.GroupStart());
  for (int b = 0; b < nb; ++b) {
    NCK(g_nccl.Send(d_send + b, sizeof(P2pHello), ncclChar, H.neighbor_ranks[b], (ncclComm_t)h->nccl_comm, h->stream));
    NCK(g_nccl.Recv(d_recv + b, sizeof(P2pHello), ncclChar, H.neighbor_ranks[b], (ncclComm_t)h->nccl_comm, h->stream));
  }
  NCK(g_nccl.GroupEnd());
  CK(cudaMemcpyAsync(hr.data(), d_recv, sizeof(P2pHello) * nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, d_send));
  CKR(lvpp_dfree(h, d_recv));
  H.peer_flag.assign((size_t)nb, nullptr);
  H.peer_rbuf[0].assign((size_t)nb, nullptr);
  H.peer_rbuf[1].assign((size_t)nb, nullptr);
  for (int b = 0; b < nb && ok; ++b) {
    if (!hr[b].ok) { ok = 0; break; }
    void* base = nullptr;
    if (cudaIpcOpenMemHandle(&base, hr[b].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    H.peer_base.push_back(base);
    H.peer_flag[b] = (unsigned long long*)base + hr[b].slot;
    double* r0 = (double*)((char*)base + flag_bytes);
    H.peer_rbuf[0][b] = r0 + hr[b].seg_off;
    H.peer_rbuf[1][b] = r0 + 2 * hr[b].nrecv + hr[b].seg_off;
  }
  // the decision must be the same on every rank: any failure anywhere -> everybody stays on NCCL send/recv
  double fails = ok ? 0.0 : 1.0;
  double* d_f = nullptr;
  CKR(lvpp_dalloc(h, &d_f, 1));
  CK(cudaMemcpyAsync(d_f, &fails, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CKR(lvpp_allreduce_sum(h, d_f, 1));
  CK(cudaMemcpyAsync(&fails, d_f, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, d_f));
  if (fails > 0.0 || nb == 0) return 0;
  H.flags = (unsigned long long*)H.arena;
  H.rbuf[0] = (double*)((char*)H.arena + flag_bytes);
  H.rbuf[1] = H.rbuf[0] + 2 * (nr > 0 ? nr : 1);
  CKR(lvpp_dalloc(h, &H.counters, (size_t)nb));
  H.seq = 0;
  H.p2p = true;
  return 0;
}

static int halo_forward_p2p(lvpp_problem* h, LevelHalo& H, double* d_v) {
  const unsigned long long seq = ++H.seq;
  const int par = (int)(seq & 1);
  HaloPush hp;
  hp.nb = H.num_neighbors;
  for (int b = 0; b < H.num_neighbors; ++b) {
    hp.send_ptr[b] = H.send_ptr[b];
    hp.peer_buf[b] = (double2*)H.peer_rbuf[par][b];
    hp.peer_flag[b] = H.peer_flag[b];
  }
  hp.send_ptr[H.num_neighbors] = H.send_ptr[H.num_neighbors];
  const int64_t nr = H.recv_ptr.back(), ns = H.send_ptr.back();
  int64_t blocks = (std::max(nr, ns) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > LVPP_NUM_SMS) blocks = LVPP_NUM_SMS;  // all blocks resident: see k_halo_p2p
  LAUNCH(h, k_halo_p2p, (int)blocks, 256, 0, hp, H.send_nodes, H.recv_nodes, nr, (double2*)d_v, (const double2*)H.rbuf[par],
         H.flags, H.counters, seq, h->p2p_err, h->halo_timeout_ticks);
  CK(cudaGetLastError());
  return 0;
}

int lvpp_halo_forward_level(lvpp_problem* h, LevelHalo& H, double* d_v) {
  if (h->nranks <= 1 || H.num_neighbors == 0) return 0;
  if (!h->nccl_comm) { lvpp_set_error("communicator not initialised"); return LVPP_E_COMM; }
  if (H.p2p) return halo_forward_p2p(h, H, d_v);
  const int64_t ns = H.send_ptr.back(), nr = H.recv_ptr.back();
  if (ns > 0) {
    LAUNCH(h, k_pack, lvpp_grid(ns, 256, 4), 256, 0, ns, H.send_nodes, (const double2*)d_v, (double2*)H.send_buf);
    CK(cudaGetLastError());
  }
  NCK(g_nccl.GroupStart());
  for (int b = 0; b < H.num_neighbors; ++b) {
    const int64_t s0 = H.send_ptr[b], s1 = H.send_ptr[b + 1], r0 = H.recv_ptr[b], r1 = H.recv_ptr[b + 1];
    if (s1 > s0)
      NCK(g_nccl.Send(H.send_buf + 2 * s0, (size_t)(2 * (s1 - s0)), ncclDouble, H.neighbor_ranks[b],
                      (ncclComm_t)h->nccl_comm, h->stream));
    if (r1 > r0)
      NCK(g_nccl.Recv(H.recv_buf + 2 * r0, (size_t)(2 * (r1 - r0)), ncclDouble, H.neighbor_ranks[b],
                      (ncclComm_t)h->nccl_comm, h->stream));
  }
  NCK(g_nccl.GroupEnd());
  if (nr > 0) {
    LAUNCH(h, k_unpack, lvpp_grid(nr, 256, 4), 256, 0, nr, H.recv_nodes, (const double2*)H.recv_buf, (double2*)d_v);
    CK(cudaGetLastError());
  }
  return 0;
}

int lvpp_halo_forward_impl(lvpp_problem* h, double* d_v) { return lvpp_halo_forward_level(h, h->halo, d_v); }

int lvpp_halo_exchange_i32(lvpp_problem* h, const LevelHalo& H, const int32_t* d_send, int32_t* d_recv) {
  if (h->nranks <= 1 || H.num_neighbors == 0) return 0;
  if (!h->nccl_comm) { lvpp_set_error("communicator not initialised"); return LVPP_E_COMM; }
  NCK(g_nccl.GroupStart());
  for (int b = 0; b < H.num_neighbors; ++b) {
    const int64_t s0 = H.send_ptr[b], s1 = H.send_ptr[b + 1], r0 = H.recv_ptr[b], r1 = H.recv_ptr[b + 1];
    if (s1 > s0)
      NCK(g_nccl.Send(d_send + s0, (size_t)(s1 - s0), ncclInt32, H.neighbor_ranks[b], (ncclComm_t)h->nccl_comm, h->stream));
    if (r1 > r0)
      NCK(g_nccl.Recv(d_recv + r0, (size_t)(r1 - r0), ncclInt32, H.neighbor_ranks[b], (ncclComm_t)h->nccl_comm, h->stream));
  }
  NCK(g_nccl.GroupEnd());
  return 0;
}

extern "C" int lvpp_halo_forward(lvpp_handle h, double* d_v) {
  if (!h || !d_v) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  CKR(lvpp_halo_forward_impl(h, d_v));
  CKR(lvpp_sync_check_comm(h));
  return LVPP_OK;
}
