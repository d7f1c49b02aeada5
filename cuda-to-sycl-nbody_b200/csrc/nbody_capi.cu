// nbody_capi.cu -- C ABI (include/nbody_b200.h): device memory, stream layout, multi-GPU
// sharding and the NCCL position exchange around the kernels of nbody_kernels.cu.
//
// Memory / stream layout per GPU (replaces ParticleData_d + sendToDevice/recvFromDevice,
// src/simulator.cuh:87-98, src/simulator.cu:79-129):
//   pos[2]   full-N float4 (x,y,z,mass) replicas, double buffered like pos_d/pos_next_d
//   vel      float4 per OWNED body (contiguous i-shard [i_begin, i_begin+i_count))
//   acc      float4 per owned body: accumulators carried between j-chunk launches
//   stage[3] N floats each (+ vstage[3], owned bodies): SoA staging for the reference's ParticleData host layout
//   compute stream: force+integrate launches; comm stream: NCCL broadcasts of the new shard;
//   copy stream: device->host copies of a read-back, overlapped with the velocity de-interleave
//
// Multi-GPU iteration (bit-exact w.r.t. one GPU): the j-loop is cut into `world` chunks in
// ascending rank order = ascending j order; chunk c is launched as soon as the broadcast of
// rank c's new positions (previous iteration) has landed (one event per chunk), accumulators
// are carried through `acc`, so the per-body sum is still ONE fp32 chain over j = 0..N-1.
// Broadcasts of iteration t overlap the chunk launches of iteration t+1.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only; the library is dlopen()ed when world > 1
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/nbody_b200.h"
#include "nbody_kernels.cuh"

namespace {

thread_local char g_err[512] = "no error";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(expr)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail((int)e_, "CUDA error %d (%s) at %s:%d: %s", (int)e_, cudaGetErrorString(e_), \
                  __FILE__, __LINE__, #expr);                                               \
  } while (0)

// ---- NCCL, bound at run time -------------------------------------------------------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names)
    if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!lib) return fail(NBODY_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                    \
  *(void **)(&g_nccl.field) = dlsym(lib, name);                             \
  if (!g_nccl.field) return fail(NBODY_E_NCCL, "NCCL symbol %s missing", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Broadcast, "ncclBroadcast")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(CommGetAsyncError, "ncclCommGetAsyncError")
#undef SYM
  g_nccl.lib = lib;
  return 0;
}

#define NK(expr)                                                                         \
  do {                                                                                   \
    ncclResult_t r_ = (expr);                                                            \
    if (r_ != ncclSuccess)                                                               \
      return fail(NBODY_E_NCCL, "NCCL error %d (%s) at %s:%d: %s", (int)r_,              \
                  g_nccl.GetErrorString(r_), __FILE__, __LINE__, #expr);                 \
  } while (0)

struct DeviceCtx {
  int device = 0;
  int rank = 0;  // global rank of this shard
  uint32_t i_begin = 0, i_count = 0;
  int sms = 0;
  cudaStream_t compute = nullptr, comm = nullptr, copy = nullptr;
  float4 *pos[2] = {nullptr, nullptr};
  float4 *vel = nullptr, *acc = nullptr;
  float4 *gather = nullptr;  // lazily allocated full-N scratch (velocity / accel gathers)
  float *stage[3] = {nullptr, nullptr, nullptr};
  float *vstage[3] = {nullptr, nullptr, nullptr};  // second staging triple (owned bodies): velocity read-back
  float *mass = nullptr;  // optional staging for masses
  ncclComm_t nccl = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_computed = nullptr, ev_comm_done = nullptr;
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};  // staging triple filled (read-back pipeline)
  std::vector<cudaEvent_t> chunk_ready;
  // peer-push exchange: every rank's two position replicas as seen from this device
  // (own pointers, cudaDeviceEnablePeerAccess mappings, or cudaIpcOpenMemHandle mappings)
  std::vector<float4 *> peer_pos[2];
  std::vector<void *> ipc_opened;
  cudaEvent_t iter_done[2] = {nullptr, nullptr};
  float *barrier_word = nullptr;  // 1 float, NCCL all-reduce used as a device-side barrier (rank mode)
  nbody::SegSync sync;               // j-segment hand-off words + ticket counter + error word of this device
  unsigned int *error_host = nullptr;  // host side of sync.error (mapped pinned memory)
  nbody::KernelConfig cfg{};
  std::string name;
};

}  // namespace

struct nbody_handle {
  nbody_params p{};
  uint32_t n = 0;
  int world = 1;
  std::vector<DeviceCtx> devs;
  std::vector<uint32_t> shard_begin;  // world+1 entries
  int cur = 0;                        // index of the position buffer holding the current state
  bool replicas_fresh = true;         // current positions complete on every device, no pending events
  bool has_mass = false;
  int exchange = 0;  // 0 = chunked NCCL broadcasts, 1 = peer push from the kernel epilogue
  uint64_t iter_count = 0;
  int kernel = NBODY_KERNEL_AUTO;
  float last_ms = 0.0f, last_dev_ms = 0.0f;
  uint64_t launches = 0;
  char kname[128] = "";
};

namespace {

// first body of rank r's shard: equal tile-aligned (128-body) shards, the last one ragged
uint64_t shard_start(uint64_t n, int world, int r) {
  uint64_t per = (n + world - 1) / world;
  per = (per + 127u) / 128u * 128u;
  uint64_t b = per * (uint64_t)r;
  return b < n ? b : n;
}

void plan_shards(nbody_handle *h) {
  h->shard_begin.resize(h->world + 1);
  for (int r = 0; r <= h->world; r++) h->shard_begin[r] = (uint32_t)shard_start(h->n, h->world, r);
}

void refresh_configs(nbody_handle *h) {
  for (auto &d : h->devs)
    d.cfg = nbody::choose_config(h->kernel, h->p.calc_method, h->p.dist_eps, d.i_count, d.sms, h->has_mass);
}

int alloc_device(nbody_handle *h, DeviceCtx &d) {
  CK(cudaSetDevice(d.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, d.device));
  d.sms = prop.multiProcessorCount;
  d.name = prop.name;
  CK(cudaStreamCreateWithFlags(&d.compute, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&d.comm, cudaStreamNonBlocking));
  const size_t n = h->n;
  const size_t own = d.i_count ? d.i_count : 1;
  CK(cudaMalloc(&d.pos[0], n * sizeof(float4)));
  CK(cudaMalloc(&d.pos[1], n * sizeof(float4)));
  CK(cudaMalloc(&d.vel, own * sizeof(float4)));
  CK(cudaMalloc(&d.acc, own * sizeof(float4)));
  // hand-off state of the j-segmented launches: [0] ticket counter, then one word per 64 owned bodies
  d.sync.n_groups = (uint32_t)(own / 64 + 2);
  CK(cudaMalloc(&d.sync.words, ((size_t)d.sync.n_groups + 1) * sizeof(unsigned int)));
  CK(cudaMemset(d.sync.words, 0, ((size_t)d.sync.n_groups + 1) * sizeof(unsigned int)));
  if (const char *e = getenv("NBODY_HANDOFF_TIMEOUT_S")) d.sync.timeout_ns = (unsigned long long)(atof(e) * 1e9);
  CK(cudaHostAlloc(&d.error_host, sizeof(unsigned int), cudaHostAllocMapped));
  *d.error_host = 0;
  CK(cudaHostGetDevicePointer(&d.sync.error, d.error_host, 0));
  for (int k = 0; k < 3; k++) CK(cudaMalloc(&d.stage[k], n * sizeof(float)));
  for (int k = 0; k < 3; k++) CK(cudaMalloc(&d.vstage[k], own * sizeof(float)));
  CK(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
  for (auto &e : d.ev_stage) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaEventCreate(&d.ev_start));
  CK(cudaEventCreate(&d.ev_stop));
  CK(cudaEventCreateWithFlags(&d.ev_computed, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&d.ev_comm_done, cudaEventDisableTiming));
  d.chunk_ready.resize(h->world);
  for (auto &e : d.chunk_ready) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return 0;
}

int sync_all(nbody_handle *h) {
  for (auto &d : h->devs) {
    CK(cudaSetDevice(d.device));
    CK(cudaStreamSynchronize(d.compute));
    CK(cudaStreamSynchronize(d.comm));
    if (d.copy) CK(cudaStreamSynchronize(d.copy));
  }
  return 0;
}

// every rank's shard of a full-N float4 array is sent to all ranks, roots in ascending order;
// when `events` is set, chunk_ready[c] is recorded after root c's broadcast on each device
int exchange_shards(nbody_handle *h, int which /*0 = pos[buf], 1 = gather*/, int buf,
                    bool record_events) {
  if (h->world == 1) return 0;
  for (int c = 0; c < h->world; c++) {
    const uint32_t b = h->shard_begin[c], cnt = h->shard_begin[c + 1] - b;
    if (cnt == 0) continue;
    NK(g_nccl.GroupStart());
    for (auto &d : h->devs) {
      float4 *base = which == 0 ? d.pos[buf] : d.gather;
      NK(g_nccl.Broadcast(base + b, base + b, (size_t)cnt * 4, ncclFloat, c, d.nccl, d.comm));
    }
    NK(g_nccl.GroupEnd());
    if (record_events)
      for (auto &d : h->devs) {
        CK(cudaSetDevice(d.device));
        CK(cudaEventRecord(d.chunk_ready[c], d.comm));
      }
  }
  return 0;
}

// enqueue one force(+integrate) pass over all j on every local device.  flags_last carries
// kLastChunk and optionally kAccelOut.  Reads pos[src]; integrating writes pos[src^1].
int enqueue_pass(nbody_handle *h, int src, int flags_last, bool wait_chunks) {
  for (auto &d : h->devs) {
    if (d.i_count == 0) continue;
    CK(cudaSetDevice(d.device));
    nbody::StepArgs a;
    a.pos = d.pos[src];
    a.pos_next = d.pos[src ^ 1];
    a.vel = d.vel;
    a.acc = d.acc;
    a.n = h->n;
    a.i_begin = d.i_begin;
    a.i_count = d.i_count;
    a.eps = h->p.dist_eps;
    a.dt = h->p.dt;
    a.G = h->p.G;
    a.damping = h->p.damping;
    a.n_peers = 0;
    a.sync = &d.sync;
    if (h->world == 1 || h->exchange == 1) {
      // one launch over all j.  Peer push: the epilogue stores the new positions into every
      // other rank's next-position replica as well (not for the accel dump, which moves nothing)
      a.j_begin = 0;
      a.j_end = h->n;
      a.flags = nbody::kFirstChunk | flags_last;
      if (h->world > 1 && !(flags_last & nbody::kAccelOut))
        for (int r = 0; r < h->world; r++)
          if (r != d.rank) a.peer_next[a.n_peers++] = d.peer_pos[src ^ 1][r];
      CK(nbody::launch_step(d.cfg, a, d.compute));
      h->launches++;
      continue;
    }
    int first = 1;
    int last_nonempty = -1;
    for (int c = 0; c < h->world; c++)
      if (h->shard_begin[c + 1] > h->shard_begin[c]) last_nonempty = c;
    for (int c = 0; c < h->world; c++) {
      a.j_begin = h->shard_begin[c];
      a.j_end = h->shard_begin[c + 1];
      if (a.j_end == a.j_begin) continue;
      if (wait_chunks) CK(cudaStreamWaitEvent(d.compute, d.chunk_ready[c], 0));
      a.flags = (first ? nbody::kFirstChunk : 0) | (c == last_nonempty ? flags_last : 0);
      first = 0;
      CK(nbody::launch_step(d.cfg, a, d.compute));
      h->launches++;
    }
  }
  return 0;
}

int upload_soa(nbody_handle *h, DeviceCtx &d, const float *x, const float *y, const float *z,
               const float *m, float w, float4 *dst, uint32_t begin, uint32_t count) {
  if (count == 0) return 0;
  CK(cudaMemcpyAsync(d.stage[0], x + begin, count * sizeof(float), cudaMemcpyHostToDevice, d.compute));
  CK(cudaMemcpyAsync(d.stage[1], y + begin, count * sizeof(float), cudaMemcpyHostToDevice, d.compute));
  CK(cudaMemcpyAsync(d.stage[2], z + begin, count * sizeof(float), cudaMemcpyHostToDevice, d.compute));
  const float *md = nullptr;
  if (m) {
    if (!d.mass) CK(cudaMalloc(&d.mass, (size_t)h->n * sizeof(float)));
    CK(cudaMemcpyAsync(d.mass, m + begin, count * sizeof(float), cudaMemcpyHostToDevice, d.compute));
    md = d.mass;
  }
  CK(nbody::launch_interleave(d.stage[0], d.stage[1], d.stage[2], md, w, dst, count, d.compute));
  h->launches++;
  return 0;
}

// float4 device array -> three host arrays.  The de-interleave runs on the compute stream into one of
// the two staging triples (which = 0: N floats each; 1: owned-body count each), the three copies go to
// the COPY stream behind an event, so the de-interleave of the next array overlaps them.  Asynchronous:
// the caller ends with sync_all().  With registered / pinned host memory (nbody_host_register) the
// copies are true DMA transfers; with pageable memory the driver stages them.
int download_soa(nbody_handle *h, DeviceCtx &d, const float4 *src, uint32_t count, float *x, float *y,
                 float *z, int which = 0) {
  if (count == 0) return 0;
  float *const *st = which ? d.vstage : d.stage;
  CK(nbody::launch_deinterleave(src, st[0], st[1], st[2], count, d.compute));
  h->launches++;
  CK(cudaEventRecord(d.ev_stage[which], d.compute));
  CK(cudaStreamWaitEvent(d.copy, d.ev_stage[which], 0));
  CK(cudaMemcpyAsync(x, st[0], count * sizeof(float), cudaMemcpyDeviceToHost, d.copy));
  CK(cudaMemcpyAsync(y, st[1], count * sizeof(float), cudaMemcpyDeviceToHost, d.copy));
  CK(cudaMemcpyAsync(z, st[2], count * sizeof(float), cudaMemcpyDeviceToHost, d.copy));
  return 0;
}

// gathers a shard-local float4 array (vel or acc) of every rank into host SoA / AoS output
int read_sharded(nbody_handle *h, int which /*0 = vel, 1 = acc*/, float *x, float *y, float *z,
                 float *f4) {
  const bool remote = h->world > (int)h->devs.size();  // other processes own some shards
  if (remote) {
    for (auto &d : h->devs) {
      CK(cudaSetDevice(d.device));
      if (!d.gather) CK(cudaMalloc(&d.gather, (size_t)h->n * sizeof(float4)));
      if (d.i_count)
        CK(cudaMemcpyAsync(d.gather + d.i_begin, which == 0 ? d.vel : d.acc,
                           (size_t)d.i_count * sizeof(float4), cudaMemcpyDeviceToDevice, d.compute));
      CK(cudaEventRecord(d.ev_computed, d.compute));
      CK(cudaStreamWaitEvent(d.comm, d.ev_computed, 0));
    }
    int rc = exchange_shards(h, 1, 0, false);
    if (rc) return rc;
    DeviceCtx &d = h->devs[0];
    CK(cudaSetDevice(d.device));
    CK(cudaEventRecord(d.ev_comm_done, d.comm));
    CK(cudaStreamWaitEvent(d.compute, d.ev_comm_done, 0));
    if (f4)
      CK(cudaMemcpyAsync(f4, d.gather, (size_t)h->n * sizeof(float4), cudaMemcpyDeviceToHost, d.compute));
    else if ((rc = download_soa(h, d, d.gather, h->n, x, y, z)))
      return rc;
    return sync_all(h);
  }
  for (auto &d : h->devs) {
    CK(cudaSetDevice(d.device));
    const float4 *src = which == 0 ? d.vel : d.acc;
    if (d.i_count == 0) continue;
    if (f4) {
      CK(cudaMemcpyAsync(f4 + 4 * (size_t)d.i_begin, src, (size_t)d.i_count * sizeof(float4),
                         cudaMemcpyDeviceToHost, d.compute));
    } else {
      int rc = download_soa(h, d, src, d.i_count, x + d.i_begin, y + d.i_begin, z + d.i_begin, 1);
      if (rc) return rc;
    }
  }
  return sync_all(h);
}


// Chooses how new positions reach the other GPUs each iteration and maps peer memory.
//   NBODY_EXCHANGE=p2p  : peer push -- the integrate epilogue stores into every GPU's replica over
//                         NVLink (single process: cudaDeviceEnablePeerAccess; one process per GPU:
//                         CUDA IPC handles exchanged with one ncclAllGather), one launch per iteration,
//                         a device-side barrier between iterations, no data-path collective
//   NBODY_EXCHANGE=nccl : rank-ordered ncclBroadcast per iteration, overlapped with the j-chunk kernels
//   unset / auto        : p2p when every pair of GPUs has peer access, else nccl
int setup_exchange(nbody_handle *h) {
  const char *env = getenv("NBODY_EXCHANGE");
  const bool want_nccl = env && !strcmp(env, "nccl");
  const bool want_p2p = env && !strcmp(env, "p2p");
  const bool single_process = (int)h->devs.size() == h->world;
  const bool rank_mode = h->devs.size() == 1 && h->world > 1;
  h->exchange = 0;
  if (want_nccl || (!single_process && !rank_mode)) return 0;
  if (h->world - 1 > nbody::kMaxPeers) {  // StepArgs carries at most kMaxPeers peer replicas
    if (want_p2p) return fail(NBODY_E_INVALID, "NBODY_EXCHANGE=p2p supports at most %d GPUs", nbody::kMaxPeers + 1);
    return 0;
  }

  if (single_process) {
    for (auto &d : h->devs)
      for (auto &e : h->devs) {
        if (d.device == e.device) continue;
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, d.device, e.device));
        if (!can) {
          if (want_p2p) return fail(NBODY_E_INVALID, "NBODY_EXCHANGE=p2p but device %d cannot access device %d", d.device, e.device);
          return 0;
        }
      }
    for (auto &d : h->devs) {
      CK(cudaSetDevice(d.device));
      for (auto &e : h->devs) {
        if (d.device == e.device) continue;
        cudaError_t pe = cudaDeviceEnablePeerAccess(e.device, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (pe != cudaSuccess) return fail((int)pe, "cudaDeviceEnablePeerAccess(%d -> %d): %s", d.device, e.device, cudaGetErrorString(pe));
      }
      for (int b = 0; b < 2; b++) {
        d.peer_pos[b].assign(h->world, nullptr);
        for (auto &e : h->devs) d.peer_pos[b][e.rank] = e.pos[b];
      }
      for (int b = 0; b < 2; b++) CK(cudaEventCreateWithFlags(&d.iter_done[b], cudaEventDisableTiming));
    }
    h->exchange = 1;
    return 0;
  }

  // one process per GPU: exchange CUDA IPC handles of the two replicas through NCCL
  DeviceCtx &d = h->devs[0];
  CK(cudaSetDevice(d.device));
  struct Handles { cudaIpcMemHandle_t mem[2]; int ok; int pad[15]; };
  static_assert(sizeof(Handles) % 16 == 0, "all-gather payload alignment");
  Handles mine;
  memset(&mine, 0, sizeof mine);
  mine.ok = 1;
  for (int b = 0; b < 2; b++)
    if (cudaIpcGetMemHandle(&mine.mem[b], d.pos[b]) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
  Handles *dev_all = nullptr;
  std::vector<Handles> all(h->world);
  CK(cudaMalloc(&dev_all, sizeof(Handles) * h->world));
  CK(cudaMemcpyAsync(dev_all + d.rank, &mine, sizeof mine, cudaMemcpyHostToDevice, d.comm));
  NK(g_nccl.AllGather(dev_all + d.rank, dev_all, sizeof(Handles), ncclChar, d.nccl, d.comm));
  CK(cudaMemcpyAsync(all.data(), dev_all, sizeof(Handles) * h->world, cudaMemcpyDeviceToHost, d.comm));
  CK(cudaStreamSynchronize(d.comm));
  CK(cudaFree(dev_all));
  int ok = 1;
  for (auto &a : all) ok &= a.ok;
  for (int b = 0; b < 2; b++) d.peer_pos[b].assign(h->world, nullptr);
  for (int r = 0; r < h->world && ok; r++) {
    for (int b = 0; b < 2 && ok; b++) {
      if (r == d.rank) { d.peer_pos[b][r] = d.pos[b]; continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].mem[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      d.ipc_opened.push_back(p);
      d.peer_pos[b][r] = (float4 *)p;
    }
  }
  // every rank must take the same decision: agree with a MIN all-reduce
  float *flag = nullptr;
  CK(cudaMalloc(&flag, sizeof(float)));
  float fv = ok ? 1.0f : 0.0f;
  CK(cudaMemcpyAsync(flag, &fv, sizeof fv, cudaMemcpyHostToDevice, d.comm));
  NK(g_nccl.AllReduce(flag, flag, 1, ncclFloat, ncclMin, d.nccl, d.comm));
  CK(cudaMemcpyAsync(&fv, flag, sizeof fv, cudaMemcpyDeviceToHost, d.comm));
  CK(cudaStreamSynchronize(d.comm));
  d.barrier_word = flag;
  if (fv < 0.5f) {
    for (void *p : d.ipc_opened) cudaIpcCloseMemHandle(p);
    d.ipc_opened.clear();
    cudaGetLastError();
    if (want_p2p) return fail(NBODY_E_INVALID, "NBODY_EXCHANGE=p2p but CUDA IPC peer mapping is not available between all ranks");
    return 0;
  }
  h->exchange = 1;
  return 0;
}

int create_common(const nbody_params *p, const std::vector<int> &devices, int first_rank, int world,
                  const ncclUniqueId *uid, nbody_handle **out) {
  if (!p || !out) return fail(NBODY_E_INVALID, "null argument");
  *out = nullptr;
  if (p->num_particles == 0 || p->num_particles >= (1ull << 31))
    return fail(NBODY_E_INVALID, "num_particles must be in [1, 2^31)");
  if (p->iters_per_frame < 0) return fail(NBODY_E_INVALID, "iters_per_frame < 0");
  if (p->calc_method != NBODY_CALC_BRANCH && p->calc_method != NBODY_CALC_PREDICATED &&
      p->calc_method != NBODY_CALC_PREDICATED_FIXED)
    return fail(NBODY_E_INVALID, "calc_method must be BRANCH, PREDICATED or PREDICATED_FIXED");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(NBODY_E_NOGPU, "no CUDA device visible: this library has no CPU fallback");
  }
  for (int dv : devices)
    if (dv < 0 || dv >= ndev) return fail(NBODY_E_INVALID, "device %d not present (%d visible)", dv, ndev);

  nbody_handle *h = new nbody_handle;
  h->p = *p;
  h->n = (uint32_t)p->num_particles;
  h->world = world;
  plan_shards(h);
  h->devs.resize(devices.size());
  for (size_t k = 0; k < devices.size(); k++) {
    DeviceCtx &d = h->devs[k];
    d.device = devices[k];
    d.rank = first_rank + (int)k;
    d.i_begin = h->shard_begin[d.rank];
    d.i_count = h->shard_begin[d.rank + 1] - d.i_begin;
    int rc = alloc_device(h, d);
    if (rc) { nbody_destroy(h); return rc; }
  }
  if (world > 1) {
    int rc = load_nccl();
    if (rc) { nbody_destroy(h); return rc; }
    ncclUniqueId id;
    if (uid) id = *uid;
    else if (g_nccl.GetUniqueId(&id) != ncclSuccess) { nbody_destroy(h); return fail(NBODY_E_NCCL, "ncclGetUniqueId failed"); }
    ncclResult_t r = g_nccl.GroupStart();
    for (auto &d : h->devs) {
      if (r != ncclSuccess) break;
      cudaSetDevice(d.device);
      r = g_nccl.CommInitRank(&d.nccl, world, id, d.rank);
    }
    if (r == ncclSuccess) r = g_nccl.GroupEnd();
    if (r != ncclSuccess) {
      int rc2 = fail(NBODY_E_NCCL, "NCCL communicator init failed: %s", g_nccl.GetErrorString(r));
      nbody_destroy(h);
      return rc2;
    }
  }
  if (world > 1) {
    int rc = setup_exchange(h);
    if (rc) { nbody_destroy(h); return rc; }
  }
  refresh_configs(h);
  *out = h;
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int nbody_abi_version(void) { return NBODY_B200_ABI_VERSION; }
const char *nbody_last_error(void) { return g_err; }

int nbody_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void nbody_default_params(nbody_params *out) {
  if (!out) return;
  out->G = 2.0f;
  out->dt = 0.005f;
  out->num_particles = 50 * 256;
  out->iters_per_frame = 4;
  out->damping = 0.999998f;
  out->dist_eps = 1.0e-7f;
  out->gw_size = 64;
  out->calc_method = NBODY_CALC_BRANCH;
}

int nbody_plan_shard(uint64_t n, int world, int rank, uint64_t *begin, uint64_t *count) {
  if (!begin || !count || world < 1 || rank < 0 || rank >= world) return fail(NBODY_E_INVALID, "bad shard query");
  *begin = shard_start(n, world, rank);
  *count = shard_start(n, world, rank + 1) - *begin;
  return 0;
}

int nbody_nccl_unique_id(void *out128) {
  if (!out128) return fail(NBODY_E_INVALID, "null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return 0;
}

int nbody_create(const nbody_params *p, int n_gpus, nbody_handle **out) {
  if (n_gpus <= 0) {
    const char *e = getenv("NBODY_GPUS");
    n_gpus = e ? atoi(e) : 1;
    if (n_gpus <= 0) n_gpus = 1;
  }
  std::vector<int> devs(n_gpus);
  for (int i = 0; i < n_gpus; i++) devs[i] = i;
  nbody_handle *h = nullptr;
  int rc = create_common(p, devs, 0, n_gpus, nullptr, &h);
  if (rc) return rc;
  // reference constructor: generate the galaxy, upload it (src/simulator.cu:31-33)
  const size_t n = h->n;
  std::vector<float> s(6 * n);
  rc = nbody_generate_disk_galaxy(n, &s[0], &s[n], &s[2 * n], &s[3 * n], &s[4 * n], &s[5 * n]);
  if (!rc) rc = nbody_set_state(h, &s[0], &s[n], &s[2 * n], &s[3 * n], &s[4 * n], &s[5 * n]);
  if (rc) { nbody_destroy(h); return rc; }
  *out = h;
  return 0;
}

int nbody_create_rank(const nbody_params *p, int device, int rank, int world,
                      const void *nccl_unique_id, nbody_handle **out) {
  if (world < 1 || rank < 0 || rank >= world) return fail(NBODY_E_INVALID, "bad rank/world");
  if (world > 1 && !nccl_unique_id) return fail(NBODY_E_INVALID, "nccl_unique_id required when world > 1");
  ncclUniqueId id;
  if (nccl_unique_id) memcpy(&id, nccl_unique_id, sizeof id);
  nbody_handle *h = nullptr;
  int rc = create_common(p, std::vector<int>{device}, rank, world, world > 1 ? &id : nullptr, &h);
  if (rc) return rc;
  const size_t n = h->n;
  std::vector<float> s(6 * n);
  rc = nbody_generate_disk_galaxy(n, &s[0], &s[n], &s[2 * n], &s[3 * n], &s[4 * n], &s[5 * n]);
  if (!rc) rc = nbody_set_state(h, &s[0], &s[n], &s[2 * n], &s[3 * n], &s[4 * n], &s[5 * n]);
  if (rc) { nbody_destroy(h); return rc; }
  *out = h;
  return 0;
}

int nbody_destroy(nbody_handle *h) {
  if (!h) return 0;
  for (auto &d : h->devs) {
    cudaSetDevice(d.device);
    if (d.compute) cudaStreamSynchronize(d.compute);
    if (d.comm) cudaStreamSynchronize(d.comm);
    if (d.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(d.nccl);
    cudaFree(d.pos[0]);
    cudaFree(d.pos[1]);
    cudaFree(d.vel);
    cudaFree(d.acc);
    cudaFree(d.gather);
    cudaFree(d.mass);
    cudaFree(d.barrier_word);
    cudaFree(d.sync.words);
    if (d.error_host) cudaFreeHost(d.error_host);
    for (int k = 0; k < 3; k++) cudaFree(d.vstage[k]);
    for (cudaEvent_t e : d.ev_stage)
      if (e) cudaEventDestroy(e);
    if (d.copy) cudaStreamDestroy(d.copy);
    for (void *p : d.ipc_opened) cudaIpcCloseMemHandle(p);
    for (cudaEvent_t e : d.iter_done)
      if (e) cudaEventDestroy(e);
    for (int k = 0; k < 3; k++) cudaFree(d.stage[k]);
    for (cudaEvent_t e : {d.ev_start, d.ev_stop, d.ev_computed, d.ev_comm_done})
      if (e) cudaEventDestroy(e);
    for (auto e : d.chunk_ready)
      if (e) cudaEventDestroy(e);
    if (d.compute) cudaStreamDestroy(d.compute);
    if (d.comm) cudaStreamDestroy(d.comm);
  }
  cudaGetLastError();
  delete h;
  return 0;
}

int nbody_set_kernel(nbody_handle *h, int kernel) {
  if (!h) return fail(NBODY_E_INVALID, "null handle");
  if (kernel < NBODY_KERNEL_AUTO || kernel > NBODY_KERNEL_SCALAR) return fail(NBODY_E_INVALID, "unknown kernel id");
  if ((kernel == NBODY_KERNEL_PACKED || kernel == NBODY_KERNEL_SCALAR) &&
      (h->p.calc_method != NBODY_CALC_BRANCH || !nbody::eps_allows_unpredicated(h->p.dist_eps)))
    return fail(NBODY_E_INVALID,
                "unpredicated kernels are not bit-exact for this distEps / calcMethod; use AUTO or GENERIC");
  if ((kernel == NBODY_KERNEL_PACKED || kernel == NBODY_KERNEL_SCALAR) && h->has_mass)
    return fail(NBODY_E_STATE, "per-body masses need the AUTO or GENERIC kernel");
  if ((kernel == NBODY_KERNEL_PACKED || kernel == NBODY_KERNEL_SCALAR) && !nbody::variants_built())
    return fail(NBODY_E_INVALID, "the CTA-tiled comparison kernels are only in libnbody_b200_variants.so (make VARIANTS=1)");
  h->kernel = kernel;
  refresh_configs(h);
  return 0;
}

// rewritten in the built library by tools/sass_post.py with the number of production-kernel instantiations whose
// unrolled tile body it re-ordered (sass_sched.py) / regenerated from scratch (sass_gen.py)
extern "C" const volatile char nbody_sass_sched_marker[] = "NBODY_SASS_SCHED=00";
extern "C" const volatile char nbody_sass_gen_marker[] = "NBODY_SASS_GEN=00";
extern "C" const volatile char nbody_sass_genm_marker[] = "NBODY_SASS_GENM=00";  // ... of the per-body-mass instantiations
extern "C" const volatile char nbody_sass_gens_marker[] = "NBODY_SASS_GENS=00";  // ... the scalar small-shard kernel

const char *nbody_kernel_name(nbody_handle *h) {
  if (!h || h->devs.empty()) return "";
  char base[96];
  nbody::config_name(h->devs[0].cfg, base, sizeof base);
  const bool sched = nbody_sass_sched_marker[17] != '0' || nbody_sass_sched_marker[18] != '0';
  const bool gen = nbody_sass_gen_marker[15] != '0' || nbody_sass_gen_marker[16] != '0';
  const bool genm = nbody_sass_genm_marker[16] != '0' || nbody_sass_genm_marker[17] != '0';
  const nbody::KernelConfig &kc = h->devs[0].cfg;
  if (kc.family == nbody::kFamSegmented || kc.family == nbody::kFamUnsegmented) {
    if (kc.mass ? genm : gen) strncat(base, "+sass-gen", sizeof base - strlen(base) - 1);
    else if (sched) strncat(base, "+sass-sched", sizeof base - strlen(base) - 1);
  } else if (kc.family == nbody::kFamSmall && kc.r == 1 && !kc.mass &&
             (nbody_sass_gens_marker[16] != '0' || nbody_sass_gens_marker[17] != '0')) {
    strncat(base, "+sass-gen", sizeof base - strlen(base) - 1);
  }
  if (h->world > 1)
    snprintf(h->kname, sizeof h->kname, "%s|x%d:%s", base, h->world, h->exchange == 1 ? "peer-push" : "nccl-bcast");
  else
    snprintf(h->kname, sizeof h->kname, "%s", base);
  return h->kname;
}

int nbody_describe_auto(const nbody_params *p, uint64_t shard_bodies, int sms, int has_mass, char *buf, size_t len) {
  if (!p || !buf || len == 0 || sms < 1 || shard_bodies == 0 || shard_bodies > 0xffffffffull)
    return fail(NBODY_E_INVALID, "nbody_describe_auto: bad argument");
  const nbody::KernelConfig kc = nbody::choose_config(0 /*AUTO*/, p->calc_method, p->dist_eps, (uint32_t)shard_bodies, sms, has_mass != 0);
  nbody::config_name(kc, buf, len);
  return 0;
}

int nbody_set_state(nbody_handle *h, const float *x, const float *y, const float *z, const float *vx,
                    const float *vy, const float *vz) {
  if (!h || !x || !y || !z || !vx || !vy || !vz) return fail(NBODY_E_INVALID, "null argument");
  int rc = sync_all(h);
  if (rc) return rc;
  h->cur = 0;
  for (auto &d : h->devs) {
    CK(cudaSetDevice(d.device));
    rc = upload_soa(h, d, x, y, z, nullptr, 1.0f, d.pos[0], 0, h->n);
    if (rc) return rc;
    if (h->has_mass) {  // masses set earlier stay with their bodies
      CK(nbody::launch_set_w(d.pos[0], d.mass, 1.0f, h->n, d.compute));
      h->launches++;
    }
    // the staging arrays are reused for the velocities: same stream, so ordered after the interleave
    rc = upload_soa(h, d, vx, vy, vz, nullptr, 0.0f, d.vel, d.i_begin, d.i_count);
    if (rc) return rc;
  }
  h->replicas_fresh = true;
  return sync_all(h);
}

int nbody_set_mass(nbody_handle *h, const float *m) {
  if (!h) return fail(NBODY_E_INVALID, "null handle");
  if (m && (h->kernel == NBODY_KERNEL_PACKED || h->kernel == NBODY_KERNEL_SCALAR))
    return fail(NBODY_E_STATE, "per-body masses need the AUTO or GENERIC kernel");
  int rc = sync_all(h);
  if (rc) return rc;
  for (auto &d : h->devs) {
    CK(cudaSetDevice(d.device));
    if (m) {
      if (!d.mass) CK(cudaMalloc(&d.mass, (size_t)h->n * sizeof(float)));
      CK(cudaMemcpyAsync(d.mass, m, (size_t)h->n * sizeof(float), cudaMemcpyHostToDevice, d.compute));
    }
    CK(nbody::launch_set_w(d.pos[h->cur], m ? d.mass : nullptr, 1.0f, h->n, d.compute));
    h->launches++;
  }
  h->has_mass = m != nullptr;
  refresh_configs(h);
  return sync_all(h);
}

int nbody_step(nbody_handle *h) {
  if (!h) return fail(NBODY_E_INVALID, "null handle");
  const int iters = h->p.iters_per_frame;
  auto t0 = std::chrono::steady_clock::now();
  for (auto &d : h->devs) {
    CK(cudaSetDevice(d.device));
    CK(cudaEventRecord(d.ev_start, d.compute));
  }
  if (iters > 0 && h->world > 1 && h->exchange == 1 && (int)h->devs.size() < h->world) {
    // one process per GPU, peer push: this rank's first kernel stores into every OTHER rank's
    // next-position replica.  Those ranks may still be reading that buffer (a read-back or state
    // dump of the previous frame, or nbody_set_state), so every rank joins a device-side barrier
    // before the frame's first launch -- nbody_step is collective over the ranks.
    DeviceCtx &d = h->devs[0];
    CK(cudaSetDevice(d.device));
    NK(g_nccl.AllReduce(d.barrier_word, d.barrier_word, 1, ncclFloat, ncclMin, d.nccl, d.compute));
  }
  for (int it = 0; it < iters; it++) {
    int rc = enqueue_pass(h, h->cur, nbody::kLastChunk, !h->replicas_fresh);
    if (rc) return rc;
    h->cur ^= 1;
    if (h->world > 1 && h->exchange == 1) {
      // peer push: the kernels already delivered the shards; what is left is the barrier that
      // keeps iteration t+1 (which overwrites the buffer iteration t read) behind every GPU's t
      const int slot = (int)(h->iter_count & 1);
      if ((int)h->devs.size() == h->world) {
        for (auto &d : h->devs) {
          CK(cudaSetDevice(d.device));
          CK(cudaEventRecord(d.iter_done[slot], d.compute));
        }
        for (auto &d : h->devs) {
          CK(cudaSetDevice(d.device));
          for (auto &e : h->devs)
            if (e.device != d.device) CK(cudaStreamWaitEvent(d.compute, e.iter_done[slot], 0));
        }
      } else {
        DeviceCtx &d = h->devs[0];
        CK(cudaSetDevice(d.device));
        NK(g_nccl.AllReduce(d.barrier_word, d.barrier_word, 1, ncclFloat, ncclMin, d.nccl, d.compute));
      }
      h->iter_count++;
    } else if (h->world > 1) {
      for (auto &d : h->devs) {
        CK(cudaSetDevice(d.device));
        CK(cudaEventRecord(d.ev_computed, d.compute));
        CK(cudaStreamWaitEvent(d.comm, d.ev_computed, 0));
      }
      rc = exchange_shards(h, 0, h->cur, true);
      if (rc) return rc;
      h->replicas_fresh = false;
    }
  }
  for (auto &d : h->devs) {
    CK(cudaSetDevice(d.device));
    if (h->world > 1 && iters > 0 && h->exchange == 0) {
      CK(cudaEventRecord(d.ev_comm_done, d.comm));
      CK(cudaStreamWaitEvent(d.compute, d.ev_comm_done, 0));
    }
    CK(cudaEventRecord(d.ev_stop, d.compute));
  }
  int rc = sync_all(h);
  if (rc) return rc;
  auto t1 = std::chrono::steady_clock::now();
  h->last_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
  h->replicas_fresh = true;  // everything drained: no event waits needed for the next pass
  float mx = 0.0f;
  for (auto &d : h->devs) {
    float ms = 0.0f;
    CK(cudaSetDevice(d.device));
    CK(cudaEventElapsedTime(&ms, d.ev_start, d.ev_stop));
    if (ms > mx) mx = ms;
    if (d.nccl) {
      ncclResult_t ar = ncclSuccess;
      if (g_nccl.CommGetAsyncError(d.nccl, &ar) != ncclSuccess || ar != ncclSuccess)
        return fail(NBODY_E_NCCL, "NCCL asynchronous error: %s", g_nccl.GetErrorString(ar));
    }
  }
  h->last_dev_ms = mx;
  for (auto &d : h->devs)
    if (d.error_host && *(volatile unsigned int *)d.error_host) {
      *d.error_host = 0;
      return fail(NBODY_E_STATE, "device %d: a j-segment hand-off wait exceeded %.0f s (GPU time-sliced, halted by a debugger or "
                                 "slowed by a sanitizer? NBODY_HANDOFF_TIMEOUT_S=0 waits for ever); the state of this handle is no "
                                 "longer valid -- reload it with nbody_set_state",
                  d.device, d.sync.timeout_ns * 1e-9);
    }
  return 0;
}

float nbody_last_step_ms(nbody_handle *h) { return h ? h->last_ms : 0.0f; }
float nbody_last_step_device_ms(nbody_handle *h) { return h ? h->last_dev_ms : 0.0f; }
uint64_t nbody_launch_count(nbody_handle *h) { return h ? h->launches : 0; }

int nbody_read_pos(nbody_handle *h, float *x, float *y, float *z) {
  if (!h || !x || !y || !z) return fail(NBODY_E_INVALID, "null argument");
  DeviceCtx &d = h->devs[0];
  CK(cudaSetDevice(d.device));
  int rc = download_soa(h, d, d.pos[h->cur], h->n, x, y, z);
  if (rc) return rc;
  return sync_all(h);
}

int nbody_read_pos_f4(nbody_handle *h, float *xyzw) {
  if (!h || !xyzw) return fail(NBODY_E_INVALID, "null argument");
  DeviceCtx &d = h->devs[0];
  CK(cudaSetDevice(d.device));
  CK(cudaMemcpyAsync(xyzw, d.pos[h->cur], (size_t)h->n * sizeof(float4), cudaMemcpyDeviceToHost, d.compute));
  return sync_all(h);
}

int nbody_read_vel(nbody_handle *h, float *vx, float *vy, float *vz) {
  if (!h || !vx || !vy || !vz) return fail(NBODY_E_INVALID, "null argument");
  return read_sharded(h, 0, vx, vy, vz, nullptr);
}

int nbody_read_vel_f4(nbody_handle *h, float *xyzw) {
  if (!h || !xyzw) return fail(NBODY_E_INVALID, "null argument");
  return read_sharded(h, 0, nullptr, nullptr, nullptr, xyzw);
}

int nbody_read_state(nbody_handle *h, float *x, float *y, float *z, float *vx, float *vy, float *vz) {
  if (!h || !x || !y || !z || !vx || !vy || !vz) return fail(NBODY_E_INVALID, "null argument");
  if (h->world > (int)h->devs.size()) {  // other processes own some shards: gather the velocities first
    int rc = nbody_read_pos(h, x, y, z);
    return rc ? rc : nbody_read_vel(h, vx, vy, vz);
  }
  DeviceCtx &d0 = h->devs[0];
  CK(cudaSetDevice(d0.device));
  int rc = download_soa(h, d0, d0.pos[h->cur], h->n, x, y, z, 0);
  if (rc) return rc;
  for (auto &d : h->devs) {  // velocity de-interleave overlaps the position copies
    CK(cudaSetDevice(d.device));
    if ((rc = download_soa(h, d, d.vel, d.i_count, vx + d.i_begin, vy + d.i_begin, vz + d.i_begin, 1))) return rc;
  }
  return sync_all(h);
}

int nbody_local_range(nbody_handle *h, uint64_t *begin, uint64_t *count) {
  if (!h || !begin || !count || h->devs.empty()) return fail(NBODY_E_INVALID, "null argument");
  *begin = h->devs.front().i_begin;
  *count = (uint64_t)h->devs.back().i_begin + h->devs.back().i_count - *begin;
  return 0;
}

int nbody_read_local(nbody_handle *h, float *x, float *y, float *z, float *vx, float *vy, float *vz) {
  if (!h || !x || !y || !z || !vx || !vy || !vz) return fail(NBODY_E_INVALID, "null argument");
  const size_t b0 = h->devs.front().i_begin;
  for (auto &d : h->devs) {  // every device serves its own shard from its own replica: parallel PCIe links
    CK(cudaSetDevice(d.device));
    const size_t o = d.i_begin - b0;
    int rc = download_soa(h, d, d.pos[h->cur] + d.i_begin, d.i_count, x + o, y + o, z + o, 0);
    if (!rc) rc = download_soa(h, d, d.vel, d.i_count, vx + o, vy + o, vz + o, 1);
    if (rc) return rc;
  }
  return sync_all(h);
}

int nbody_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(NBODY_E_INVALID, "null argument");
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return 0;
}

int nbody_host_unregister(void *ptr) {
  if (!ptr) return fail(NBODY_E_INVALID, "null argument");
  CK(cudaHostUnregister(ptr));
  return 0;
}

int nbody_compute_accel(nbody_handle *h, float *ax, float *ay, float *az) {
  if (!h || !ax || !ay || !az) return fail(NBODY_E_INVALID, "null argument");
  int rc = sync_all(h);
  if (rc) return rc;
  rc = enqueue_pass(h, h->cur, nbody::kLastChunk | nbody::kAccelOut, false);
  if (rc) return rc;
  return read_sharded(h, 1, ax, ay, az, nullptr);
}

namespace {
struct CkptHeader {
  char magic[8];
  uint64_t n;
  float G, dt, damping, dist_eps;
  int32_t iters_per_frame, calc_method, has_mass, reserved;
};
const char kCkptMagic[8] = {'N', 'B', 'B', '2', '0', '0', 0, 1};
}  // namespace

int nbody_save_state(nbody_handle *h, const char *path) {
  if (!h || !path) return fail(NBODY_E_INVALID, "null argument");
  const size_t n = h->n;
  std::vector<float> buf(7 * n);
  int rc = nbody_read_pos(h, &buf[0], &buf[n], &buf[2 * n]);
  if (!rc) rc = nbody_read_vel(h, &buf[3 * n], &buf[4 * n], &buf[5 * n]);
  if (rc) return rc;
  if (h->has_mass) {
    std::vector<float> p4(4 * n);
    if ((rc = nbody_read_pos_f4(h, p4.data()))) return rc;
    for (size_t i = 0; i < n; i++) buf[6 * n + i] = p4[4 * i + 3];
  }
  CkptHeader hd;
  memset(&hd, 0, sizeof hd);
  memcpy(hd.magic, kCkptMagic, 8);
  hd.n = n;
  hd.G = h->p.G;
  hd.dt = h->p.dt;
  hd.damping = h->p.damping;
  hd.dist_eps = h->p.dist_eps;
  hd.iters_per_frame = h->p.iters_per_frame;
  hd.calc_method = h->p.calc_method;
  hd.has_mass = h->has_mass ? 1 : 0;
  FILE *f = fopen(path, "wb");
  if (!f) return fail(NBODY_E_INVALID, "cannot open %s for writing", path);
  const size_t count = (h->has_mass ? 7 : 6) * n;
  bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && fwrite(buf.data(), sizeof(float), count, f) == count;
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : fail(NBODY_E_INVALID, "short write to %s", path);
}

int nbody_load_state(nbody_handle *h, const char *path) {
  if (!h || !path) return fail(NBODY_E_INVALID, "null argument");
  FILE *f = fopen(path, "rb");
  if (!f) return fail(NBODY_E_INVALID, "cannot open %s", path);
  CkptHeader hd;
  if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, kCkptMagic, 8) != 0) {
    fclose(f);
    return fail(NBODY_E_INVALID, "%s is not an nbody-b200 checkpoint", path);
  }
  if (hd.n != h->n) {
    fclose(f);
    return fail(NBODY_E_INVALID, "checkpoint holds %llu bodies, handle %u", (unsigned long long)hd.n, h->n);
  }
  const size_t n = h->n, count = (hd.has_mass ? 7 : 6) * n;
  std::vector<float> buf(count);
  const bool ok = fread(buf.data(), sizeof(float), count, f) == count;
  fclose(f);
  if (!ok) return fail(NBODY_E_INVALID, "%s is truncated", path);
  int rc = nbody_set_mass(h, hd.has_mass ? &buf[6 * n] : nullptr);
  if (!rc) rc = nbody_set_state(h, &buf[0], &buf[n], &buf[2 * n], &buf[3 * n], &buf[4 * n], &buf[5 * n]);
  return rc;
}

const char *nbody_device_name(nbody_handle *h) { return (h && !h->devs.empty()) ? h->devs[0].name.c_str() : "Unknown Device"; }
uint64_t nbody_num_particles(nbody_handle *h) { return h ? h->n : 0; }
int nbody_num_gpus(nbody_handle *h) { return h ? (int)h->devs.size() : 0; }
int nbody_world_size(nbody_handle *h) { return h ? h->world : 0; }

int nbody_launch_step_device(const nbody_params *p, const void *pos4, void *vel4, void *pos4_next,
                             uint64_t i_begin, uint64_t i_count, int kernel, int flags, void *cuda_stream) {
  if (!p || !pos4 || !vel4 || !pos4_next) return fail(NBODY_E_INVALID, "null argument");
  if (p->num_particles == 0 || p->num_particles >= (1ull << 31) || i_begin + i_count > p->num_particles)
    return fail(NBODY_E_INVALID, "bad body range");
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if ((kernel == NBODY_KERNEL_PACKED || kernel == NBODY_KERNEL_SCALAR) &&
      (p->calc_method != NBODY_CALC_BRANCH || !nbody::eps_allows_unpredicated(p->dist_eps)))
    return fail(NBODY_E_INVALID, "unpredicated kernels are not bit-exact for this distEps / calcMethod");
  if ((kernel == NBODY_KERNEL_PACKED || kernel == NBODY_KERNEL_SCALAR) && !nbody::variants_built())
    return fail(NBODY_E_INVALID, "the CTA-tiled comparison kernels are only in libnbody_b200_variants.so (make VARIANTS=1)");
  if (flags & ~NBODY_DEVSTEP_MASS) return fail(NBODY_E_INVALID, "unknown flags");
  const bool has_mass = (flags & NBODY_DEVSTEP_MASS) != 0;
  if (has_mass && (kernel == NBODY_KERNEL_PACKED || kernel == NBODY_KERNEL_SCALAR))
    return fail(NBODY_E_INVALID, "per-body masses need the AUTO or GENERIC kernel");
  nbody::KernelConfig cfg = nbody::choose_config(kernel, p->calc_method, p->dist_eps, (uint32_t)i_count, sms, has_mass);
  nbody::StepArgs a;
  a.pos = (const float4 *)pos4;
  a.pos_next = (float4 *)pos4_next;
  a.vel = (float4 *)vel4;
  a.acc = nullptr;
  a.n = (uint32_t)p->num_particles;
  a.i_begin = (uint32_t)i_begin;
  a.i_count = (uint32_t)i_count;
  a.j_begin = 0;
  a.j_end = a.n;
  a.eps = p->dist_eps;
  a.dt = p->dt;
  a.G = p->G;
  a.damping = p->damping;
  a.flags = nbody::kFirstChunk | nbody::kLastChunk;
  a.n_peers = 0;
  a.sync = nullptr;  // no hand-off buffers with caller-owned memory: one j-segment per body group
  CK(nbody::launch_step(cfg, a, (cudaStream_t)cuda_stream));
  return 0;
}

}  // extern "C"
