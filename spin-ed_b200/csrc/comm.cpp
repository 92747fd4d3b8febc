// comm.cpp -- one process per GPU; collectives are NCCL over NVLink/NVSwitch.
//
// The reference is a single shared-memory process (OpenMP inside liblattice_symmetries,
// /root/reference/configure:63); sharding rows over GPUs is new.  NCCL is bound at run time with
// dlopen so that libsped.so itself loads on a machine without NCCL (symbol-export tests) and so
// that, inside a PyTorch process, the NCCL already mapped by torch is the one used.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"

#if !defined(SPED_CE_DEFAULT)
#define SPED_CE_DEFAULT 1  // measured on 4 B200: chain_40 20.5 ms per matvec against 22.2 with NCCL send/recv rounds
#endif

namespace sped {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api() {
  static NcclApi a;
  if (a.handle) return a;
  char const* names[] = {"libnccl.so.2", "libnccl.so"};
  for (char const* n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) fail(SPED_NCCL_ERROR, std::string("cannot load NCCL: ") + dlerror());
  auto load = [&](char const* sym) {
    void* p = dlsym(a.handle, sym);
    if (!p) fail(SPED_NCCL_ERROR, std::string("NCCL symbol missing: ") + sym);
    return p;
  };
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(load("ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(load("ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(load("ncclCommDestroy"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(load("ncclAllGather"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(load("ncclAllReduce"));
  a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(load("ncclBroadcast"));
  a.Send = reinterpret_cast<decltype(a.Send)>(load("ncclSend"));
  a.Recv = reinterpret_cast<decltype(a.Recv)>(load("ncclRecv"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(load("ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(load("ncclGroupEnd"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(load("ncclGetErrorString"));
  return a;
}

void nccl_check(ncclResult_t r, char const* what) {
  if (r != ncclSuccess) fail(SPED_NCCL_ERROR, std::string(what) + ": " + api().GetErrorString(r));
}

Comm g_comm;

constexpr int kCeStreams = 4;

struct CeExchange {
  size_t bytes = 0;                  // capacity of each send buffer
  void* send[2] = {nullptr, nullptr};
  std::vector<void*> peer[2];        // peer[q][p]: send buffer q of rank p mapped into this process (own: send[q])
  cudaStream_t streams[kCeStreams] = {};
  cudaEvent_t ev_start = nullptr, ev_done[kCeStreams] = {};
  unsigned long long* d_word = nullptr;  // the barrier's all-reduce operand
  unsigned char* d_handles = nullptr;    // staging of the IPC handles for their all-gather
  unsigned long long counter = 0;
  bool streams_ready = false;
  bool unavailable = false;  // set on every rank alike when the IPC set-up failed somewhere
};
CeExchange g_ce;

// Peers' mappings are closed before anybody frees the memory behind them: with `collective` a
// barrier separates the two steps on all ranks (the callers guarantee every rank is here).
void ce_release_buffers(bool collective);

}  // namespace

Comm& comm() { return g_comm; }

static_assert(sizeof(ncclUniqueId) == 128, "unique id is exchanged as 128 bytes");

void comm_unique_id(void* out128) {
  ncclUniqueId id;
  nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, &id, 128);
}

void comm_init(int world, int rank, void const* id128) {
  if (world < 1 || rank < 0 || rank >= world) fail(LS_INVALID_ARGUMENT, "invalid world size / rank");
  if (g_comm.nccl) fail(LS_INVALID_ARGUMENT, "communicator already initialised");
  g_comm.world = world;
  g_comm.rank = rank;
  if (world == 1) return;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t c;
  nccl_check(api().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  g_comm.nccl = c;
  CUDA_CHECK(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
  int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&g_comm.gather_stream, cudaStreamNonBlocking, prio_hi));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_comm.ev_ready, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_comm.ev_gathered, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_comm.ev_round1, cudaEventDisableTiming));
  // NCCL sets its channels up at the first collective of each kind (about a second on 8 GPUs): do that
  // here, as part of bringing the communicator up, not inside the first solve
  {
    unsigned char* d_tmp = nullptr;
    CUDA_CHECK(cudaMalloc((void**)&d_tmp, (size_t)world * 16));
    CUDA_CHECK(cudaMemset(d_tmp, 0, (size_t)world * 16));
    nccl_check(api().AllReduce(d_tmp, d_tmp, 2, ncclDouble, ncclSum, c, g_comm.stream), "ncclAllReduce (warm-up)");
    CUDA_CHECK(cudaStreamSynchronize(g_comm.stream));
    nccl_check(api().AllGather(d_tmp + (size_t)rank * 16, d_tmp, 16, ncclChar, c, g_comm.gather_stream), "ncclAllGather (warm-up)");
    CUDA_CHECK(cudaStreamSynchronize(g_comm.gather_stream));
    cudaFree(d_tmp);
  }
}

void comm_finalize() {
  if (g_ce.streams_ready) {
    cudaDeviceSynchronize();
    ce_release_buffers(false);  // teardown: other ranks may already be gone
    for (auto& st : g_ce.streams) cudaStreamDestroy(st);
    cudaEventDestroy(g_ce.ev_start);
    for (auto& ev : g_ce.ev_done) cudaEventDestroy(ev);
    cudaFree(g_ce.d_word);
    cudaFree(g_ce.d_handles);
    g_ce = CeExchange{};
  }
  if (g_comm.nccl) {
    api().CommDestroy(static_cast<ncclComm_t>(g_comm.nccl));
    g_comm.nccl = nullptr;
  }
  if (g_comm.stream) {
    cudaStreamDestroy(g_comm.stream);
    g_comm.stream = nullptr;
  }
  if (g_comm.gather_stream) {
    cudaStreamDestroy(g_comm.gather_stream);
    g_comm.gather_stream = nullptr;
  }
  if (g_comm.ev_ready) cudaEventDestroy(g_comm.ev_ready);
  if (g_comm.ev_gathered) cudaEventDestroy(g_comm.ev_gathered);
  if (g_comm.ev_round1) cudaEventDestroy(g_comm.ev_round1);
  g_comm.ev_ready = g_comm.ev_gathered = g_comm.ev_round1 = nullptr;
  g_comm.world = 1;
  g_comm.rank = 0;
}

void comm_allgather_inplace(void* buf, size_t chunk_bytes, cudaStream_t s) {
  if (!g_comm.active()) return;
  char* base = static_cast<char*>(buf);
  nccl_check(api().AllGather(base + (size_t)g_comm.rank * chunk_bytes, base, chunk_bytes, ncclChar,
                             static_cast<ncclComm_t>(g_comm.nccl), s),
             "ncclAllGather");
}

void comm_exchange_round(void* buf, size_t chunk_bytes, int d_lo, int d_hi, cudaStream_t s) {
  if (!g_comm.active() || d_lo > d_hi) return;
  char* base = static_cast<char*>(buf);
  int const P = g_comm.world, r = g_comm.rank;
  ncclComm_t c = static_cast<ncclComm_t>(g_comm.nccl);
  nccl_check(api().GroupStart(), "ncclGroupStart");
  for (int d = d_lo; d <= d_hi; ++d) {
    int const to = ((r - d) % P + P) % P, from = (r + d) % P;
    nccl_check(api().Send(base + (size_t)r * chunk_bytes, chunk_bytes, ncclChar, to, c, s), "ncclSend");
    nccl_check(api().Recv(base + (size_t)from * chunk_bytes, chunk_bytes, ncclChar, from, c, s), "ncclRecv");
  }
  nccl_check(api().GroupEnd(), "ncclGroupEnd");
}

// ---- exchange by the copy engines over peer memory -------------------------------------------
namespace {
void ce_release_buffers(bool collective) {
  bool const any = g_ce.send[0] || g_ce.send[1];
  for (int q = 0; q < 2; ++q) {
    for (size_t p = 0; p < g_ce.peer[q].size(); ++p)
      if ((int)p != g_comm.rank && g_ce.peer[q][p]) cudaIpcCloseMemHandle(g_ce.peer[q][p]);
    g_ce.peer[q].clear();
  }
  if (collective && g_comm.nccl && g_ce.d_word) {
    nccl_check(api().AllReduce(g_ce.d_word, g_ce.d_word, 1, ncclUint64, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), g_comm.stream),
               "ncclAllReduce (barrier)");
    CUDA_CHECK(cudaStreamSynchronize(g_comm.stream));
  }
  (void)any;
  for (int q = 0; q < 2; ++q) {
    if (g_ce.send[q]) cudaFree(g_ce.send[q]);
    g_ce.send[q] = nullptr;
  }
  g_ce.bytes = 0;
}
}  // namespace

// Long shards only (the 40/42-spin chains: hundreds of MB per shard): the one-word all-reduce and
// the per-peer copy launches cost tens of microseconds, which a 16-32 MB shard exchange (6x6 over 4-8
// ranks, 0.2 ms with one NCCL all-gather) cannot spare.  SPED_EXCHANGE=ce / nccl forces either.
bool comm_ce_wanted(size_t chunk_bytes) {
  if (!g_comm.active()) return false;
  char const* e = std::getenv("SPED_EXCHANGE");
  if (e && e[0] == 'c') return true;
  if (e && e[0] == 'n') return false;
  return SPED_CE_DEFAULT && chunk_bytes >= ((size_t)64 << 20);
}

// Collective.  False -- on every rank alike -- when some rank could not set its side up (no memory for
// the send buffers, no IPC / peer access between the GPUs): the exchange then stays on NCCL for good.
bool comm_ce_prepare(size_t chunk_bytes) {
  if (g_ce.unavailable) return false;
  if (chunk_bytes <= g_ce.bytes) return true;
  ncclComm_t c = static_cast<ncclComm_t>(g_comm.nccl);
  int const P = g_comm.world;
  if (!g_ce.streams_ready) {
    for (auto& st : g_ce.streams) CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_ce.ev_start, cudaEventDisableTiming));
    for (auto& ev : g_ce.ev_done) CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaMalloc((void**)&g_ce.d_word, 8));
    CUDA_CHECK(cudaMemset(g_ce.d_word, 0, 8));
    CUDA_CHECK(cudaMalloc((void**)&g_ce.d_handles, (size_t)P * 2 * sizeof(cudaIpcMemHandle_t)));
    g_ce.streams_ready = true;
  }
  // nobody may still be reading the old buffers: everything queued so far completes on every rank
  CUDA_CHECK(cudaDeviceSynchronize());
  nccl_check(api().AllReduce(g_ce.d_word, g_ce.d_word, 1, ncclUint64, ncclSum, c, g_comm.stream), "ncclAllReduce (barrier)");
  CUDA_CHECK(cudaStreamSynchronize(g_comm.stream));
  ce_release_buffers(true);
  size_t const cap = chunk_bytes + chunk_bytes / 8;  // some slack: a wider storage type next time need not remap
  // Every rank goes through the same collectives whatever fails locally; the failures are counted at the end.
  unsigned long long failed = 0;
  if (char const* e = std::getenv("SPED_EXCHANGE_TEST_FAIL"))  // test hook: pretend this rank cannot set its side up
    if (*e && std::atoi(e) == g_comm.rank) failed = 1;
  cudaIpcMemHandle_t mine[2];
  std::memset(mine, 0, sizeof mine);
  for (int q = 0; q < 2 && !failed; ++q) {
    if (cudaMalloc(&g_ce.send[q], cap) != cudaSuccess || cudaMemset(g_ce.send[q], 0, cap) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine[q], g_ce.send[q]) != cudaSuccess) {
      cudaGetLastError();
      failed = 1;
    }
  }
  size_t const hb = 2 * sizeof(cudaIpcMemHandle_t);
  CUDA_CHECK(cudaMemcpy(g_ce.d_handles + (size_t)g_comm.rank * hb, mine, hb, cudaMemcpyHostToDevice));
  nccl_check(api().AllGather(g_ce.d_handles + (size_t)g_comm.rank * hb, g_ce.d_handles, hb, ncclChar, c, g_comm.stream),
             "ncclAllGather (IPC handles)");
  unsigned long long* d_failed = nullptr;
  CUDA_CHECK(cudaMalloc((void**)&d_failed, 8));
  CUDA_CHECK(cudaMemcpyAsync(d_failed, &failed, 8, cudaMemcpyHostToDevice, g_comm.stream));
  nccl_check(api().AllReduce(d_failed, d_failed, 1, ncclUint64, ncclSum, c, g_comm.stream), "ncclAllReduce (IPC set-up)");
  unsigned long long failed_anywhere = 0;
  CUDA_CHECK(cudaMemcpyAsync(&failed_anywhere, d_failed, 8, cudaMemcpyDeviceToHost, g_comm.stream));
  CUDA_CHECK(cudaStreamSynchronize(g_comm.stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)P * 2);
  CUDA_CHECK(cudaMemcpy(all.data(), g_ce.d_handles, (size_t)P * hb, cudaMemcpyDeviceToHost));
  if (!failed_anywhere) {  // every rank has its buffers: map the peers', then agree once more
    for (int q = 0; q < 2; ++q) {
      g_ce.peer[q].assign(P, nullptr);
      for (int p = 0; p < P; ++p) {
        if (p == g_comm.rank) {
          g_ce.peer[q][p] = g_ce.send[q];
        } else if (cudaIpcOpenMemHandle(&g_ce.peer[q][p], all[(size_t)p * 2 + q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          g_ce.peer[q][p] = nullptr;
          failed = 1;
        }
      }
    }
    CUDA_CHECK(cudaMemcpyAsync(d_failed, &failed, 8, cudaMemcpyHostToDevice, g_comm.stream));
    nccl_check(api().AllReduce(d_failed, d_failed, 1, ncclUint64, ncclSum, c, g_comm.stream), "ncclAllReduce (IPC mapping)");
    CUDA_CHECK(cudaMemcpyAsync(&failed_anywhere, d_failed, 8, cudaMemcpyDeviceToHost, g_comm.stream));
    CUDA_CHECK(cudaStreamSynchronize(g_comm.stream));
  }
  cudaFree(d_failed);
  if (failed_anywhere) {
    ce_release_buffers(true);
    g_ce.unavailable = true;
    SPED_LOG("copy-engine exchange not available (%llu rank(s) could not allocate or map the send buffers): the exchange stays on NCCL",
             failed_anywhere);
    return false;
  }
  g_ce.bytes = cap;
  SPED_LOG("copy-engine exchange: two send buffers of %.1f MB mapped from %d peers", cap / 1e6, P - 1);
  return true;
}

void comm_ce_publish(void const* shard, size_t bytes, cudaStream_t s) {
  if (bytes) CUDA_CHECK(cudaMemcpyAsync(g_ce.send[g_ce.counter & 1], shard, bytes, cudaMemcpyDeviceToDevice, s));
}

void comm_ce_barrier(cudaStream_t g) {
  nccl_check(api().AllReduce(g_ce.d_word, g_ce.d_word, 1, ncclUint64, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), g),
             "ncclAllReduce (exchange barrier)");
}

void comm_ce_pull_round(void* buf, size_t chunk_bytes, int d_lo, int d_hi, cudaStream_t g) {
  if (d_lo > d_hi) return;
  int const P = g_comm.world, r = g_comm.rank, q = (int)(g_ce.counter & 1);
  char* base = static_cast<char*>(buf);
  CUDA_CHECK(cudaEventRecord(g_ce.ev_start, g));
  int used = 0;
  for (int d = d_lo; d <= d_hi; ++d) {
    int const from = (r + d) % P, k = (d - d_lo) % kCeStreams;
    if (k >= used) {
      CUDA_CHECK(cudaStreamWaitEvent(g_ce.streams[k], g_ce.ev_start, 0));
      used = k + 1;
    }
    CUDA_CHECK(cudaMemcpyAsync(base + (size_t)from * chunk_bytes, g_ce.peer[q][from], chunk_bytes, cudaMemcpyDeviceToDevice,
                               g_ce.streams[k]));
  }
  for (int k = 0; k < used; ++k) {
    CUDA_CHECK(cudaEventRecord(g_ce.ev_done[k], g_ce.streams[k]));
    CUDA_CHECK(cudaStreamWaitEvent(g, g_ce.ev_done[k], 0));
  }
}

void comm_ce_advance() { ++g_ce.counter; }

void comm_allreduce_sum_f64(double* dev, size_t count, cudaStream_t s) {
  if (!g_comm.active() || count == 0) return;
  nccl_check(api().AllReduce(dev, dev, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), s),
             "ncclAllReduce");
}

void comm_allreduce_sum_u64(unsigned long long* dev, size_t count, cudaStream_t s) {
  if (!g_comm.active() || count == 0) return;
  nccl_check(api().AllReduce(dev, dev, count, ncclUint64, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), s),
             "ncclAllReduce");
}

void comm_broadcast_bytes(void* dev, size_t bytes, int root, cudaStream_t s) {
  if (!g_comm.active() || bytes == 0) return;
  nccl_check(api().Broadcast(dev, dev, bytes, ncclChar, root, static_cast<ncclComm_t>(g_comm.nccl), s),
             "ncclBroadcast");
}

void comm_group_start() {
  if (g_comm.active()) nccl_check(api().GroupStart(), "ncclGroupStart");
}
void comm_group_end() {
  if (g_comm.active()) nccl_check(api().GroupEnd(), "ncclGroupEnd");
}

// Block-cyclic row distribution (see RowDist in device_types.h): blocks of 2^log2b rows are dealt
// round-robin, the block size shrinking for small problems so that every rank still gets work.
RowDist make_row_dist(u64 n, int world, int rank) {
  RowDist d{};
  d.n = n;
  d.world = (u32)world;
  d.rank = (u32)rank;
  // large blocks keep the near-diagonal elements (flips of low sites that leave the word canonical)
  // on the owning rank: at 2^16 rows about half of a chain's elements have a local source, which
  // the streaming kernel handles while the all-gather is still in flight; >= 256 blocks per rank
  // keep the ranks balanced
  u32 lb = 16;
  if (char const* e = std::getenv("SPED_LOG2_BLOCK")) lb = (u32)std::max(5, std::min(24, std::atoi(e)));
  while (lb > 5 && (n >> lb) < (u64)world * 256) --lb;
  d.log2b = lb;
  d.chunk = 0;
  for (int r = 0; r < world; ++r) d.chunk = std::max(d.chunk, dist_rows_of(n, (u32)world, (u32)r, lb));
  d.n_local = dist_rows_of(n, (u32)world, (u32)rank, lb);
  if (world == 1) d.chunk = n;
  return d;
}

}  // namespace sped
