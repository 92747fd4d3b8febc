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

#include "internal.h"

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
}

void comm_finalize() {
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
