// internal.h -- host-side object model of libsped (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <complex>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sped.h"
#include "permprog.h"

namespace sped {
using u64 = std::uint64_t;
using i64 = std::int64_t;
using u32 = std::uint32_t;
using cplx = std::complex<double>;
}  // namespace sped

#include "device_types.h"

namespace sped {

extern bool g_logging;
extern std::uint64_t g_launches;
// NVTX ranges around the phases of the hot path (basis build, cache fill, matvec passes, exchange,
// solver phases): visible in Nsight Systems / ncu --nvtx, free when no tool is attached (nvtx3 is
// header-only and resolves its injection library lazily).
struct NvtxRange {
  explicit NvtxRange(char const* name);
  ~NvtxRange();
  NvtxRange(NvtxRange const&) = delete;
  NvtxRange& operator=(NvtxRange const&) = delete;
};
#define SPED_NVTX_CAT2(a, b) a##b
#define SPED_NVTX_CAT(a, b) SPED_NVTX_CAT2(a, b)
#define SPED_NVTX(name) ::sped::NvtxRange SPED_NVTX_CAT(nvtx_range_, __LINE__)(name)

#define SPED_LOG(...)                               \
  do {                                              \
    if (::sped::g_logging) {                        \
      std::fprintf(stderr, "[sped] " __VA_ARGS__);  \
      std::fprintf(stderr, "\n");                   \
    }                                               \
  } while (0)

struct Error {
  int code;
  std::string what;
};
[[noreturn]] void fail(int code, std::string what);
void cuda_check(cudaError_t e, char const* expr, char const* file, int line);
#define CUDA_CHECK(expr) ::sped::cuda_check((expr), #expr, __FILE__, __LINE__)
#define KERNEL_LAUNCHED() (++::sped::g_launches)

// Runs fn(), mapping exceptions to status codes; remembers the message for ls_error_to_string.
int guarded(void (*thunk)(void*), void* ctx);
template <class F>
int guard(F&& f) {
  auto thunk = [](void* p) { (*static_cast<F*>(p))(); };
  return guarded(thunk, &f);
}

// ---- simple RAII device buffer ----
template <class T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t n) { alloc(n); }
  DeviceBuffer(DeviceBuffer const&) = delete;
  DeviceBuffer& operator=(DeviceBuffer const&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : ptr(o.ptr), count(o.count) { o.ptr = nullptr; o.count = 0; }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    if (this != &o) { release(); ptr = o.ptr; count = o.count; o.ptr = nullptr; o.count = 0; }
    return *this;
  }
  ~DeviceBuffer() { release(); }
  void alloc(size_t n) {
    release();
    if (n) CUDA_CHECK(cudaMalloc((void**)&ptr, n * sizeof(T)));
    count = n;
  }
  void release() {
    // destructors may run from a GC finalizer at process exit: ignore errors (context may be gone)
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  void upload(T const* host, size_t n) {
    if (n > count) alloc(n);
    if (n) CUDA_CHECK(cudaMemcpy(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
  }
  void upload(std::vector<T> const& v) { upload(v.data(), v.size()); }
  std::vector<T> download() const {
    std::vector<T> v(count);
    if (count) CUDA_CHECK(cudaMemcpy(v.data(), ptr, count * sizeof(T), cudaMemcpyDeviceToHost));
    return v;
  }
};

// ---- symmetry / group (host only; group.cpp) ----
struct Symmetry {
  std::vector<unsigned> perm;
  unsigned sector = 0;
  unsigned periodicity = 1;
};

struct GroupElement {
  std::vector<int> perm;  // (g.x)[i] = x[perm[i]]
  i64 phase = 0;          // numerator over Group::denom
};

struct Group {
  unsigned n = 0;  // permutation length; 0 = trivial group without generators
  i64 denom = 2;   // common denominator of all phases (always even)
  std::vector<GroupElement> elems;  // elems[0] = identity
  bool real_characters() const;
};

std::shared_ptr<Symmetry> make_symmetry(unsigned n, unsigned const* perm, unsigned sector);
std::shared_ptr<Group> make_group(std::vector<Symmetry const*> const& gens);

// Compile the traversal of all group elements into a PermProgram (permprog.cpp).
HostProgram compile_program(Group const& g, unsigned n_spins, int spin_inversion);

// Sector dimension by character-weighted Burnside counting (exact integers); group.cpp.
u64 burnside_dimension(Group const& g, unsigned n_spins, int hamming_weight, int spin_inversion);

// ---- communicator (comm.cpp) ----
struct Comm {
  int world = 1;
  int rank = 0;
  void* nccl = nullptr;  // ncclComm_t
  cudaStream_t stream = nullptr;
  cudaStream_t gather_stream = nullptr;           // the all-gather of a sharded matvec runs here ...
  cudaEvent_t ev_ready = nullptr, ev_gathered = nullptr;  // ... between these two events
  cudaEvent_t ev_round1 = nullptr;                // end of the first exchange round (three source classes)
  bool active() const { return world > 1; }
};
Comm& comm();
void comm_unique_id(void* out128);
void comm_init(int world, int rank, void const* id128);
void comm_finalize();
// in-place all-gather: every rank owns chunk `rank` of `chunk_bytes` inside buf (P chunks)
void comm_allgather_inplace(void* buf, size_t chunk_bytes, cudaStream_t s);
// one round of the shard exchange: this rank's chunk goes to ranks rank-d and the chunks of ranks
// rank+d arrive in place, for d in [d_lo, d_hi] (mod world); grouped NCCL send/recv
void comm_exchange_round(void* buf, size_t chunk_bytes, int d_lo, int d_hi, cudaStream_t s);
// ---- shard exchange by the copy engines over peer memory (comm.cpp) ----
// Every rank publishes its shard in an IPC-shared send buffer (two of them, used in turn), a
// one-word all-reduce tells everybody that all shards are in place, and each rank PULLS the shards of
// its peers with cudaMemcpyAsync on a few copy streams: the copy engines move the data over NVLink,
// no SM runs a transfer kernel beside the gather-bound passes of the product.
bool comm_ce_wanted(size_t chunk_bytes);      // same answer on every rank: world, shard size, environment
bool comm_ce_prepare(size_t chunk_bytes);     // collective: (re)allocate and map the send buffers when they grow;
                                              // false on every rank alike when that is not possible (stay on NCCL)
void comm_ce_publish(void const* shard, size_t bytes, cudaStream_t s);
void comm_ce_barrier(cudaStream_t g);         // all shards of this exchange are published
// shards of ranks rank+d, d in [d_lo, d_hi], into their places of buf; complete in stream order on g
void comm_ce_pull_round(void* buf, size_t chunk_bytes, int d_lo, int d_hi, cudaStream_t g);
void comm_ce_advance();                       // next exchange uses the other send buffer
void comm_allreduce_sum_f64(double* dev, size_t count, cudaStream_t s);
void comm_allreduce_sum_u64(unsigned long long* dev, size_t count, cudaStream_t s);
void comm_broadcast_bytes(void* dev, size_t bytes, int root, cudaStream_t s);
void comm_group_start();
void comm_group_end();
RowDist make_row_dist(u64 n, int world, int rank);  // block-cyclic row distribution (device_types.h)

// ---- basis (basis.cu) ----
struct Basis {
  std::shared_ptr<Group> group;
  unsigned n_spins = 0;
  int hamming_weight = -1;
  int spin_inversion = 0;
  HostProgram program;                // canonicalisation program (empty for the trivial group)
  DeviceBuffer<unsigned char> d_program;  // packed device image of `program`
  DeviceBuffer<double> d_norm_table;  // norm_table[s] = sqrt(s / |G'|)
  DeviceBuffer<double> d_chi_table;   // (cos, sin)(2 pi k / denom), k < denom

  std::mutex mutex;
  bool built = false;
  u64 n_states = 0;
  DeviceBuffer<u64> d_reps;
  DeviceBuffer<std::uint16_t> d_stab;
  DeviceBuffer<unsigned char> d_bucket;
  BasisIndex index;
  double build_seconds = 0.0;
  std::shared_ptr<std::vector<u64>> host_reps;  // lazily mirrored for ls_get_states
  std::shared_ptr<void> jit_cache;              // run-time specialised kernels (jit.cpp)
  u64 generation = 0;                           // bumped by every (re)build

  bool trivial() const { return group->elems.size() <= 1 && spin_inversion == 0; }
  u64 group_order() const { return (u64)group->elems.size() * (spin_inversion != 0 ? 2 : 1); }
  bool use32() const { return n_spins <= 32; }
  u64 expected_dimension() const;
  void ensure_device_tables();
  void build();                                     // K1
  void adopt(u64 size, u64 const* host_reps_in);    // K1b
  void finish_build();                              // stabilisers + bucket table
  std::shared_ptr<std::vector<u64>> states_host();
  RowDist dist() const { return make_row_dist(n_states, comm().world, comm().rank); }
};

std::shared_ptr<Basis> make_basis(std::shared_ptr<Group> g, unsigned n_spins, int hw, int inv);
void jit_prefetch(Basis& b);  // jit.cpp: start compiling the specialised cache-fill kernel in the background

// ---- interactions / operator (operator.cu) ----
struct Interaction {
  int k = 0;
  std::vector<cplx> matrix;            // row-major 2^k x 2^k
  std::vector<std::uint16_t> sites;    // count * k
  bool is_real() const;
};

struct EighStats {
  u64 matvecs = 0;
  int iterations = 0, restarts = 0;
  double seconds_total = 0, seconds_matvec = 0, seconds_ortho = 0;
  double seconds_residual = 0, seconds_restart = 0, seconds_project = 0;
};

struct Operator {
  std::shared_ptr<Basis> basis;
  std::vector<Interaction> terms;
  bool real_matrices = true;
  bool real_diagonal = true;

  // device image (prepared lazily once the basis is built; re-prepared if the basis is rebuilt)
  std::mutex mutex;
  u64 prepared_generation = ~0ull;
  DeviceBuffer<unsigned char> d_terms;  // packed bonds + matrices (see operator.cu)
  DeviceBuffer<double> d_diag;          // local rows (real part) [+ imaginary part if !real_diagonal]
  DeviceBuffer<unsigned char> stage_x, stage_y;  // grow-only device staging of the host-pointer entry
  DeviceBuffer<unsigned char> block_x;           // grow-only: interleaved copy of a block of vectors (opcache.cu)
  // grow-only device workspace of sped_eigh (Krylov basis, H V, residuals, ...), kept between calls:
  // cudaFree of multi-gigabyte buffers costs 0.1-0.8 s of wall time, more than a warm 6x6 solve
  DeviceBuffer<unsigned char> eigh_ws[12];
  void* workspace(int slot, size_t bytes) {
    if (eigh_ws[slot].count < bytes) eigh_ws[slot].alloc(bytes);
    return eigh_ws[slot].ptr;
  }
  void release_workspace() {
    for (auto& b : eigh_ws) b.release();
  }

  // operator cache (opcache.cu): off-diagonal elements of the local rows resident in HBM
  int cache_mode = -1;        // -1: auto (build when it fits), 0: never, 1: always try
  bool cache_ready = false;
  bool cache_rejected = false;  // decided not to (or failed to) build for this basis generation
  DeviceBuffer<u64> c_slice_off;
  DeviceBuffer<u32> c_idx;
  DeviceBuffer<unsigned char> c_code;  // u8 or u16 per CODED element (compact, see CacheView::code_off)
  DeviceBuffer<u64> c_code_off;        // [slices * classes + 1]
  int c_code_wide = 0;
  DeviceBuffer<std::uint16_t> c_len;   // [2 * c_classes][local rows]: default-coefficient / coded elements per source class
  DeviceBuffer<u32> c_slice_start;     // [slices][kClassStride] first slot of classes 1, 2 (several classes only)
  u32 c_classes = 1, c_near = 0, c_default_code = 0, c_rounds = 0;
  DeviceBuffer<double> c_table;
  u64 c_slices = 0, c_slots = 0, cache_bytes = 0;
  double cache_build_seconds = 0;
  bool cache_usable();        // true once the cache is (or has just been) built
  void drop_cache();
  // local rows [row_lo, row_hi) only (row_lo a multiple of 32); y is still indexed by local row
  // phase: 0 all classes, 1 + c: source class c only (class 0 starts from the diagonal, later ones add to y)
  void cached_matmat(int dtype, u64 block, void const* x, u64 xs, void* y, u64 ys, cudaStream_t s, u64 row_lo = 0,
                     u64 row_hi = ~(u64)0, int phase = 0, bool beside_transfer = false);
  // y_local = H x for one column given only this rank's shard of x: all-gather into `xfull`
  // ([rank][local] layout, world * chunk entries) overlapped with the local-source pass
  void matvec_sharded(int dtype, void const* x_local, void* y_local, void* xfull, cudaStream_t s);
  cudaStream_t pipe_compute = nullptr, pipe_copy = nullptr;  // host-pointer entry: kernel / D2H pipeline
  cudaEvent_t pipe_events[8] = {};
  ~Operator() {  // may run from a GC finalizer at process exit: errors are ignored
    for (auto e : pipe_events)
      if (e) cudaEventDestroy(e);
    if (pipe_compute) cudaStreamDestroy(pipe_compute);
    if (pipe_copy) cudaStreamDestroy(pipe_copy);
  }
  void cached_count(unsigned long long* d_out);  // adds the local element count to *d_out
  RowDist dist{};  // rows of this rank (fixed at prepare())
  bool counted = false;
  u64 n_offdiag = 0;
  EighStats last_stats;

  bool is_real() const { return real_matrices && basis->group->real_characters(); }
  void prepare();
  void matmat_device(int dtype, u64 block, void const* x, u64 xs, void* y, u64 ys, cudaStream_t s);
  void matmat_host(int dtype, u64 size, u64 block, void const* x, u64 xs, void* y, u64 ys);
  // host blocks of this rank's rows only (sped_operator_matmat_local)
  void matmat_host_local(int dtype, u64 block, void const* x, u64 xs, void* y, u64 ys);
  void expectation_host(int dtype, u64 size, u64 block, void const* x, u64 xs, cplx* out);
  void count_elements(u64& rows, u64& offdiag);
};

int exchange_rounds(unsigned world, u64 chunk);  // opcache.cu
unsigned exchange_near(unsigned world, u64 chunk);

// Code maps of the operator cache (opcache.cu, build_code_maps)
struct CodeMaps {
  std::vector<double> values_re, values_im;                            // distinct off-diagonal matrix values
  std::vector<std::uint16_t> hid_map, sid_map, sid_stab, pid_map, pid_phase;
  u32 n_pid = 1, default_code = 0;
  u64 n_codes = 0;
};
char const* build_code_maps(Operator const& op, CodeMaps& out);
// host copy of the packed terms (operator.cu): bonds in the order the kernels visit them
void packed_terms_host(std::vector<Interaction> const& terms, std::vector<DevBond>& bonds, std::vector<double>& pool_re,
                       std::vector<double>& pool_im, std::vector<std::uint16_t>& masks);

std::shared_ptr<Interaction> make_interaction(int k, void const* matrix, unsigned count, std::uint16_t const* sites);
std::shared_ptr<Operator> make_operator(std::shared_ptr<Basis> b, std::vector<Interaction const*> const& terms);

inline size_t dtype_size(int dtype) {
  switch (dtype) {
    case SPED_F32: return 4;
    case SPED_F64: return 8;
    case SPED_C64: return 8;
    case SPED_C128: return 16;
  }
  fail(LS_INVALID_DATATYPE, "unknown datatype tag");
}
inline bool dtype_is_complex(int dtype) { return dtype == SPED_C64 || dtype == SPED_C128; }

// ---- eigensolver (eigh.cu) ----
int eigh(Operator& op, int dtype, u64 n_evals, double eps, int max_basis, int max_block, int min_restart,
         double* evals, void* evecs, double* rnorms, sped_monitor_fn monitor, void* ctx);

}  // namespace sped
