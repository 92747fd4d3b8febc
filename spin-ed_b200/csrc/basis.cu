// basis.cu -- representative-basis construction on the device (kernels K1, K1b of SURVEY 2.2).
//
// Replaces ls_build / ls_build_unsafe / ls_get_states of liblattice_symmetries as called from
// /root/reference/src/SpinED/Internal.hs:178-196,230-244.  Candidates of the sector are walked in
// increasing integer order (Gosper's next-bit-permutation from an unranked start), each candidate
// is tested with the canonicalisation program (early exit on the first smaller image), survivors
// are compacted in order:  mark (one bit per candidate) -> block counts -> scan -> write.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "device_common.cuh"

namespace sped {

namespace {

constexpr int kCandPerThread = 64;                       // one 64-bit survivor mask per thread
constexpr u64 kCandPerBlock = (u64)kThreads * kCandPerThread;

struct Binomials {
  u64 c[65 * 65];
};

__device__ __forceinline__ u64 binom_at(u64 const* table, int n, int k) { return __ldg(table + n * 65 + k); }

// rank (increasing integer order among words with k bits set) -> word
__device__ u64 unrank_word(u64 const* table, u64 r, int k, int n) {
  u64 x = 0;
  int c = n;
  for (int i = k; i >= 1; --i) {
    u64 b;
    do {
      --c;
      b = binom_at(table, c, i);
    } while (b > r);
    x |= 1ull << c;
    r -= b;
  }
  return x;
}

__device__ __forceinline__ u64 next_same_popcount(u64 x) {
  u64 t = x | (x - 1);
  return (t + 1) | (((~t & (0 - ~t)) - 1) >> (__ffsll((long long)x)));
}

struct EnumParams {
  u64 rank_lo;     // first candidate rank of this launch
  u64 rank_hi;     // one past the last
  int hamming_weight;
  int n_spins;
  u64 const* binom;
};

__device__ __forceinline__ u64 first_candidate(EnumParams const& p, u64 rank) {
  return p.hamming_weight >= 0 ? (p.hamming_weight == 0 ? 0ull : unrank_word(p.binom, rank, p.hamming_weight, p.n_spins))
                               : rank;
}
__device__ __forceinline__ u64 next_candidate(EnumParams const& p, u64 x) {
  return p.hamming_weight >= 0 ? next_same_popcount(x) : x + 1;
}

// Pass 1: survivor mask per thread + survivor count per block.
template <class W>
__global__ void __launch_bounds__(kThreads) mark_kernel(EnumParams p, ProgramView<W> prog, bool staged, u64* marks,
                                                        u32* block_counts) {
  extern __shared__ __align__(16) unsigned char smem[];
  ProgramView<W> P = stage_program<W>(prog, smem);
  u64 tid = (u64)blockIdx.x * kThreads + threadIdx.x;
  u64 r0 = p.rank_lo + tid * kCandPerThread;
  u64 mask = 0;
  if (r0 < p.rank_hi) {
    u64 cnt = min((u64)kCandPerThread, p.rank_hi - r0);
    u64 x = first_candidate(p, r0);
    for (u64 j = 0; j < cnt; ++j) {
      if (stabilizer_scan<W>(P, (W)x, true) > 0) mask |= 1ull << j;
      if (j + 1 < cnt) x = next_candidate(p, x);
    }
  }
  marks[tid] = mask;
  int c = __popcll(mask);
  __shared__ int warp_sums[kThreads / 32];
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += warp_sums[w];
    block_counts[blockIdx.x] = (u32)s;
  }
}

// Pass 2: exclusive scan of the block counts, offset by the running total (single block).
__global__ void __launch_bounds__(1024) scan_kernel(u32 const* counts, u64* offsets, u32 n, u64* running_total,
                                                    u64* tile_count) {
  __shared__ u64 partial[1024];
  u32 per = (n + 1023) / 1024;
  u32 lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  u64 s = 0;
  for (u32 i = lo; i < hi; ++i) s += counts[i];
  partial[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 run = *running_total;
    for (int i = 0; i < 1024; ++i) {
      u64 v = partial[i];
      partial[i] = run;
      run += v;
    }
    *tile_count = run - *running_total;
    *running_total = run;
  }
  __syncthreads();
  u64 run = partial[threadIdx.x];
  for (u32 i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
}

// Pass 3: re-walk the candidates and write the marked ones at their final positions.
__global__ void __launch_bounds__(kThreads) write_kernel(EnumParams p, u64 const* marks, u64 const* block_offsets,
                                                         u64* out, u64 capacity, int* overflow) {
  u64 tid = (u64)blockIdx.x * kThreads + threadIdx.x;
  u64 mask = marks[tid];
  int c = __popcll(mask);
  // exclusive scan of c over the block
  __shared__ int warp_sums[kThreads / 32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_sums[w];
  u64 pos = block_offsets[blockIdx.x] + (u64)(before + incl - c);
  if (mask == 0) return;
  u64 r0 = p.rank_lo + tid * kCandPerThread;
  u64 x = first_candidate(p, r0);
  int last = 63 - __clzll((long long)mask);
  for (int j = 0; j <= last; ++j) {
    if ((mask >> j) & 1ull) {
      if (pos < capacity) out[pos] = x;
      else *overflow = 1;
      ++pos;
    }
    if (j < last) x = next_candidate(p, x);
  }
}

// Trivial group with a hamming weight: every candidate is a representative.
__global__ void __launch_bounds__(kThreads) enumerate_all_kernel(EnumParams p, u64* out) {
  u64 tid = (u64)blockIdx.x * kThreads + threadIdx.x;
  u64 r0 = p.rank_lo + tid * kCandPerThread;
  if (r0 >= p.rank_hi) return;
  u64 cnt = min((u64)kCandPerThread, p.rank_hi - r0);
  u64 x = first_candidate(p, r0);
  for (u64 j = 0; j < cnt; ++j) {
    out[r0 + j] = x;
    if (j + 1 < cnt) x = next_candidate(p, x);
  }
}

// |Stab| of every representative (0 would mean zero norm: impossible for a valid basis).
template <class W>
__global__ void __launch_bounds__(kThreads) stabilizer_kernel(ProgramView<W> prog, bool staged, u64 const* reps,
                                                              u64 n, std::uint16_t* stab, int* invalid) {
  extern __shared__ __align__(16) unsigned char smem[];
  ProgramView<W> P = stage_program<W>(prog, smem);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    int s = stabilizer_scan<W>(P, (W)reps[i], true);
    if (s <= 0) {
      *invalid = 1;
      s = 0;
    }
    stab[i] = (std::uint16_t)s;
  }
}

// bucket[q] = first index whose prefix is >= q, for q in [0, bucket_count]: one thread per bucket,
// a binary search each (neighbouring buckets walk neighbouring paths, so the probes stay in L2).
// Orbit minima crowd the low end of the word range, so most buckets are empty: filling them from the
// element side (one thread writing every empty bucket up to the next occupied one) serialised
// millions of stores in single threads -- 38 ms for 6x6, a fifth of the basis build.
template <class I>
__global__ void __launch_bounds__(kThreads) bucket_kernel(u64 const* reps, u64 n, int shift, u32 bucket_count, I* bucket) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q <= bucket_count; q += (u64)gridDim.x * blockDim.x) {
    u64 lo = 0, hi = n;
    while (lo < hi) {
      u64 const mid = lo + ((hi - lo) >> 1);
      if ((__ldg(reps + mid) >> shift) < q) lo = mid + 1;
      else hi = mid;
    }
    bucket[q] = (I)lo;
  }
}

__global__ void __launch_bounds__(kThreads) sorted_check_kernel(u64 const* reps, u64 n, int* unsorted) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (u64)gridDim.x * blockDim.x)
    if (reps[i - 1] >= reps[i]) *unsorted = 1;
}

template <class W>
__global__ void __launch_bounds__(kThreads) state_info_kernel(ProgramView<W> prog, bool staged, u64 const* states,
                                                              u64 n, u64* reps, std::int32_t* phases, int* stabs) {
  extern __shared__ __align__(16) unsigned char smem[];
  ProgramView<W> P = stage_program<W>(prog, smem);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    W rep;
    u32 step, flipped;
    canonicalize<W>(P, (W)states[i], rep, step, flipped);
    reps[i] = rep;
    phases[i] = element_phase<W>(P, step, flipped);
    stabs[i] = stabilizer_scan<W>(P, rep, false);
  }
}

Binomials make_binomials() {
  Binomials b;
  for (int i = 0; i < 65; ++i)
    for (int j = 0; j < 65; ++j) b.c[i * 65 + j] = 0;
  for (int i = 0; i < 65; ++i) {
    b.c[i * 65] = 1;
    for (int j = 1; j <= i; ++j) {
      // saturate instead of overflowing (only C(64, 28..36)-ish could, and they do not: max 1.8e18)
      b.c[i * 65 + j] = b.c[(i - 1) * 65 + j - 1] + (j <= i - 1 ? b.c[(i - 1) * 65 + j] : 0);
    }
  }
  return b;
}

template <class W>
struct DeviceProgram {
  ProgramView<W> view;
  size_t smem;
  bool staged;
};

constexpr size_t kMaxStagedProgram = 160 * 1024;

}  // namespace

// device image of the program: [steps][phase][ops32][ops64][fast32][fast64]
namespace {
struct ProgramLayout {
  size_t steps, phase, ops32, ops64, fast32, fast64, total;
};
ProgramLayout program_layout(HostProgram const& P) {
  auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
  ProgramLayout L;
  L.steps = 0;
  L.phase = L.steps + up(P.steps.size() * sizeof(PermStep));
  L.ops32 = L.phase + up(P.phase.size() * sizeof(std::int32_t));
  L.ops64 = L.ops32 + up(P.ops.size() * sizeof(PermOp<u32>));
  L.fast32 = L.ops64 + up(P.ops.size() * sizeof(PermOp<u64>));
  L.fast64 = L.fast32 + up(P.fast.size() * sizeof(FastStep<u32>));
  L.total = L.fast64 + up(P.fast.size() * sizeof(FastStep<u64>));
  return L;
}
}  // namespace

template <class W>
static DeviceProgram<W> device_program(Basis const& b) {
  DeviceProgram<W> d;
  auto const& P = b.program;
  ProgramLayout L = program_layout(P);
  unsigned char* base = b.d_program.ptr;
  d.view.fast = reinterpret_cast<FastStep<W> const*>(base + (sizeof(W) == 4 ? L.fast32 : L.fast64));
  d.view.steps = reinterpret_cast<PermStep const*>(base + L.steps);
  d.view.ops = reinterpret_cast<PermOp<W> const*>(base + (sizeof(W) == 4 ? L.ops32 : L.ops64));
  d.view.phase = reinterpret_cast<std::int32_t const*>(base + L.phase);
  d.view.n_steps = (u32)P.steps.size();
  d.view.n_ops = (u32)P.ops.size();
  d.view.n_spins = P.n_spins;
  d.view.shift = sizeof(W) == 4 ? 0 : P.shift;
  d.view.inversion = P.inversion;
  d.view.denom = P.denom;
  d.smem = program_smem_bytes<W>(d.view.n_steps, d.view.n_ops);
  if (d.smem > kMaxStagedProgram) fail(LS_INVALID_ARGUMENT, "symmetry group too large to stage its program in shared memory");
  d.staged = true;
  return d;
}
template DeviceProgram<u32> device_program<u32>(Basis const&);
template DeviceProgram<u64> device_program<u64>(Basis const&);

// exported to operator.cu
ProgramView<u32> program_view32(Basis const& b, size_t& smem, bool& staged) {
  auto d = device_program<u32>(b);
  smem = d.smem;
  staged = d.staged;
  return d.view;
}
ProgramView<u64> program_view64(Basis const& b, size_t& smem, bool& staged) {
  auto d = device_program<u64>(b);
  smem = d.smem;
  staged = d.staged;
  return d.view;
}

u64 Basis::expected_dimension() const {
  return burnside_dimension(*group, n_spins, hamming_weight, spin_inversion);
}

void Basis::ensure_device_tables() {
  if (d_norm_table.ptr) return;
  auto const& P = program;
  ProgramLayout L = program_layout(P);
  std::vector<PermOp<u32>> ops32;
  for (auto const& o : P.ops) ops32.push_back(PermOp<u32>{(u32)o.mask, o.amount});
  std::vector<FastStep<u32>> fast32;
  for (auto const& f : P.fast) fast32.push_back(FastStep<u32>{(u32)f.mask, f.ctl});
  std::vector<unsigned char> img(L.total, 0);
  std::memcpy(img.data() + L.steps, P.steps.data(), P.steps.size() * sizeof(PermStep));
  std::memcpy(img.data() + L.phase, P.phase.data(), P.phase.size() * sizeof(std::int32_t));
  if (!P.ops.empty()) {
    std::memcpy(img.data() + L.ops32, ops32.data(), ops32.size() * sizeof(PermOp<u32>));
    std::memcpy(img.data() + L.ops64, P.ops.data(), P.ops.size() * sizeof(PermOp<u64>));
  }
  std::memcpy(img.data() + L.fast32, fast32.data(), fast32.size() * sizeof(FastStep<u32>));
  std::memcpy(img.data() + L.fast64, P.fast.data(), P.fast.size() * sizeof(FastStep<u64>));
  d_program.upload(img);
  u64 order = group_order();
  std::vector<double> norms(order + 1);
  for (u64 s = 0; s <= order; ++s) norms[s] = std::sqrt((double)s / (double)order);
  d_norm_table.upload(norms);
  i64 D = group->denom;
  std::vector<double> chi(2 * D);
  for (i64 k = 0; k < D; ++k) {
    double ang = 2.0 * M_PI * (double)k / (double)D;
    double c = std::cos(ang), s = std::sin(ang);
    if (k == 0) { c = 1; s = 0; }
    else if (2 * k == D) { c = -1; s = 0; }
    else if (4 * k == D) { c = 0; s = 1; }
    else if (4 * k == 3 * D) { c = 0; s = -1; }
    chi[2 * k] = c;
    chi[2 * k + 1] = s;
  }
  d_chi_table.upload(chi);
}

static u64 binom_host(int n, int k) {
  static Binomials const b = make_binomials();
  return (k < 0 || k > n) ? 0 : b.c[n * 65 + k];
}

void Basis::build() {
  std::lock_guard<std::mutex> lock(mutex);
  SPED_NVTX("sped: ls_build (enumerate representatives)");
  jit_prefetch(*this);  // NVRTC of the cache-fill kernel runs on its own thread beside the enumeration
  ensure_device_tables();
  auto t0 = std::chrono::steady_clock::now();
  Comm& cm = comm();
  u64 const expected = expected_dimension();
  if (hamming_weight < 0 && n_spins >= 63) fail(LS_INVALID_NUMBER_SPINS, "unrestricted bases need fewer than 63 spins");
  u64 const total = hamming_weight >= 0 ? binom_host((int)n_spins, hamming_weight) : (1ull << n_spins);
  SPED_LOG("ls_build: %llu candidates, %llu expected representatives, |G'| = %llu",
           (unsigned long long)total, (unsigned long long)expected, (unsigned long long)group_order());
  static Binomials const host_binom = make_binomials();
  DeviceBuffer<u64> d_binom(65 * 65);
  d_binom.upload(host_binom.c, 65 * 65);
  EnumParams ep{0, total, hamming_weight, (int)n_spins, d_binom.ptr};

  // a failed (re)build must not leave a half-initialised basis marked as built
  built = false;
  index = BasisIndex{};
  d_reps.release();
  d_stab.release();
  host_reps.reset();
  // Sharding the enumeration over the ranks pays only for large sectors: below ~3e10 candidates one
  // GPU walks the sector in well under a second, less than the exchange of the survivors costs
  // (6x6, 9.1e9 candidates: 0.2 s on one GPU, 0.6-1.4 s sharded over 2-8).  Smaller sectors are
  // enumerated redundantly by every rank -- no communication at all.  The choice depends on the
  // sector and the environment only, so every rank makes the same one.
  u64 shard_min = 1ull << 35;
  if (char const* e = std::getenv("SPED_BUILD_SHARD_MIN")) shard_min = std::strtoull(e, nullptr, 10);
  bool const sharded = cm.active() && total >= shard_min;
  int const b_world = sharded ? cm.world : 1, b_rank = sharded ? cm.rank : 0;
  if (trivial()) {
    if (hamming_weight >= 0) {
      d_reps.alloc(total);
      u64 threads = (total + kCandPerThread - 1) / kCandPerThread;
      u64 blocks = (threads + kThreads - 1) / kThreads;
      u64 const max_blocks = 1u << 30;
      for (u64 b0 = 0; b0 < blocks; b0 += max_blocks) {
        EnumParams e = ep;
        e.rank_lo = b0 * kCandPerBlock;
        u64 nb = std::min(max_blocks, blocks - b0);
        enumerate_all_kernel<<<(unsigned)nb, kThreads>>>(e, d_reps.ptr);
        KERNEL_LAUNCHED();
      }
      CUDA_CHECK(cudaGetLastError());
    }
    n_states = total;
  } else {
    d_reps.alloc(std::max<u64>(expected, 1));
    // tiles of candidate ranks, interleaved over ranks
    u64 tile = 1ull << 30;
    u64 want_tiles = (u64)b_world * 8;
    while (tile > kCandPerBlock * 16 && (total + tile - 1) / tile < want_tiles) tile >>= 1;
    u64 n_tiles = (total + tile - 1) / tile;
    u64 blocks_per_tile = (tile + kCandPerBlock - 1) / kCandPerBlock;
    DeviceBuffer<u64> d_marks(blocks_per_tile * kThreads);
    DeviceBuffer<u32> d_counts(blocks_per_tile);
    DeviceBuffer<u64> d_offsets(blocks_per_tile);
    DeviceBuffer<u64> d_scalars(2 + n_tiles);  // [0] running total, [1] scratch, [2+t] tile counts
    DeviceBuffer<int> d_flag(1);
    CUDA_CHECK(cudaMemset(d_scalars.ptr, 0, (2 + n_tiles) * sizeof(u64)));
    CUDA_CHECK(cudaMemset(d_flag.ptr, 0, sizeof(int)));
    DeviceBuffer<u64> d_staging;
    u64* out = d_reps.ptr;
    if (sharded) {
      d_staging.alloc(std::max<u64>(expected, 1));
      out = d_staging.ptr;
    }
    bool use32_ = use32();
    auto dp32 = device_program<u32>(*this);
    auto dp64 = device_program<u64>(*this);
    size_t smem = use32_ ? dp32.smem : dp64.smem;
    if (smem > 48 * 1024) {
      if (use32_) CUDA_CHECK(cudaFuncSetAttribute(mark_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else CUDA_CHECK(cudaFuncSetAttribute(mark_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    for (u64 t = 0; t < n_tiles; ++t) {
      if ((int)(t % (u64)b_world) != b_rank) continue;
      EnumParams e = ep;
      e.rank_lo = t * tile;
      e.rank_hi = std::min(total, (t + 1) * tile);
      u64 nblocks = (e.rank_hi - e.rank_lo + kCandPerBlock - 1) / kCandPerBlock;
      if (use32_) mark_kernel<u32><<<(unsigned)nblocks, kThreads, smem>>>(e, dp32.view, dp32.staged, d_marks.ptr, d_counts.ptr);
      else mark_kernel<u64><<<(unsigned)nblocks, kThreads, smem>>>(e, dp64.view, dp64.staged, d_marks.ptr, d_counts.ptr);
      KERNEL_LAUNCHED();
      scan_kernel<<<1, 1024>>>(d_counts.ptr, d_offsets.ptr, (u32)nblocks, d_scalars.ptr, d_scalars.ptr + 2 + t);
      KERNEL_LAUNCHED();
      write_kernel<<<(unsigned)nblocks, kThreads>>>(e, d_marks.ptr, d_offsets.ptr, out, expected, d_flag.ptr);
      KERNEL_LAUNCHED();
    }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
    int overflow = 0;
    CUDA_CHECK(cudaMemcpy(&overflow, d_flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (sharded) {
      // every rank learns every tile's count, then each tile is broadcast from its owner into place
      comm_allreduce_sum_u64(reinterpret_cast<unsigned long long*>(d_scalars.ptr + 2), n_tiles, cm.stream);
      CUDA_CHECK(cudaStreamSynchronize(cm.stream));
      std::vector<u64> sc = d_scalars.download();
      std::vector<u64> local_off(cm.world, 0);
      u64 global_off = 0;
      u64 sum = 0;
      for (u64 t = 0; t < n_tiles; ++t) sum += sc[2 + t];
      if (sum != expected) overflow = 1;
      if (!overflow) {
        comm_group_start();
        for (u64 t = 0; t < n_tiles; ++t) {
          if (t && t % 64 == 0) {  // keep NCCL groups to a moderate number of operations
            comm_group_end();
            comm_group_start();
          }
          int owner = (int)(t % (u64)cm.world);
          u64 cnt = sc[2 + t];
          if (cnt) {
            // root sends from its staging area; everyone (root included) receives into d_reps
            void* buf = d_reps.ptr + global_off;
            if (owner == cm.rank)
              CUDA_CHECK(cudaMemcpyAsync(buf, d_staging.ptr + local_off[owner], cnt * 8, cudaMemcpyDeviceToDevice, cm.stream));
            comm_broadcast_bytes(buf, cnt * 8, owner, cm.stream);
          }
          local_off[owner] += cnt;
          global_off += cnt;
        }
        comm_group_end();
        CUDA_CHECK(cudaStreamSynchronize(cm.stream));
      }
      n_states = sum;
    } else {
      u64 produced = 0;
      CUDA_CHECK(cudaMemcpy(&produced, d_scalars.ptr, sizeof(u64), cudaMemcpyDeviceToHost));
      n_states = produced;
    }
    if (overflow || n_states != expected)
      fail(SPED_INTERNAL_ERROR, "representative count disagrees with the Burnside dimension of the sector");
  }
  finish_build();
  CUDA_CHECK(cudaDeviceSynchronize());
  build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  SPED_LOG("ls_build: %llu representatives in %.3f s", (unsigned long long)n_states, build_seconds);
}

void Basis::adopt(u64 size, u64 const* reps_in) {
  std::lock_guard<std::mutex> lock(mutex);
  SPED_NVTX("sped: ls_build_unsafe (adopt representatives)");
  jit_prefetch(*this);
  ensure_device_tables();
  auto t0 = std::chrono::steady_clock::now();
  built = false;  // a failing ls_build_unsafe must not leave the previous basis marked as built
  index = BasisIndex{};
  host_reps.reset();
  d_stab.release();
  if (trivial() && hamming_weight < 0) {
    if (size != (1ull << n_spins)) fail(LS_DIMENSION_MISMATCH, "representatives do not span the full space");
    d_reps.release();
  } else {
    d_reps.alloc(std::max<u64>(size, 1));
    d_reps.upload(reps_in, size);
  }
  n_states = size;
  finish_build();
  CUDA_CHECK(cudaDeviceSynchronize());
  build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void Basis::finish_build() {
  SPED_NVTX("sped: stabilisers + prefix buckets");
  index = BasisIndex{};
  index.n_states = n_states;
  if (trivial() && hamming_weight < 0) {
    index.direct = 1;
    built = true;
    ++generation;
    return;
  }
  DeviceBuffer<int> d_flag(2);
  CUDA_CHECK(cudaMemset(d_flag.ptr, 0, 2 * sizeof(int)));
  if (!trivial() && n_states) {
    d_stab.alloc(n_states);
    int grid = persistent_grid(n_states, kThreads, 8);
    if (use32()) {
      auto d = device_program<u32>(*this);
      if (d.smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(stabilizer_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem));
      stabilizer_kernel<u32><<<grid, kThreads, d.smem>>>(d.view, d.staged, d_reps.ptr, n_states, d_stab.ptr, d_flag.ptr);
    } else {
      auto d = device_program<u64>(*this);
      if (d.smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(stabilizer_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem));
      stabilizer_kernel<u64><<<grid, kThreads, d.smem>>>(d.view, d.staged, d_reps.ptr, n_states, d_stab.ptr, d_flag.ptr);
    }
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
  }
  // prefix bucket table over the significant bits of the largest representative
  u64 max_rep = 0;
  if (n_states) CUDA_CHECK(cudaMemcpy(&max_rep, d_reps.ptr + (n_states - 1), 8, cudaMemcpyDeviceToHost));
  int top = 1;
  while (top < 64 && (max_rep >> top) != 0) ++top;
  int want = 4;
  while (want < 27 && (1ull << want) < 2 * n_states) ++want;
  int pbits = std::min(top, want);
  index.bucket_shift = top - pbits;
  index.bucket_count = 1u << pbits;
  index.bucket_wide = n_states >= 0xffffffffull ? 1 : 0;
  size_t entry = index.bucket_wide ? 8 : 4;
  d_bucket.alloc(((size_t)index.bucket_count + 1) * entry);
  int grid = persistent_grid((u64)index.bucket_count + 1, kThreads, 8);
  if (index.bucket_wide)
    bucket_kernel<u64><<<grid, kThreads>>>(d_reps.ptr, n_states, index.bucket_shift, index.bucket_count,
                                           reinterpret_cast<u64*>(d_bucket.ptr));
  else
    bucket_kernel<u32><<<grid, kThreads>>>(d_reps.ptr, n_states, index.bucket_shift, index.bucket_count,
                                           reinterpret_cast<u32*>(d_bucket.ptr));
  KERNEL_LAUNCHED();
  sorted_check_kernel<<<persistent_grid(n_states + 1, kThreads, 8), kThreads>>>(d_reps.ptr, n_states, d_flag.ptr + 1);
  KERNEL_LAUNCHED();
  CUDA_CHECK(cudaGetLastError());
  int flags[2] = {0, 0};
  CUDA_CHECK(cudaMemcpy(flags, d_flag.ptr, sizeof(flags), cudaMemcpyDeviceToHost));
  if (flags[1]) fail(LS_INVALID_STATE, "representatives are not strictly increasing");
  if (flags[0]) fail(LS_NOT_A_REPRESENTATIVE, "a supplied state is not a representative of non-zero norm");
  index.reps = d_reps.ptr;
  index.stab = d_stab.ptr;
  index.bucket = d_bucket.ptr;
  built = true;
  ++generation;
}

std::shared_ptr<std::vector<u64>> Basis::states_host() {
  std::lock_guard<std::mutex> lock(mutex);
  if (!built) fail(LS_CACHE_NOT_BUILT, "basis has not been built");
  if (!host_reps) {
    auto v = std::make_shared<std::vector<u64>>(n_states);
    if (index.direct) {
      for (u64 i = 0; i < n_states; ++i) (*v)[i] = i;
    } else if (n_states) {
      CUDA_CHECK(cudaMemcpy(v->data(), d_reps.ptr, n_states * 8, cudaMemcpyDeviceToHost));
    }
    host_reps = v;
  }
  return host_reps;
}

std::shared_ptr<Basis> make_basis(std::shared_ptr<Group> g, unsigned n_spins, int hw, int inv) {
  if (n_spins == 0 || n_spins > 64) fail(LS_INVALID_NUMBER_SPINS, "number_spins must be in 1..64");
  if (hw < -1 || hw > (int)n_spins) fail(LS_INVALID_HAMMING_WEIGHT, "hamming_weight must be in 0..number_spins");
  if (inv != 0 && inv != 1 && inv != -1) fail(LS_INVALID_SPIN_INVERSION, "spin_inversion must be -1, 0 or +1");
  if (inv != 0 && hw >= 0 && 2 * hw != (int)n_spins)
    fail(LS_INVALID_SPIN_INVERSION, "spin inversion requires hamming_weight == number_spins / 2");
  if (g->n != 0 && g->n != n_spins)
    fail(LS_INVALID_ARGUMENT, "symmetry permutations must act on exactly number_spins sites");
  auto b = std::make_shared<Basis>();
  b->group = std::move(g);
  b->n_spins = n_spins;
  b->hamming_weight = hw;
  b->spin_inversion = inv;
  b->program = compile_program(*b->group, n_spins, inv);
  return b;
}

// state_info for tests / host tools
void basis_state_info(Basis& b, u64 count, u64 const* states, u64* reps, double* chars, double* norms) {
  if (b.trivial()) {
    for (u64 i = 0; i < count; ++i) {
      reps[i] = states[i];
      chars[2 * i] = 1;
      chars[2 * i + 1] = 0;
      norms[i] = 1;
    }
    return;
  }
  b.ensure_device_tables();
  DeviceBuffer<u64> d_in(count), d_rep(count);
  DeviceBuffer<std::int32_t> d_ph(count);
  DeviceBuffer<int> d_st(count);
  d_in.upload(states, count);
  int grid = persistent_grid(count, kThreads, 4);
  if (b.use32()) {
    auto d = device_program<u32>(b);
    if (d.smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(state_info_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem));
    state_info_kernel<u32><<<grid, kThreads, d.smem>>>(d.view, d.staged, d_in.ptr, count, d_rep.ptr, d_ph.ptr, d_st.ptr);
  } else {
    auto d = device_program<u64>(b);
    if (d.smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(state_info_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem));
    state_info_kernel<u64><<<grid, kThreads, d.smem>>>(d.view, d.staged, d_in.ptr, count, d_rep.ptr, d_ph.ptr, d_st.ptr);
  }
  KERNEL_LAUNCHED();
  CUDA_CHECK(cudaGetLastError());
  auto r = d_rep.download();
  auto ph = d_ph.download();
  auto st = d_st.download();
  i64 D = b.group->denom;
  for (u64 i = 0; i < count; ++i) {
    reps[i] = r[i];
    i64 k = ph[i];
    double ang = 2.0 * M_PI * (double)k / (double)D;
    double c = std::cos(ang), s = std::sin(ang);
    if (k == 0) { c = 1; s = 0; }
    else if (2 * k == D) { c = -1; s = 0; }
    else if (4 * k == D) { c = 0; s = 1; }
    else if (4 * k == 3 * D) { c = 0; s = -1; }
    chars[2 * i] = c;
    chars[2 * i + 1] = s;
    norms[i] = std::sqrt((double)st[i] / (double)b.group_order());
  }
}

}  // namespace sped
