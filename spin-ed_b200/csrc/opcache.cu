// opcache.cu -- the operator cache: after one matrix-free pass the non-zero off-diagonal elements
// of the local rows stay resident in HBM (180 GB per B200 is what makes this possible), and every
// later application of the operator is a coalesced stream of (column, coefficient code) plus the
// gather of x -- HBM-bound instead of integer-ALU-bound.
//
// The reference recomputes every matrix element in every ls_operator_matmat call
// (/root/reference/src/SpinED/Internal.hs:411-429 is called once per PRIMME block per iteration).
// Results agree with the matrix-free kernel to rounding: same elements and the same coefficient
// factors, w = v * (norm_s * (1 / norm_r)), but the elements that carry the default coefficient
// (first matrix value, chi = 1, trivial stabiliser: almost all of them) are stored first, without a
// code, and their x entries are summed before the one multiplication.
//
// Layout (sliced ELL): rows in slices of 32 = one warp; slot j of lane l of slice s is at
// slice_off[s] + 32 j + l.  Positions are u32 (N < 2^32); coded elements carry a u8 code (u16 when
// there are more than 256 distinct (value, phase, stabiliser) triples) into a table of
// (Re v, Im v, norm_s), v = M[a][b] * chi(g').  4 bytes per default element, 5 or 6 per coded one.
#include <chrono>
#include <cstdlib>
#include <map>

#include "canon.cuh"

namespace sped {

ProgramView<u32> program_view32(Basis const& b, size_t& smem, bool& staged);
ProgramView<u64> program_view64(Basis const& b, size_t& smem, bool& staged);
void* jit_cache_fill_kernel(Basis& b, double images);
MatvecParams operator_params(Operator& op);

namespace {

// Upper bound of the row lengths (transitions with a non-zero matrix element, whether or not the
// target survives the projection) and its maximum over each slice.
__global__ void __launch_bounds__(kThreads) slice_width_kernel(RowContext ctx, TermsView terms_g, u32* widths) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<false>(terms_g, smem);
  u64 const n_local = ctx.dist.n_local;
  u64 const n_padded = (n_local + 31) & ~(u64)31;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_padded; i += (u64)gridDim.x * blockDim.x) {
    u32 ub = 0;
    if (i < n_local) {
      u64 const row = dist_local_to_global(ctx.dist, i);
      u64 const r = ctx.index.direct ? row : __ldg(ctx.index.reps + row);
      for (u32 bnd = 0; bnd < terms.n_bonds; ++bnd) {
        DevBond const bd = terms.bonds[bnd];
        u32 a = 0;
        for (u32 j = 0; j < bd.k; ++j) a |= (u32)((r >> ((bd.sites >> (8 * j)) & 0xffu)) & 1ull) << (bd.k - 1 - j);
        ub += __popc((u32)terms.masks[bd.zoff + a]);
      }
    }
    u32 mx = __reduce_max_sync(0xffffffffu, ub);
    if ((i & 31) == 0) widths[i >> 5] = mx;
  }
}

// Several classes: per slice class c takes max_lanes len_c slots per lane; widths[s] is their sum
// and slice_start[s] = first slot of classes 1, 2, 3.
// `split`: len holds (default, coded) counts per class (after a staging traversal) instead of class totals.
__global__ void __launch_bounds__(kThreads) class_width_kernel(std::uint16_t const* len, u64 n_local, u32 n_classes,
                                                               u32* widths, u32* slice_start, bool split) {
  u64 const n_padded = (n_local + 31) & ~(u64)31;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_padded; i += (u64)gridDim.x * blockDim.x) {
    u32 w[kMaxClasses] = {0u, 0u, 0u};
#pragma unroll
    for (u32 c = 0; c < (u32)kMaxClasses; ++c) {
      u32 v = (c < n_classes && i < n_local) ? len[(u64)(2 * c) * n_local + i] : 0u;  // count pass: class totals
      if (split && c < n_classes && i < n_local) v += len[(u64)(2 * c + 1) * n_local + i];
      w[c] = __reduce_max_sync(0xffffffffu, v);
    }
    if ((i & 31) == 0) {
      widths[i >> 5] = w[0] + w[1] + w[2];
      slice_start[kClassStride * (i >> 5)] = w[0];
      slice_start[kClassStride * (i >> 5) + 1] = w[0] + w[1];
    }
  }
}

// slice_off[s] = 32 * sum_{t < s} widths[t]   (single block; slices are few millions at most)
__global__ void __launch_bounds__(1024) slice_scan_kernel(u32 const* widths, u64* slice_off, u64 n) {
  __shared__ u64 partial[1024];
  u64 per = (n + 1023) / 1024;
  u64 lo = min(n, (u64)threadIdx.x * per), hi = min(n, lo + per);
  u64 s = 0;
  for (u64 i = lo; i < hi; ++i) s += widths[i];
  partial[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 run = 0;
    for (int i = 0; i < 1024; ++i) {
      u64 v = partial[i];
      partial[i] = run;
      run += v;
    }
    slice_off[n] = run * 32;
  }
  __syncthreads();
  u64 run = partial[threadIdx.x];
  for (u64 i = lo; i < hi; ++i) {
    slice_off[i] = run * 32;
    run += widths[i];
  }
}

template <class W, bool SYM>
__global__ void __launch_bounds__(kThreads) cache_fill_kernel(FillParams p, ProgramView<W> prog) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<false>(p.terms, smem);
  if constexpr (SYM) {
    ProgramCanon<W> canon{stage_program<W>(prog, smem + terms_smem_bytes(p.terms, false))};
    cache_fill_rows(p, terms, canon);
  } else {
    cache_fill_rows(p, terms, TrivialCanon());
  }
}

// table[c] = (Re v, Im v, norm_s) for code c = (hid * n_pid + pid) * n_sid + sid, pid numbering the
// phase numerators that occur in the group (pid_phase[pid]); built on the device so that v is
// rounded exactly like in the matrix-free kernel.
__global__ void table_kernel(double const* values_re, double const* values_im, double const* chi_table,
                             double const* norm_table, std::uint16_t const* sid_stab, std::uint16_t const* pid_phase,
                             u32 n_hid, u32 n_pid, u32 n_sid, bool cplx, bool sym, double* table) {
  u32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_hid * n_pid * n_sid) return;
  u32 sid = c % n_sid, pid = (c / n_sid) % n_pid, hid = c / (n_sid * n_pid);
  u32 ph = sym ? pid_phase[pid] : 0u;
  double2 v = make_double2(values_re[hid], values_im[hid]);
  if (sym) {
    if (cplx) {
      v = cmul(v, make_double2(chi_table[2 * ph], chi_table[2 * ph + 1]));
    } else {
      v.x = ph == 0 ? v.x : -v.x;
      v.y = 0.0;
    }
  }
  table[3 * c] = v.x;
  table[3 * c + 1] = v.y;
  table[3 * c + 2] = sym ? norm_table[sid_stab[sid]] : 1.0;
}

}  // namespace
}  // namespace sped

#include "cached_kernel.cuh"

namespace sped {
namespace {

// SPED_CACHED_VARIANT (tuning knob; unset = the measured best, chosen by row length): 0 plain loads,
// 19 / 33 the two default kernels, other values only in builds with -DSPED_CACHED_SWEEP.
int cached_variant() {
  static int v = [] {
    char const* e = std::getenv("SPED_CACHED_VARIANT");
    return e && *e ? std::atoi(e) : -1;
  }();
  return v;
}

// Persistent launch: one wave of exactly as many blocks as are resident (occupancy x 148 SMs).
template <void (*Kernel)(CachedParams)>
void launch_cached_kernel(CachedParams const& p, cudaStream_t s) {
  static int const per_sm = [] {
    int n = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, Kernel, kThreads, 0));
    char const* e = std::getenv("SPED_CACHED_BLOCKS_PER_SM");
    if (e && *e) n = std::min(n, std::max(1, std::atoi(e)));
    return std::max(n, 1);
  }();
  int grid = persistent_grid(p.row_hi - p.row_lo, kThreads, per_sm);
  // A pass that runs beside one of NCCL's transfer kernels uses short-lived blocks (four rows per
  // thread): they keep freeing SM resources, so the transfer's CTAs -- launched on a higher-priority
  // stream -- become resident at once instead of waiting for a persistent wave to drain.
  if (p.beside_transfer) grid = (int)std::min<u64>(((p.row_hi - p.row_lo) + 4 * kThreads - 1) / (4 * kThreads), 1u << 30);
  Kernel<<<std::max(grid, 1), kThreads, 0, s>>>(p);
}

template <class T, int NB, class Code, bool SYM>
void launch_cached_variant(CachedParams const& p, cudaStream_t s) {
  switch (cached_variant()) {
    case 0: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, false, 4, 6>>(p, s); break;
#if defined(SPED_CACHED_SWEEP)  // tuning builds only (profiles/r02_kernel_variants.md): (elements in flight, blocks per SM[, pipelined])
    case 3: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 8, 5>>(p, s); break;
    case 5: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 4, 8>>(p, s); break;
    case 7: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 2, 8>>(p, s); break;
    case 9: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 4, 6>>(p, s); break;
    case 11: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 6, 5>>(p, s); break;
    case 15: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 6, 5, true>>(p, s); break;
    case 17: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 4, 6, true>>(p, s); break;
    case 23: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 4, 5, true>>(p, s); break;
    case 25: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 8, 5, true>>(p, s); break;
    case 27: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 10, 4, true>>(p, s); break;
    case 29: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 12, 3, true>>(p, s); break;
    case 31: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 8, 3, true>>(p, s); break;
    case 35: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 8, 4, false>>(p, s); break;
#endif
    case 19: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 8, 4, true>>(p, s); break;
    case 33: launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 6, 4, true>>(p, s); break;
    default:
      // software-pipelined position loads, 4 blocks of 256 per SM (64 registers); batches of 6 for long
      // rows (6x6: 37 elements per row, 1.31 ms), of 8 for short ones (chain_36: 18.5, 2.57 ms) -- measured
      if (p.mean_row_length >= 28.0f) launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 6, 4, true>>(p, s);
      else launch_cached_kernel<cached_matvec_kernel<T, NB, Code, SYM, true, 8, 4, true>>(p, s);
      break;
  }
}

template <class T, int NB>
void launch_cached_nb(CachedParams const& p, cudaStream_t s) {
  bool const wide = p.cache.code_wide != 0, sym = p.sym != 0;
  if (wide && sym) launch_cached_variant<T, NB, std::uint16_t, true>(p, s);
  else if (wide) launch_cached_variant<T, NB, std::uint16_t, false>(p, s);
  else if (sym) launch_cached_variant<T, NB, std::uint8_t, true>(p, s);
  else launch_cached_variant<T, NB, std::uint8_t, false>(p, s);
}

template <class T, int NB>
void launch_block(CachedParams const& p, T const* x, u64 xs, u64 n_entries, T* scratch, cudaStream_t s) {
  interleave_kernel<T, NB><<<persistent_grid(n_entries, kThreads, 8), kThreads, 0, s>>>(x, xs, p.ncols, n_entries, scratch);
  KERNEL_LAUNCHED();
  CachedParams q = p;
  q.x = scratch;
  constexpr int U = NB == 2 ? 4 : 2;
  bool const wide = p.cache.code_wide != 0, sym = p.sym != 0;
  if (wide && sym) launch_cached_kernel<cached_block_kernel<T, NB, std::uint16_t, true, U>>(q, s);
  else if (wide) launch_cached_kernel<cached_block_kernel<T, NB, std::uint16_t, false, U>>(q, s);
  else if (sym) launch_cached_kernel<cached_block_kernel<T, NB, std::uint8_t, true, U>>(q, s);
  else launch_cached_kernel<cached_block_kernel<T, NB, std::uint8_t, false, U>>(q, s);
}

// `scratch` holds the interleaved copy of up to four columns (n_entries * 4 values)
template <class T>
void launch_cached(CachedParams p, u64 block, u64 xs, u64 ys, u64 n_entries, void* scratch, cudaStream_t s) {
  T const* x = static_cast<T const*>(p.x);
  T* y = static_cast<T*>(p.y);
  for (u64 c0 = 0; c0 < block;) {
    u64 left = block - c0;
    p.x = x + c0 * xs;
    p.y = y + c0 * ys;
    p.ncols = (u32)std::min<u64>(left, 4);
    if (p.ncols == 1) launch_cached_nb<T, 1>(p, s);
    else if (p.ncols == 2) launch_block<T, 2>(p, x + c0 * xs, xs, n_entries, static_cast<T*>(scratch), s);
    else launch_block<T, 4>(p, x + c0 * xs, xs, n_entries, static_cast<T*>(scratch), s);
    KERNEL_LAUNCHED();
    c0 += p.ncols;
  }
  CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(kThreads) sum_len_kernel(std::uint16_t const* len, u64 n, unsigned long long* out) {
  unsigned long long mine = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) mine += len[i];
  for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

// SPED_FILL_STAGED=0: several-class caches are built with a counting and a filling traversal (the
// fallback when the staging area does not fit) instead of one traversal plus a placement pass.
bool staged_fill_enabled() {
  char const* e = std::getenv("SPED_FILL_STAGED");
  return !(e && e[0] == '0');
}

int env_cache_mode() {
  char const* e = std::getenv("SPED_OPERATOR_CACHE");
  if (!e || !*e) return -1;
  return e[0] == '0' ? 0 : 1;
}

}  // namespace

// Number of exchange rounds of a sharded matvec (and remote source classes of the cache): a function
// of the world size, the shard length and the environment only, so that every rank makes the same
// choice.  Two rounds (grouped NCCL send/recv, near peers first) let the pass over the first group's
// elements overlap the second transfer, which pays when the transfers are long (40-spin chains:
// gigabytes per shard); shards below 2^23 entries (6x6 over 4-8 ranks: 16-32 MB) are latency-bound
// and go through ONE NCCL all-gather, which is much faster than grouped point-to-point calls there
// (measured on 4 B200: 0.21 ms for the all-gather against 0.22 + 0.25 ms for the two rounds).
// SPED_REMOTE_GROUPS=1 / 2 forces one / two rounds.
int exchange_rounds(unsigned world, u64 chunk) {
  if (world <= 1) return 0;
  if (world == 2) return 1;
  char const* e = std::getenv("SPED_REMOTE_GROUPS");
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '2') return 2;
  return chunk >= ((u64)1 << 23) ? 2 : 1;
}

// Coefficient codes of the operator cache: code = (hid * n_pid + pid) * n_sid + sid with hid the
// distinct off-diagonal matrix value, pid the phase numerator (among those that occur in G') and
// sid the stabiliser size of the target (a divisor of |G'|).  Returns null, or why there can be no
// cache.  Host only; shared with the host emulation of the kernels (emul.cpp).
char const* build_code_maps(Operator const& op, CodeMaps& out) {
  // distinct off-diagonal matrix values -> hid; stabiliser sizes -> sid
  std::vector<double>&values_re = out.values_re, &values_im = out.values_im;
  std::vector<std::uint16_t>& hid_map = out.hid_map;
  auto const& terms = op.terms;
  Basis& b = *op.basis;
  {
    std::map<std::pair<double, double>, u32> seen;
    for (auto const& t : terms) {
      u32 dim = 1u << t.k;
      for (u32 a = 0; a < dim; ++a)
        for (u32 c = 0; c < dim; ++c) {
          cplx v = t.matrix[a * dim + c];
          u32 id = 0;
          if (a != c && v != cplx(0, 0)) {
            auto key = std::make_pair(v.real(), v.imag());
            auto it = seen.find(key);
            if (it == seen.end()) {
              it = seen.emplace(key, (u32)values_re.size()).first;
              values_re.push_back(v.real());
              values_im.push_back(v.imag());
            }
            id = it->second;
          }
          hid_map.push_back((std::uint16_t)id);
        }
    }
  }
  if (values_re.empty()) return "operator is diagonal";
  bool const sym = !b.trivial();
  u64 const order = b.group_order();
  std::vector<std::uint16_t>&sid_map = out.sid_map, &sid_stab = out.sid_stab;
  sid_map.assign(order + 1, 0);
  if (sym) {
    for (u64 s = 1; s <= order; ++s)
      if (order % s == 0) {
        sid_map[s] = (std::uint16_t)sid_stab.size();
        sid_stab.push_back((std::uint16_t)s);
      }
  } else {
    sid_stab.push_back(1);
  }
  // phase numerators that some element of G' carries (the canonicalisation never reports others)
  std::vector<std::uint16_t>&pid_map = out.pid_map, &pid_phase = out.pid_phase;
  if (sym) {
    u32 const D = (u32)b.group->denom;
    if (D > 65535) return "more than 65535 distinct phases";
    std::vector<char> occurs(D, 0);
    for (auto const& e : b.group->elems) {
      occurs[(u32)e.phase % D] = 1;
      if (b.spin_inversion < 0) occurs[((u32)e.phase + D / 2) % D] = 1;
    }
    pid_map.assign(D, 0);
    for (u32 ph = 0; ph < D; ++ph)
      if (occurs[ph]) {
        pid_map[ph] = (std::uint16_t)pid_phase.size();
        pid_phase.push_back((std::uint16_t)ph);
      }
  } else {
    pid_phase.push_back(0);
  }
  u32 const n_pid = out.n_pid = (u32)pid_phase.size();
  u64 const n_codes = out.n_codes = (u64)values_re.size() * n_pid * sid_stab.size();
  if (n_codes > 65536) return "more than 65536 distinct coefficients";

  // the coefficient almost every element carries: first off-diagonal value, chi = 1, trivial stabiliser
  out.default_code = ((0u * n_pid + (sym ? (u32)pid_map[0] : 0u)) * (u32)sid_stab.size()) + (sym ? (u32)sid_map[1] : 0u);
  if (char const* e = std::getenv("SPED_DEFAULT_CLASS"))
    if (e[0] == '0') out.default_code = ~0u;  // tuning / diagnosis: no element matches, everything is stored coded
  return nullptr;
}

// Peers of the first exchange round: ranks r+1 .. r+near (a function of the world size and the
// environment only, like exchange_rounds: every rank must agree on who talks in which round).
unsigned exchange_near(unsigned world, u64 chunk) { return exchange_rounds(world, chunk) == 2 ? world / 2 : world - 1; }

void Operator::drop_cache() {
  cache_ready = false;
  cache_rejected = false;
  c_slice_off.release();
  c_idx.release();
  c_code.release();
  c_code_off.release();
  c_len.release();
  c_slice_start.release();
  c_classes = 1;
  c_near = 0;
  c_rounds = 0;
  c_table.release();
  c_slices = c_slots = cache_bytes = 0;
}

// Builds the cache on first use when allowed and when it fits; afterwards just reports it.
bool Operator::cache_usable() {
  if (cache_ready) return true;
  int mode = cache_mode >= 0 ? cache_mode : env_cache_mode();
  if (mode == 0 || cache_rejected) return false;
  Basis& b = *basis;
  auto reject = [&](char const* why) {
    SPED_LOG("operator cache not used: %s", why);
    cache_rejected = true;
    drop_cache();
    cache_rejected = true;
    return false;
  };
  if (dist.chunk * dist.world >= 0xffffffffull) return reject("replicated vector longer than 2^32 - 1 entries");
  u64 const n_local = dist.n_local;
  if (n_local == 0) return false;
  auto t0 = std::chrono::steady_clock::now();
  SPED_NVTX("sped: operator cache fill (matrix-free traversal)");

  CodeMaps cm_;
  if (char const* why = build_code_maps(*this, cm_)) return reject(why);
  std::vector<double>&values_re = cm_.values_re, &values_im = cm_.values_im;
  std::vector<std::uint16_t>&hid_map = cm_.hid_map, &sid_map = cm_.sid_map, &sid_stab = cm_.sid_stab, &pid_map = cm_.pid_map,
                           &pid_phase = cm_.pid_phase;
  bool const sym = !b.trivial();
  u32 const n_pid = cm_.n_pid;
  u64 const n_codes = cm_.n_codes;

  // maps and coefficient table
  DeviceBuffer<double> d_vre, d_vim;
  DeviceBuffer<std::uint16_t> d_hid, d_sid_map, d_sid_stab, d_pid_map, d_pid_phase;
  d_vre.upload(values_re);
  d_vim.upload(values_im);
  d_hid.upload(hid_map);
  d_sid_map.upload(sid_map);
  d_sid_stab.upload(sid_stab);
  d_pid_map.upload(pid_map);
  d_pid_phase.upload(pid_phase);
  c_table.alloc(n_codes * 3);
  bool const cplx_table = !is_real();
  table_kernel<<<(unsigned)((n_codes + 127) / 128), 128>>>(d_vre.ptr, d_vim.ptr, b.d_chi_table.ptr, b.d_norm_table.ptr,
                                                          d_sid_stab.ptr, d_pid_phase.ptr, (u32)values_re.size(),
                                                          n_pid, (u32)sid_stab.size(), cplx_table, sym, c_table.ptr);
  KERNEL_LAUNCHED();

  c_slices = (n_local + 31) / 32;
  c_code_wide = n_codes > 256 ? 1 : 0;
  u64 const code_bytes = c_code_wide ? 2 : 1;
  // source classes (see CacheView): local / peers of the first exchange round / of the second
  u32 const world = dist.world;
  c_rounds = (u32)exchange_rounds(world, dist.chunk);
  c_classes = 1 + c_rounds;
  c_near = exchange_near(world, dist.chunk);  // 8 ranks: 4 peers in the first round, 3 in the second
  bool const two = c_rounds > 0;  // several ranks: exact class sizes from a counting traversal
  MatvecParams mp = operator_params(*this);
  size_t tsm = terms_smem_bytes(mp.terms, false);
  c_len.alloc(n_local * 2 * c_classes);
  DeviceBuffer<int> d_flag(1);
  CUDA_CHECK(cudaMemset(d_flag.ptr, 0, sizeof(int)));
  FillParams fp{};
  fp.ctx = mp.ctx;
  fp.terms = mp.terms;
  fp.len = c_len.ptr;
  fp.n_classes = c_classes;
  fp.near = c_near;
  fp.rounds = c_rounds;
  c_default_code = cm_.default_code;
  fp.default_code = c_default_code;
  fp.hid_map = d_hid.ptr;
  fp.sid_map = sym ? d_sid_map.ptr : nullptr;
  fp.pid_map = sym ? d_pid_map.ptr : nullptr;
  fp.denom = n_pid;
  fp.n_sid = (u32)sid_stab.size();
  fp.code_wide = c_code_wide;
  fp.overflow = d_flag.ptr;
  // the matrix-free traversal (run-time specialised kernel when available)
  fp.row_lo = 0;
  fp.row_hi = n_local;
  // one choice for the whole build (one or two traversals of the local rows): the specialised kernel if
  // it is ready or worth waiting for, else the interpreted one -- the two give identical results
  double const images = (double)n_local * (double)b.group_order() * 0.5 * (double)mp.terms.n_bonds * (two ? 2.0 : 1.0);
  void* const jit = sym ? jit_cache_fill_kernel(b, images) : nullptr;
  auto launch_fill = [&]() {
    int grid = persistent_grid(fp.row_hi - fp.row_lo, kThreads, 8);
    if (jit) {
      void* args[] = {&fp};
      if (tsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(jit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
      CUDA_CHECK(cudaLaunchKernel(jit, dim3(grid), dim3(kThreads), args, tsm, nullptr));
    } else if (!sym) {
      if (tsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(cache_fill_kernel<u64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
      cache_fill_kernel<u64, false><<<grid, kThreads, tsm>>>(fp, ProgramView<u64>{});
    } else if (b.use32()) {
      size_t psm; bool staged;
      auto prog = program_view32(b, psm, staged);
      if (tsm + psm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(cache_fill_kernel<u32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tsm + psm)));
      cache_fill_kernel<u32, true><<<grid, kThreads, tsm + psm>>>(fp, prog);
    } else {
      size_t psm; bool staged;
      auto prog = program_view64(b, psm, staged);
      if (tsm + psm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(cache_fill_kernel<u64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tsm + psm)));
      cache_fill_kernel<u64, true><<<grid, kThreads, tsm + psm>>>(fp, prog);
    }
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
  };

  // slice widths: one class -- the cheap upper bound (transitions with a non-zero matrix element);
  // two classes -- exact per-class counts from a first traversal
  DeviceBuffer<u32> d_widths(c_slices);
  // Several classes: ONE traversal into a staging area (every element in the next slot of its lane,
  // slots from the cheap width bound), then a placement kernel that lays the elements out by class --
  // when the staging area (4 + code bytes per slot) fits beside the cache; else count + fill traversals.
  bool staged_done = false;
  if (two && staged_fill_enabled()) {
    DeviceBuffer<u32> d_ub(c_slices);
    if (tsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(slice_width_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
    slice_width_kernel<<<persistent_grid(n_local, kThreads, 8), kThreads, tsm>>>(mp.ctx, mp.terms, d_ub.ptr);
    KERNEL_LAUNCHED();
    DeviceBuffer<u64> d_stage_off(c_slices + 1);
    slice_scan_kernel<<<1, 1024>>>(d_ub.ptr, d_stage_off.ptr, c_slices);
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
    u64 stage_slots = 0;
    CUDA_CHECK(cudaMemcpy(&stage_slots, d_stage_off.ptr + c_slices, 8, cudaMemcpyDeviceToHost));
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    // the final layout pads every class separately: allow for 1.35 x the staging slots
    u64 const stage_bytes = stage_slots * (4 + code_bytes);
    u64 const final_guess = (u64)(1.35 * (double)stage_slots) * 4 + n_local * 8 * c_classes;
    if (stage_bytes + final_guess < free_b - free_b / 8) {
      DeviceBuffer<u32> d_stage_idx(std::max<u64>(stage_slots, 1));
      DeviceBuffer<unsigned char> d_stage_code(std::max<u64>(stage_slots, 1) * code_bytes);
      fp.stage = 1;
      fp.count_only = 0;
      fp.slice_off = d_stage_off.ptr;
      fp.slice_start = nullptr;
      fp.idx = d_stage_idx.ptr;
      fp.code = d_stage_code.ptr;
      fp.code_slot0 = 0;
      launch_fill();
      fp.stage = 0;
      CUDA_CHECK(cudaDeviceSynchronize());
      int overflow = 0;
      CUDA_CHECK(cudaMemcpy(&overflow, d_flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
      if (overflow) return reject("internal: a row exceeded its staging slot bound");
      // final layout from the per-class counts: class widths, offsets, compact code offsets
      c_slice_start.alloc(c_slices * kClassStride);
      class_width_kernel<<<persistent_grid(c_slices * 32, kThreads, 8), kThreads>>>(c_len.ptr, n_local, c_classes, d_widths.ptr,
                                                                                   c_slice_start.ptr, true);
      KERNEL_LAUNCHED();
      c_slice_off.alloc(c_slices + 1);
      slice_scan_kernel<<<1, 1024>>>(d_widths.ptr, c_slice_off.ptr, c_slices);
      KERNEL_LAUNCHED();
      u64 const n_regions = c_slices * c_classes;
      DeviceBuffer<u32> d_cw(std::max<u64>(n_regions, 1));
      code_width_kernel<<<persistent_grid(c_slices, kThreads, 8), kThreads>>>(c_len.ptr, n_local, c_classes, 0, c_slices, d_cw.ptr);
      KERNEL_LAUNCHED();
      c_code_off.alloc(n_regions + 1);
      slice_scan_kernel<<<1, 1024>>>(d_cw.ptr, c_code_off.ptr, n_regions);
      KERNEL_LAUNCHED();
      CUDA_CHECK(cudaGetLastError());
      u64 code_slots = 0;
      CUDA_CHECK(cudaMemcpy(&c_slots, c_slice_off.ptr + c_slices, 8, cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(&code_slots, c_code_off.ptr + n_regions, 8, cudaMemcpyDeviceToHost));
      u64 const need = c_slots * 4 + code_slots * code_bytes + n_local * 4 * c_classes + (c_slices + 1) * 16 + (n_regions + 1) * 8 + n_codes * 24;
      CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
      u64 held = 0;
      for (auto const& w : eigh_ws) held += w.count;
      u64 const wanted = 4 * n_local * 16;
      u64 const reserve = std::max<u64>((u64)2 << 30, wanted > held ? wanted - held : 0);
      // (the staging area is released before the solver allocates anything: it does not count against the reserve)
      if (mode != 1 && need + reserve > free_b + stage_bytes) return reject("does not fit in the free device memory (with room for the solver's vectors)");
      if (c_slots * 4 + code_slots * code_bytes > free_b - free_b / 16) return reject("does not fit in device memory");
      c_idx.alloc(std::max<u64>(c_slots, 1));
      c_code.alloc(std::max<u64>(code_slots, 1) * code_bytes);
      CUDA_CHECK(cudaMemsetAsync(c_code.ptr, 0, std::max<u64>(code_slots, 1) * code_bytes));
      PlaceParams q{};
      q.dist = dist;
      q.stage_off = d_stage_off.ptr;
      q.stage_idx = d_stage_idx.ptr;
      q.stage_code = d_stage_code.ptr;
      q.out.slice_off = c_slice_off.ptr;
      q.out.len = c_len.ptr;
      q.out.slice_start = c_slice_start.ptr;
      q.out.n_slices = c_slices;
      q.out.n_classes = c_classes;
      q.out.near = c_near;
      q.out.rounds = c_rounds;
      q.out.default_code = c_default_code;
      q.out.code_off = c_code_off.ptr;
      q.idx = c_idx.ptr;
      q.code = c_code.ptr;
      int const grid = persistent_grid(n_local, kThreads, 8);
      if (c_code_wide) cache_place_kernel<std::uint16_t><<<grid, kThreads>>>(q);
      else cache_place_kernel<std::uint8_t><<<grid, kThreads>>>(q);
      KERNEL_LAUNCHED();
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaDeviceSynchronize());
      cache_bytes = need;
      cache_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      cache_ready = true;
      SPED_LOG("operator cache: %llu slots, %.2f GB, built in %.3f s (one traversal + placement)", (unsigned long long)c_slots,
               need / 1e9, cache_build_seconds);
      staged_done = true;
    } else {
      SPED_LOG("operator cache: no room for the staging area (%.1f GB): counting and filling in two traversals", stage_bytes / 1e9);
    }
  }
  if (staged_done) return true;
  if (two) {
    fp.count_only = 1;
    launch_fill();
    c_slice_start.alloc(c_slices * kClassStride);
    class_width_kernel<<<persistent_grid(c_slices * 32, kThreads, 8), kThreads>>>(c_len.ptr, n_local, c_classes, d_widths.ptr,
                                                                                 c_slice_start.ptr, false);
    fp.count_only = 0;
    fp.slice_start = c_slice_start.ptr;
  } else {
    if (tsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(slice_width_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
    slice_width_kernel<<<persistent_grid(n_local, kThreads, 8), kThreads, tsm>>>(mp.ctx, mp.terms, d_widths.ptr);
  }
  KERNEL_LAUNCHED();
  c_slice_off.alloc(c_slices + 1);
  slice_scan_kernel<<<1, 1024>>>(d_widths.ptr, c_slice_off.ptr, c_slices);
  KERNEL_LAUNCHED();
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpy(&c_slots, c_slice_off.ptr + c_slices, 8, cudaMemcpyDeviceToHost));
  // What stays resident (the codes of the few coded elements come on top, counted below).  The fill
  // writes one code per slot into a temporary and only the coded parts are kept (CacheView::code_off);
  // it runs in row chunks so that the temporary stays below 2 GB instead of a fifth of the cache
  // (chain_40 on one GPU: 20 GB that would not fit beside an 80 GB cache and the solver's vectors).
  u64 const meta = n_local * 4 * c_classes + (c_slices + 1) * (c_classes > 1 ? 16 : 8) + (c_slices * c_classes + 1) * 8 + n_codes * 24;
  u64 need = c_slots * 4 + meta;
  u64 chunk_bytes = (u64)1 << 31;
  if (char const* e = std::getenv("SPED_FILL_CHUNK_BYTES"))  // tests: force several chunks on small decks
    if (*e) chunk_bytes = std::max<u64>(1, std::strtoull(e, nullptr, 10));
  u64 const n_chunks = std::max<u64>(1, std::min<u64>(c_slices, (c_slots * code_bytes + chunk_bytes - 1) / chunk_bytes));
  u64 const slices_per_chunk = (c_slices + n_chunks - 1) / n_chunks;
  std::vector<u64> chunk_slot0(n_chunks + 1, 0);  // first slot of every chunk
  u64 temp_slots = 0;
  for (u64 k = 0; k <= n_chunks; ++k) {
    u64 const sl = std::min(c_slices, k * slices_per_chunk);
    CUDA_CHECK(cudaMemcpy(&chunk_slot0[k], c_slice_off.ptr + sl, 8, cudaMemcpyDeviceToHost));
    if (k) temp_slots = std::max(temp_slots, chunk_slot0[k] - chunk_slot0[k - 1]);
  }
  u64 const peak = need + temp_slots * code_bytes + slices_per_chunk * c_classes * 8;
  size_t free_b = 0, total_b = 0;
  CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  // automatic mode keeps room for what a solver typically allocates afterwards (a few more vectors of
  // the local rows); mode 1 only insists on the cache itself fitting
  // (what sped_eigh already holds on this operator -- it takes its workspace before the first
  // application -- counts towards that room)
  u64 held = 0;
  for (auto const& w : eigh_ws) held += w.count;
  u64 const wanted = 4 * n_local * 16;
  u64 const reserve = std::max<u64>((u64)2 << 30, wanted > held ? wanted - held : 0);
  if (mode != 1 && need + reserve > free_b) return reject("does not fit in the free device memory (with room for the solver's vectors)");
  if (peak > free_b - free_b / 16) return reject("does not fit in device memory");
  c_idx.alloc(std::max<u64>(c_slots, 1));
  DeviceBuffer<unsigned char> code_temp(std::max<u64>(temp_slots, 1) * code_bytes);
  u64 const n_regions = c_slices * c_classes;
  DeviceBuffer<u32> d_cw(std::max<u64>(n_regions, 1));
  DeviceBuffer<u64> d_chunk_off(slices_per_chunk * c_classes + 1);
  std::vector<DeviceBuffer<unsigned char>> chunk_codes(n_chunks);
  std::vector<u64> chunk_code_slots(n_chunks, 0);

  fp.slice_off = c_slice_off.ptr;
  fp.idx = c_idx.ptr;
  fp.code = code_temp.ptr;
  CacheView v{};  // what the compaction kernel reads
  v.slice_off = c_slice_off.ptr;
  v.code = code_temp.ptr;
  v.len = c_len.ptr;
  v.slice_start = c_slice_start.ptr;
  v.n_slices = c_slices;
  v.n_classes = c_classes;
  for (u64 k = 0; k < n_chunks; ++k) {
    u64 const s_lo = std::min(c_slices, k * slices_per_chunk), s_hi = std::min(c_slices, s_lo + slices_per_chunk);
    if (s_lo == s_hi) continue;
    fp.row_lo = 32 * s_lo;
    fp.row_hi = std::min<u64>(32 * s_hi, n_local);
    fp.code_slot0 = chunk_slot0[k];
    launch_fill();
    // widths of this chunk's coded regions, their offsets inside the chunk, the compact codes
    u64 const regions = (s_hi - s_lo) * c_classes;
    code_width_kernel<<<persistent_grid(s_hi - s_lo, kThreads, 8), kThreads>>>(c_len.ptr, n_local, c_classes, s_lo, s_hi, d_cw.ptr);
    KERNEL_LAUNCHED();
    slice_scan_kernel<<<1, 1024>>>(d_cw.ptr + s_lo * c_classes, d_chunk_off.ptr, regions);
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(&chunk_code_slots[k], d_chunk_off.ptr + regions, 8, cudaMemcpyDeviceToHost));
    if (chunk_code_slots[k]) {
      chunk_codes[k].alloc(chunk_code_slots[k] * code_bytes);
      // (entries past a lane's own coded elements are padding nobody reads; zeroed so that the copy
      // below moves initialised memory only)
      CUDA_CHECK(cudaMemsetAsync(chunk_codes[k].ptr, 0, chunk_code_slots[k] * code_bytes));
      int const grid = persistent_grid(fp.row_hi - fp.row_lo, kThreads, 8);
      if (c_code_wide)
        code_compact_kernel<std::uint16_t><<<grid, kThreads>>>(v, n_local, fp.row_lo, fp.row_hi, fp.code_slot0, d_chunk_off.ptr,
                                                              reinterpret_cast<std::uint16_t*>(chunk_codes[k].ptr));
      else
        code_compact_kernel<std::uint8_t><<<grid, kThreads>>>(v, n_local, fp.row_lo, fp.row_hi, fp.code_slot0, d_chunk_off.ptr,
                                                             reinterpret_cast<std::uint8_t*>(chunk_codes[k].ptr));
      KERNEL_LAUNCHED();
      CUDA_CHECK(cudaGetLastError());
    }
  }
  CUDA_CHECK(cudaDeviceSynchronize());
  int overflow = 0;
  CUDA_CHECK(cudaMemcpy(&overflow, d_flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
  if (overflow) return reject("internal: a row exceeded its slot bound");
  {  // global offsets of the coded regions; the chunks' codes, concatenated, are the compact stream
    code_temp.release();
    c_code_off.alloc(n_regions + 1);
    slice_scan_kernel<<<1, 1024>>>(d_cw.ptr, c_code_off.ptr, n_regions);
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
    u64 code_slots = 0;
    CUDA_CHECK(cudaMemcpy(&code_slots, c_code_off.ptr + n_regions, 8, cudaMemcpyDeviceToHost));
    c_code.alloc(std::max<u64>(code_slots, 1) * code_bytes);
    u64 at = 0;
    for (u64 k = 0; k < n_chunks; ++k) {
      if (chunk_code_slots[k])
        CUDA_CHECK(cudaMemcpy(c_code.ptr + at * code_bytes, chunk_codes[k].ptr, chunk_code_slots[k] * code_bytes, cudaMemcpyDeviceToDevice));
      at += chunk_code_slots[k];
      chunk_codes[k].release();
    }
    if (at != code_slots) return reject("internal: the compact code stream does not add up");
    need += code_slots * code_bytes;
  }
  cache_bytes = need;
  cache_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  cache_ready = true;
  SPED_LOG("operator cache: %llu slots, %.2f GB, built in %.3f s", (unsigned long long)c_slots, need / 1e9,
           cache_build_seconds);
  return true;
}

// Number of stored elements of the local rows (integer sum: order does not matter).
void Operator::cached_count(unsigned long long* d_out) {
  u64 n_local = dist.n_local;
  if (!n_local) return;
  sum_len_kernel<<<persistent_grid(n_local, kThreads, 4), kThreads>>>(c_len.ptr, n_local * 2 * c_classes, d_out);
  KERNEL_LAUNCHED();
  CUDA_CHECK(cudaGetLastError());
}

void Operator::cached_matmat(int dtype, u64 block, void const* x, u64 xs, void* y, u64 ys, cudaStream_t s, u64 row_lo,
                             u64 row_hi, int phase, bool beside_transfer) {
  SPED_NVTX(phase == 0 ? "sped: cached matvec (all classes)" : phase == 1 ? "sped: cached matvec (local class)" : "sped: cached matvec (remote class)");
  static bool const fetch_set = [] {  // tuning knob: DRAM->L2 fetch granularity (32, 64 or 128 bytes)
    char const* e = std::getenv("SPED_L2_FETCH");
    if (e && *e) {
      cudaError_t rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(e));
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      SPED_LOG("L2 fetch granularity: asked %s, rc %d, now %zu", e, (int)rc, got);
    }
    return true;
  }();
  (void)fetch_set;
  Basis& b = *basis;
  MatvecParams mp = operator_params(*this);
  CachedParams p{};
  p.cache = CacheView{c_slice_off.ptr, c_idx.ptr, c_code.ptr, c_len.ptr, c_slice_start.ptr, c_table.ptr, c_slices,
                      c_code_wide, (u32)(c_table.count / 3), c_classes, c_near, c_default_code, c_rounds, c_code_off.ptr};
  p.phase = phase;
  p.beside_transfer = beside_transfer ? 1 : 0;
  p.mean_row_length = dist.n_local ? (float)((double)c_slots / (double)dist.n_local) : 0.0f;
  p.ctx = mp.ctx;
  p.diag_re = mp.diag_re;
  p.diag_im = mp.diag_im;
  p.x = x;
  p.y = y;
  p.xs = xs;
  p.ys = ys;
  p.sym = b.trivial() ? 0 : 1;
  p.row_lo = row_lo;
  p.row_hi = std::min<u64>(row_hi, dist.n_local);
  if (p.row_lo >= p.row_hi) return;
  if (p.row_lo & 31) fail(SPED_INTERNAL_ERROR, "row ranges of the cached matvec start at a multiple of 32");
  u64 const n_entries = dist.chunk * dist.world;  // entries of one replicated column
  void* scratch = nullptr;
  if (block > 1) {
    if (phase != kPhaseAll) fail(SPED_INTERNAL_ERROR, "block applications handle all source classes in one pass");
    size_t const need_bytes = n_entries * 4 * dtype_size(dtype);
    if (block_x.count < need_bytes) block_x.alloc(need_bytes);
    scratch = block_x.ptr;
  }
  switch (dtype) {
    case SPED_F32: launch_cached<float>(p, block, xs, ys, n_entries, scratch, s); break;
    case SPED_F64: launch_cached<double>(p, block, xs, ys, n_entries, scratch, s); break;
    case SPED_C64: launch_cached<float2>(p, block, xs, ys, n_entries, scratch, s); break;
    case SPED_C128: launch_cached<double2>(p, block, xs, ys, n_entries, scratch, s); break;
    default: fail(LS_INVALID_DATATYPE, "unknown datatype tag");
  }
}

}  // namespace sped
