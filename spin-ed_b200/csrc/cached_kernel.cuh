// cached_kernel.cuh -- device code of the streaming (operator-cache) matvec, shared by opcache.cu and
// by the single-threaded host emulation of the test suite (emul.cpp: the same source compiled by the
// host compiler with the CUDA built-ins shimmed, so that slot arithmetic and summation logic are
// checked against the oracle without a GPU).  Needs Traits / acc_fma from matvec_kernel.cuh.
#pragma once
#include "matvec_kernel.cuh"

namespace sped {

#if !defined(SPED_KERNEL_THREADS)
#define SPED_KERNEL_THREADS 256
#endif
// The host emulation defines this as `static`: its host-compiled kernel bodies must not be merged
// (same mangled name, weak symbol) with the launch stubs nvcc emits for the real kernels.
#if !defined(SPED_KERNEL_LINKAGE)
#define SPED_KERNEL_LINKAGE
#endif

struct CachedParams {
  CacheView cache;
  RowContext ctx;
  double const* diag_re;
  double const* diag_im;
  void const* x;
  void* y;
  u64 xs, ys;
  u32 ncols;
  int sym;
  u64 row_lo, row_hi;  // local rows handled by this launch (row_lo is a multiple of 32)
  int beside_transfer; // this pass overlaps an exchange round (see launch_cached_kernel)
  int phase;           // kPhaseAll, or 1 + class (CacheView): class 0 starts from the diagonal, later classes add to y
};
constexpr int kPhaseAll = 0;     // diagonal + every stored element
constexpr int kPhaseLocal = 1;   // diagonal + elements whose source this rank owns (needs no exchange); 2, 3: exchange rounds

// ---- cache-policy loads (PTX): the (index, code) stream is read exactly once per application, so
// it bypasses L1 and is marked evict-first in L2; the gathers of x are marked evict-last so that
// as much of the vector as possible stays resident in the 126 MB L2 between touches.
template <bool HINT>
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 p = 0;
  if constexpr (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool HINT>
__device__ __forceinline__ u64 l2_policy_evict_last() {
  u64 p = 0;
  if constexpr (HINT) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool HINT>
__device__ __forceinline__ u32 load_stream(u32 const* a, u64 pol) {
  if constexpr (HINT) {
    u32 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
  } else {
    return __ldg(a);
  }
}
template <bool HINT>
__device__ __forceinline__ u32 load_stream(dev_u16 const* a, u64 pol) {
  if constexpr (HINT) {
    u32 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
  } else {
    return (u32)__ldg(a);
  }
}
template <bool HINT>
__device__ __forceinline__ u32 load_stream(dev_u8 const* a, u64 pol) {
  if constexpr (HINT) {
    u32 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
  } else {
    return (u32)__ldg(a);
  }
}
// gather of one vector entry as the accumulator type
template <bool HINT> __device__ __forceinline__ double load_x(float const* a, u64 pol) {
  if constexpr (HINT) {
    float v;
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return (double)v;
  } else {
    return (double)__ldg(a);
  }
}
template <bool HINT> __device__ __forceinline__ double load_x(double const* a, u64 pol) {
  if constexpr (HINT) {
    double v;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
    return v;
  } else {
    return __ldg(a);
  }
}
template <bool HINT> __device__ __forceinline__ double2 load_x(float2 const* a, u64 pol) {
  if constexpr (HINT) {
    float x, y;
    asm("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(x), "=f"(y) : "l"(a), "l"(pol));
    return make_double2(x, y);
  } else {
    float2 v = __ldg(a);
    return make_double2(v.x, v.y);
  }
}
template <bool HINT> __device__ __forceinline__ double2 load_x(double2 const* a, u64 pol) {
  if constexpr (HINT) {
    double2 v;
    asm("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol));
    return v;
  } else {
    return __ldg(a);
  }
}

// y = H x from the cache: one warp per slice, coalesced index/code loads, read-only gathers of x,
// accumulation in the stored (= matrix-free) order.  Elements are taken U at a time: all U index
// and code loads are issued first, then the U gathers, then the U multiply-adds in order, so every
// thread keeps U independent gathers in flight.
// classes handled by pass `phase` (see cache_first_remote)
__device__ __forceinline__ void phase_classes(CacheView const& c, int phase, u32& lo, u32& hi) {
  u32 const fr = cache_first_remote(c.window);
  if (phase == kPhaseAll) { lo = 0; hi = c.n_classes; }
  else if (phase == kPhaseLocal) { lo = 0; hi = fr < c.n_classes ? fr : c.n_classes; }
  else { lo = fr + (u32)phase - 2u; hi = lo + 1u; }
}

template <class T, int NB, class Code, bool SYM, bool HINT, int U>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS, Traits<T>::cplx ? 1 : (U == 8 ? 5 : 6)) cached_matvec_kernel(CachedParams p) {
  typedef Traits<T> TR;
  typedef typename TR::Acc Acc;
  constexpr bool CPLX = TR::cplx;
  static_assert(NB == 1, "several columns go through cached_block_kernel");
  T const* __restrict__ x = static_cast<T const*>(p.x);
  T* __restrict__ y = static_cast<T*>(p.y);
  u32 const* __restrict__ cidx = p.cache.idx;
  Code const* __restrict__ ccode = static_cast<Code const*>(p.cache.code);
  double const* __restrict__ table = p.cache.table;
  if constexpr (sizeof(Code) == 1) {  // at most 256 codes: the coefficient table lives in shared memory
    __shared__ double s_table[3 * 256];
    for (u32 k = threadIdx.x; k < 3 * p.cache.n_codes; k += blockDim.x) s_table[k] = p.cache.table[k];
    __syncthreads();
    table = s_table;
  }
  // per warp: the window of x around the slice it is working on (CacheView, window class)
  __shared__ T s_window[SPED_KERNEL_THREADS / 32][kWindowEntries];
  u32 const warp_lanes = blockDim.x < 32u ? blockDim.x : 32u;  // 1 in the single-threaded host emulation
  u32 const lane = threadIdx.x % warp_lanes;
  T* const win = s_window[threadIdx.x / warp_lanes % (SPED_KERNEL_THREADS / 32)];
  u64 const pol_stream = l2_policy_evict_first<HINT>();
  u64 const pol_x = l2_policy_evict_last<HINT>();
  u64 const self0 = (u64)p.ctx.dist.rank * p.ctx.dist.chunk;  // this rank's shard inside the replicated x
  u64 const n_rows = p.ctx.dist.n_local;
  u32 cls_lo, cls_hi;
  phase_classes(p.cache, p.phase, cls_lo, cls_hi);
  bool const use_window = p.cache.window != 0 && cls_lo == 0;
  // warp-uniform loop over slices: i0 is the first row of the warp's slice (row_lo is a multiple of 32)
  for (u64 i0 = p.row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < p.row_hi; i0 += (u64)gridDim.x * blockDim.x) {
    u64 const i = i0 + lane;
    bool const active = i < p.row_hi;
    if (use_window) {  // stage x[s - W, s + 32 + W) of this rank's shard, s = first row of the slice (zeros outside the shard)
      u64 const sbase = i0 & ~(u64)31;  // == i0 on the GPU (blocks and row ranges are multiples of 32 rows)
      __syncwarp();
      for (u32 k = lane; k < kWindowEntries; k += warp_lanes) {
        u64 const local = sbase + k - kWindow;  // wraps below zero: caught by the bound
        T v{};
        if (local < n_rows) v = x[self0 + local];
        win[k] = v;
      }
      __syncwarp();
    }
    if (!active) continue;
    double inv_nr = 1.0;
    if constexpr (SYM) {
      u64 const row = dist_local_to_global(p.ctx.dist, i);
      inv_nr = 1.0 / __ldg(p.ctx.norm_table + __ldg(p.ctx.index.stab + row));
    }
    Acc acc;
    if (cls_lo == 0) {  // start from the diagonal term
      double dre = __ldg(p.diag_re + i);
      acc = acc_zero(Acc());
      Acc xv = load_x<HINT>(x + self0 + i, pol_x);
      if constexpr (CPLX) {
        double dim_ = p.diag_im ? __ldg(p.diag_im + i) : 0.0;
        acc_fma(acc, make_double2(dre, dim_), xv);
      } else {
        acc_fma(acc, dre, xv);
      }
    } else {  // continue from what the passes over the earlier classes stored
      acc = TR::load(y + i);
    }
    u64 const slice_base = __ldg(p.cache.slice_off + (i >> 5)) + (i & 31);
    // the stored elements of this lane: per source class first the elements that carry the default
    // coefficient (front of the class region, no code: their x entries are summed and multiplied
    // once), then the coded ones (back of the region, downwards).  One copy of the loop body: a
    // second inlined copy costs 15 registers and with them a resident block per SM.
    u32 const width = (u32)((__ldg(p.cache.slice_off + (i >> 5) + 1) - __ldg(p.cache.slice_off + (i >> 5))) >> 5);
#pragma unroll 1
    for (u32 seg = 2 * cls_lo; seg < 2 * cls_hi; ++seg) {
      u32 const cls = seg >> 1;
      bool const coded = (seg & 1u) != 0;
      u32 const len = __ldg(p.cache.len + (u64)seg * n_rows + i);
      u32 const lo = cls == 0 ? 0u : __ldg(p.cache.slice_start + 3 * (i >> 5) + (cls - 1));
      u32 const hi = cls + 1 == p.cache.n_classes ? width : __ldg(p.cache.slice_start + 3 * (i >> 5) + cls);
      // element j sits at slot lo + j (default coefficient) or hi - 1 - j (coded)
      u32 const slot0 = coded ? hi - 1u : lo;
      int const sstep = coded ? -1 : 1;
      Acc part = acc_zero(Acc());  // this segment: sum of w x (coded) or of x (default coefficient)
      if (use_window && cls == 0) {
        // window class: the slot holds an offset into the staged window
        for (u32 j = 0; j < len; ++j) {
          u64 const pos = slice_base + (u64)(slot0 + (u32)(sstep * (int)j)) * 32;
          u32 const off = load_stream<HINT>(cidx + pos, pol_stream);
          Acc const xv = TR::to_acc(win[off]);
          if constexpr (CPLX) {
            double2 w = make_double2(1.0, 0.0);
            if (coded) {
              double const* t = table + 3 * load_stream<HINT>(ccode + pos, pol_stream);
              w = make_double2(t[0], t[1]);
              if constexpr (SYM) {
                double const scale = t[2] * inv_nr;
                w.x *= scale;
                w.y *= scale;
              }
            }
            acc_fma(part, w, xv);
          } else {
            double w = 1.0;
            if (coded) {
              double const* t = table + 3 * load_stream<HINT>(ccode + pos, pol_stream);
              w = t[0];
              if constexpr (SYM) w = w * (t[2] * inv_nr);
            }
            acc_fma(part, w, xv);
          }
        }
      } else {
        for (u32 j0 = 0; j0 < len; j0 += U) {
          // `coded` is laundered through a volatile move so that the compiler does not unswitch the
          // loop on it: two specialised copies of the body cost 30 registers (80 instead of 48)
          u32 coded_i = seg & 1u;
#if defined(__CUDA_ARCH__)
          asm volatile("mov.u32 %0, %1;" : "=r"(coded_i) : "r"(seg & 1u));
#endif
          bool const coded_l = coded_i != 0;
          u32 idx[U], code[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            bool const live = j0 + u < len;
            u64 const pos = slice_base + (u64)(slot0 + (u32)(sstep * (int)(j0 + u))) * 32;
            idx[u] = live ? load_stream<HINT>(cidx + pos, pol_stream) : (u32)(self0 + i);
            code[u] = (live && coded_l) ? load_stream<HINT>(ccode + pos, pol_stream) : 0u;
          }
          Acc xv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) xv[u] = load_x<HINT>(x + idx[u], pol_x);  // dead slots re-read x[self]: an L1 hit
          // one accumulate path for both kinds: coded elements look their coefficient up, the
          // others add x itself (the common coefficient is applied once, after the loop)
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (j0 + u < len) {
              if constexpr (CPLX) {
                double2 w = make_double2(1.0, 0.0);
                if (coded_l) {
                  double const* t = table + 3 * code[u];
                  w = make_double2(t[0], t[1]);
                  if constexpr (SYM) {
                    double const scale = t[2] * inv_nr;
                    w.x *= scale;
                    w.y *= scale;
                  }
                }
                acc_fma(part, w, xv[u]);
              } else {
                double w = 1.0;
                if (coded_l) {
                  double const* t = table + 3 * code[u];
                  w = t[0];
                  if constexpr (SYM) w = w * (t[2] * inv_nr);
                }
                acc_fma(part, w, xv[u]);
              }
            }
          }
        }
      }
      if (len) {  // acc += part (coded) or w_default * part
        double const* t = table + 3 * p.cache.default_code;  // only dereferenced for a non-empty default segment
        if constexpr (CPLX) {
          double2 w = make_double2(1.0, 0.0);
          if (!coded) {
            w = make_double2(t[0], t[1]);
            if constexpr (SYM) {
              double const scale = t[2] * inv_nr;
              w.x *= scale;
              w.y *= scale;
            }
          }
          acc_fma(acc, w, part);
        } else {
          double w = 1.0;
          if (!coded) {
            w = t[0];
            if constexpr (SYM) w = w * (t[2] * inv_nr);
          }
          acc_fma(acc, w, part);
        }
      }
    }
    TR::store(y + i, acc);
  }
}

#if defined(__CUDA_ARCH__)
constexpr bool kBlockHint = true;   // cache-policy loads (PTX)
#else
constexpr bool kBlockHint = false;  // host pass / host emulation: plain loads
#endif

// ---- block applications (ncols > 1) ----------------------------------------------------------
// The gathers, not the bytes, bound the streaming kernel, so a block of vectors is first
// interleaved ([position][column], NB = 2 or 4 columns, missing ones zero): the 16- or 32-byte
// vector load that fetches an element's source for one column then brings the other columns with
// it, and the index/code stream is read once for the whole block.
template <class T, int NB>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) interleave_kernel(T const* x, u64 xs, u32 ncols, u64 n, T* out) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      T v{};
      if (c < (int)ncols) v = x[(u64)c * xs + i];
      out[i * NB + c] = v;
    }
  }
}

// NB consecutive entries of the interleaved block as accumulator values.  `volatile`: the compiler
// must not sink these loads into the per-element conditionals (that would serialise the gathers).
#if defined(__CUDA_ARCH__)
template <int NB>
__device__ __forceinline__ void load_xrow(double const* a, u64 pol, double (&out)[NB]) {
#pragma unroll
  for (int k = 0; k < NB; k += 2)
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(out[k]), "=d"(out[k + 1]) : "l"(a + k), "l"(pol));
}
template <int NB>
__device__ __forceinline__ void load_xrow(float const* a, u64 pol, double (&out)[NB]) {
  if constexpr (NB == 4) {
    float v0, v1, v2, v3;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "l"(a), "l"(pol));
    out[0] = v0; out[1] = v1; out[2] = v2; out[3] = v3;
  } else {
    float v0, v1;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v0), "=f"(v1) : "l"(a), "l"(pol));
    out[0] = v0; out[1] = v1;
  }
}
template <int NB>
__device__ __forceinline__ void load_xrow(float2 const* a, u64 pol, double2 (&out)[NB]) {
#pragma unroll
  for (int k = 0; k < NB; k += 2) {
    float v0, v1, v2, v3;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "l"(a + k), "l"(pol));
    out[k] = make_double2(v0, v1);
    out[k + 1] = make_double2(v2, v3);
  }
}
template <int NB>
__device__ __forceinline__ void load_xrow(double2 const* a, u64 pol, double2 (&out)[NB]) {
#pragma unroll
  for (int k = 0; k < NB; ++k)
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(out[k].x), "=d"(out[k].y) : "l"(a + k), "l"(pol));
}

#else  // host emulation: plain loads
template <int NB, class T, class A>
inline void load_xrow(T const* a, u64, A (&out)[NB]) {
  for (int k = 0; k < NB; ++k) out[k] = Traits<T>::load(a + k);
}
#endif

template <class T, int NB, class Code, bool SYM, int U>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) cached_block_kernel(CachedParams p) {
  typedef Traits<T> TR;
  typedef typename TR::Acc Acc;
  constexpr bool CPLX = TR::cplx;
  T const* __restrict__ xt = static_cast<T const*>(p.x);  // interleaved: entry (pos, c) at pos * NB + c
  T* __restrict__ y = static_cast<T*>(p.y);
  u32 const* __restrict__ cidx = p.cache.idx;
  Code const* __restrict__ ccode = static_cast<Code const*>(p.cache.code);
  double const* __restrict__ table = p.cache.table;
  if constexpr (sizeof(Code) == 1) {
    __shared__ double s_table[3 * 256];
    for (u32 k = threadIdx.x; k < 3 * p.cache.n_codes; k += blockDim.x) s_table[k] = p.cache.table[k];
    __syncthreads();
    table = s_table;
  }
  u64 const pol_stream = l2_policy_evict_first<kBlockHint>();
  u64 const pol_x = l2_policy_evict_last<kBlockHint>();
  u64 const self0 = (u64)p.ctx.dist.rank * p.ctx.dist.chunk;
  u64 const n_rows = p.ctx.dist.n_local;
  for (u64 i = p.row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.row_hi; i += (u64)gridDim.x * blockDim.x) {
    double inv_nr = 1.0;
    if constexpr (SYM) {
      u64 const row = dist_local_to_global(p.ctx.dist, i);
      inv_nr = 1.0 / __ldg(p.ctx.norm_table + __ldg(p.ctx.index.stab + row));
    }
    Acc acc[NB];
    {
      Acc xv[NB];
      load_xrow<NB>(xt + (self0 + i) * NB, pol_x, xv);
      double const dre = __ldg(p.diag_re + i);
      double const dim_ = (CPLX && p.diag_im) ? __ldg(p.diag_im + i) : 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        acc[c] = acc_zero(Acc());
        if constexpr (CPLX) acc_fma(acc[c], make_double2(dre, dim_), xv[c]);
        else acc_fma(acc[c], dre, xv[c]);
      }
    }
    u64 const slice_base = __ldg(p.cache.slice_off + (i >> 5)) + (i & 31);
    u32 const width = (u32)((__ldg(p.cache.slice_off + (i >> 5) + 1) - __ldg(p.cache.slice_off + (i >> 5))) >> 5);
#pragma unroll 1
    for (u32 seg = 0; seg < 2 * p.cache.n_classes; ++seg) {  // (class, default-coefficient | coded), see CacheView
      u32 const cls = seg >> 1;
      bool const coded = (seg & 1u) != 0;
      u32 const len = __ldg(p.cache.len + (u64)seg * n_rows + i);
      u32 const lo = cls == 0 ? 0u : __ldg(p.cache.slice_start + 3 * (i >> 5) + (cls - 1));
      u32 const hi = cls + 1 == p.cache.n_classes ? width : __ldg(p.cache.slice_start + 3 * (i >> 5) + cls);
      // window class: the slot holds an offset into the slice's window; here it is turned back into a position
      bool const windowed = cls == 0 && p.cache.window != 0;
      u32 const window_base = (u32)(self0 + (i & ~(u64)31)) - kWindow;
      long long const step = coded ? -32ll : 32ll;
      u64 const base = slice_base + (u64)(coded ? hi - 1u : lo) * 32;
      for (u32 j0 = 0; j0 < len; j0 += U) {
        u32 idx[U], code[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          bool const live = j0 + u < len;
          u64 const pos = (u64)((long long)base + step * (long long)(j0 + u));
          idx[u] = live ? load_stream<kBlockHint>(cidx + pos, pol_stream) + (windowed ? window_base : 0u) : (u32)(self0 + i);
          code[u] = live ? (coded ? load_stream<kBlockHint>(ccode + pos, pol_stream) : p.cache.default_code) : 0u;
        }
        Acc xv[U][NB];
#pragma unroll
        for (int u = 0; u < U; ++u) load_xrow<NB>(xt + (u64)idx[u] * NB, pol_x, xv[u]);
        // NOTE (SASS, not yet measured): ptxas software-pipelines this loop two gathers deep at 40
        // registers (6 blocks/SM) instead of issuing all U first; a warp fence is hoisted above the
        // gathers and a block fence costs a MEMBAR.SC -- to be tuned on hardware.
        // branch-free: a dead slot multiplies x[self] by zero, so that the gathers above cannot be
        // sunk into per-element conditionals (which serialises them)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          bool const live = j0 + u < len;
          double const* t = table + 3 * code[u];
          if constexpr (CPLX) {
            double2 w = make_double2(live ? t[0] : 0.0, live ? t[1] : 0.0);
            if constexpr (SYM) {
              double const scale = t[2] * inv_nr;
              w.x *= scale;
              w.y *= scale;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) acc_fma(acc[c], w, xv[u][c]);
          } else {
            double w = live ? t[0] : 0.0;
            if constexpr (SYM) w = w * (t[2] * inv_nr);
#pragma unroll
            for (int c = 0; c < NB; ++c) acc_fma(acc[c], w, xv[u][c]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (c < (int)p.ncols) TR::store(y + (u64)c * p.ys + i, acc[c]);
  }
}


}  // namespace sped
