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
  int phase;           // kPhaseAll, or one class of a two-class cache (CacheView)
};
constexpr int kPhaseAll = 0;     // diagonal + every stored element
constexpr int kPhaseLocal = 1;   // diagonal + elements whose source this rank owns (needs no all-gather)
constexpr int kPhaseRemote = 2;  // y += remote-source elements

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
template <class T, int NB, class Code, bool SYM, bool HINT, int U>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) cached_matvec_kernel(CachedParams p) {
  typedef Traits<T> TR;
  typedef typename TR::Acc Acc;
  constexpr bool CPLX = TR::cplx;
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
  u64 const pol_stream = l2_policy_evict_first<HINT>();
  u64 const pol_x = l2_policy_evict_last<HINT>();
  u64 const self0 = (u64)p.ctx.dist.rank * p.ctx.dist.chunk;  // this rank's shard inside the replicated x
  bool const two = p.cache.len_remote != nullptr;
  for (u64 i = p.row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.row_hi; i += (u64)gridDim.x * blockDim.x) {
    double inv_nr = 1.0;
    if constexpr (SYM) {
      u64 const row = dist_local_to_global(p.ctx.dist, i);
      inv_nr = 1.0 / __ldg(p.ctx.norm_table + __ldg(p.ctx.index.stab + row));
    }
    Acc acc[NB];
    if (p.phase != kPhaseRemote) {  // start from the diagonal term
      double dre = __ldg(p.diag_re + i);
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        acc[c] = acc_zero(Acc());
        if (c < (int)p.ncols) {
          Acc xv = load_x<HINT>(x + (u64)c * p.xs + self0 + i, pol_x);
          if constexpr (CPLX) {
            double dim_ = p.diag_im ? __ldg(p.diag_im + i) : 0.0;
            acc_fma(acc[c], make_double2(dre, dim_), xv);
          } else {
            acc_fma(acc[c], dre, xv);
          }
        }
      }
    } else {  // continue from what the local pass stored
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        acc[c] = acc_zero(Acc());
        if (c < (int)p.ncols) acc[c] = TR::load(y + (u64)c * p.ys + i);
      }
    }
    u64 const slice_base = __ldg(p.cache.slice_off + (i >> 5)) + (i & 31);
    // the stored elements of this lane, class by class, in stored order (one copy of the loop body:
    // a second inlined copy costs 15 registers and with them a resident block per SM)
    u32 first = 0, len = p.phase != kPhaseRemote ? __ldg(p.cache.len + i) : 0u;
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      if (seg == 1) {
        if (!two || p.phase == kPhaseLocal) break;
        first = __ldg(p.cache.slice_wl + (i >> 5));
        len = __ldg(p.cache.len_remote + i);
      }
      u64 const base = slice_base + (u64)first * 32;
      for (u32 j0 = 0; j0 < len; j0 += U) {
        u32 idx[U], code[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          bool const live = j0 + u < len;
          u64 const pos = base + (u64)(j0 + u) * 32;
          idx[u] = live ? load_stream<HINT>(cidx + pos, pol_stream) : (u32)(self0 + i);
          code[u] = live ? load_stream<HINT>(ccode + pos, pol_stream) : 0u;
        }
        if constexpr (NB == 1) {
          Acc xv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) xv[u] = load_x<HINT>(x + idx[u], pol_x);  // dead slots re-read x[self]: an L1 hit
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (j0 + u < len) {
              double const* t = table + 3 * code[u];
              if constexpr (CPLX) {
                double2 w = make_double2(t[0], t[1]);
                if constexpr (SYM) {
                  double const scale = t[2] * inv_nr;
                  w.x *= scale;
                  w.y *= scale;
                }
                acc_fma(acc[0], w, xv[u]);
              } else {
                double w = t[0];
                if constexpr (SYM) w = w * (t[2] * inv_nr);
                acc_fma(acc[0], w, xv[u]);
              }
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (j0 + u < len) {
              double const* t = table + 3 * code[u];
              if constexpr (CPLX) {
                double2 w = make_double2(t[0], t[1]);
                if constexpr (SYM) {
                  double const scale = t[2] * inv_nr;
                  w.x *= scale;
                  w.y *= scale;
                }
#pragma unroll
                for (int c = 0; c < NB; ++c)
                  if (c < (int)p.ncols) acc_fma(acc[c], w, load_x<HINT>(x + (u64)c * p.xs + idx[u], pol_x));
              } else {
                double w = t[0];
                if constexpr (SYM) w = w * (t[2] * inv_nr);
#pragma unroll
                for (int c = 0; c < NB; ++c)
                  if (c < (int)p.ncols) acc_fma(acc[c], w, load_x<HINT>(x + (u64)c * p.xs + idx[u], pol_x));
              }
            }
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
