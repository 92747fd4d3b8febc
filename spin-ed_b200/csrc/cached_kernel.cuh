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
  float mean_row_length;  // slots per local row (host side: picks the batch length of the streaming kernel)
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

// Predicated forms: a dead slot (past the end of the lane's list) issues no memory request at all
// and yields zero.  Written as predicated PTX inside one volatile asm so that the compiler can
// neither sink the load into a later conditional (which would serialise the gathers) nor
// speculate it.
template <bool HINT>
__device__ __forceinline__ u32 load_stream_if(bool live, u32 const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  u32 v;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%2], %3;\n\t}"
                 : "=r"(v) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.nc.u32 %0, [%2];\n\t}"
                 : "=r"(v) : "r"((u32)live), "l"(a));
  return v;
#else
  (void)pol;
  return live ? *a : 0u;
#endif
}
template <bool HINT>
__device__ __forceinline__ u32 load_stream_if(bool live, dev_u8 const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  u32 v;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%2], %3;\n\t}"
                 : "=r"(v) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.nc.u8 %0, [%2];\n\t}"
                 : "=r"(v) : "r"((u32)live), "l"(a));
  return v;
#else
  (void)pol;
  return live ? (u32)*a : 0u;
#endif
}
template <bool HINT>
__device__ __forceinline__ u32 load_stream_if(bool live, dev_u16 const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  u32 v;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%2], %3;\n\t}"
                 : "=r"(v) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p ld.global.nc.u16 %0, [%2];\n\t}"
                 : "=r"(v) : "r"((u32)live), "l"(a));
  return v;
#else
  (void)pol;
  return live ? (u32)*a : 0u;
#endif
}
template <bool HINT> __device__ __forceinline__ double load_x_if(bool live, double const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  double v;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.global.nc.L2::cache_hint.f64 %0, [%2], %3;\n\t}"
                 : "=d"(v) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.global.nc.f64 %0, [%2];\n\t}"
                 : "=d"(v) : "r"((u32)live), "l"(a));
  return v;
#else
  (void)pol;
  return live ? *a : 0.0;
#endif
}
template <bool HINT> __device__ __forceinline__ double load_x_if(bool live, float const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  float v;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.f32 %0, 0f00000000;\n\t@p ld.global.nc.L2::cache_hint.f32 %0, [%2], %3;\n\t}"
                 : "=f"(v) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.f32 %0, 0f00000000;\n\t@p ld.global.nc.f32 %0, [%2];\n\t}"
                 : "=f"(v) : "r"((u32)live), "l"(a));
  return (double)v;
#else
  (void)pol;
  return live ? (double)*a : 0.0;
#endif
}
template <bool HINT> __device__ __forceinline__ double2 load_x_if(bool live, double2 const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  double2 v;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t@p ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%3], %4;\n\t}"
                 : "=d"(v.x), "=d"(v.y) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t@p ld.global.nc.v2.f64 {%0, %1}, [%3];\n\t}"
                 : "=d"(v.x), "=d"(v.y) : "r"((u32)live), "l"(a));
  return v;
#else
  (void)pol;
  return live ? *a : double2{0.0, 0.0};
#endif
}
template <bool HINT> __device__ __forceinline__ double2 load_x_if(bool live, float2 const* a, u64 pol) {
#if defined(__CUDA_ARCH__)
  float x, y;
  if constexpr (HINT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t@p ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%3], %4;\n\t}"
                 : "=f"(x), "=f"(y) : "r"((u32)live), "l"(a), "l"(pol));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t@p ld.global.nc.v2.f32 {%0, %1}, [%3];\n\t}"
                 : "=f"(x), "=f"(y) : "r"((u32)live), "l"(a));
  return make_double2(x, y);
#else
  (void)pol;
  return live ? double2{(double)a->x, (double)a->y} : double2{0.0, 0.0};
#endif
}

// classes handled by pass `phase`: 0 -- all; 1 + c -- class c only
__device__ __forceinline__ void phase_classes(CacheView const& c, int phase, u32& lo, u32& hi) {
  if (phase == kPhaseAll) { lo = 0; hi = c.n_classes; }
  else { lo = (u32)phase - 1u; hi = lo + 1u; }
}

// coefficient of a coded element: table entry (Re v, Im v, norm_s) times 1 / norm_r
template <bool CPLX, bool SYM> struct Coeff;
template <bool SYM> struct Coeff<false, SYM> {
  typedef double W;
  static __device__ __forceinline__ W get(double const* t, double inv_nr) {
    double w = t[0];
    if constexpr (SYM) w = w * (t[2] * inv_nr);
    return w;
  }
};
template <bool SYM> struct Coeff<true, SYM> {
  typedef double2 W;
  static __device__ __forceinline__ W get(double const* t, double inv_nr) {
    double2 w = make_double2(t[0], t[1]);
    if constexpr (SYM) {
      double const scale = t[2] * inv_nr;
      w.x *= scale;
      w.y *= scale;
    }
    return w;
  }
};
__device__ __forceinline__ void acc_add(double& a, double x) { a += x; }
__device__ __forceinline__ void acc_add(double2& a, double2 x) { a.x += x.x; a.y += x.y; }

// y = H x from the cache: one warp per slice, coalesced index (and code) loads, read-only gathers
// of x.  Per source class a row first takes the elements that carry the default coefficient (front
// of the class region, no code: their x entries are summed and multiplied once), then the coded
// ones (back of the region, downwards; each looks its coefficient up).  Elements are taken U at a
// time: all U index loads are issued first, then the U gathers, then the U accumulations in order,
// so every thread keeps U independent gathers in flight.
template <class T, int NB, class Code, bool SYM, bool HINT, int U, int MINB, bool PIPE = false>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS, Traits<T>::cplx ? 1 : MINB) cached_matvec_kernel(CachedParams p) {
  typedef Traits<T> TR;
  typedef typename TR::Acc Acc;
  constexpr bool CPLX = TR::cplx;
  typedef Coeff<CPLX, SYM> CF;
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
  u64 const pol_stream = l2_policy_evict_first<HINT>();
  u64 const pol_x = l2_policy_evict_last<HINT>();
  u64 const self0 = (u64)p.ctx.dist.rank * p.ctx.dist.chunk;  // this rank's shard inside the replicated x
  u64 const n_rows = p.ctx.dist.n_local;
  u32 cls_lo, cls_hi;
  phase_classes(p.cache, p.phase, cls_lo, cls_hi);
  bool const has_default = p.cache.default_code < p.cache.n_codes;
  typename CF::W w_default{};
  for (u64 i = p.row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.row_hi; i += (u64)gridDim.x * blockDim.x) {
    double inv_nr = 1.0;
    if constexpr (SYM) {
      u64 const row = dist_local_to_global(p.ctx.dist, i);
      inv_nr = 1.0 / __ldg(p.ctx.norm_table + __ldg(p.ctx.index.stab + row));
    }
    if (has_default) w_default = CF::get(table + 3 * p.cache.default_code, inv_nr);
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
    u64 const slice = i >> 5;
    u64 const slice_base = __ldg(p.cache.slice_off + slice) + (i & 31);
    u32 const width = (u32)((__ldg(p.cache.slice_off + slice + 1) - __ldg(p.cache.slice_off + slice)) >> 5);
#pragma unroll 1
    for (u32 cls = cls_lo; cls < cls_hi; ++cls) {
      u32 const lo = cls == 0 ? 0u : __ldg(p.cache.slice_start + kClassStride * slice + (cls - 1));
      u32 const hi = cls + 1 == p.cache.n_classes ? width : __ldg(p.cache.slice_start + kClassStride * slice + cls);
      u32 const nd = __ldg(p.cache.len + (u64)(2 * cls) * n_rows + i);
      u32 const nx = __ldg(p.cache.len + (u64)(2 * cls + 1) * n_rows + i);
      // default coefficient: element j at slot lo + j
      if (nd) {
        Acc part = acc_zero(Acc());
        u32 const* q = cidx + slice_base + (u64)lo * 32;
        if constexpr (PIPE) {
          // software pipeline: the positions of the NEXT batch are requested right behind the gathers of
          // the current one, so their latency hides behind the gathers instead of preceding them
          u32 idx[U];
#pragma unroll
          for (int u = 0; u < U; ++u) idx[u] = load_stream_if<HINT>((u32)u < nd, q + 32 * u, pol_stream);
#pragma unroll 1
          for (u32 j0 = 0; j0 < nd; j0 += U) {
            Acc xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = load_x_if<HINT>(j0 + u < nd, x + idx[u], pol_x);
            q += 32 * U;
            u32 nxt[U];
#pragma unroll
            for (int u = 0; u < U; ++u) nxt[u] = load_stream_if<HINT>(j0 + U + u < nd, q + 32 * u, pol_stream);
#pragma unroll
            for (int u = 0; u < U; ++u) acc_add(part, xv[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) idx[u] = nxt[u];
          }
        } else {
#pragma unroll 1
          for (u32 j0 = 0; j0 < nd; j0 += U, q += 32 * U) {
            u32 idx[U];
#pragma unroll
            for (int u = 0; u < U; ++u) idx[u] = load_stream_if<HINT>(j0 + u < nd, q + 32 * u, pol_stream);
            Acc xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = load_x_if<HINT>(j0 + u < nd, x + idx[u], pol_x);  // dead slots: no request, zero
#pragma unroll
            for (int u = 0; u < U; ++u) acc_add(part, xv[u]);
          }
        }
        acc_fma(acc, w_default, part);
      }
      // coded: element j at slot hi - 1 - j
      if (nx) {
        u64 const top = slice_base + (u64)(hi - 1u) * 32;
        u64 const cbase = __ldg(p.cache.code_off + slice * p.cache.n_classes + cls) + (i & 31);  // compact code stream
#pragma unroll 1
        for (u32 j0 = 0; j0 < nx; j0 += U) {
          u32 idx[U], code[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            bool const live = j0 + u < nx;
            idx[u] = load_stream_if<HINT>(live, cidx + (top - (u64)(j0 + u) * 32), pol_stream);
            code[u] = load_stream_if<HINT>(live, ccode + (cbase + (u64)(j0 + u) * 32), pol_stream);
          }
          Acc xv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) xv[u] = load_x_if<HINT>(j0 + u < nx, x + idx[u], pol_x);
#pragma unroll
          for (int u = 0; u < U; ++u) acc_fma(acc, CF::get(table + 3 * code[u], inv_nr), xv[u]);  // dead slots add w * 0
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

// scalars of a loaded block row -> accumulator-typed entries
template <int NB, class S> __device__ __forceinline__ void unpack_row(S const* v, double (&out)[NB]) {
#pragma unroll
  for (int k = 0; k < NB; ++k) out[k] = (double)v[k];
}
template <int NB, class S> __device__ __forceinline__ void unpack_row(S const* v, double2 (&out)[NB]) {
#pragma unroll
  for (int k = 0; k < NB; ++k) out[k] = make_double2((double)v[2 * k], (double)v[2 * k + 1]);
}

// predicated block-row load: a dead slot issues no request and yields zeros
template <class T> struct IsSingle { static constexpr bool value = false; };
template <> struct IsSingle<float> { static constexpr bool value = true; };
template <> struct IsSingle<float2> { static constexpr bool value = true; };

template <int NB, class T, class A>
__device__ __forceinline__ void load_xrow_if(bool live, T const* a, u64 pol, A (&out)[NB]) {
#if defined(__CUDA_ARCH__)
  if constexpr (IsSingle<T>::value) {
    constexpr int NF = (int)(sizeof(T) * NB / 4);  // 2, 4 or 8 floats
    float v[NF];
    float const* f = reinterpret_cast<float const*>(a);
    if constexpr (NF == 2) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
                   "@p ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%3], %4;\n\t}"
                   : "=f"(v[0]), "=f"(v[1]) : "r"((u32)live), "l"(f), "l"(pol));
    } else {
#pragma unroll
      for (int h = 0; h < NF / 4; ++h)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
                     "mov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
                     "@p ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%5], %6;\n\t}"
                     : "=f"(v[4 * h]), "=f"(v[4 * h + 1]), "=f"(v[4 * h + 2]), "=f"(v[4 * h + 3])
                     : "r"((u32)live), "l"(f + 4 * h), "l"(pol));
    }
    unpack_row<NB>(v, out);
  } else {
    constexpr int ND = (int)(sizeof(T) * NB / 8);  // 2, 4 or 8 doubles, two per load
    double v[ND];
    double const* d = reinterpret_cast<double const*>(a);
#pragma unroll
    for (int h = 0; h < ND / 2; ++h)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
                   "@p ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%3], %4;\n\t}"
                   : "=d"(v[2 * h]), "=d"(v[2 * h + 1]) : "r"((u32)live), "l"(d + 2 * h), "l"(pol));
    unpack_row<NB>(v, out);
  }
#else
  (void)pol;
  for (int k = 0; k < NB; ++k) out[k] = live ? Traits<T>::load(a + k) : acc_zero(A());
#endif
}

template <class T, int NB, class Code, bool SYM, int U>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS, (sizeof(T) * NB <= 16 ? 5 : 4)) cached_block_kernel(CachedParams p) {
  typedef Traits<T> TR;
  typedef typename TR::Acc Acc;
  constexpr bool CPLX = TR::cplx;
  typedef Coeff<CPLX, SYM> CF;
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
  u32 cls_lo, cls_hi;
  phase_classes(p.cache, p.phase, cls_lo, cls_hi);
  bool const has_default = p.cache.default_code < p.cache.n_codes;
  typename CF::W w_default{};
  for (u64 i = p.row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.row_hi; i += (u64)gridDim.x * blockDim.x) {
    double inv_nr = 1.0;
    if constexpr (SYM) {
      u64 const row = dist_local_to_global(p.ctx.dist, i);
      inv_nr = 1.0 / __ldg(p.ctx.norm_table + __ldg(p.ctx.index.stab + row));
    }
    if (has_default) w_default = CF::get(table + 3 * p.cache.default_code, inv_nr);
    Acc acc[NB];
    if (cls_lo == 0) {
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
    } else {
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = c < (int)p.ncols ? TR::load(y + (u64)c * p.ys + i) : acc_zero(Acc());
    }
    u64 const slice = i >> 5;
    u64 const slice_base = __ldg(p.cache.slice_off + slice) + (i & 31);
    u32 const width = (u32)((__ldg(p.cache.slice_off + slice + 1) - __ldg(p.cache.slice_off + slice)) >> 5);
#pragma unroll 1
    for (u32 cls = cls_lo; cls < cls_hi; ++cls) {
      u32 const lo = cls == 0 ? 0u : __ldg(p.cache.slice_start + kClassStride * slice + (cls - 1));
      u32 const hi = cls + 1 == p.cache.n_classes ? width : __ldg(p.cache.slice_start + kClassStride * slice + cls);
      u32 const nd = __ldg(p.cache.len + (u64)(2 * cls) * n_rows + i);
      u32 const nx = __ldg(p.cache.len + (u64)(2 * cls + 1) * n_rows + i);
      if (nd) {  // default coefficient: element j at slot lo + j; sum the block rows, multiply once
        Acc part[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) part[c] = acc_zero(Acc());
        u32 const* q = cidx + slice_base + (u64)lo * 32;
#pragma unroll 1
        for (u32 j0 = 0; j0 < nd; j0 += U, q += 32 * U) {
          u32 idx[U];
#pragma unroll
          for (int u = 0; u < U; ++u) idx[u] = load_stream_if<kBlockHint>(j0 + u < nd, q + 32 * u, pol_stream);
          Acc xv[U][NB];
#pragma unroll
          for (int u = 0; u < U; ++u) load_xrow_if<NB>(j0 + u < nd, xt + (u64)idx[u] * NB, pol_x, xv[u]);
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int c = 0; c < NB; ++c) acc_add(part[c], xv[u][c]);
          }
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) acc_fma(acc[c], w_default, part[c]);
      }
      if (nx) {  // coded: element j at slot hi - 1 - j, its code at entry j of the compact code stream
        u64 const top = slice_base + (u64)(hi - 1u) * 32;
        u64 const cbase = __ldg(p.cache.code_off + slice * p.cache.n_classes + cls) + (i & 31);
#pragma unroll 1
        for (u32 j0 = 0; j0 < nx; j0 += U) {
          u32 idx[U], code[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            bool const live = j0 + u < nx;
            idx[u] = load_stream_if<kBlockHint>(live, cidx + (top - (u64)(j0 + u) * 32), pol_stream);
            code[u] = load_stream_if<kBlockHint>(live, ccode + (cbase + (u64)(j0 + u) * 32), pol_stream);
          }
          Acc xv[U][NB];
#pragma unroll
          for (int u = 0; u < U; ++u) load_xrow_if<NB>(j0 + u < nx, xt + (u64)idx[u] * NB, pol_x, xv[u]);
#pragma unroll
          for (int u = 0; u < U; ++u) {  // dead slots add w * 0
            typename CF::W const w = CF::get(table + 3 * code[u], inv_nr);
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

// ---- compact code stream --------------------------------------------------------------------
// The fill writes the code of a coded element beside its position (one code per slot, a temporary);
// afterwards only the coded parts are kept: class c of slice s gets cw = max over its lanes of the
// coded elements, and entry j of lane l lives at code_off[s * n_classes + c] + 32 j + l.

// cw[s * n_classes + c] = longest coded list of class c among the lanes of slice s (one thread per
// slice), for the slices slice_lo .. slice_hi - 1
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) code_width_kernel(dev_u16 const* len, u64 n_local, u32 n_classes,
                                                                                             u64 slice_lo, u64 slice_hi, u32* cw) {
  for (u64 s = slice_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; s < slice_hi; s += (u64)gridDim.x * blockDim.x) {
    u64 const hi = (32 * s + 32 < n_local) ? 32 * s + 32 : n_local;
    for (u32 c = 0; c < n_classes; ++c) {
      u32 mx = 0;
      for (u64 i = 32 * s; i < hi; ++i) {
        u32 const v = __ldg(len + (u64)(2 * c + 1) * n_local + i);
        mx = v > mx ? v : mx;
      }
      cw[s * n_classes + c] = mx;
    }
  }
}

// The codes of the coded elements of the local rows row_lo .. row_hi - 1 in compact order.  `c.code`
// still is the one-per-slot temporary of the fill, which covers the slots from code_slot0 on;
// `chunk_off` holds the offsets of this chunk's regions relative to `out`:
// chunk_off[(s - row_lo / 32) * n_classes + cls].
template <class Code>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) code_compact_kernel(CacheView c, u64 n_local, u64 row_lo, u64 row_hi,
                                                                                               u64 code_slot0, u64 const* chunk_off, Code* out) {
  Code const* __restrict__ full = static_cast<Code const*>(c.code);
  for (u64 i = row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < row_hi; i += (u64)gridDim.x * blockDim.x) {
    u64 const slice = i >> 5;
    u64 const slice_base = __ldg(c.slice_off + slice) + (i & 31);
    u32 const width = (u32)((__ldg(c.slice_off + slice + 1) - __ldg(c.slice_off + slice)) >> 5);
    for (u32 cls = 0; cls < c.n_classes; ++cls) {
      u32 const nx = __ldg(c.len + (u64)(2 * cls + 1) * n_local + i);
      if (!nx) continue;
      u32 const hi = cls + 1 == c.n_classes ? width : __ldg(c.slice_start + kClassStride * slice + cls);
      u64 const top = slice_base + (u64)(hi - 1u) * 32 - code_slot0;
      u64 const cbase = __ldg(chunk_off + (slice - (row_lo >> 5)) * c.n_classes + cls) + (i & 31);
      for (u32 j = 0; j < nx; ++j) out[cbase + (u64)j * 32] = full[top - (u64)j * 32];
    }
  }
}

// ---- placement after a staging traversal (several source classes) ---------------------------
// The staging traversal (FillParams::stage) left, per lane, the elements of the row in traversal order --
// position and code -- and the per-class counts in `len`.  With the class widths known from those
// counts, this kernel lays them out exactly as the class-aware fill would: default-coefficient
// elements from the front of their class region, coded ones from its back with the code in the compact
// stream.  One canonicalisation per element instead of two (count + fill).
struct PlaceParams {
  RowDist dist;
  u64 const* stage_off;   // [n_slices + 1] staging offsets (cheap width bound)
  u32 const* stage_idx;
  void const* stage_code;  // u8 / u16 per staging slot
  CacheView out;           // slice_off, len, slice_start, n_classes, near, rounds, default_code, code_off of the final layout
  u32* idx;                // final positions
  void* code;              // final compact code stream
};

template <class Code>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) cache_place_kernel(PlaceParams q) {
  Code const* __restrict__ scode = static_cast<Code const*>(q.stage_code);
  Code* __restrict__ ocode = static_cast<Code*>(q.code);
  u64 const n_local = q.dist.n_local;
  u32 const nc = q.out.n_classes;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (u64)gridDim.x * blockDim.x) {
    u64 const slice = i >> 5, lane = i & 31;
    u64 const sbase = __ldg(q.stage_off + slice) + lane;
    u64 const fbase = __ldg(q.out.slice_off + slice) + lane;
    u32 const width = (u32)((__ldg(q.out.slice_off + slice + 1) - __ldg(q.out.slice_off + slice)) >> 5);
    u32 start[kMaxClasses + 1] = {0u, 0u, 0u, 0u};
    for (u32 c = 1; c <= (u32)kMaxClasses; ++c) start[c] = c < nc ? __ldg(q.out.slice_start + kClassStride * slice + (c - 1)) : width;
    u32 total = 0;
    for (u32 c = 0; c < nc; ++c)
      total += (u32)__ldg(q.out.len + (u64)(2 * c) * n_local + i) + (u32)__ldg(q.out.len + (u64)(2 * c + 1) * n_local + i);
    u32 cd[kMaxClasses] = {0u, 0u, 0u}, cx[kMaxClasses] = {0u, 0u, 0u};
    for (u32 t = 0; t < total; ++t) {
      u32 const pos = __ldg(q.stage_idx + sbase + (u64)t * 32);
      u32 const code = (u32)scode[sbase + (u64)t * 32];
      u32 const cls = dist_source_class(q.dist, pos, q.out.rounds, q.out.near);
      if (code == q.out.default_code) {
        q.idx[fbase + (u64)(start[cls] + cd[cls]) * 32] = pos;
        ++cd[cls];
      } else {
        u32 const j = cx[cls]++;
        q.idx[fbase + (u64)(start[cls + 1] - 1u - j) * 32] = pos;
        ocode[__ldg(q.out.code_off + slice * nc + cls) + lane + (u64)j * 32] = (Code)code;
      }
    }
  }
}

}  // namespace sped
