// matvec_kernel.cuh -- device code of the matrix-free symmetry-adapted matvec (kernel K3).
//
// Compiled twice: statically by nvcc into libsped.so (operator.cu, canonicalisation by the program
// interpreter of permprog.h) and at run time by NVRTC for sm_100a with the canonicalisation of
// one particular symmetry group emitted as straight-line code (jit.cpp).  Builtin types only.
//
// Pull form, one row per thread, no atomics, fixed summation order (terms as given, tuples as
// given, target local configuration ascending):
//   y[r] = d[r] x[r] + sum_{t, b != a} M_t[a][b] chi(g') (n_s / n_r) x[index(s)],   g'.r' = s,
// where r' is r with the tuple's bits replaced by b.  Replaces ls_operator_matmat of
// liblattice_symmetries (/root/reference/src/SpinED/Internal.hs:377-378,411-429).
#pragma once
#include "device_types.h"

namespace sped {

__host__ __device__ inline unsigned long terms_smem_bytes(TermsView const& t, bool cplx) {
  unsigned long b = ((unsigned long)t.n_bonds * sizeof(DevBond) + 15ul) & ~15ul;
  unsigned long p = ((unsigned long)t.pool_size * 8ul + 15ul) & ~15ul;
  unsigned long m = ((unsigned long)t.mask_size * 2ul + 15ul) & ~15ul;
  return b + p * (cplx ? 2ul : 1ul) + m;
}

// Cooperative copy of bonds, matrices and masks into shared memory; ends with __syncthreads().
template <bool CPLX>
__device__ __forceinline__ TermsView stage_terms(TermsView g, unsigned char* smem) {
  DevBond* bonds = reinterpret_cast<DevBond*>(smem);
  unsigned char* p = smem + (((unsigned long)g.n_bonds * sizeof(DevBond) + 15ul) & ~15ul);
  double* re = reinterpret_cast<double*>(p);
  p += ((unsigned long)g.pool_size * 8ul + 15ul) & ~15ul;
  double* im = nullptr;
  if (CPLX) {
    im = reinterpret_cast<double*>(p);
    p += ((unsigned long)g.pool_size * 8ul + 15ul) & ~15ul;
  }
  dev_u16* masks = reinterpret_cast<dev_u16*>(p);
  for (u32 i = threadIdx.x; i < g.n_bonds; i += blockDim.x) bonds[i] = g.bonds[i];
  for (u32 i = threadIdx.x; i < g.pool_size; i += blockDim.x) {
    re[i] = g.pool_re[i];
    if (CPLX) im[i] = g.pool_im[i];
  }
  for (u32 i = threadIdx.x; i < g.mask_size; i += blockDim.x) masks[i] = g.masks[i];
  __syncthreads();
  TermsView v = g;
  v.bonds = bonds;
  v.pool_re = re;
  v.pool_im = im;
  v.masks = masks;
  return v;
}

// ---- storage type T <-> accumulator (double or double2) ----
template <class T> struct Traits;
template <> struct Traits<float> {
  typedef double Acc;
  static constexpr bool cplx = false;
  static __device__ __forceinline__ Acc load(float const* p) { return (double)__ldg(p); }
  static __device__ __forceinline__ Acc to_acc(float v) { return (double)v; }
  static __device__ __forceinline__ void store(float* p, Acc v) { *p = (float)v; }
};
template <> struct Traits<double> {
  typedef double Acc;
  static constexpr bool cplx = false;
  static __device__ __forceinline__ Acc load(double const* p) { return __ldg(p); }
  static __device__ __forceinline__ Acc to_acc(double v) { return v; }
  static __device__ __forceinline__ void store(double* p, Acc v) { *p = v; }
};
template <> struct Traits<float2> {
  typedef double2 Acc;
  static constexpr bool cplx = true;
  static __device__ __forceinline__ Acc load(float2 const* p) { float2 v = __ldg(p); return make_double2(v.x, v.y); }
  static __device__ __forceinline__ Acc to_acc(float2 v) { return make_double2(v.x, v.y); }
  static __device__ __forceinline__ void store(float2* p, Acc v) { *p = make_float2((float)v.x, (float)v.y); }
};
template <> struct Traits<double2> {
  typedef double2 Acc;
  static constexpr bool cplx = true;
  static __device__ __forceinline__ Acc load(double2 const* p) { return __ldg(p); }
  static __device__ __forceinline__ Acc to_acc(double2 v) { return v; }
  static __device__ __forceinline__ void store(double2* p, Acc v) { *p = v; }
};

__device__ __forceinline__ double acc_zero(double) { return 0.0; }
__device__ __forceinline__ double2 acc_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ void acc_fma(double& acc, double w, double x) { acc += w * x; }
__device__ __forceinline__ void acc_fma(double2& acc, double2 w, double2 x) {
  acc.x += w.x * x.x - w.y * x.y;
  acc.y += w.x * x.y + w.y * x.x;
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Index of representative `rep` in the sorted array, or ~0 if absent: prefix bucket, then a
// binary search inside the bucket.
__device__ __forceinline__ u64 lookup_index(BasisIndex const& ix, u64 rep) {
  if (ix.direct) return rep;
  u64 prefix = rep >> ix.bucket_shift;
  if (prefix >= ix.bucket_count) return ~(u64)0;
  u64 lo, hi;
  if (ix.bucket_wide) {
    u64 const* b = static_cast<u64 const*>(ix.bucket);
    lo = __ldg(b + prefix);
    hi = __ldg(b + prefix + 1);
  } else {
    u32 const* b = static_cast<u32 const*>(ix.bucket);
    lo = __ldg(b + prefix);
    hi = __ldg(b + prefix + 1);
  }
  while (lo < hi) {
    u64 mid = lo + ((hi - lo) >> 1);
    u64 v = __ldg(ix.reps + mid);
    if (v < rep) lo = mid + 1;
    else hi = mid;
  }
  if (lo < ix.n_states && __ldg(ix.reps + lo) == rep) return lo;
  return ~(u64)0;
}

// Walks the non-zero off-diagonal transitions of row word r in the fixed order and calls
// sink(bond, a, b, flipped_word).  The inner search loop is cheap and may diverge; the sink call
// site is reached by all lanes that still have work, so canonicalisation runs with full warps.
template <class Sink>
__device__ __forceinline__ void for_each_transition(TermsView const& T, u64 r, Sink&& sink) {
  u32 bond = 0;
  u32 bits = 0, a = 0;
  DevBond bd;
  bd.sites = 0; bd.moff = 0; bd.zoff = 0; bd.k = 0; bd.pad_ = 0;
  for (;;) {
    while (bits == 0 && bond < T.n_bonds) {
      bd = T.bonds[bond++];
      a = 0;
      for (u32 j = 0; j < bd.k; ++j) a |= (u32)((r >> ((bd.sites >> (8 * j)) & 0xffu)) & 1ul) << (bd.k - 1 - j);
      bits = T.masks[bd.zoff + a];
    }
    if (bits == 0) break;
    u32 b = (u32)__ffs((int)bits) - 1u;
    bits &= bits - 1;
    u32 diff = a ^ b;
    u64 rp = r;
    for (u32 j = 0; j < bd.k; ++j) rp ^= (u64)((diff >> (bd.k - 1 - j)) & 1u) << ((bd.sites >> (8 * j)) & 0xffu);
    sink(bd, a, b, rp);
  }
}

// The body of the matvec kernel.  Canon maps a basis word to (representative, phase numerator of
// the element reaching it); Canon::symmetric is false for the trivial group.
template <class T, int NB, class Canon>
__device__ __forceinline__ void matvec_rows(MatvecParams const& p, TermsView const& terms, Canon const& canon) {
  typedef Traits<T> TR;
  typedef typename TR::Acc Acc;
  constexpr bool CPLX = TR::cplx;
  constexpr bool SYM = Canon::symmetric;
  BasisIndex const ix = p.ctx.index;
  T const* x = static_cast<T const*>(p.x);
  T* y = static_cast<T*>(p.y);
  RowDist const dist = p.ctx.dist;
  u64 const n_local = dist.n_local;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (u64)gridDim.x * blockDim.x) {
    u64 const row = dist_local_to_global(dist, i);
    u64 const r = ix.direct ? row : __ldg(ix.reps + row);
    double inv_nr = 1.0;
    if (SYM) inv_nr = 1.0 / __ldg(p.ctx.norm_table + __ldg(ix.stab + row));
    Acc acc[NB];
    {
      double dre = __ldg(p.diag_re + i);
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        acc[c] = acc_zero(Acc());
        if (c < (int)p.ncols) {
          Acc xv = TR::load(x + (u64)c * p.xs + (u64)dist.rank * dist.chunk + i);
          if constexpr (CPLX) {
            double dim_ = p.diag_im ? __ldg(p.diag_im + i) : 0.0;
            acc_fma(acc[c], make_double2(dre, dim_), xv);
          } else {
            acc_fma(acc[c], dre, xv);
          }
        }
      }
    }
    for_each_transition(terms, r, [&](DevBond const& bd, u32 a, u32 b, u64 rp) {
      u32 const dim = 1u << bd.k;
      u64 rep = rp;
      int ph = 0;
      if (SYM) canon(rp, rep, ph);
      u64 idx = lookup_index(ix, rep);
      if (idx == ~(u64)0) return;
      u64 const pos = dist_global_to_pos(dist, idx);  // where x[idx] sits in the replicated vector
      double hre = terms.pool_re[bd.moff + a * dim + b];
      double scale = 1.0;
      if (SYM) scale = __ldg(p.ctx.norm_table + __ldg(ix.stab + idx)) * inv_nr;
      if constexpr (CPLX) {
        double2 w = make_double2(hre, terms.pool_im[bd.moff + a * dim + b]);
        if (SYM) {
          double2 chi = make_double2(__ldg(p.ctx.chi_table + 2 * ph), __ldg(p.ctx.chi_table + 2 * ph + 1));
          w = cmul(w, chi);
          w.x *= scale;
          w.y *= scale;
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c < (int)p.ncols) acc_fma(acc[c], w, TR::load(x + (u64)c * p.xs + pos));
      } else {
        double w = hre;
        if (SYM) w = (ph == 0 ? w : -w) * scale;
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c < (int)p.ncols) acc_fma(acc[c], w, TR::load(x + (u64)c * p.xs + pos));
      }
    });
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (c < (int)p.ncols) TR::store(y + (u64)c * p.ys + i, acc[c]);
  }
}

// Fills the operator cache: the same traversal as matvec_rows, but instead of gathering x the
// (target index, coefficient code) of every element that exists is written to its slot.  With
// several classes (CacheView) a first pass with count_only set sizes the classes.
// Consecutive local rows map to consecutive lanes (blockDim and the grid stride are multiples of
// 32), so a warp owns exactly one slice at a time.
template <class Canon>
__device__ __forceinline__ void cache_fill_rows(FillParams const& p, TermsView const& terms, Canon const& canon) {
  constexpr bool SYM = Canon::symmetric;
  BasisIndex const ix = p.ctx.index;
  RowDist const dist = p.ctx.dist;
  u64 const n_local = dist.n_local;
  bool const count_only = p.count_only != 0;
  u32 const nc = p.n_classes;
  u64 const row_hi = p.row_hi < n_local ? p.row_hi : n_local;
  for (u64 i = p.row_lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < row_hi; i += (u64)gridDim.x * blockDim.x) {
    u64 const row = dist_local_to_global(dist, i);
    u64 const r = ix.direct ? row : __ldg(ix.reps + row);
    u64 const slice = i >> 5;
    u64 base = 0;
    u32 start[kMaxClasses + 1] = {0u, 0u, 0u, 0u};  // first slot of each class; start[c >= n_classes] = width
    bool const stage = p.stage != 0;
    u32 width = 0, staged = 0;  // staging traversal: slots of the lane, elements written so far
    if (!count_only) {
      base = __ldg(p.slice_off + slice) + (i & 31);
      width = (u32)((__ldg(p.slice_off + slice + 1) - __ldg(p.slice_off + slice)) >> 5);
      if (!stage)
        for (u32 c = 1; c <= (u32)kMaxClasses; ++c) start[c] = c < nc ? __ldg(p.slice_start + kClassStride * slice + (c - 1)) : width;
    }
    // per class: cd = elements with the default coefficient (stored from the front of the class
    // region), cx = coded elements (stored from its back)
    u32 cd[kMaxClasses] = {0u, 0u, 0u}, cx[kMaxClasses] = {0u, 0u, 0u};
    for_each_transition(terms, r, [&](DevBond const& bd, u32 a, u32 b, u64 rp) {
      u64 rep = rp;
      int ph = 0;
      if (SYM) canon(rp, rep, ph);
      u64 idx = lookup_index(ix, rep);
      if (idx == ~(u64)0) return;
      u64 const pos = dist_global_to_pos(dist, idx);  // stored ready for the gather
      u32 const cls = dist_source_class(dist, pos, p.rounds, p.near);
      u32 hid = p.hid_map[bd.moff + a * (1u << bd.k) + b];
      u32 sid = SYM ? (u32)__ldg(p.sid_map + __ldg(ix.stab + idx)) : 0u;
      u32 const pid = SYM ? (u32)__ldg(p.pid_map + ph) : 0u;
      u32 const code = (hid * p.denom + pid) * p.n_sid + sid;
      bool const dflt = code == p.default_code;
      u32 const nd = cd[cls], nx = cx[cls];
      if (dflt) ++cd[cls]; else ++cx[cls];
      if (count_only) return;
      if (stage) {  // traversal order, all classes together; the code of every element is kept
        if (staged >= width) {
          *p.overflow = 1;
          return;
        }
        u64 const at = base + (u64)staged * 32;
        ++staged;
        p.idx[at] = (u32)pos;
        if (p.code_wide) static_cast<dev_u16*>(p.code)[at - p.code_slot0] = (dev_u16)code;
        else static_cast<dev_u8*>(p.code)[at - p.code_slot0] = (dev_u8)code;
        return;
      }
      u32 const lo = start[cls], hi = start[cls + 1];
      if (lo + nd + nx >= hi) {  // the two ends would meet
        *p.overflow = 1;
        return;
      }
      u64 const at = base + (u64)(dflt ? lo + nd : hi - 1u - nx) * 32;
      p.idx[at] = (u32)pos;
      if (!dflt) {
        if (p.code_wide) static_cast<dev_u16*>(p.code)[at - p.code_slot0] = (dev_u16)code;
        else static_cast<dev_u8*>(p.code)[at - p.code_slot0] = (dev_u8)code;
      }
    });
    for (u32 c = 0; c < nc; ++c) {
      if (count_only) {  // class sizes only: the width pass needs cd + cx per class
        p.len[(u64)(2 * c) * n_local + i] = (dev_u16)(cd[c] + cx[c]);
      } else {
        p.len[(u64)(2 * c) * n_local + i] = (dev_u16)cd[c];
        p.len[(u64)(2 * c + 1) * n_local + i] = (dev_u16)cx[c];
      }
    }
  }
}

#if defined(SPED_JIT)
// ---- run-time specialised entry points (NVRTC): SPED_T, SPED_NB and sped_jit_canonicalize come
// from the generated header "sped_jit_program.h" ----
struct JitCanon {
  static constexpr bool symmetric = true;
  __device__ __forceinline__ void operator()(u64 x, u64& rep, int& phase) const { sped_jit_canonicalize(x, rep, phase); }
};

// One entry point per module (SPED_JIT_KIND: 0 matvec, 1 cache fill): the generated canonicalisation
// is thousands of lines of straight-line code, and compiling it into a kernel that the run never
// launches would double the NVRTC time of the first (cold) solve.
#if SPED_JIT_KIND == 0
extern "C" __global__ void __launch_bounds__(256) sped_matvec_jit(MatvecParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<Traits<SPED_T>::cplx>(p.terms, smem);
  matvec_rows<SPED_T, SPED_NB>(p, terms, JitCanon());
}
#else
extern "C" __global__ void __launch_bounds__(256) sped_cache_fill_jit(FillParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<false>(p.terms, smem);
  cache_fill_rows(p, terms, JitCanon());
}
#endif
#endif

}  // namespace sped
