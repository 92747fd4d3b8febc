namespace sped {
extern char const k_src_device_types[] = R"SPEDRAW(
// device_types.h -- plain-old-data records shared by the host code, the statically compiled
// kernels and the run-time specialised (NVRTC) kernels.  Builtin types only: this header is also
// compiled by NVRTC, which has no standard library headers.
#pragma once

namespace sped {

#if defined(SPED_JIT)
typedef unsigned long u64;   // LP64, same as std::uint64_t on the host side
typedef long i64;
typedef unsigned int u32;
#endif
typedef unsigned short dev_u16;
typedef unsigned char dev_u8;
typedef int dev_i32;

// Lookup structures of a built basis (device pointers).
struct BasisIndex {
  u64 const* reps;        // sorted representatives, global, replicated on every rank
  dev_u16 const* stab;    // |Stab| per representative (norm^2 = stab / |G'|); null for the trivial group
  void const* bucket;     // prefix table: u32 or u64 entries
  u64 n_states;
  int bucket_shift;       // prefix = rep >> bucket_shift
  u32 bucket_count;       // number of prefixes (table has bucket_count + 1 entries)
  int bucket_wide;        // 1: u64 entries
  int direct;             // 1: index == state (no hamming weight, trivial group)
};

// One term instance ("bond"): a k-site tuple and where its matrix lives in the pools.
struct DevBond {
  u32 sites;       // site j in bits [8j, 8j+8)
  dev_u16 moff;    // offset (in matrix elements) of this bond's matrix in the pool
  dev_u16 zoff;    // offset of its row masks in the mask pool
  u32 k;           // number of sites (1..4)
  u32 pad_;
};

struct TermsView {
  DevBond const* bonds;
  double const* pool_re;   // real parts of all matrices, row-major dim x dim each
  double const* pool_im;   // imaginary parts (same layout)
  dev_u16 const* masks;    // per matrix row a: bitmask of b != a with M[a][b] != 0
  u32 n_bonds;
  u32 pool_size;
  u32 mask_size;
};

#if defined(__CUDACC__) || defined(SPED_JIT)
#define SPED_DIST_FN __host__ __device__ inline
#else
#define SPED_DIST_FN inline
#endif

// Row distribution over ranks: rows are dealt round-robin in blocks of B = 2^log2b consecutive
// rows (B is a multiple of 32), so every rank holds a statistically identical mix of rows -- the
// sorted representatives are *not* homogeneous (row length and gather locality drift with the
// index), and contiguous row blocks leave the ranks badly unbalanced.  A rank stores its rows
// compactly in local order; the replicated vector is laid out [rank][local index] with every
// rank's shard padded to `chunk` entries, which is exactly what an NCCL all-gather of the local
// shards produces.  With one rank local == global.
struct RowDist {
  u64 n;        // global number of rows
  u64 n_local;  // rows owned by this rank
  u64 chunk;    // padded shard length: max over ranks of rows owned
  u32 world, rank;
  u32 log2b;
  u32 pad_;
};

SPED_DIST_FN u64 dist_local_to_global(RowDist const& d, u64 i) {
  if (d.world == 1) return i;
  u64 blk = i >> d.log2b;
  return ((blk * d.world + d.rank) << d.log2b) + (i & (((u64)1 << d.log2b) - 1));
}
// position of global row g in the replicated [rank][local] layout
SPED_DIST_FN u64 dist_global_to_pos(RowDist const& d, u64 g) {
  if (d.world == 1) return g;
  u64 blk = g >> d.log2b;
  u64 owner = blk % d.world;
  return owner * d.chunk + ((blk / d.world) << d.log2b) + (g & (((u64)1 << d.log2b) - 1));
}
SPED_DIST_FN u64 dist_rows_of(u64 n, u32 world, u32 rank, u32 log2b) {
  u64 full = n >> log2b, rem = n & (((u64)1 << log2b) - 1);
  u64 nb = full / world + (rank < full % world ? 1 : 0);
  return (nb << log2b) + (rank == full % world ? rem : 0);
}

struct RowContext {
  BasisIndex index;
  double const* norm_table;  // norm_table[s] = sqrt(s / |G'|)
  double const* chi_table;   // (cos, sin)(2 pi k / denom)
  RowDist dist;              // which rows this rank owns
};

struct MatvecParams {
  RowContext ctx;
  TermsView terms;
  double const* diag_re;  // local rows
  double const* diag_im;  // null when the diagonal is real
  void const* x;          // replicated, column-major, stride xs
  void* y;                // local rows, column-major, stride ys
  u64 xs, ys;
  u32 ncols;              // columns handled by this launch (<= NB)
  unsigned long long* counter;  // count mode only
};

// Operator cache: the non-zero off-diagonal elements of the local rows, kept in HBM once the
// first matrix-free application has found them.  Rows are grouped in slices of 32 (one warp);
// inside a slice element j of lane l sits at slice_off[s] + 32 j + l, so a warp reads its
// column indices with one coalesced 128-byte load per j.
//
// With more than one rank the elements of a row are stored in classes by where their source entry
// of x lives: class 0 -- owned by this rank (available before the exchange of the Krylov vector
// has delivered anything); class 1 -- owned by one of the `near` next ranks (first exchange round);
// class 2 -- the other ranks (second round; absent when all peers fit in one round).  Inside a slice
// class c occupies slots [start_c, start_c+1) for every lane, so every class is read coalesced; the
// streaming kernel runs once per class, each pass overlapping the transfer the next one waits for.
constexpr int kMaxClasses = 4;
// Window class (optional, class 0 when present): a warp walks the 32 consecutive local rows of a
// slice, and a fifth (6x6) to two fifths (chains) of their elements gather from local rows within
// a few hundred rows of the slice itself.  The streaming kernel stages x[32 s - W, 32 s + 32 + W)
// (local indices) in shared memory once per slice and serves these elements from there: a handful
// of bank-conflict wavefronts instead of ~13 L1 lines per warp gather.  Their slots hold the offset
// into the window instead of a position.
constexpr u32 kWindow = 128;                     // W
constexpr u32 kWindowEntries = 32 + 2 * kWindow;
struct CacheView {
  u64 const* slice_off;  // [n_slices + 1], in elements
  u32 const* idx;        // position of the target in the replicated vector ([rank][local] layout);
                         // window class: offset into the slice's window
  void const* code;      // index into `table`: u8 when there are <= 256 codes, else u16
  dev_u16 const* len;    // [2 * n_classes][local rows]: per source class the elements that carry the
                         // default coefficient (no code is read for them), then the coded ones
  u32 const* slice_start;  // [n_slices][3] first slot of classes 1, 2, 3; null with one class
  double const* table;   // [n_codes][3]: (Re v, Im v, norm_s) with v = M[a][b] * chi(g')
  u64 n_slices;
  int code_wide;         // 1: u16 codes
  u32 n_codes;           // entries of `table`
  u32 n_classes;         // [window] local [remote near] [remote far]: 1 .. 4
  u32 near;              // first remote class = owners rank+1 .. rank+near (mod world)
  u32 default_code;      // the coefficient almost every element carries (first matrix value, chi = 1,
                         // trivial stabiliser).  Inside its class region [start_c, start_c+1) a row keeps
                         // these elements from the front, s = start_c + j, and the others -- with
                         // their code -- from the back, s = start_c+1 - 1 - j ("two-ended"), so one
                         // traversal fills both without knowing their numbers in advance.
  u32 window;            // 1: class 0 is the window class
  u32 rounds;            // exchange rounds = remote classes (0 with one rank)
  u32 pad_;
};

// first remote class / classes handled by pass `phase` (0: all; 1: everything this rank owns the
// sources of; 2, 3: first / second exchange round)
SPED_DIST_FN u32 cache_first_remote(u32 window) { return window ? 2u : 1u; }

// Class of the entry at position `pos` of the replicated vector for local row i of rank d.rank;
// *window_offset receives the offset into the slice's window for the window class.
SPED_DIST_FN u32 dist_source_class(RowDist const& d, u64 pos, u64 i, u32 window, u32 rounds, u32 near, u32* window_offset) {
  u32 const owner = d.world == 1 ? 0u : (u32)(pos / d.chunk);
  if (owner == d.rank) {
    if (window) {
      u64 const local = pos - (u64)d.rank * d.chunk;
      u64 const off = local + kWindow - (i & ~(u64)31);  // wraps to a huge value when below the window
      if (off < kWindowEntries) {
        *window_offset = (u32)off;
        return 0u;
      }
    }
    return window ? 1u : 0u;
  }
  u32 const dd = owner > d.rank ? owner - d.rank : owner + d.world - d.rank;
  return cache_first_remote(window) + ((rounds == 2 && dd > near) ? 1u : 0u);
}

struct FillParams {
  RowContext ctx;
  TermsView terms;
  u64 const* slice_off;
  u32* idx;
  void* code;              // u8 or u16 per slot, see code_wide
  dev_u16* len;            // [2 * n_classes][local rows] (see CacheView)
  u32 default_code;
  u32 pad1_;
  u32 const* slice_start;  // several classes, fill pass: [n_slices][3] (see CacheView); null otherwise
  int count_only;          // exact class sizes wanted: first pass, only `len` is written
  u32 n_classes;
  u32 near;
  u32 window;              // see CacheView
  u32 rounds;
  u32 pad0_;
  dev_u16 const* hid_map;  // [pool_size] matrix element -> distinct-value id
  dev_u16 const* sid_map;  // [|G'| + 1] stabiliser size -> id (null for the trivial group)
  dev_u16 const* pid_map;  // [denom] phase numerator -> id among the phases that occur (null: trivial group)
  u32 denom;               // number of distinct phases that occur (1 for the trivial group)
  u32 n_sid;               // number of distinct stabiliser sizes (1 for the trivial group)
  int code_wide;           // 1: u16 codes
  int* overflow;
};

}  // namespace sped
)SPEDRAW";
extern char const k_src_matvec_kernel[] = R"SPEDRAW(
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
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (u64)gridDim.x * blockDim.x) {
    u64 const row = dist_local_to_global(dist, i);
    u64 const r = ix.direct ? row : __ldg(ix.reps + row);
    u64 const slice = i >> 5;
    u64 base = 0;
    u32 start[kMaxClasses + 1] = {0u, 0u, 0u, 0u, 0u};  // first slot of each class; start[c >= n_classes] = width
    if (!count_only) {
      base = __ldg(p.slice_off + slice) + (i & 31);
      u32 const width = (u32)((__ldg(p.slice_off + slice + 1) - __ldg(p.slice_off + slice)) >> 5);
      for (u32 c = 1; c <= (u32)kMaxClasses; ++c) start[c] = c < nc ? __ldg(p.slice_start + 3 * slice + (c - 1)) : width;
    }
    // per class: cd = elements with the default coefficient (stored from the front of the class
    // region), cx = coded elements (stored from its back)
    u32 cd[kMaxClasses] = {0u, 0u, 0u, 0u}, cx[kMaxClasses] = {0u, 0u, 0u, 0u};
    for_each_transition(terms, r, [&](DevBond const& bd, u32 a, u32 b, u64 rp) {
      u64 rep = rp;
      int ph = 0;
      if (SYM) canon(rp, rep, ph);
      u64 idx = lookup_index(ix, rep);
      if (idx == ~(u64)0) return;
      u64 const pos = dist_global_to_pos(dist, idx);  // stored ready for the gather
      u32 woff = 0;
      u32 cls = dist_source_class(dist, pos, i, p.window, p.rounds, p.near, &woff);
      // a full window region (its width is a guess when the classes are not counted first) sends
      // the element to the plain local class, where every local source can live
      if (!count_only && cls == 0 && p.window && cd[0] + cx[0] >= start[1] - start[0]) cls = 1;
      u32 hid = p.hid_map[bd.moff + a * (1u << bd.k) + b];
      u32 sid = SYM ? (u32)__ldg(p.sid_map + __ldg(ix.stab + idx)) : 0u;
      u32 const pid = SYM ? (u32)__ldg(p.pid_map + ph) : 0u;
      u32 const code = (hid * p.denom + pid) * p.n_sid + sid;
      bool const dflt = code == p.default_code;
      u32 const nd = cd[cls], nx = cx[cls];
      if (dflt) ++cd[cls]; else ++cx[cls];
      if (count_only) return;
      u32 const lo = start[cls], hi = start[cls + 1];
      if (lo + nd + nx >= hi) {  // the two ends would meet
        *p.overflow = 1;
        return;
      }
      u64 const at = base + (u64)(dflt ? lo + nd : hi - 1u - nx) * 32;
      p.idx[at] = (cls == 0 && p.window) ? woff : (u32)pos;
      if (!dflt) {
        if (p.code_wide) static_cast<dev_u16*>(p.code)[at] = (dev_u16)code;
        else static_cast<dev_u8*>(p.code)[at] = (dev_u8)code;
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

extern "C" __global__ void __launch_bounds__(256) sped_matvec_jit(MatvecParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<Traits<SPED_T>::cplx>(p.terms, smem);
  matvec_rows<SPED_T, SPED_NB>(p, terms, JitCanon());
}

extern "C" __global__ void __launch_bounds__(256) sped_cache_fill_jit(FillParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<false>(p.terms, smem);
  cache_fill_rows(p, terms, JitCanon());
}
#endif

}  // namespace sped
)SPEDRAW";
}
