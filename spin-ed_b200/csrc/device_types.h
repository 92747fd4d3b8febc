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
constexpr int kMaxClasses = 3;
constexpr int kClassStride = kMaxClasses - 1;  // slice_start entries per slice
struct CacheView {
  u64 const* slice_off;  // [n_slices + 1], in elements
  u32 const* idx;        // position of the target in the replicated vector ([rank][local] layout)
  void const* code;      // index into `table`: u8 when there are <= 256 codes, else u16; COMPACT: only the
                         // coded elements have an entry, see code_off
  dev_u16 const* len;    // [2 * n_classes][local rows]: per source class the elements that carry the
                         // default coefficient (no code is read for them), then the coded ones
  u32 const* slice_start;  // [n_slices][kClassStride] first slot of classes 1, 2; null with one class
  double const* table;   // [n_codes][3]: (Re v, Im v, norm_s) with v = M[a][b] * chi(g')
  u64 n_slices;
  int code_wide;         // 1: u16 codes
  u32 n_codes;           // entries of `table`
  u32 n_classes;         // local [remote near] [remote far]: 1 .. 3
  u32 near;              // first remote class = owners rank+1 .. rank+near (mod world)
  u32 default_code;      // the coefficient almost every element carries (first matrix value, chi = 1,
                         // trivial stabiliser).  Inside its class region [start_c, start_c+1) a row keeps
                         // these elements from the front, s = start_c + j, and the others -- with
                         // their code -- from the back, s = start_c+1 - 1 - j ("two-ended"), so one
                         // traversal fills both without knowing their numbers in advance.
  u32 rounds;            // exchange rounds = remote classes (0 with one rank)
  u64 const* code_off;   // [n_slices * n_classes + 1]: the codes of class c of slice s start at
                         // code_off[s * n_classes + c]; coded element j of lane l at + 32 j + l.  (Almost every
                         // element carries the default coefficient, so a code per SLOT -- one byte in five of
                         // the cache -- would be memory that is never read.)
};

// Class of the entry at position `pos` of the replicated vector for the rows of rank d.rank:
// 0 -- this rank owns it; 1 -- one of the `near` next ranks does (first exchange round); 2 -- the rest.
SPED_DIST_FN u32 dist_source_class(RowDist const& d, u64 pos, u32 rounds, u32 near) {
  u32 const owner = d.world == 1 ? 0u : (u32)(pos / d.chunk);
  if (owner == d.rank) return 0u;
  u32 const dd = owner > d.rank ? owner - d.rank : owner + d.world - d.rank;
  return 1u + ((rounds == 2 && dd > near) ? 1u : 0u);
}

struct FillParams {
  RowContext ctx;
  TermsView terms;
  u64 const* slice_off;
  u32* idx;
  void* code;              // u8 or u16 per slot (see code_wide) of the slots from code_slot0 on: a temporary
                           // that covers the rows of this launch only; the codes of the coded elements are
                           // compacted afterwards (CacheView::code_off)
  dev_u16* len;            // [2 * n_classes][local rows] (see CacheView)
  u32 default_code;
  u32 pad1_;
  u64 row_lo, row_hi;      // local rows of this launch (multiples of 32, or the end): the fill runs in row
                           // chunks so that the code temporary stays small
  u64 code_slot0;          // first slot `code` covers = slice_off[row_lo / 32]
  int stage;               // 1: staging traversal of a several-class build -- every element, whatever its
                           // class, goes to the next slot of its lane (slice_off = staging offsets from the
                           // cheap width bound) with its position in idx and its code in code (for ALL elements);
                           // len gets the per-class counts.  cache_place_kernel then lays the elements out by
                           // class without a second canonicalisation (cached_kernel.cuh).
  int pad2_;
  u32 const* slice_start;  // several classes, fill pass: [n_slices][kClassStride] (see CacheView); null otherwise
  int count_only;          // exact class sizes wanted: first pass, only `len` is written
  u32 n_classes;
  u32 near;
  u32 rounds;
  dev_u16 const* hid_map;  // [pool_size] matrix element -> distinct-value id
  dev_u16 const* sid_map;  // [|G'| + 1] stabiliser size -> id (null for the trivial group)
  dev_u16 const* pid_map;  // [denom] phase numerator -> id among the phases that occur (null: trivial group)
  u32 denom;               // number of distinct phases that occur (1 for the trivial group)
  u32 n_sid;               // number of distinct stabiliser sizes (1 for the trivial group)
  int code_wide;           // 1: u16 codes
  int* overflow;
};

}  // namespace sped
