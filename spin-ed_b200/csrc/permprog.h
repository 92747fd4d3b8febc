// permprog.h -- the canonicalisation program: a traversal of the symmetry group compiled on the
// host into cheap word operations, interpreted by the sm_100a kernels (and by the host, for
// verification at compile time only).
//
// The reference's back end (liblattice_symmetries, bound at src/SpinED/Internal.hs:172-179,377)
// evaluates every group element from scratch with a full Benes network.  Here the group is walked
// as a path  g_0 = id -> g_1 -> ... -> g_{|G|-1}:  step k turns the previous image g_{k-1}.x into
// g_k.x by the permutation q_k = g_{k-1}^{-1} g_k, and the path is chosen so that most q_k are
// "few-displacement" permutations (a lattice translation is two masked rotates).  A step is
//   fast     y' = ((rotl(y, r1) & m) | (rotl(y, r2) & ~m)) & all          (one 8/16-byte record)
//   general  kind 0: y' = OR_j ( rotl(y, amount_j) & mask_j )
//            kind 1: Benes delta swaps  t = ((y >> d) ^ y) & m;  y ^= t ^ (t << d)
// whichever is cheapest for q_k.  Spin inversion costs one conditional XOR per image.
//
// Word layouts: number_spins <= 32 -> 32-bit words; <= 48 -> 64-bit words carried shifted left by
// kKeyShift bits, so that (image | path position) is a single 64-bit key and the running minimum
// over the orbit is one compare-select; > 48 -> plain 64-bit words with separate bookkeeping.
#pragma once
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define SPED_HD __host__ __device__ __forceinline__
#else
#define SPED_HD inline
#endif

namespace sped {

constexpr unsigned kKeyShift = 16;   // room for (2 * path position + flip) below a shifted image
constexpr unsigned kFastGeneral = 1u << 16;  // FastStep::ctl flag: use the general step record

struct PermStep {
  std::uint32_t first_op;
  std::uint16_t n_ops;
  std::uint16_t kind;  // 0 = rotate-mask, 1 = Benes delta swaps
};

template <class W>
struct PermOp {
  W mask;
  std::uint32_t amount;
};
template <>
struct PermOp<std::uint64_t> {
  std::uint64_t mask;
  std::uint32_t amount;
  std::uint32_t pad_;
};

template <class W>
struct FastStep {
  W mask;
  std::uint32_t ctl;  // bits 0-7: r1, bits 8-15: r2, bit 16: general step
};
template <>
struct FastStep<std::uint64_t> {
  std::uint64_t mask;
  std::uint32_t ctl;
  std::uint32_t pad_;
};

// Device-visible view (all pointers address the same memory space: shared once staged).
template <class W>
struct ProgramView {
  FastStep<W> const* fast;    // [n_steps]
  PermStep const* steps;      // [n_steps]; steps[0] is the identity (n_ops = 0)
  PermOp<W> const* ops;       // [n_ops]
  std::int32_t const* phase;  // [n_steps] character phase numerator of g_k (mod denom)
  std::uint32_t n_steps;
  std::uint32_t n_ops;
  std::uint32_t n_spins;
  std::uint32_t shift;        // words are carried shifted left by this many bits (0 or kKeyShift)
  std::int32_t inversion;     // 0, +1, -1
  std::int32_t denom;         // even
};

struct HostProgram {
  std::vector<PermStep> steps;
  std::vector<PermOp<std::uint64_t>> ops;  // kept 64-bit on the host; masks already shifted
  std::vector<FastStep<std::uint64_t>> fast;
  std::vector<std::int32_t> phase;
  std::vector<std::uint32_t> element;      // index into Group::elems of g_k
  std::uint32_t n_spins = 0;
  std::uint32_t word_bits = 64;            // 32 or 64
  std::uint32_t shift = 0;
  std::int32_t inversion = 0;
  std::int32_t denom = 2;
  std::uint32_t rot_ops = 0, benes_ops = 0, fast_steps = 0;
  bool empty() const { return steps.size() <= 1 && inversion == 0; }
};

template <class W>
SPED_HD W word_rotl(W y, unsigned r);

template <>
SPED_HD std::uint32_t word_rotl<std::uint32_t>(std::uint32_t y, unsigned r) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(y, y, r);
#else
  r &= 31u;
  return r ? (y << r) | (y >> (32u - r)) : y;
#endif
}

template <>
SPED_HD std::uint64_t word_rotl<std::uint64_t>(std::uint64_t y, unsigned r) {
#if defined(__CUDA_ARCH__)
  std::uint32_t lo = (std::uint32_t)y, hi = (std::uint32_t)(y >> 32);
  if (r & 32u) {
    std::uint32_t t = lo;
    lo = hi;
    hi = t;
  }
  std::uint32_t nh = __funnelshift_l(lo, hi, r);
  std::uint32_t nl = __funnelshift_l(hi, lo, r);
  return ((std::uint64_t)nh << 32) | nl;
#else
  r &= 63u;
  return r ? (y << r) | (y >> (64u - r)) : y;
#endif
}

template <class W>
SPED_HD W apply_step(W y, PermStep st, PermOp<W> const* ops) {
  PermOp<W> const* o = ops + st.first_op;
  if (st.kind == 0) {
    W acc = 0;
    for (unsigned j = 0; j < st.n_ops; ++j) acc |= word_rotl<W>(y, o[j].amount) & o[j].mask;
    return acc;
  }
  for (unsigned j = 0; j < st.n_ops; ++j) {
    unsigned d = o[j].amount;
    W t = ((y >> d) ^ y) & o[j].mask;
    y ^= t ^ (t << d);
  }
  return y;
}

// all-ones over the (shifted) spin positions
template <class W>
SPED_HD W full_mask(unsigned n_spins, unsigned shift) {
  W m = n_spins >= sizeof(W) * 8 ? ~(W)0 : (((W)1 << n_spins) - 1);
  return (W)(m << shift);
}

// step k of the program applied to the (shifted) image y
template <class W>
SPED_HD W advance(ProgramView<W> const& P, std::uint32_t k, W y, W all) {
  FastStep<W> const s = P.fast[k];
  if (s.ctl & kFastGeneral) return apply_step<W>(y, P.steps[k], P.ops);
  W a = word_rotl<W>(y, s.ctl & 63u);
  W b = word_rotl<W>(y, (s.ctl >> 8) & 63u);
  return ((a & s.mask) | (b & ~s.mask)) & all;
}

// min(y, flip(y)) and whether the flip was taken: flip(y) < y iff the top spin bit of y is set.
template <class W>
SPED_HD W fold_inversion(W y, unsigned top_bit, W all, unsigned& flipped) {
  flipped = (unsigned)(y >> top_bit) & 1u;
  return y ^ (all & (W)(0 - (W)flipped));
}

// Representative of x and the element reaching it (first minimiser along the path).
// x and rep are plain (unshifted) words.
template <class W>
SPED_HD void canonicalize(ProgramView<W> const& P, W x, W& rep, std::uint32_t& step, std::uint32_t& flipped) {
  W const all = full_mask<W>(P.n_spins, P.shift);
  unsigned const top = P.n_spins - 1 + P.shift;
  bool const inv = P.inversion != 0;
  unsigned f = 0;
  if (sizeof(W) == 4 || P.shift != 0) {
    // packed key: (image, 2 * position + flip) ordered lexicographically in one 64-bit integer
    unsigned const ks = sizeof(W) == 4 ? 32u : 0u;  // 32-bit images sit in the high word
    W y = (W)(x << P.shift);
    W z = inv ? fold_inversion<W>(y, top, all, f) : y;
    std::uint64_t best = ((std::uint64_t)z << ks) | f;
    for (std::uint32_t k = 1; k < P.n_steps; ++k) {
      y = advance<W>(P, k, y, all);
      z = inv ? fold_inversion<W>(y, top, all, f) : y;
      std::uint64_t key = ((std::uint64_t)z << ks) | (std::uint64_t)(2u * k + f);
      best = key < best ? key : best;
    }
    if (sizeof(W) == 4) {
      rep = (W)(best >> 32);
      step = (std::uint32_t)(best & 0xffffffffu) >> 1;
    } else {
      rep = (W)(best >> kKeyShift);
      step = (std::uint32_t)(best & ((1u << kKeyShift) - 1u)) >> 1;
    }
    flipped = (std::uint32_t)(best & 1u);
    return;
  }
  W y = x;
  W best = inv ? fold_inversion<W>(y, top, all, f) : y;
  std::uint32_t bstep = 0, bflip = f;
  for (std::uint32_t k = 1; k < P.n_steps; ++k) {
    y = advance<W>(P, k, y, all);
    W z = inv ? fold_inversion<W>(y, top, all, f) : y;
    if (z < best) {
      best = z;
      bstep = k;
      bflip = f;
    }
  }
  rep = best;
  step = bstep;
  flipped = bflip;
}

// Phase numerator (mod denom) of the element (step, flipped).
template <class W>
SPED_HD std::int32_t element_phase(ProgramView<W> const& P, std::uint32_t step, std::uint32_t flipped) {
  std::int32_t ph = P.phase[step];
  if (flipped && P.inversion < 0) {
    ph += P.denom / 2;
    if (ph >= P.denom) ph -= P.denom;
  }
  return ph;
}

// Stabiliser scan of the plain word x.  Returns -1 as soon as some image is smaller than x (x is
// not an orbit minimum; only when early_exit), 0 if a stabiliser element has a non-trivial
// character (norm 0), otherwise |Stab(x)| >= 1.
template <class W>
SPED_HD int stabilizer_scan(ProgramView<W> const& P, W x, bool early_exit) {
  W const all = full_mask<W>(P.n_spins, P.shift);
  W const xs = (W)(x << P.shift);
  W y = xs;
  int stab = 0;
  bool bad = false;
  for (std::uint32_t k = 0; k < P.n_steps; ++k) {
    if (k) y = advance<W>(P, k, y, all);
    if (early_exit && y < xs) return -1;
    if (y == xs) {
      ++stab;
      bad = bad || (P.phase[k] != 0);
    }
    if (P.inversion != 0) {
      W z = y ^ all;
      if (early_exit && z < xs) return -1;
      if (z == xs) {
        ++stab;
        bad = bad || (element_phase<W>(P, k, 1u) != 0);
      }
    }
  }
  return bad ? 0 : stab;
}

}  // namespace sped
