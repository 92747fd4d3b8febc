// permprog.h -- the canonicalisation program: a traversal of the symmetry group compiled on the
// host into cheap word operations, interpreted by the sm_100a kernels (and by the host, for
// verification at compile time only).
//
// The reference's back end (liblattice_symmetries, bound at src/SpinED/Internal.hs:172-179,377)
// evaluates every group element from scratch with a full Benes network.  Here the group is walked
// as a path  g_0 = id -> g_1 -> ... -> g_{|G|-1}:  step k turns the previous image g_{k-1}.x into
// g_k.x by the permutation q_k = g_{k-1}^{-1} g_k, and the path is chosen so that most q_k are
// "few-displacement" permutations (translations are two masked rotates).  Each step is either
//   kind 0  rotate-mask:  y' = OR_j ( rotl(y, amount_j) & mask_j )
//   kind 1  Benes:        for each delta swap:  t = ((y >> d) ^ y) & m;  y ^= t ^ (t << d)
// whichever is cheaper for q_k.  Spin inversion costs one conditional XOR per image.
#pragma once
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define SPED_HD __host__ __device__ __forceinline__
#else
#define SPED_HD inline
#endif

namespace sped {

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

// Device-visible view (pointers may address global or shared memory).
template <class W>
struct ProgramView {
  PermStep const* steps;      // [n_steps]; steps[0] is the identity (n_ops = 0)
  PermOp<W> const* ops;       // [n_ops]
  std::int32_t const* phase;  // [n_steps] character phase numerator of g_k (mod denom)
  std::uint32_t n_steps;
  std::uint32_t n_ops;
  std::uint32_t n_spins;
  std::int32_t inversion;     // 0, +1, -1
  std::int32_t denom;         // even
};

struct HostProgram {
  std::vector<PermStep> steps;
  std::vector<PermOp<std::uint64_t>> ops;  // always kept 64-bit on the host
  std::vector<std::int32_t> phase;
  std::vector<std::uint32_t> element;      // index into Group::elems of g_k
  std::uint32_t n_spins = 0;
  std::int32_t inversion = 0;
  std::int32_t denom = 2;
  std::uint32_t rot_ops = 0, benes_ops = 0;
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

template <class W>
SPED_HD W full_mask(unsigned n_spins) {
  return n_spins >= sizeof(W) * 8 ? ~(W)0 : (((W)1 << n_spins) - 1);
}

// min(y, flip(y)) and whether the flip was taken: flip(y) < y iff bit n-1 of y is set.
template <class W>
SPED_HD W fold_inversion(W y, unsigned n_spins, W all, unsigned& flipped) {
  flipped = (unsigned)(y >> (n_spins - 1)) & 1u;
  return y ^ (all & (W)(0 - (W)flipped));
}

struct CanonResult64 {
  std::uint64_t rep;
  std::uint32_t step;     // path position of the first minimising element
  std::uint32_t flipped;  // 1 if the spin flip was applied on top of it
};

// Representative of x and the element reaching it (first minimiser along the path).
template <class W>
SPED_HD void canonicalize(ProgramView<W> const& P, W x, W& rep, std::uint32_t& step, std::uint32_t& flipped) {
  W const all = full_mask<W>(P.n_spins);
  W y = x;
  W best = x;
  std::uint32_t bstep = 0, bflip = 0;
  if (P.inversion != 0) {
    unsigned f;
    best = fold_inversion<W>(y, P.n_spins, all, f);
    bflip = f;
    for (std::uint32_t k = 1; k < P.n_steps; ++k) {
      y = apply_step<W>(y, P.steps[k], P.ops);
      W z = fold_inversion<W>(y, P.n_spins, all, f);
      if (z < best) {
        best = z;
        bstep = k;
        bflip = f;
      }
    }
  } else {
    for (std::uint32_t k = 1; k < P.n_steps; ++k) {
      y = apply_step<W>(y, P.steps[k], P.ops);
      if (y < best) {
        best = y;
        bstep = k;
      }
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

// Stabiliser scan of x.  Returns -1 as soon as some image is smaller than x (x is not an orbit
// minimum; only when early_exit), 0 if a stabiliser element has a non-trivial character (norm 0),
// otherwise |Stab(x)| >= 1.
template <class W>
SPED_HD int stabilizer_scan(ProgramView<W> const& P, W x, bool early_exit) {
  W const all = full_mask<W>(P.n_spins);
  W y = x;
  int stab = 0;
  bool bad = false;
  for (std::uint32_t k = 0; k < P.n_steps; ++k) {
    if (k) y = apply_step<W>(y, P.steps[k], P.ops);
    if (early_exit && y < x) return -1;
    if (y == x) {
      ++stab;
      bad = bad || (P.phase[k] != 0);
    }
    if (P.inversion != 0) {
      W z = y ^ all;
      if (early_exit && z < x) return -1;
      if (z == x) {
        ++stab;
        bad = bad || (element_phase<W>(P, k, 1u) != 0);
      }
    }
  }
  return bad ? 0 : stab;
}

}  // namespace sped
