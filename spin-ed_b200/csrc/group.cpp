// group.cpp -- symmetries, group closure, Burnside sector dimension and the compiler that turns
// a group into a canonicalisation program (see permprog.h).  Host only.
//
// Replaces the setup half of liblattice_symmetries that the reference reaches through
// ls_create_symmetry / ls_create_group (src/SpinED/Internal.hs:69-70,120-121).
#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <random>

#include "internal.h"

namespace sped {

// ------------------------------------------------------------------------------------------
// symmetry + group closure
// ------------------------------------------------------------------------------------------
static unsigned permutation_period(std::vector<unsigned> const& p) {
  // lcm of cycle lengths
  std::vector<char> seen(p.size(), 0);
  u64 period = 1;
  for (size_t i = 0; i < p.size(); ++i) {
    if (seen[i]) continue;
    u64 len = 0;
    for (size_t j = i; !seen[j]; j = p[j]) {
      seen[j] = 1;
      ++len;
    }
    period = std::lcm(period, len);
  }
  return (unsigned)period;
}

std::shared_ptr<Symmetry> make_symmetry(unsigned n, unsigned const* perm, unsigned sector) {
  if (n == 0) fail(LS_INVALID_PERMUTATION, "empty permutation");
  if (n > 64) fail(LS_PERMUTATION_TOO_LONG, "permutations longer than 64 sites are not supported");
  std::vector<char> seen(n, 0);
  for (unsigned i = 0; i < n; ++i) {
    if (perm[i] >= n || seen[perm[i]]) fail(LS_INVALID_PERMUTATION, "permutation is not a bijection of 0..n-1");
    seen[perm[i]] = 1;
  }
  auto s = std::make_shared<Symmetry>();
  s->perm.assign(perm, perm + n);
  s->periodicity = permutation_period(s->perm);
  if (sector >= s->periodicity) fail(LS_INVALID_SECTOR, "sector must be smaller than the periodicity");
  s->sector = sector;
  return s;
}

bool Group::real_characters() const {
  for (auto const& e : elems)
    if (e.phase != 0 && 2 * e.phase != denom) return false;
  return true;
}

std::shared_ptr<Group> make_group(std::vector<Symmetry const*> const& gens) {
  auto g = std::make_shared<Group>();
  if (gens.empty()) {
    g->n = 0;
    g->elems.push_back(GroupElement{});
    return g;
  }
  unsigned n = (unsigned)gens[0]->perm.size();
  i64 denom = 2;
  for (auto* s : gens) {
    if (s->perm.size() != n) fail(LS_INCOMPATIBLE_SYMMETRIES, "generators act on different numbers of sites");
    denom = std::lcm(denom, (i64)s->periodicity);
  }
  g->n = n;
  g->denom = denom;
  std::vector<i64> gphase;
  for (auto* s : gens) gphase.push_back((i64)s->sector * (denom / s->periodicity) % denom);
  std::vector<int> ident(n);
  std::iota(ident.begin(), ident.end(), 0);
  std::map<std::vector<int>, size_t> where;
  where[ident] = 0;
  g->elems.push_back(GroupElement{ident, 0});
  for (size_t head = 0; head < g->elems.size(); ++head) {
    for (size_t k = 0; k < gens.size(); ++k) {
      std::vector<int> q(n);
      for (unsigned i = 0; i < n; ++i) q[i] = g->elems[head].perm[gens[k]->perm[i]];
      i64 ph = (g->elems[head].phase + gphase[k]) % denom;
      auto it = where.find(q);
      if (it == where.end()) {
        where[q] = g->elems.size();
        g->elems.push_back(GroupElement{q, ph});
        if (g->elems.size() > 65535) fail(LS_INVALID_ARGUMENT, "symmetry group larger than 65535 elements");
      } else if (g->elems[it->second].phase != ph) {
        fail(LS_INCOMPATIBLE_SYMMETRIES, "the same permutation is reached with two different phases");
      }
    }
  }
  return g;
}

// ------------------------------------------------------------------------------------------
// Burnside counting
// ------------------------------------------------------------------------------------------
static std::vector<unsigned> cycle_lengths(std::vector<int> const& p, unsigned n_spins) {
  std::vector<char> seen(n_spins, 0);
  std::vector<unsigned> out;
  for (unsigned i = 0; i < n_spins; ++i) {
    if (seen[i]) continue;
    unsigned len = 0;
    for (unsigned j = i; !seen[j]; j = (unsigned)p[j]) {
      seen[j] = 1;
      ++len;
    }
    out.push_back(len);
  }
  return out;
}

u64 burnside_dimension(Group const& g, unsigned n_spins, int hw, int inv) {
  std::vector<int> ident(n_spins);
  std::iota(ident.begin(), ident.end(), 0);
  long double total = 0;
  for (auto const& e : g.elems) {
    auto const& perm = e.perm.empty() ? ident : e.perm;
    auto cyc = cycle_lengths(perm, n_spins);
    // states fixed by g: constant on cycles; count those of the right hamming weight
    long double fixed;
    if (hw < 0) {
      fixed = std::ldexp(1.0L, (int)cyc.size());
    } else {
      std::vector<long double> poly(n_spins + 1, 0.0L);
      poly[0] = 1;
      for (unsigned len : cyc)
        for (int d = (int)n_spins; d >= (int)len; --d) poly[d] += poly[d - len];
      fixed = poly[hw];
    }
    long double ang = 2.0L * M_PIl * (long double)e.phase / (long double)g.denom;
    total += fixed * std::cos(ang);
    if (inv != 0) {
      // states with g.x = flip(x): alternate along every cycle -> all cycles even, 2 per cycle
      bool all_even = true;
      for (unsigned len : cyc) all_even = all_even && (len % 2 == 0);
      long double fx = 0;
      if (all_even && (hw < 0 || 2 * hw == (int)n_spins)) fx = std::ldexp(1.0L, (int)cyc.size());
      total += fx * std::cos(ang) * (long double)inv;
    }
  }
  long double dim = total / (long double)(g.elems.size() * (inv != 0 ? 2 : 1));
  if (dim < 0) dim = 0;
  return (u64)std::llroundl(dim);
}

// ------------------------------------------------------------------------------------------
// program compiler
// ------------------------------------------------------------------------------------------
namespace {

// Benes routing of `src` (out[i] = in[src[i]]) over a W-wide butterfly, level by level.
// Returns delta swaps (mask, delta) in application order with empty stages dropped.
std::vector<std::pair<u64, unsigned>> benes_stages(std::vector<int> src, unsigned W) {
  unsigned levels = 0;
  while ((1u << levels) < W) ++levels;
  std::vector<u64> in_mask(levels, 0), out_mask(levels, 0);
  for (unsigned level = 0; level + 1 < levels; ++level) {
    unsigned width = W >> level, d = width / 2;
    std::vector<int> next(W);
    for (unsigned lo = 0; lo < W; lo += width) {
      auto partner = [&](int pos) { return (unsigned)pos < lo + d ? pos + (int)d : pos - (int)d; };
      std::vector<int> feeds(W, -1);  // feeds[input position] = output position
      for (unsigned j = lo; j < lo + width; ++j) feeds[src[j]] = (int)j;
      std::vector<int> net(W, -1);  // 0: lower half-network, 1: upper
      for (unsigned s = lo; s < lo + width; ++s) {
        if (net[s] != -1) continue;
        int e = (int)s, c = 0;
        while (net[e] == -1) {
          net[e] = c;
          int ip = partner(e);
          net[ip] = 1 - c;
          e = src[partner(feeds[ip])];
          // e shares an output pair with ip, so it must use the half-network ip does not: c stays
        }
      }
      std::vector<int> moved(W, -1);  // position of each input element after the input column
      for (unsigned i = lo; i < lo + d; ++i) {
        bool swap = net[i] == 1;
        if (swap) in_mask[level] |= 1ull << i;
        moved[i] = swap ? (int)(i + d) : (int)i;
        moved[i + d] = swap ? (int)i : (int)(i + d);
      }
      for (unsigned j = lo; j < lo + d; ++j) {
        int e_lo = src[j], e_hi = src[j + d];
        bool swap = net[e_lo] == 1;
        if (swap) out_mask[level] |= 1ull << j;
        next[j] = moved[swap ? e_hi : e_lo];
        next[j + d] = moved[swap ? e_lo : e_hi];
      }
    }
    src = next;
  }
  for (unsigned lo = 0; lo < W; lo += 2)
    if (src[lo] == (int)lo + 1) in_mask[levels - 1] |= 1ull << lo;
  std::vector<std::pair<u64, unsigned>> out;
  for (unsigned l = 0; l < levels; ++l)
    if (in_mask[l]) out.push_back({in_mask[l], (W >> l) / 2});
  for (int l = (int)levels - 2; l >= 0; --l)
    if (out_mask[l]) out.push_back({out_mask[l], (W >> l) / 2});
  return out;
}

struct StepCode {
  unsigned kind = 0;
  std::vector<PermOp<u64>> ops;
  unsigned cost = 0;
  bool fast = false;
  FastStep<u64> fs{};
};

// Cheapest encoding of q (spin i <- spin q[i]) on a W-bit word whose spins sit at bit positions
// [shift, shift + n_spins).
StepCode encode_step(std::vector<int> const& q, unsigned n_spins, unsigned W, unsigned shift) {
  std::map<unsigned, u64> classes;
  for (unsigned i = 0; i < n_spins; ++i) classes[(i + W - (unsigned)q[i]) % W] |= 1ull << (i + shift);
  unsigned per_rot = W == 32 ? 2 : 4, per_swap = W == 32 ? 6 : 11, fast_cost = W == 32 ? 5 : 9;
  std::vector<int> src(W);
  for (unsigned i = 0; i < W; ++i) src[i] = (int)i;
  for (unsigned i = 0; i < n_spins; ++i) src[i + shift] = q[i] + (int)shift;
  auto stages = benes_stages(src, W);
  StepCode rot;
  rot.kind = 0;
  rot.cost = (unsigned)classes.size() * per_rot + 6;  // + general-step interpretation overhead
  for (auto const& c : classes) rot.ops.push_back(PermOp<u64>{c.second, c.first, 0});
  if (classes.size() <= 2) {
    rot.fast = true;
    rot.cost = fast_cost;
    auto it = classes.begin();
    u64 all = (n_spins == 64 ? ~0ull : ((1ull << n_spins) - 1)) << shift;
    unsigned r1 = it->first, r2 = it->first;
    u64 m = all;
    if (classes.size() == 2) {
      m = it->second;
      ++it;
      r2 = it->first;
    }
    rot.fs.mask = m;
    rot.fs.ctl = r1 | (r2 << 8);
    return rot;
  }
  if (stages.size() * per_swap + 6 < rot.cost) {
    StepCode b;
    b.kind = 1;
    b.cost = (unsigned)stages.size() * per_swap + 6;
    for (auto const& s : stages) b.ops.push_back(PermOp<u64>{s.first, s.second, 0});
    return b;
  }
  return rot;
}

u64 permute_naive(std::vector<int> const& p, u64 x) {
  u64 y = 0;
  for (size_t i = 0; i < p.size(); ++i) y |= ((x >> p[i]) & 1ull) << i;
  return y;
}

}  // namespace

HostProgram compile_program(Group const& g, unsigned n_spins, int inv) {
  HostProgram P;
  P.n_spins = n_spins;
  P.inversion = inv;
  P.denom = (std::int32_t)g.denom;
  size_t const m = g.elems.size();
  unsigned const W = n_spins <= 32 ? 32 : 64;
  P.word_bits = W;
  P.shift = (W == 64 && n_spins + kKeyShift <= 64 && 2 * m < (1u << kKeyShift)) ? kKeyShift : 0;
  P.steps.push_back(PermStep{0, 0, 0});
  P.fast.push_back(FastStep<u64>{0, kFastGeneral, 0});
  P.phase.push_back(0);
  P.element.push_back(0);
  if (m <= 1) return P;
  if (g.n != n_spins) fail(LS_INVALID_ARGUMENT, "symmetry permutations must act on exactly number_spins sites");

  std::map<std::vector<int>, unsigned> where;
  for (size_t k = 0; k < m; ++k) where[g.elems[k].perm] = (unsigned)k;
  std::vector<std::vector<int>> inverse(m, std::vector<int>(n_spins));
  for (size_t k = 0; k < m; ++k)
    for (unsigned i = 0; i < n_spins; ++i) inverse[k][g.elems[k].perm[i]] = (int)i;
  // cost of realising each group element as a single step
  std::vector<StepCode> code(m);
  for (size_t k = 1; k < m; ++k) code[k] = encode_step(g.elems[k].perm, n_spins, W, P.shift);

  // greedy nearest-neighbour path through the group: from image g.x the image h.x is reached by
  // q = g^{-1} h  (q[i] = g^{-1}[h[i]]), itself a group element.
  std::vector<char> visited(m, 0);
  visited[0] = 1;
  unsigned cur = 0;
  std::vector<int> q(n_spins);
  for (size_t count = 1; count < m; ++count) {
    unsigned best_h = 0, best_q = 0, best_cost = ~0u;
    for (unsigned h = 1; h < m; ++h) {
      if (visited[h]) continue;
      for (unsigned i = 0; i < n_spins; ++i) q[i] = inverse[cur][g.elems[h].perm[i]];
      auto it = where.find(q);
      if (it == where.end()) fail(SPED_INTERNAL_ERROR, "group is not closed");
      unsigned c = code[it->second].cost;
      if (c < best_cost) {
        best_cost = c;
        best_h = h;
        best_q = it->second;
      }
    }
    StepCode const& sc = code[best_q];
    if (sc.fast) {
      P.steps.push_back(PermStep{(std::uint32_t)P.ops.size(), 0, 0});
      P.fast.push_back(sc.fs);
      ++P.fast_steps;
    } else {
      P.steps.push_back(PermStep{(std::uint32_t)P.ops.size(), (std::uint16_t)sc.ops.size(), (std::uint16_t)sc.kind});
      for (auto const& o : sc.ops) P.ops.push_back(o);
      (sc.kind == 0 ? P.rot_ops : P.benes_ops) += (std::uint32_t)sc.ops.size();
      P.fast.push_back(FastStep<u64>{0, kFastGeneral, 0});
    }
    P.phase.push_back((std::int32_t)g.elems[best_h].phase);
    P.element.push_back(best_h);
    visited[best_h] = 1;
    cur = best_h;
  }

  // verify the program against bit-by-bit application of every group element
  std::vector<PermOp<std::uint32_t>> ops32;
  std::vector<FastStep<std::uint32_t>> fast32;
  for (auto const& o : P.ops) ops32.push_back(PermOp<std::uint32_t>{(std::uint32_t)o.mask, o.amount});
  for (auto const& f : P.fast) fast32.push_back(FastStep<std::uint32_t>{(std::uint32_t)f.mask, f.ctl});
  ProgramView<u64> v64{P.fast.data(), P.steps.data(), P.ops.data(), P.phase.data(), (u32)P.steps.size(),
                       (u32)P.ops.size(), n_spins, P.shift, inv, P.denom};
  ProgramView<std::uint32_t> v32{fast32.data(), P.steps.data(), ops32.data(), P.phase.data(), (u32)P.steps.size(),
                                 (u32)P.ops.size(), n_spins, 0, inv, P.denom};
  std::mt19937_64 rng(0x5EED5EEDull);
  u64 const all = n_spins == 64 ? ~0ull : ((1ull << n_spins) - 1);
  for (int trial = 0; trial < 16; ++trial) {
    u64 x = rng() & all;
    u64 y64 = x << P.shift;
    std::uint32_t y32 = (std::uint32_t)x;
    for (size_t k = 1; k < P.steps.size(); ++k) {
      u64 got;
      if (W == 64) {
        y64 = advance<u64>(v64, (u32)k, y64, full_mask<u64>(n_spins, P.shift));
        got = y64 >> P.shift;
      } else {
        y32 = advance<std::uint32_t>(v32, (u32)k, y32, full_mask<std::uint32_t>(n_spins, 0));
        got = y32;
      }
      if (got != permute_naive(g.elems[P.element[k]].perm, x))
        fail(SPED_INTERNAL_ERROR, "canonicalisation program failed self-verification");
    }
  }
  return P;
}

}  // namespace sped
