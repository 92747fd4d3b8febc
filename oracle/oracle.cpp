// oracle.cpp -- CPU restatement of SpinED's hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (libsped.so) never links, loads or calls it.
//
// PARITY UNPINNED: the reference tree (/root/reference) holds no numerics.  Everything this file
// computes is done upstream inside two un-vendored, un-pinned dependencies that are absent here:
//   * twesterhout/lattice-symmetries (cloned from master by /root/reference/configure:26-30)
//   * PRIMME via twesterhout/primme-hs (/root/reference/cabal.project:13-15)
// so the restatement follows the *published* algorithm of lattice-symmetries (symmetry-adapted
// basis of orbit representatives, Benes-network bit permutations, sorted representatives with a
// prefix bucket + binary search, row-parallel matrix-free operator application) and is anchored on
// the reference's own call sites:
//   ls_create_symmetry / ls_create_group      src/SpinED/Internal.hs:69-70,120-121
//   ls_create_spin_basis, ls_build            src/SpinED/Internal.hs:172-179
//   ls_get_states                             src/SpinED/Internal.hs:187-196
//   ls_create_interaction{1..4}               src/SpinED/Internal.hs:260-270
//   ls_operator_matmat / ls_operator_expectation  src/SpinED/Internal.hs:377-381
// and on the field semantics of /root/reference/template.yaml (hamming weight = number of up
// spins :5-8, sector k <-> eigenvalue exp(2 pi i k / N) :30-34, spin_inversion = +-1 :37-43,
// matrices are 2^k x 2^k :52-70).  It is pinned against (a) the README known answer (4-ring: 16
// states, E0 = -8, README.md:56-95), (b) the structural facts of test/Spec.hs:38-43,72-84 and
// (c) an independent numpy dense-projector construction (oracle/dense_truth.py), see
// tests/test_oracle.py.
//
// Conventions (stated once, used by the oracle and the product alike):
//   * bit i of a basis word = spin at site i, 1 = up.
//   * a symmetry with permutation p acts as  (g.x)[i] = x[p[i]].
//   * character chi(g) = exp(+2 pi i phase(g)), phase = sector / periodicity; spin inversion
//     contributes the factor `spin_inversion` (+-1) when the flip is used.
//   * representative = minimum (as unsigned integer) over the orbit under G x {1, flip}.
//   * norm(x)^2 = (1/|G'|) sum_{g in Stab(x)} conj(chi(g)).
//   * local matrix index of a k-site tuple (s0..s_{k-1}): a = sum_j bit(s_j) << (k-1-j)
//     (first listed site is the most significant bit, Kronecker order).
//   * pull form: y[r] = sum_t sum_b M_t[a_r][b] * chi(g') * (n_s / n_r) * x[index(s)],
//     where r' = r with the tuple's bits replaced by b and g'.r' = s is r's representative.
//   * fixed summation order: terms in the order given, tuples in the order given, b ascending;
//     the diagonal (b == a) is accumulated in the same sweep.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using u64 = std::uint64_t;
using i64 = std::int64_t;
using cplx = std::complex<double>;

enum {
  ORC_OK = 0,
  ORC_INVALID_ARGUMENT = 2,
  ORC_INVALID_HAMMING_WEIGHT = 3,
  ORC_INVALID_SPIN_INVERSION = 4,
  ORC_INVALID_NUMBER_SPINS = 5,
  ORC_INVALID_PERMUTATION = 6,
  ORC_INVALID_SECTOR = 7,
  ORC_INVALID_DATATYPE = 9,
  ORC_INCOMPATIBLE_SYMMETRIES = 11,
  ORC_NOT_BUILT = 14,
  ORC_DIMENSION_MISMATCH = 19,
};

// ---------------------------------------------------------------------------------------------
// Bit permutations: naive (ground truth) and Benes network (fast path, as lattice-symmetries does)
// ---------------------------------------------------------------------------------------------
u64 apply_naive(const std::vector<int>& p, u64 x) {
  u64 y = 0;
  for (size_t i = 0; i < p.size(); ++i) y |= ((x >> p[i]) & 1ull) << i;
  return y;
}

struct Benes {
  // stages applied in order; each is a delta swap: t = ((x >> d) ^ x) & m; x ^= t ^ (t << d)
  std::vector<std::pair<u64, int>> stages;
  u64 operator()(u64 x) const {
    for (auto const& s : stages) {
      u64 t = ((x >> s.second) ^ x) & s.first;
      x ^= t ^ (t << s.second);
    }
    return x;
  }
};

// Route src (out[i] = in[src[i]], a permutation of 0..63) through a 64-wide Benes network.
// Recursive halving: at width 2d the input and output columns of delta-d swaps are chosen by
// 2-colouring the constraint cycles (partners at the input and at the output must take different
// sub-networks).
void benes_route(std::vector<int> src, int lo, int width, std::vector<u64>& in_masks,
                 std::vector<u64>& out_masks, int level) {
  // src is indexed by absolute output position; only [lo, lo+width) is ours, values in the same
  // range.
  if (width == 1) return;
  int d = width / 2;
  if (width == 2) {
    // single swap stage (the middle column)
    if (src[lo] == lo + 1) in_masks[level] |= 1ull << lo;
    return;
  }
  std::vector<int> inv(64, -1);  // inv[input position] = output position
  for (int j = lo; j < lo + width; ++j) inv[src[j]] = j;
  // colour[input position]: 0 = goes through lower sub-network, 1 = upper
  std::vector<int> colour(64, -1);
  auto partner = [&](int pos) { return pos < lo + d ? pos + d : pos - d; };
  for (int start = lo; start < lo + width; ++start) {
    if (colour[start] != -1) continue;
    int e = start;
    int c = 0;
    while (colour[e] == -1) {
      colour[e] = c;
      int ip = partner(e);  // input partner takes the other sub-network
      colour[ip] = 1 - c;
      // the output partner of ip's destination must take the sub-network ip did not take
      int out_ip = inv[ip];
      int out_partner = partner(out_ip);
      e = src[out_partner];  // input element feeding that output
      c = 1 - colour[ip];
    }
  }
  // input swaps: pair (i, i+d) swapped iff the element at the lower position goes to the upper net
  std::vector<int> after_in(64, -1);  // after_in[position after input stage] = original input pos
  for (int i = lo; i < lo + d; ++i) {
    if (colour[i] == 1) {
      in_masks[level] |= 1ull << i;
      after_in[i] = i + d;
      after_in[i + d] = i;
    } else {
      after_in[i] = i;
      after_in[i + d] = i + d;
    }
  }
  std::vector<int> pos_after_in(64, -1);
  for (int i = lo; i < lo + width; ++i) pos_after_in[after_in[i]] = i;
  // output swaps: output pair (j, j+d); lower output takes from the lower net unless swapped
  std::vector<int> sub(64, -1);  // sub[position before output stage] = position after input stage
  for (int j = lo; j < lo + d; ++j) {
    int e_lo = src[j], e_hi = src[j + d];
    if (colour[e_lo] == 1) {
      out_masks[level] |= 1ull << j;
      sub[j] = pos_after_in[e_hi];
      sub[j + d] = pos_after_in[e_lo];
    } else {
      sub[j] = pos_after_in[e_lo];
      sub[j + d] = pos_after_in[e_hi];
    }
  }
  std::vector<int> next = src;
  for (int j = lo; j < lo + width; ++j) next[j] = sub[j];
  benes_route(next, lo, d, in_masks, out_masks, level + 1);
  benes_route(next, lo + d, d, in_masks, out_masks, level + 1);
}

Benes benes_compile(const std::vector<int>& p) {
  std::vector<int> src(64);
  for (int i = 0; i < 64; ++i) src[i] = i < (int)p.size() ? p[i] : i;
  std::vector<u64> in_masks(6, 0), out_masks(6, 0);
  benes_route(src, 0, 64, in_masks, out_masks, 0);
  Benes b;
  for (int l = 0; l < 6; ++l)
    if (in_masks[l]) b.stages.push_back({in_masks[l], 32 >> l});
  for (int l = 4; l >= 0; --l)
    if (out_masks[l]) b.stages.push_back({out_masks[l], 32 >> l});
  return b;
}

// ---------------------------------------------------------------------------------------------
// Group
// ---------------------------------------------------------------------------------------------
struct Element {
  std::vector<int> perm;
  i64 phase;  // numerator over Group::denom
  Benes net;
  cplx chi;
};

struct Group {
  int n = 0;  // permutation length (0 for the trivial group of unspecified length)
  i64 denom = 1;
  std::vector<Element> elems;
};

int periodicity_of(const std::vector<int>& p) {
  std::vector<int> cur(p.size());
  std::iota(cur.begin(), cur.end(), 0);
  for (int k = 1;; ++k) {
    std::vector<int> nxt(p.size());
    for (size_t i = 0; i < p.size(); ++i) nxt[i] = cur[p[i]];
    cur = nxt;
    bool id = true;
    for (size_t i = 0; i < p.size(); ++i) id = id && cur[i] == (int)i;
    if (id) return k;
  }
}

bool is_bijection(const std::vector<int>& p) {
  std::vector<char> seen(p.size(), 0);
  for (int v : p) {
    if (v < 0 || v >= (int)p.size() || seen[v]) return false;
    seen[v] = 1;
  }
  return true;
}

int make_group(int n, int n_gens, const int* perms, const int* sectors, Group& g) {
  g.n = n;
  std::vector<std::vector<int>> gens;
  std::vector<int> periods;
  i64 denom = 1;
  for (int k = 0; k < n_gens; ++k) {
    std::vector<int> p(perms + (size_t)k * n, perms + (size_t)(k + 1) * n);
    if (!is_bijection(p)) return ORC_INVALID_PERMUTATION;
    int per = periodicity_of(p);
    if (sectors[k] < 0 || sectors[k] >= per) return ORC_INVALID_SECTOR;
    gens.push_back(p);
    periods.push_back(per);
    denom = std::lcm(denom, (i64)per);
  }
  if (denom % 2) denom *= 2;  // so that a sign (-1) is representable as denom/2
  g.denom = denom;
  std::vector<i64> gphase(n_gens);
  for (int k = 0; k < n_gens; ++k) gphase[k] = (i64)sectors[k] * (denom / periods[k]) % denom;
  std::vector<int> id(n);
  std::iota(id.begin(), id.end(), 0);
  std::map<std::vector<int>, i64> seen;
  std::vector<std::pair<std::vector<int>, i64>> queue;
  seen[id] = 0;
  queue.push_back({id, 0});
  for (size_t head = 0; head < queue.size(); ++head) {
    auto cur = queue[head];
    for (int k = 0; k < n_gens; ++k) {
      // apply generator after cur:  (gen . cur . x)[i] = (cur.x)[gen[i]] = x[cur[gen[i]]]
      std::vector<int> q(n);
      for (int i = 0; i < n; ++i) q[i] = cur.first[gens[k][i]];
      i64 ph = (cur.second + gphase[k]) % denom;
      auto it = seen.find(q);
      if (it == seen.end()) {
        seen[q] = ph;
        queue.push_back({q, ph});
      } else if (it->second != ph) {
        return ORC_INCOMPATIBLE_SYMMETRIES;
      }
    }
  }
  for (auto& e : queue) {
    Element el;
    el.perm = e.first;
    el.phase = e.second;
    el.net = benes_compile(e.first);
    double ang = 2.0 * M_PI * (double)e.second / (double)denom;
    // exact values on the axes keep real sectors exactly real
    if (e.second == 0) el.chi = cplx(1, 0);
    else if (2 * e.second == denom) el.chi = cplx(-1, 0);
    else if (4 * e.second == denom) el.chi = cplx(0, 1);
    else if (4 * e.second == 3 * denom) el.chi = cplx(0, -1);
    else el.chi = cplx(std::cos(ang), std::sin(ang));
    g.elems.push_back(std::move(el));
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------------
// Basis
// ---------------------------------------------------------------------------------------------
struct Basis {
  Group group;
  int n = 0;
  int hw = -1;
  int inv = 0;
  bool built = false;
  bool use_naive = false;  // test hook: bit-by-bit permutations instead of Benes
  std::vector<u64> reps;
  std::vector<double> norms;
  bool lazy_norms = false;  // adopted without the O(N |G|) norm pass (bounded samples of huge sectors)
  // lookup: prefix buckets on the top `pbits` bits of the n-bit word
  int pbits = 0, pshift = 0;
  std::vector<u64> bucket;

  u64 full_mask() const { return n == 64 ? ~0ull : ((1ull << n) - 1); }
  bool trivial() const { return group.elems.size() <= 1 && inv == 0; }
  size_t gsize() const { return group.elems.size() * (inv != 0 ? 2 : 1); }

  inline u64 image(const Element& e, u64 x) const {
    return use_naive ? apply_naive(e.perm, x) : e.net(x);
  }

  // representative, character of the element mapping x to it, sum over the stabiliser of
  // conj(chi) expressed exactly: returns stab count if every stabiliser element has chi == 1,
  // else 0 (a sum of a non-trivial character over a subgroup vanishes).
  void state_info(u64 x, u64& rep, cplx& chi, int& stab_ok_count) const {
    if (trivial()) {
      rep = x;
      chi = 1.0;
      stab_ok_count = 1;
      return;
    }
    const i64 D = group.denom;
    const u64 M = full_mask();
    rep = x;
    i64 best_phase = 0;
    int stab = 0;
    bool bad = false;
    for (auto const& e : group.elems) {
      u64 y = image(e, x);
      for (int f = 0; f < (inv != 0 ? 2 : 1); ++f) {
        u64 z = f ? (y ^ M) : y;
        i64 ph = e.phase;
        if (f && inv == -1) ph = (ph + D / 2) % D;
        if (z < rep) {
          rep = z;
          best_phase = ph;
        }
        if (z == x) {
          ++stab;
          if (ph != 0) bad = true;
        }
      }
    }
    if (best_phase == 0) chi = cplx(1, 0);
    else if (2 * best_phase == D) chi = cplx(-1, 0);
    else if (4 * best_phase == D) chi = cplx(0, 1);
    else if (4 * best_phase == 3 * D) chi = cplx(0, -1);
    else {
      double ang = 2.0 * M_PI * (double)best_phase / (double)D;
      chi = cplx(std::cos(ang), std::sin(ang));
    }
    stab_ok_count = bad ? 0 : stab;
  }

  double norm_from_stab(int stab) const { return std::sqrt((double)stab / (double)gsize()); }

  // norm of representative number idx (computed on demand when the basis was adopted lazily)
  double norm_at(u64 idx) const {
    if (!lazy_norms) return norms[idx];
    if (trivial()) return 1.0;
    u64 rep; cplx chi; int stab;
    state_info(reps[idx], rep, chi, stab);
    return norm_from_stab(stab);
  }

  // is x an orbit minimum with non-zero norm?  (early exit on the first smaller image)
  bool is_representative(u64 x, int& stab_out) const {
    const i64 D = group.denom;
    const u64 M = full_mask();
    int stab = 0;
    bool bad = false;
    for (auto const& e : group.elems) {
      u64 y = image(e, x);
      if (y < x) return false;
      if (y == x) {
        ++stab;
        if (e.phase != 0) bad = true;
      }
      if (inv != 0) {
        u64 z = y ^ M;
        if (z < x) return false;
        if (z == x) {
          ++stab;
          i64 ph = inv == -1 ? (e.phase + D / 2) % D : e.phase;
          if (ph != 0) bad = true;
        }
      }
    }
    stab_out = stab;
    return !bad;
  }

  void build_index() {
    int top = n;
    if (!reps.empty()) {
      u64 mx = reps.back();
      top = mx == 0 ? 1 : 64 - __builtin_clzll(mx);
    }
    pbits = std::min(top, 16);
    pshift = top - pbits;
    bucket.assign(((size_t)1 << pbits) + 1, 0);
    for (u64 r : reps) bucket[(r >> pshift) + 1]++;
    for (size_t i = 1; i < bucket.size(); ++i) bucket[i] += bucket[i - 1];
  }

  // index of representative r, or -1
  i64 index_of(u64 r) const {
    u64 pre = r >> pshift;
    if (pre >= ((u64)1 << pbits)) return -1;
    auto b = reps.begin() + bucket[pre], e = reps.begin() + bucket[pre + 1];
    auto it = std::lower_bound(b, e, r);
    if (it == e || *it != r) return -1;
    return it - reps.begin();
  }
};

u64 binom(int n, int k) {
  if (k < 0 || k > n) return 0;
  static u64 table[65][65];
  static bool init = false;
  if (!init) {
    for (int i = 0; i <= 64; ++i) {
      table[i][0] = 1;
      for (int j = 1; j <= i; ++j)
        table[i][j] = table[i - 1][j - 1] + (j <= i - 1 ? table[i - 1][j] : 0);
    }
    init = true;
  }
  return table[n][k];
}

// rank (in increasing integer order) -> word with `k` bits set
u64 unrank(u64 r, int k) {
  u64 x = 0;
  for (int i = k; i >= 1; --i) {
    int c = i - 1;
    while (binom(c + 1, i) <= r) ++c;
    x |= 1ull << c;
    r -= binom(c, i);
  }
  return x;
}

inline u64 next_same_popcount(u64 x) {
  u64 t = x | (x - 1);
  return (t + 1) | (((~t & -~t) - 1) >> (__builtin_ctzll(x) + 1));
}

int build_basis(Basis& b) {
  b.lazy_norms = false;
  b.reps.clear();
  b.norms.clear();
  const int n = b.n;
  u64 total = b.hw >= 0 ? binom(n, b.hw) : (n == 64 ? 0 : (1ull << n));
  if (b.hw < 0 && n > 40) return ORC_INVALID_ARGUMENT;  // oracle bound, not a reference rule
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  size_t nchunks = std::max<size_t>(1, std::min<u64>(total / 4096 + 1, (u64)nthreads * 64));
  std::vector<std::vector<u64>> out_r(nchunks);
  std::vector<std::vector<int>> out_s(nchunks);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < nchunks; ++c) {
    u64 lo = (u64)((__uint128_t)total * c / nchunks), hi = (u64)((__uint128_t)total * (c + 1) / nchunks);
    if (lo >= hi) continue;
    u64 x = b.hw >= 0 ? (b.hw == 0 ? 0 : unrank(lo, b.hw)) : lo;
    for (u64 r = lo; r < hi; ++r) {
      int stab = 1;
      if (b.trivial() || b.is_representative(x, stab)) {
        out_r[c].push_back(x);
        out_s[c].push_back(stab);
      }
      if (r + 1 < hi) x = b.hw >= 0 ? next_same_popcount(x) : x + 1;
    }
  }
  for (size_t c = 0; c < nchunks; ++c) {
    for (size_t i = 0; i < out_r[c].size(); ++i) {
      b.reps.push_back(out_r[c][i]);
      b.norms.push_back(b.trivial() ? 1.0 : b.norm_from_stab(out_s[c][i]));
    }
  }
  b.build_index();
  b.built = true;
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------------
// Operator
// ---------------------------------------------------------------------------------------------
struct Term {
  int k;
  std::vector<cplx> m;  // row-major 2^k x 2^k
  std::vector<int> sites;  // count * k
};

struct Operator {
  Basis* basis;
  std::vector<Term> terms;
  bool real_matrices() const {
    for (auto& t : terms)
      for (auto& v : t.m)
        if (v.imag() != 0.0) return false;
    return true;
  }
  bool is_real() const {
    if (!real_matrices()) return false;
    for (auto const& e : basis->group.elems)
      if (e.chi.imag() != 0.0) return false;
    return true;
  }
};

template <class T> struct Scalar;
template <> struct Scalar<float> { static cplx load(const float* p) { return cplx(*p, 0); } static void store(float* p, cplx v) { *p = (float)v.real(); } };
template <> struct Scalar<double> { static cplx load(const double* p) { return cplx(*p, 0); } static void store(double* p, cplx v) { *p = v.real(); } };
template <> struct Scalar<std::complex<float>> { static cplx load(const std::complex<float>* p) { return cplx(p->real(), p->imag()); } static void store(std::complex<float>* p, cplx v) { *p = std::complex<float>((float)v.real(), (float)v.imag()); } };
template <> struct Scalar<cplx> { static cplx load(const cplx* p) { return *p; } static void store(cplx* p, cplx v) { *p = v; } };

// y = H x (column-major blocks), optionally also counts matrix elements
// rows row_lo, row_lo + stride, ... < row_hi; or, with row_list != null, exactly the n_list listed rows,
// whose results are then stored compactly (y[c * ys + k] for the k-th listed row)
template <class T>
int matmat(const Operator& op, u64 size, u64 block, const T* x, u64 xs, T* y, u64 ys, u64* n_offdiag,
           u64 row_lo = 0, u64 row_hi = ~0ull, u64 row_stride = 1, const u64* row_list = nullptr, u64 n_list = 0) {
  const Basis& B = *op.basis;
  if (!B.built) return ORC_NOT_BUILT;
  if (size != B.reps.size()) return ORC_DIMENSION_MISMATCH;
  u64 count = 0;
  if (row_hi > size) row_hi = size;
  if (row_stride == 0) row_stride = 1;
  i64 n_rows = row_hi > row_lo ? (i64)((row_hi - row_lo + row_stride - 1) / row_stride) : 0;
  if (row_list) {
    n_rows = (i64)n_list;
    for (u64 k = 0; k < n_list; ++k)
      if (row_list[k] >= size) return ORC_INVALID_ARGUMENT;
  }
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : count)
  for (i64 it = 0; it < n_rows; ++it) {
    const i64 row = row_list ? (i64)row_list[it] : (i64)(row_lo + (u64)it * row_stride);
    const u64 r = B.reps[row];
    const double nr = B.norm_at(row);
    std::vector<cplx> acc(block, cplx(0, 0));
    for (auto const& t : op.terms) {
      const int k = t.k, dim = 1 << k;
      const size_t ntuples = t.sites.size() / k;
      for (size_t s = 0; s < ntuples; ++s) {
        const int* st = &t.sites[s * k];
        int a = 0;
        for (int j = 0; j < k; ++j) a |= (int)((r >> st[j]) & 1ull) << (k - 1 - j);
        for (int bcfg = 0; bcfg < dim; ++bcfg) {
          cplx h = t.m[(size_t)a * dim + bcfg];
          if (h == cplx(0, 0)) continue;
          if (bcfg == a) {
            for (u64 c = 0; c < block; ++c) acc[c] += h * Scalar<T>::load(x + c * xs + row);
            continue;
          }
          u64 rp = r;
          for (int j = 0; j < k; ++j) {
            u64 bit = (u64)((bcfg >> (k - 1 - j)) & 1);
            rp = (rp & ~(1ull << st[j])) | (bit << st[j]);
          }
          u64 rep;
          cplx chi;
          int stab;
          B.state_info(rp, rep, chi, stab);
          if (stab == 0) continue;
          i64 idx = B.index_of(rep);
          if (idx < 0) continue;
          ++count;
          cplx w = h * chi * (B.norm_at(idx) / nr);
          for (u64 c = 0; c < block; ++c) acc[c] += w * Scalar<T>::load(x + c * xs + idx);
        }
      }
    }
    if (y)
      for (u64 c = 0; c < block; ++c) Scalar<T>::store(y + c * ys + (row_list ? it : row), acc[c]);
  }
  if (n_offdiag) *n_offdiag = count;
  return ORC_OK;
}

}  // namespace

extern "C" {

void* orc_basis_new(int n, int hw, int inv, int n_gens, const int* perms, const int* sectors, int* err) {
  *err = ORC_OK;
  if (n <= 0 || n > 64) { *err = ORC_INVALID_NUMBER_SPINS; return nullptr; }
  if (hw < -1 || hw > n) { *err = ORC_INVALID_HAMMING_WEIGHT; return nullptr; }
  if (inv != 0 && inv != 1 && inv != -1) { *err = ORC_INVALID_SPIN_INVERSION; return nullptr; }
  if (inv != 0 && hw >= 0 && 2 * hw != n) { *err = ORC_INVALID_SPIN_INVERSION; return nullptr; }
  auto* b = new Basis;
  b->n = n; b->hw = hw; b->inv = inv;
  int rc = make_group(n, n_gens, perms, sectors, b->group);
  if (rc != ORC_OK) { *err = rc; delete b; return nullptr; }
  return b;
}
void orc_basis_free(void* b) { delete (Basis*)b; }
void orc_basis_use_naive(void* b, int flag) { ((Basis*)b)->use_naive = flag != 0; }
u64 orc_basis_group_size(void* b) { return ((Basis*)b)->group.elems.size(); }
void orc_basis_group_element(void* b, u64 idx, int* perm, double* phase) {
  auto& e = ((Basis*)b)->group.elems[idx];
  for (size_t i = 0; i < e.perm.size(); ++i) perm[i] = e.perm[i];
  *phase = (double)e.phase / (double)((Basis*)b)->group.denom;
}
int orc_periodicity(int n, const int* perm) {
  std::vector<int> p(perm, perm + n);
  if (!is_bijection(p)) return -1;
  return periodicity_of(p);
}
int orc_basis_build(void* b) { return build_basis(*(Basis*)b); }
int orc_basis_build_unsafe(void* bp, u64 size, const u64* reps) {
  Basis& b = *(Basis*)bp;
  b.lazy_norms = false;
  b.reps.assign(reps, reps + size);
  b.norms.resize(size);
#pragma omp parallel for
  for (i64 i = 0; i < (i64)size; ++i) {
    u64 rep; cplx chi; int stab;
    b.state_info(b.reps[i], rep, chi, stab);
    b.norms[i] = b.trivial() ? 1.0 : b.norm_from_stab(stab);
  }
  b.build_index();
  b.built = true;
  return ORC_OK;
}
// adopt without computing the norms (they are derived per use): for bounded samples of sectors
// with ~10^9 representatives, where the eager pass alone would take minutes of CPU time
int orc_basis_adopt_lazy(void* bp, u64 size, const u64* reps) {
  Basis& b = *(Basis*)bp;
  b.reps.assign(reps, reps + size);
  b.norms.clear();
  b.lazy_norms = true;
  b.build_index();
  b.built = true;
  return ORC_OK;
}
// representatives among the candidates of combinatorial rank [rank_lo, rank_hi) of the sector
// (independent enumeration of a window of a huge sector); returns their number, writes at most cap
// of them to out, and the first / one-past-last candidate WORD of the window to word_range[2]
// (word_range[1] = ~0 when the window reaches the end of the sector).
u64 orc_basis_build_range(void* bp, u64 rank_lo, u64 rank_hi, u64 cap, u64* out, u64* word_range) {
  Basis& b = *(Basis*)bp;
  const int n = b.n;
  u64 total = b.hw >= 0 ? binom(n, b.hw) : (n == 64 ? ~0ull : (1ull << n));
  rank_hi = std::min(rank_hi, total);
  if (rank_lo >= rank_hi) { word_range[0] = word_range[1] = ~0ull; return 0; }
  auto word = [&](u64 r) { return b.hw >= 0 ? (b.hw == 0 ? 0ull : unrank(r, b.hw)) : r; };
  word_range[0] = word(rank_lo);
  word_range[1] = rank_hi < total ? word(rank_hi) : ~0ull;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  const u64 span = rank_hi - rank_lo;
  size_t nchunks = std::max<size_t>(1, std::min<u64>(span / 4096 + 1, (u64)nthreads * 16));
  std::vector<std::vector<u64>> found(nchunks);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < nchunks; ++c) {
    u64 lo = rank_lo + (u64)((__uint128_t)span * c / nchunks), hi = rank_lo + (u64)((__uint128_t)span * (c + 1) / nchunks);
    if (lo >= hi) continue;
    u64 x = word(lo);
    for (u64 r = lo; r < hi; ++r) {
      int stab = 1;
      if (b.trivial() || b.is_representative(x, stab)) found[c].push_back(x);
      if (r + 1 < hi) x = b.hw >= 0 ? next_same_popcount(x) : x + 1;
    }
  }
  u64 count = 0;
  for (auto const& v : found)
    for (u64 x : v) {
      if (count < cap) out[count] = x;
      ++count;
    }
  return count;
}
u64 orc_sector_candidates(void* bp) {
  Basis& b = *(Basis*)bp;
  return b.hw >= 0 ? binom(b.n, b.hw) : (b.n == 64 ? ~0ull : (1ull << b.n));
}
u64 orc_basis_size(void* b) { return ((Basis*)b)->reps.size(); }
void orc_basis_states(void* b, u64* out) { auto& r = ((Basis*)b)->reps; std::memcpy(out, r.data(), r.size() * 8); }
void orc_basis_norms(void* b, double* out) { auto& r = ((Basis*)b)->norms; std::memcpy(out, r.data(), r.size() * 8); }
void orc_state_info(void* bp, u64 x, u64* rep, double* chi, double* norm) {
  Basis& b = *(Basis*)bp;
  cplx c; int stab;
  b.state_info(x, *rep, c, stab);
  chi[0] = c.real(); chi[1] = c.imag();
  *norm = b.trivial() ? 1.0 : b.norm_from_stab(stab);
}
long long orc_basis_index(void* b, u64 rep) { return ((Basis*)b)->index_of(rep); }
u64 orc_apply_permutation(int n, const int* perm, u64 x, int naive) {
  std::vector<int> p(perm, perm + n);
  return naive ? apply_naive(p, x) : benes_compile(p)(x);
}

void* orc_operator_new(void* basis) { auto* o = new Operator; o->basis = (Basis*)basis; return o; }
void orc_operator_free(void* o) { delete (Operator*)o; }
int orc_operator_add_term(void* op, int k, const double* matrix, int count, const int* sites) {
  if (k < 1 || k > 4) return ORC_INVALID_ARGUMENT;
  Operator& o = *(Operator*)op;
  Term t;
  t.k = k;
  int dim = 1 << k;
  t.m.resize((size_t)dim * dim);
  for (int i = 0; i < dim * dim; ++i) t.m[i] = cplx(matrix[2 * i], matrix[2 * i + 1]);
  t.sites.assign(sites, sites + (size_t)count * k);
  for (int s : t.sites)
    if (s < 0 || s >= o.basis->n) return ORC_INVALID_ARGUMENT;
  o.terms.push_back(std::move(t));
  return ORC_OK;
}
int orc_operator_is_real(void* op) { return ((Operator*)op)->is_real() ? 1 : 0; }

// rows row_lo, row_lo + stride, ... < row_hi only (bounded samples for CPU timing); y keeps the
// full layout, untouched rows are left as they are.
int orc_operator_matmat_rows(void* op, int dtype, u64 size, u64 block, const void* x, u64 xs, void* y, u64 ys,
                             u64* n_offdiag, u64 row_lo, u64 row_hi, u64 row_stride) {
  Operator& o = *(Operator*)op;
  if ((dtype == 0 || dtype == 1) && !o.is_real()) return ORC_INVALID_DATATYPE;
  switch (dtype) {
    case 0: return matmat<float>(o, size, block, (const float*)x, xs, (float*)y, ys, n_offdiag, row_lo, row_hi, row_stride);
    case 1: return matmat<double>(o, size, block, (const double*)x, xs, (double*)y, ys, n_offdiag, row_lo, row_hi, row_stride);
    case 2: return matmat<std::complex<float>>(o, size, block, (const std::complex<float>*)x, xs, (std::complex<float>*)y, ys, n_offdiag, row_lo, row_hi, row_stride);
    case 3: return matmat<cplx>(o, size, block, (const cplx*)x, xs, (cplx*)y, ys, n_offdiag, row_lo, row_hi, row_stride);
  }
  return ORC_INVALID_DATATYPE;
}
// exactly the listed rows; y_out[k] = (H x)[rows[k]] (one column, compact)
int orc_operator_matmat_list(void* op, int dtype, u64 size, const void* x, u64 n_rows, const u64* rows, void* y_out,
                             u64* n_offdiag) {
  Operator& o = *(Operator*)op;
  if ((dtype == 0 || dtype == 1) && !o.is_real()) return ORC_INVALID_DATATYPE;
  switch (dtype) {
    case 0: return matmat<float>(o, size, 1, (const float*)x, size, (float*)y_out, n_rows, n_offdiag, 0, ~0ull, 1, rows, n_rows);
    case 1: return matmat<double>(o, size, 1, (const double*)x, size, (double*)y_out, n_rows, n_offdiag, 0, ~0ull, 1, rows, n_rows);
    case 2: return matmat<std::complex<float>>(o, size, 1, (const std::complex<float>*)x, size, (std::complex<float>*)y_out, n_rows, n_offdiag, 0, ~0ull, 1, rows, n_rows);
    case 3: return matmat<cplx>(o, size, 1, (const cplx*)x, size, (cplx*)y_out, n_rows, n_offdiag, 0, ~0ull, 1, rows, n_rows);
  }
  return ORC_INVALID_DATATYPE;
}
int orc_operator_matmat(void* op, int dtype, u64 size, u64 block, const void* x, u64 xs, void* y, u64 ys, u64* n_offdiag) {
  return orc_operator_matmat_rows(op, dtype, size, block, x, xs, y, ys, n_offdiag, 0, ~0ull, 1);
}

// out[k] = <x_k | O | x_k>  (complex128), via a c128 matvec
int orc_operator_expectation(void* op, int dtype, u64 size, u64 block, const void* x, u64 xs, double* out) {
  Operator& o = *(Operator*)op;
  std::vector<cplx> xc(size * block), yc(size * block);
  for (u64 c = 0; c < block; ++c)
    for (u64 i = 0; i < size; ++i) {
      switch (dtype) {
        case 0: xc[c * size + i] = ((const float*)x)[c * xs + i]; break;
        case 1: xc[c * size + i] = ((const double*)x)[c * xs + i]; break;
        case 2: xc[c * size + i] = cplx(((const std::complex<float>*)x)[c * xs + i]); break;
        case 3: xc[c * size + i] = ((const cplx*)x)[c * xs + i]; break;
        default: return ORC_INVALID_DATATYPE;
      }
    }
  int rc = matmat<cplx>(o, size, block, xc.data(), size, yc.data(), size, nullptr);
  if (rc) return rc;
  for (u64 c = 0; c < block; ++c) {
    cplx acc(0, 0);
    for (u64 i = 0; i < size; ++i) acc += std::conj(xc[c * size + i]) * yc[c * size + i];
    out[2 * c] = acc.real();
    out[2 * c + 1] = acc.imag();
  }
  return ORC_OK;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"
