// emul.cpp -- single-threaded HOST emulation of the matvec kernels, for the test-suite only.
//
// Compiled by the host compiler (never by nvcc): the CUDA built-ins the device code uses are
// shimmed below (one "thread", one "block"), and the very same sources -- matvec_kernel.cuh
// (matrix-free rows, cache fill) and cached_kernel.cuh (streaming kernel) -- run on small problems
// so that slot arithmetic, source classes and summation logic can be checked against the oracle
// without a GPU (tests/test_emulation.py).  The orchestration around the kernels (code maps,
// slice widths, offsets) mirrors Operator::cache_usable with std::vector buffers.  This is a
// verification hook built into the test-only library libsped_emul.so (emul.h); libsped.so neither
// contains nor calls it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include "emul.h"
#include "internal.h"

// ---- shims for the device built-ins (one thread, one block) ----
#if !defined(__launch_bounds__)
#define __launch_bounds__(...)
#endif
#define SPED_KERNEL_LINKAGE static
#define SPED_KERNEL_THREADS 32  // one emulated warp per block: block_reduce_store sums blockDim / 32 warp results
namespace {
struct Dim1 { unsigned x, y, z; };
}
static const Dim1 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
template <class T> static inline T __ldg(T const* p) { return *p; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }

// the other lanes of the emulated warp hold nothing: a shuffle from them contributes zero
template <class T> static inline T __shfl_down_sync(unsigned, T, int) { return T(0); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

#include "canon.cuh"          // ProgramCanon, TrivialCanon (+ matvec_kernel.cuh)
#include "cached_kernel.cuh"  // the streaming kernel
#include "eigh_kernels.cuh"   // fused kernels of the single-pair eigensolver iteration

namespace sped {
namespace {

struct Problem {
  Operator& op;
  Basis& b;
  RowDist dist;
  std::vector<u64> bucket;
  std::vector<double> norm_table, chi_table, diag_re, diag_im;
  std::vector<DevBond> bonds;
  std::vector<double> pool_re, pool_im;
  std::vector<std::uint16_t> masks;
  std::vector<FastStep<u32>> fast32;
  std::vector<PermOp<u32>> ops32;
  BasisIndex ix{};
  TermsView terms{};
  RowContext ctx{};
  bool sym;

  Problem(Operator& o, u64 n, u64 const* reps, std::uint16_t const* stab, int world, int rank) : op(o), b(*o.basis) {
    sym = !b.trivial();
    dist = make_row_dist(n, world, rank);
    bucket = {0, n, n};  // one prefix bucket holding everything (representatives are < 2^63)
    ix.reps = reps;
    ix.stab = sym ? stab : nullptr;
    ix.bucket = bucket.data();
    ix.n_states = n;
    ix.bucket_shift = 63;
    ix.bucket_count = 2;
    ix.bucket_wide = 1;
    ix.direct = 0;
    u64 const order = b.group_order();
    norm_table.resize(order + 1);
    for (u64 s = 0; s <= order; ++s) norm_table[s] = std::sqrt((double)s / (double)order);
    i64 const D = b.group->denom;
    chi_table.resize(2 * (size_t)D);
    for (i64 k = 0; k < D; ++k) {
      long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)D;
      chi_table[2 * k] = (double)std::cos(ang);
      chi_table[2 * k + 1] = (double)std::sin(ang);
      if (2 * k == D) { chi_table[2 * k] = -1.0; chi_table[2 * k + 1] = 0.0; }
      if (k == 0) { chi_table[0] = 1.0; chi_table[1] = 0.0; }
    }
    packed_terms_host(op.terms, bonds, pool_re, pool_im, masks);
    terms.bonds = bonds.data();
    terms.pool_re = pool_re.data();
    terms.pool_im = pool_im.data();
    terms.masks = masks.data();
    terms.n_bonds = (u32)bonds.size();
    terms.pool_size = (u32)pool_re.size();
    terms.mask_size = (u32)masks.size();
    ctx = RowContext{ix, norm_table.data(), chi_table.data(), dist};
    // diagonal of the local rows (what diagonal_kernel computes)
    diag_re.assign(std::max<u64>(dist.n_local, 1), 0.0);
    diag_im.assign(std::max<u64>(dist.n_local, 1), 0.0);
    for (u64 i = 0; i < dist.n_local; ++i) {
      u64 const r = reps[dist_local_to_global(dist, i)];
      double re = 0, im = 0;
      for (auto const& bd : bonds) {
        u32 a = 0;
        for (u32 j = 0; j < bd.k; ++j) a |= (u32)((r >> ((bd.sites >> (8 * j)) & 0xffu)) & 1ull) << (bd.k - 1 - j);
        u32 const dim = 1u << bd.k;
        re += pool_re[bd.moff + a * dim + a];
        im += pool_im[bd.moff + a * dim + a];
      }
      diag_re[i] = re;
      diag_im[i] = im;
    }
    auto const& P = b.program;
    for (auto const& f : P.fast) fast32.push_back(FastStep<u32>{(u32)f.mask, f.ctl});
    for (auto const& q : P.ops) ops32.push_back(PermOp<u32>{(u32)q.mask, q.amount});
  }
  ProgramView<u64> view64() const {
    auto const& P = b.program;
    return ProgramView<u64>{P.fast.data(), P.steps.data(), P.ops.data(), P.phase.data(), (u32)P.steps.size(), (u32)P.ops.size(),
                            P.n_spins, P.shift, P.inversion, P.denom};
  }
  ProgramView<u32> view32() const {
    auto const& P = b.program;
    return ProgramView<u32>{fast32.data(), P.steps.data(), ops32.data(), P.phase.data(), (u32)P.steps.size(), (u32)P.ops.size(),
                            P.n_spins, 0, P.inversion, P.denom};
  }
  // calls f(canon) with the canonicalisation functor of this basis
  template <class F>
  void with_canon(F&& f) const {
    if (!sym) f(TrivialCanon());
    else if (b.use32()) f(ProgramCanon<u32>{view32()});
    else f(ProgramCanon<u64>{view64()});
  }
};

template <class T>
void run(Problem& pr, T const* x_global, T* y_free, T* y_all, T* y_phased, u64* stats, u32 ncols, T* y_block) {
  RowDist const& d = pr.dist;
  u64 const n_local = d.n_local, padded = std::max<u64>(d.chunk * d.world, 1);
  std::vector<T> xfull(padded, T{});
  for (u64 g = 0; g < d.n; ++g) xfull[dist_global_to_pos(d, g)] = x_global[g];
  bool const cplx_diag = !pr.op.real_diagonal;

  // ---- matrix-free rows ----
  MatvecParams mp{};
  mp.ctx = pr.ctx;
  mp.terms = pr.terms;
  mp.diag_re = pr.diag_re.data();
  mp.diag_im = cplx_diag ? pr.diag_im.data() : nullptr;
  mp.x = xfull.data();
  mp.y = y_free;
  mp.xs = padded;
  mp.ys = std::max<u64>(n_local, 1);
  mp.ncols = 1;
  pr.with_canon([&](auto const& canon) { matvec_rows<T, 1>(mp, pr.terms, canon); });

  // ---- operator cache: maps, table, widths, offsets, fill (mirrors Operator::cache_usable) ----
  CodeMaps cm;
  if (char const* why = build_code_maps(pr.op, cm)) fail(LS_INVALID_ARGUMENT, std::string("no operator cache: ") + why);
  u32 const n_sid = (u32)cm.sid_stab.size();
  std::vector<double> table(3 * cm.n_codes);
  for (u64 c = 0; c < cm.n_codes; ++c) {  // table_kernel
    u32 sid = (u32)(c % n_sid), pid = (u32)((c / n_sid) % cm.n_pid), hid = (u32)(c / ((u64)n_sid * cm.n_pid));
    u32 ph = pr.sym ? cm.pid_phase[pid] : 0u;
    double vx = cm.values_re[hid], vy = cm.values_im[hid];
    if (pr.sym) {
      if (!pr.op.is_real()) {
        double cx = pr.chi_table[2 * ph], cy = pr.chi_table[2 * ph + 1];
        double nx = vx * cx - vy * cy, ny = vx * cy + vy * cx;
        vx = nx;
        vy = ny;
      } else {
        vx = ph == 0 ? vx : -vx;
        vy = 0.0;
      }
    }
    table[3 * c] = vx;
    table[3 * c + 1] = vy;
    table[3 * c + 2] = pr.sym ? pr.norm_table[cm.sid_stab[sid]] : 1.0;
  }
  u32 const rounds = (u32)exchange_rounds(d.world, d.chunk);
  u32 const n_classes = 1u + rounds;
  u32 const near = exchange_near(d.world, d.chunk);
  bool const wide = cm.n_codes > 256;
  u64 const n_slices = (n_local + 31) / 32;
  std::vector<std::uint16_t> len(std::max<u64>(n_local, 1) * 2 * n_classes, 0);
  std::vector<u32> widths(std::max<u64>(n_slices, 1), 0), slice_start(std::max<u64>(n_slices, 1) * kClassStride, 0);
  int overflow = 0;
  FillParams fp{};
  fp.ctx = pr.ctx;
  fp.terms = pr.terms;
  fp.len = len.data();
  fp.n_classes = n_classes;
  fp.near = near;
  fp.rounds = rounds;
  fp.default_code = cm.default_code;
  fp.hid_map = cm.hid_map.data();
  fp.sid_map = pr.sym ? cm.sid_map.data() : nullptr;
  fp.pid_map = pr.sym ? cm.pid_map.data() : nullptr;
  fp.denom = cm.n_pid;
  fp.n_sid = n_sid;
  fp.code_wide = wide ? 1 : 0;
  fp.overflow = &overflow;
  fp.row_lo = 0;
  fp.row_hi = n_local;
  if (rounds > 0) {  // several ranks: exact class sizes from a counting traversal
    fp.count_only = 1;
    pr.with_canon([&](auto const& canon) { cache_fill_rows(fp, pr.terms, canon); });
    for (u64 s = 0; s < n_slices; ++s) {  // class_width_kernel
      u32 w[kMaxClasses] = {0, 0, 0};
      for (u64 i = 32 * s; i < std::min<u64>(32 * s + 32, n_local); ++i)
        for (u32 c = 0; c < n_classes; ++c) w[c] = std::max<u32>(w[c], len[(u64)(2 * c) * n_local + i]);
      widths[s] = w[0] + w[1] + w[2];
      slice_start[kClassStride * s] = w[0];
      slice_start[kClassStride * s + 1] = w[0] + w[1];
    }
    fp.count_only = 0;
    fp.slice_start = slice_start.data();
  } else {
    for (u64 s = 0; s < n_slices; ++s) {  // slice_width_kernel: cheap upper bound
      u32 mx = 0;
      for (u64 i = 32 * s; i < std::min<u64>(32 * s + 32, n_local); ++i) {
        u64 const r = pr.ix.reps[dist_local_to_global(d, i)];
        u32 ub = 0;
        for (auto const& bd : pr.bonds) {
          u32 a = 0;
          for (u32 j = 0; j < bd.k; ++j) a |= (u32)((r >> ((bd.sites >> (8 * j)) & 0xffu)) & 1ull) << (bd.k - 1 - j);
          ub += (u32)__builtin_popcount((unsigned)pr.masks[bd.zoff + a]);
        }
        mx = std::max(mx, ub);
      }
      widths[s] = mx;
    }
  }
  std::vector<u64> slice_off(n_slices + 1, 0);
  for (u64 s = 0; s < n_slices; ++s) slice_off[s + 1] = slice_off[s] + 32ull * widths[s];
  u64 const slots = slice_off[n_slices];
  std::vector<u32> idx(std::max<u64>(slots, 1), 0xdeadbeefu);
  fp.slice_off = slice_off.data();
  fp.idx = idx.data();
  std::fill(len.begin(), len.end(), 0);
  // The fill runs in row chunks (here: 3 slices each, so that even small decks take several); its
  // one-code-per-slot temporary covers one chunk, and the codes of the coded elements are compacted
  // chunk by chunk (code_width_kernel, slice_scan_kernel, code_compact_kernel), then concatenated.
  u64 const slices_per_chunk = 3;
  std::vector<u32> cw(std::max<u64>(n_slices * n_classes, 1), 0);
  std::vector<unsigned char> compact;
  CacheView v{};
  v.slice_off = slice_off.data();
  v.len = len.data();
  v.slice_start = n_classes > 1 ? slice_start.data() : nullptr;
  v.n_slices = n_slices;
  v.n_classes = n_classes;
  for (u64 s_lo = 0; s_lo < n_slices; s_lo += slices_per_chunk) {
    u64 const s_hi = std::min(n_slices, s_lo + slices_per_chunk);
    std::vector<unsigned char> temp(std::max<u64>(slice_off[s_hi] - slice_off[s_lo], 1) * (wide ? 2 : 1), 0xee);
    fp.row_lo = 32 * s_lo;
    fp.row_hi = std::min<u64>(32 * s_hi, n_local);
    fp.code_slot0 = slice_off[s_lo];
    fp.code = temp.data();
    pr.with_canon([&](auto const& canon) { cache_fill_rows(fp, pr.terms, canon); });
    code_width_kernel(len.data(), n_local, n_classes, s_lo, s_hi, cw.data());
    u64 const regions = (s_hi - s_lo) * n_classes;
    std::vector<u64> chunk_off(regions + 1, 0);
    for (u64 k = 0; k < regions; ++k) chunk_off[k + 1] = chunk_off[k] + 32ull * cw[s_lo * n_classes + k];
    std::vector<unsigned char> part(std::max<u64>(chunk_off[regions], 1) * (wide ? 2 : 1), 0xdd);
    v.code = temp.data();
    if (wide)
      code_compact_kernel<std::uint16_t>(v, n_local, fp.row_lo, fp.row_hi, fp.code_slot0, chunk_off.data(),
                                         reinterpret_cast<std::uint16_t*>(part.data()));
    else
      code_compact_kernel<std::uint8_t>(v, n_local, fp.row_lo, fp.row_hi, fp.code_slot0, chunk_off.data(), part.data());
    compact.insert(compact.end(), part.begin(), part.begin() + chunk_off[regions] * (wide ? 2 : 1));
  }
  if (overflow) fail(SPED_INTERNAL_ERROR, "emulated cache fill: a row exceeded its slot bound");
  std::vector<u64> code_off(n_slices * n_classes + 1, 0);
  for (u64 k = 0; k < n_slices * n_classes; ++k) code_off[k + 1] = code_off[k] + 32ull * cw[k];
  u64 const code_slots = code_off[n_slices * n_classes];
  if (compact.size() != code_slots * (wide ? 2 : 1)) fail(SPED_INTERNAL_ERROR, "emulation: the compact code stream does not add up");
  compact.resize(std::max<u64>(code_slots, 1) * (wide ? 2 : 1), 0xdd);
  u64 elements = 0, dflt = 0;
  for (u32 seg = 0; seg < 2 * n_classes; ++seg)
    for (u64 i = 0; i < n_local; ++i) {
      elements += len[(u64)seg * n_local + i];
      if (!(seg & 1)) dflt += len[(u64)seg * n_local + i];
    }
  stats[0] = slots;
  stats[1] = elements;
  stats[2] = dflt;
  stats[3] = n_classes;
  stats[4] = code_slots;
  stats[5] = wide ? 2 : 1;  // bytes per coefficient code

  // ---- several classes: the one-traversal build (staging + cache_place_kernel) must produce the very
  //      same arrays as the counting + filling traversals above ----
  if (rounds > 0 && n_local) {
    std::vector<u32> ub(n_slices, 0);
    for (u64 s = 0; s < n_slices; ++s) {  // slice_width_kernel: cheap upper bound
      u32 mx = 0;
      for (u64 i = 32 * s; i < std::min<u64>(32 * s + 32, n_local); ++i) {
        u64 const r = pr.ix.reps[dist_local_to_global(d, i)];
        u32 w = 0;
        for (auto const& bd : pr.bonds) {
          u32 a = 0;
          for (u32 j = 0; j < bd.k; ++j) a |= (u32)((r >> ((bd.sites >> (8 * j)) & 0xffu)) & 1ull) << (bd.k - 1 - j);
          w += (u32)__builtin_popcount((unsigned)pr.masks[bd.zoff + a]);
        }
        mx = std::max(mx, w);
      }
      ub[s] = mx;
    }
    std::vector<u64> stage_off(n_slices + 1, 0);
    for (u64 s = 0; s < n_slices; ++s) stage_off[s + 1] = stage_off[s] + 32ull * ub[s];
    std::vector<u32> stage_idx(std::max<u64>(stage_off[n_slices], 1), 0xabababab);
    std::vector<unsigned char> stage_code(std::max<u64>(stage_off[n_slices], 1) * (wide ? 2 : 1), 0xcc);
    std::vector<std::uint16_t> len2(len.size(), 0);
    FillParams sp = fp;
    sp.stage = 1;
    sp.count_only = 0;
    sp.slice_off = stage_off.data();
    sp.slice_start = nullptr;
    sp.idx = stage_idx.data();
    sp.code = stage_code.data();
    sp.code_slot0 = 0;
    sp.len = len2.data();
    sp.row_lo = 0;
    sp.row_hi = n_local;
    pr.with_canon([&](auto const& canon) { cache_fill_rows(sp, pr.terms, canon); });
    if (overflow) fail(SPED_INTERNAL_ERROR, "emulated staging traversal: a row exceeded its slot bound");
    if (len2 != len) fail(SPED_INTERNAL_ERROR, "emulation: staged and two-traversal class counts differ");
    // class widths (class_width_kernel with split counts), offsets, compact code offsets
    std::vector<u32> widths2(n_slices, 0), slice_start2(n_slices * kClassStride, 0), cw2(n_slices * n_classes, 0);
    for (u64 s = 0; s < n_slices; ++s) {
      u32 w[kMaxClasses] = {0, 0, 0};
      for (u64 i = 32 * s; i < std::min<u64>(32 * s + 32, n_local); ++i)
        for (u32 c = 0; c < n_classes; ++c)
          w[c] = std::max<u32>(w[c], (u32)len2[(u64)(2 * c) * n_local + i] + (u32)len2[(u64)(2 * c + 1) * n_local + i]);
      widths2[s] = w[0] + w[1] + w[2];
      slice_start2[kClassStride * s] = w[0];
      slice_start2[kClassStride * s + 1] = w[0] + w[1];
    }
    std::vector<u64> slice_off2(n_slices + 1, 0), code_off2(n_slices * n_classes + 1, 0);
    for (u64 s = 0; s < n_slices; ++s) slice_off2[s + 1] = slice_off2[s] + 32ull * widths2[s];
    code_width_kernel(len2.data(), n_local, n_classes, 0, n_slices, cw2.data());
    for (u64 k = 0; k < n_slices * n_classes; ++k) code_off2[k + 1] = code_off2[k] + 32ull * cw2[k];
    std::vector<u32> idx2(std::max<u64>(slice_off2[n_slices], 1), 0xdeadbeefu);
    std::vector<unsigned char> compact2(std::max<u64>(code_off2[n_slices * n_classes], 1) * (wide ? 2 : 1), 0xdd);
    PlaceParams q{};
    q.dist = d;
    q.stage_off = stage_off.data();
    q.stage_idx = stage_idx.data();
    q.stage_code = stage_code.data();
    q.out.slice_off = slice_off2.data();
    q.out.len = len2.data();
    q.out.slice_start = slice_start2.data();
    q.out.n_slices = n_slices;
    q.out.n_classes = n_classes;
    q.out.near = near;
    q.out.rounds = rounds;
    q.out.default_code = cm.default_code;
    q.out.code_off = code_off2.data();
    q.idx = idx2.data();
    q.code = compact2.data();
    if (wide) cache_place_kernel<std::uint16_t>(q);
    else cache_place_kernel<std::uint8_t>(q);
    if (slice_off2 != slice_off || slice_start2 != slice_start || code_off2 != code_off)
      fail(SPED_INTERNAL_ERROR, "emulation: staged and two-traversal layouts differ");
    if (idx2 != idx) fail(SPED_INTERNAL_ERROR, "emulation: staged and two-traversal positions differ");
    if (compact2 != compact) fail(SPED_INTERNAL_ERROR, "emulation: staged and two-traversal code streams differ");
  }

  // ---- streaming kernel: all classes in one pass, then class by class ----
  CachedParams cp{};
  cp.cache = CacheView{slice_off.data(), idx.data(), compact.data(), len.data(), n_classes > 1 ? slice_start.data() : nullptr,
                       table.data(), n_slices, wide ? 1 : 0, (u32)cm.n_codes, n_classes, near, cm.default_code, rounds,
                       code_off.data()};
  cp.ctx = pr.ctx;
  cp.diag_re = pr.diag_re.data();
  cp.diag_im = cplx_diag ? pr.diag_im.data() : nullptr;
  cp.x = xfull.data();
  cp.xs = padded;
  cp.ys = std::max<u64>(n_local, 1);
  cp.ncols = 1;
  cp.sym = pr.sym ? 1 : 0;
  cp.row_lo = 0;
  cp.row_hi = n_local;
  // one pass over all classes: the pipelined batch-of-8 kernel; class by class: the pipelined batch-of-6
  // kernel (the two the product launches, with plain loads); a third, unpipelined run checks the plain loop
  auto launch = [&](T* y, int phase) {
    cp.y = y;
    cp.phase = phase;
    auto go = [&](auto code_tag, auto sym_tag) {
      using Code = decltype(code_tag);
      constexpr bool S = decltype(sym_tag)::value;
      if (phase == 0) cached_matvec_kernel<T, 1, Code, S, false, 8, 4, true>(cp);
      else cached_matvec_kernel<T, 1, Code, S, false, 6, 4, true>(cp);
    };
    if (wide && pr.sym) go(std::uint16_t{}, std::true_type{});
    else if (wide) go(std::uint16_t{}, std::false_type{});
    else if (pr.sym) go(std::uint8_t{}, std::true_type{});
    else go(std::uint8_t{}, std::false_type{});
  };
  if (n_local) {
    {  // the unpipelined loop must give the same sums (same order): checked here, bit for bit
      std::vector<T> plain(n_local);
      cp.y = plain.data();
      cp.phase = 0;
      if (wide && pr.sym) cached_matvec_kernel<T, 1, std::uint16_t, true, false, 4, 6, false>(cp);
      else if (wide) cached_matvec_kernel<T, 1, std::uint16_t, false, false, 4, 6, false>(cp);
      else if (pr.sym) cached_matvec_kernel<T, 1, std::uint8_t, true, false, 4, 6, false>(cp);
      else cached_matvec_kernel<T, 1, std::uint8_t, false, false, 4, 6, false>(cp);
      launch(y_all, 0);
      if (std::memcmp(plain.data(), y_all, n_local * sizeof(T)) != 0)
        fail(SPED_INTERNAL_ERROR, "emulation: pipelined and plain streaming loops disagree");
    }
    launch(y_all, 0);
    for (u32 ph = 1; ph <= 1 + rounds; ++ph) launch(y_phased, (int)ph);  // local pass, then one pass per exchange round
  }
  // ---- block kernel: columns c > 0 are x shifted cyclically by c rows (in global order) ----
  if (ncols > 1 && n_local) {
    std::vector<T> xcols(padded * ncols, T{});
    for (u32 c = 0; c < ncols; ++c)
      for (u64 g = 0; g < d.n; ++g) xcols[(u64)c * padded + dist_global_to_pos(d, g)] = x_global[(g + c) % d.n];
    u32 const nb = ncols == 2 ? 2u : 4u;
    std::vector<T> xt(padded * nb, T{});
    if (nb == 2) interleave_kernel<T, 2>(xcols.data(), padded, ncols, padded, xt.data());
    else interleave_kernel<T, 4>(xcols.data(), padded, ncols, padded, xt.data());
    cp.x = xt.data();
    cp.y = y_block;
    cp.ncols = ncols;
    cp.phase = 0;
    auto go = [&](auto nbtag) {
      constexpr int NB = decltype(nbtag)::value;
      constexpr int U = NB == 2 ? 4 : 2;
      if (wide && pr.sym) cached_block_kernel<T, NB, std::uint16_t, true, U>(cp);
      else if (wide) cached_block_kernel<T, NB, std::uint16_t, false, U>(cp);
      else if (pr.sym) cached_block_kernel<T, NB, std::uint8_t, true, U>(cp);
      else cached_block_kernel<T, NB, std::uint8_t, false, U>(cp);
    };
    if (nb == 2) go(std::integral_constant<int, 2>());
    else go(std::integral_constant<int, 4>());
  }
}

}  // namespace
}  // namespace sped

using namespace sped;

extern "C" int sped_selftest_emulate_matvec(void const* op_handle, uint64_t n, uint64_t const* reps, uint16_t const* stab,
                                            int world, int rank, int dtype, void const* x_global, void* y_free, void* y_all,
                                            void* y_phased, uint64_t* stats, unsigned ncols, void* y_block) {
  return guard([&] {
    auto& op = *static_cast<std::shared_ptr<Operator>*>(const_cast<void*>(op_handle));
    if (dtype != SPED_F64 && dtype != SPED_C128) fail(LS_INVALID_DATATYPE, "the emulation handles f64 and c128");
    if (dtype == SPED_F64 && !op->is_real()) fail(LS_OPERATOR_IS_COMPLEX, "operator is complex but a real datatype was requested");
    Problem pr(*op, n, reps, stab, world, rank);
    if (dtype == SPED_F64)
      run<double>(pr, static_cast<double const*>(x_global), static_cast<double*>(y_free), static_cast<double*>(y_all),
                  static_cast<double*>(y_phased), stats, ncols, static_cast<double*>(y_block));
    else
      run<double2>(pr, static_cast<double2 const*>(x_global), static_cast<double2*>(y_free), static_cast<double2*>(y_all),
                   static_cast<double2*>(y_phased), stats, ncols, static_cast<double2*>(y_block));
  });
}

namespace sped {
namespace {
template <class T>
void run_restart(uint64_t n, int m, int p, T* V, T* W, uint64_t ld, double2 const* C, double theta, double* out) {
  std::vector<double2> partial(1 + (size_t)p + 1, make_double2(0, 0)), partial2(2, make_double2(0, 0));
  if (m <= 4) restart_residual_kernel<T, 4>(V, W, ld, m, p, C, theta, n, partial.data());
  else restart_residual_kernel<T, 8>(V, W, ld, m, p, C, theta, n, partial.data());
  for (int j = 0; j < 1 + p; ++j) {  // finish_dot_kernel with one block: the partial is the sum
    out[2 * j] = partial[j].x;
    out[2 * j + 1] = partial[j].y;
  }
  T* w = V + (uint64_t)p * ld;
  axpy_norm_kernel<T>(V, ld, p, partial.data() + 1, w, n, partial2.data());
  out[2 * (1 + p)] = partial2[0].x;
  double kept = 0;
  int flag = 0;
  scale_rel_kernel<T>(w, n, partial2.data(), partial.data(), 1e-24, &kept, &flag);
  out[2 * (1 + p) + 1] = kept;
  out[2 * (1 + p) + 2] = (double)flag;
}
}  // namespace
}  // namespace sped

extern "C" int sped_selftest_emulate_restart(int dtype, uint64_t n, int m, int p, void* V, void* W, uint64_t ld,
                                             double const* C_re_im, double theta, double* out) {
  return guard([&] {
    if (m < 2 || m > 8 || p < 1 || p >= m) fail(LS_INVALID_ARGUMENT, "emulated restart: 2 <= m <= 8, 1 <= p < m");
    std::vector<double2> C((size_t)m * p);
    for (size_t i = 0; i < C.size(); ++i) C[i] = make_double2(C_re_im[2 * i], C_re_im[2 * i + 1]);
    switch (dtype) {
      case SPED_F32: run_restart<float>(n, m, p, static_cast<float*>(V), static_cast<float*>(W), ld, C.data(), theta, out); break;
      case SPED_F64: run_restart<double>(n, m, p, static_cast<double*>(V), static_cast<double*>(W), ld, C.data(), theta, out); break;
      case SPED_C64: run_restart<float2>(n, m, p, static_cast<float2*>(V), static_cast<float2*>(W), ld, C.data(), theta, out); break;
      case SPED_C128: run_restart<double2>(n, m, p, static_cast<double2*>(V), static_cast<double2*>(W), ld, C.data(), theta, out); break;
      default: fail(LS_INVALID_DATATYPE, "unknown datatype tag");
    }
  });
}
