// operator.cu -- host side of the matrix-free operator and the statically compiled kernels
// (K2 diagonal, K3 matvec with the interpreted canonicalisation program, K5 expectation).
//
// Replaces ls_create_interaction{1..4} / ls_create_operator / ls_operator_matmat /
// ls_operator_expectation of liblattice_symmetries as called from
// /root/reference/src/SpinED/Internal.hs:260-270,371-381,411-452.  The device code of the matvec
// lives in matvec_kernel.cuh; for symmetric sectors the kernel is normally the run-time
// specialised one from jit.cpp, the interpreted one below is the same algorithm with the group
// program read from shared memory.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "canon.cuh"

namespace sped {

ProgramView<u32> program_view32(Basis const& b, size_t& smem, bool& staged);
ProgramView<u64> program_view64(Basis const& b, size_t& smem, bool& staged);
void* jit_matvec_kernel(Basis& b, int dtype, int nb, double images);  // jit.cpp; nullptr = not available (or not worth waiting for)

namespace {

static_assert(sizeof(DevBond) == 16, "bond records are staged as 16-byte words");

// SYM: non-trivial group (canonicalise with the interpreted program).  NB: columns per pass.
template <class W, class T, int NB, bool SYM>
__global__ void __launch_bounds__(kThreads) matvec_kernel(MatvecParams p, ProgramView<W> prog) {
  constexpr bool CPLX = Traits<T>::cplx;
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<CPLX>(p.terms, smem);
  if constexpr (SYM) {
    ProgramCanon<W> canon{stage_program<W>(prog, smem + terms_smem_bytes(p.terms, CPLX))};
    matvec_rows<T, NB>(p, terms, canon);
  } else {
    matvec_rows<T, NB>(p, terms, TrivialCanon());
  }
}

// Counts off-diagonal transitions whose target is in the basis (E of SURVEY 8d).
template <class W, bool SYM>
__global__ void __launch_bounds__(kThreads) count_kernel(MatvecParams p, ProgramView<W> prog) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<false>(p.terms, smem);
  ProgramView<W> P = prog;
  if (SYM) P = stage_program<W>(prog, smem + terms_smem_bytes(p.terms, false));
  BasisIndex const ix = p.ctx.index;
  u64 const n_local = p.ctx.dist.n_local;
  unsigned long long mine = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (u64)gridDim.x * blockDim.x) {
    u64 const row = dist_local_to_global(p.ctx.dist, i);
    u64 const r = ix.direct ? row : __ldg(ix.reps + row);
    for_each_transition(terms, r, [&](DevBond const&, u32, u32, u64 rp) {
      u64 rep = rp;
      if (SYM) {
        W rw;
        u32 step, flipped;
        canonicalize<W>(P, (W)rp, rw, step, flipped);
        rep = rw;
      }
      if (lookup_index(ix, rep) != ~0ull) ++mine;
    });
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(p.counter, mine);
}

// K2: diagonal of the operator on the local rows.
__global__ void __launch_bounds__(kThreads) diagonal_kernel(RowContext ctx, TermsView terms_g, double* diag_re,
                                                            double* diag_im) {
  extern __shared__ __align__(16) unsigned char smem[];
  TermsView terms = stage_terms<true>(terms_g, smem);
  u64 const n_local = ctx.dist.n_local;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (u64)gridDim.x * blockDim.x) {
    u64 const row = dist_local_to_global(ctx.dist, i);
    u64 const r = ctx.index.direct ? row : __ldg(ctx.index.reps + row);
    double re = 0, im = 0;
    for (u32 bnd = 0; bnd < terms.n_bonds; ++bnd) {
      DevBond const bd = terms.bonds[bnd];
      u32 a = 0;
      for (u32 j = 0; j < bd.k; ++j) a |= (u32)((r >> ((bd.sites >> (8 * j)) & 0xffu)) & 1ull) << (bd.k - 1 - j);
      u32 const dim = 1u << bd.k;
      re += terms.pool_re[bd.moff + a * dim + a];
      im += terms.pool_im[bd.moff + a * dim + a];
    }
    diag_re[i] = re;
    if (diag_im) diag_im[i] = im;
  }
}

// per-block partial sums of conj(x) * y for each column (deterministic two-stage reduction)
template <class T>
__global__ void __launch_bounds__(kThreads) dot_partial_kernel(T const* x, u64 xs, u64 x_row0, T const* y, u64 ys,
                                                               u64 n, u32 ncols, double2* partial) {
  using TR = Traits<T>;
  __shared__ double2 red[kThreads / 32];
  for (u32 c = 0; c < ncols; ++c) {
    double sr = 0, si = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
      auto a = TR::load(x + (u64)c * xs + x_row0 + i);
      auto b = TR::load(y + (u64)c * ys + i);
      if constexpr (TR::cplx) {
        sr += a.x * b.x + a.y * b.y;
        si += a.x * b.y - a.y * b.x;
      } else {
        sr += a * b;
      }
    }
    for (int o = 16; o; o >>= 1) {
      sr += __shfl_down_sync(0xffffffffu, sr, o);
      si += __shfl_down_sync(0xffffffffu, si, o);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_double2(sr, si);
    __syncthreads();
    if (threadIdx.x == 0) {
      double2 s = make_double2(0, 0);
      for (int w = 0; w < kThreads / 32; ++w) {
        s.x += red[w].x;
        s.y += red[w].y;
      }
      partial[(u64)c * gridDim.x + blockIdx.x] = s;
    }
    __syncthreads();
  }
}

struct PackedTerms {
  std::vector<DevBond> bonds;
  std::vector<double> pool_re, pool_im;
  std::vector<std::uint16_t> masks;
};

bool overlap_enabled() {
  char const* e = std::getenv("SPED_OVERLAP");
  return !(e && e[0] == '0');
}

bool bond_order_descending() {
  char const* e = std::getenv("SPED_BOND_ORDER");
  return !(e && e[0] == '0');
}

PackedTerms pack_terms(std::vector<Interaction> const& terms) {
  PackedTerms pk;
  for (auto const& t : terms) {
    u32 dim = 1u << t.k;
    size_t moff = pk.pool_re.size(), zoff = pk.masks.size();
    if (moff + dim * dim > 65535 || zoff + dim > 65535) fail(LS_INVALID_ARGUMENT, "too many distinct interaction matrices");
    for (u32 a = 0; a < dim; ++a) {
      std::uint16_t m = 0;
      for (u32 b = 0; b < dim; ++b) {
        cplx v = t.matrix[a * dim + b];
        pk.pool_re.push_back(v.real());
        pk.pool_im.push_back(v.imag());
        if (a != b && v != cplx(0, 0)) m |= (std::uint16_t)(1u << b);
      }
      pk.masks.push_back(m);
    }
    size_t count = t.sites.size() / t.k;
    for (size_t s = 0; s < count; ++s) {
      DevBond b{};
      b.k = (u32)t.k;
      for (int j = 0; j < t.k; ++j) b.sites |= (u32)t.sites[s * t.k + j] << (8 * j);
      b.moff = (std::uint16_t)moff;
      b.zoff = (std::uint16_t)zoff;
      pk.bonds.push_back(b);
    }
  }
  // Bonds are visited from the most significant site down.  Rows are sorted representatives, so
  // the 32 rows of a warp share their high bits: visiting the high bonds first keeps the lanes on
  // the same bond -- and, when the flipped word is already canonical, on neighbouring entries of x
  // -- for as many elements as possible (measured on a 28-site chain: 0.47 instead of 0.60
  // 32-byte sectors per gathered element).  The order only fixes the summation order; it is the
  // same for the matrix-free and the cached path.
  if (bond_order_descending()) {
    auto top = [](DevBond const& b) {
      u32 m = 0;
      for (u32 j = 0; j < b.k; ++j) m = std::max(m, (b.sites >> (8 * j)) & 0xffu);
      return m;
    };
    std::stable_sort(pk.bonds.begin(), pk.bonds.end(), [&](DevBond const& a, DevBond const& b) { return top(a) > top(b); });
  }
  return pk;
}

// layout of Operator::d_terms: [bonds][pool_re][pool_im][masks]
TermsView terms_view(Operator const& op, u32 n_bonds, u32 pool, u32 masks) {
  auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
  unsigned char* base = op.d_terms.ptr;
  TermsView v;
  v.bonds = reinterpret_cast<DevBond const*>(base);
  size_t off = up((size_t)n_bonds * sizeof(DevBond));
  v.pool_re = reinterpret_cast<double const*>(base + off);
  off += up((size_t)pool * 8);
  v.pool_im = reinterpret_cast<double const*>(base + off);
  off += up((size_t)pool * 8);
  v.masks = reinterpret_cast<std::uint16_t const*>(base + off);
  v.n_bonds = n_bonds;
  v.pool_size = pool;
  v.mask_size = masks;
  return v;
}

struct OpShape {
  u32 n_bonds = 0, pool = 0, masks = 0;
};
OpShape shape_of(Operator const& op) {
  OpShape s;
  for (auto const& t : op.terms) {
    u32 dim = 1u << t.k;
    s.n_bonds += (u32)(t.sites.size() / t.k);
    s.pool += dim * dim;
    s.masks += dim;
  }
  return s;
}

template <class K>
void set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024)
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

template <class W, class T, int NB, bool SYM>
void launch_matvec_t(MatvecParams const& p, ProgramView<W> prog, size_t smem, int grid, cudaStream_t s) {
  auto k = matvec_kernel<W, T, NB, SYM>;
  set_smem(k, smem);
  k<<<grid, kThreads, smem, s>>>(p, prog);
  KERNEL_LAUNCHED();
}

template <class T>
void launch_matvec(Operator& op, MatvecParams p, int dtype, u64 block, u64 xs, u64 ys, cudaStream_t s) {
  Basis& b = *op.basis;
  bool const sym = !b.trivial();
  constexpr bool CPLX = Traits<T>::cplx;
  size_t tsm = terms_smem_bytes(p.terms, CPLX);
  u64 n_local = p.ctx.dist.n_local;
  int grid = persistent_grid(n_local, kThreads, 8);
  T const* x = static_cast<T const*>(p.x);
  T* y = static_cast<T*>(p.y);
  for (u64 c0 = 0; c0 < block;) {
    u64 left = block - c0;
    p.x = x + c0 * xs;
    p.y = y + c0 * ys;
    bool wide = left > 1;
    p.ncols = (u32)std::min<u64>(left, wide ? 4 : 1);
    // canonicalisation steps of this launch (about half of the bonds of a row have a transition)
    double const images = (double)n_local * (double)b.group_order() * 0.5 * (double)p.terms.n_bonds;
    void* jit = sym ? jit_matvec_kernel(b, dtype, wide ? 4 : 1, images) : nullptr;
    if (jit) {
      // run-time specialised kernel: same device code, canonicalisation emitted as straight-line code
      void* args[] = {&p};
      if (tsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(jit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
      CUDA_CHECK(cudaLaunchKernel(jit, dim3(grid), dim3(kThreads), args, tsm, s));
      KERNEL_LAUNCHED();
    } else if (!sym) {
      ProgramView<u64> none{};
      if (wide) launch_matvec_t<u64, T, 4, false>(p, none, tsm, grid, s);
      else launch_matvec_t<u64, T, 1, false>(p, none, tsm, grid, s);
    } else if (b.use32()) {
      size_t psm; bool staged;
      auto prog = program_view32(b, psm, staged);
      if (wide) launch_matvec_t<u32, T, 4, true>(p, prog, tsm + psm, grid, s);
      else launch_matvec_t<u32, T, 1, true>(p, prog, tsm + psm, grid, s);
    } else {
      size_t psm; bool staged;
      auto prog = program_view64(b, psm, staged);
      if (wide) launch_matvec_t<u64, T, 4, true>(p, prog, tsm + psm, grid, s);
      else launch_matvec_t<u64, T, 1, true>(p, prog, tsm + psm, grid, s);
    }
    c0 += p.ncols;
  }
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace

bool Interaction::is_real() const {
  for (auto const& v : matrix)
    if (v.imag() != 0.0) return false;
  return true;
}

std::shared_ptr<Interaction> make_interaction(int k, void const* matrix, unsigned count, std::uint16_t const* sites) {
  if (k < 1 || k > 4) fail(LS_INVALID_ARGUMENT, "only 1- to 4-site interactions are supported");
  auto t = std::make_shared<Interaction>();
  t->k = k;
  size_t dim = (size_t)1 << k;
  auto const* m = static_cast<double const*>(matrix);
  t->matrix.resize(dim * dim);
  for (size_t i = 0; i < dim * dim; ++i) t->matrix[i] = cplx(m[2 * i], m[2 * i + 1]);
  t->sites.assign(sites, sites + (size_t)count * k);
  for (size_t s = 0; s < count; ++s)
    for (int i = 0; i < k; ++i)
      for (int j = i + 1; j < k; ++j)
        if (t->sites[s * k + i] == t->sites[s * k + j]) fail(LS_INVALID_ARGUMENT, "an interaction tuple repeats a site");
  return t;
}

std::shared_ptr<Operator> make_operator(std::shared_ptr<Basis> b, std::vector<Interaction const*> const& terms) {
  auto op = std::make_shared<Operator>();
  for (auto* t : terms) {
    for (auto s : t->sites)
      if (s >= b->n_spins) fail(LS_INVALID_ARGUMENT, "interaction site index out of range");
    op->terms.push_back(*t);
    op->real_matrices = op->real_matrices && t->is_real();
    size_t dim = (size_t)1 << t->k;
    for (size_t a = 0; a < dim; ++a)
      if (t->matrix[a * dim + a].imag() != 0.0) op->real_diagonal = false;
  }
  op->basis = std::move(b);
  return op;
}

void Operator::prepare() {
  std::lock_guard<std::mutex> lock(mutex);
  Basis& b = *basis;
  if (!b.built) fail(LS_CACHE_NOT_BUILT, "basis has not been built (call ls_build first)");
  if (prepared_generation == b.generation) return;
  b.ensure_device_tables();
  PackedTerms pk = pack_terms(terms);
  auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t off_re = up(pk.bonds.size() * sizeof(DevBond));
  size_t off_im = off_re + up(pk.pool_re.size() * 8);
  size_t off_mask = off_im + up(pk.pool_im.size() * 8);
  size_t total = off_mask + up(pk.masks.size() * 2);
  std::vector<unsigned char> img(std::max<size_t>(total, 16), 0);
  if (!pk.bonds.empty()) std::memcpy(img.data(), pk.bonds.data(), pk.bonds.size() * sizeof(DevBond));
  if (!pk.pool_re.empty()) {
    std::memcpy(img.data() + off_re, pk.pool_re.data(), pk.pool_re.size() * 8);
    std::memcpy(img.data() + off_im, pk.pool_im.data(), pk.pool_im.size() * 8);
    std::memcpy(img.data() + off_mask, pk.masks.data(), pk.masks.size() * 2);
  }
  d_terms.upload(img);
  dist = b.dist();
  u64 n_local = dist.n_local;
  d_diag.alloc(std::max<u64>(1, n_local * (real_diagonal ? 1 : 2)));
  OpShape sh = shape_of(*this);
  TermsView tv = terms_view(*this, sh.n_bonds, sh.pool, sh.masks);
  RowContext ctx{b.index, b.d_norm_table.ptr, b.d_chi_table.ptr, dist};
  size_t smem = terms_smem_bytes(tv, true);
  set_smem(diagonal_kernel, smem);
  if (n_local) {
    diagonal_kernel<<<persistent_grid(n_local, kThreads, 8), kThreads, smem>>>(
        ctx, tv, d_diag.ptr, real_diagonal ? nullptr : d_diag.ptr + n_local);
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
  }
  CUDA_CHECK(cudaDeviceSynchronize());
  counted = false;
  drop_cache();
  prepared_generation = b.generation;
}

static MatvecParams make_params(Operator& op);
MatvecParams operator_params(Operator& op) { return make_params(op); }

static MatvecParams make_params(Operator& op) {
  Basis& b = *op.basis;
  OpShape sh = shape_of(op);
  MatvecParams p{};
  p.ctx = RowContext{b.index, b.d_norm_table.ptr, b.d_chi_table.ptr, op.dist};
  p.terms = terms_view(op, sh.n_bonds, sh.pool, sh.masks);
  u64 n_local = op.dist.n_local;
  p.diag_re = op.d_diag.ptr;
  p.diag_im = op.real_diagonal ? nullptr : op.d_diag.ptr + n_local;
  return p;
}

void Operator::matmat_device(int dtype, u64 block, void const* x, u64 xs, void* y, u64 ys, cudaStream_t s) {
  SPED_NVTX("sped: matmat (device buffers)");
  prepare();
  if (!dtype_is_complex(dtype) && !is_real()) fail(LS_OPERATOR_IS_COMPLEX, "operator is complex but a real datatype was requested");
  if (block == 0 || dist.n_local == 0) return;
  // steady state: the elements found by the first matrix-free pass are resident in HBM
  if (cache_usable()) {
    cached_matmat(dtype, block, x, xs, y, ys, s);
    return;
  }
  MatvecParams p = make_params(*this);
  p.x = x;
  p.y = y;
  p.xs = xs;
  p.ys = ys;
  switch (dtype) {
    case SPED_F32: launch_matvec<float>(*this, p, dtype, block, xs, ys, s); break;
    case SPED_F64: launch_matvec<double>(*this, p, dtype, block, xs, ys, s); break;
    case SPED_C64: launch_matvec<float2>(*this, p, dtype, block, xs, ys, s); break;
    case SPED_C128: launch_matvec<double2>(*this, p, dtype, block, xs, ys, s); break;
    default: fail(LS_INVALID_DATATYPE, "unknown datatype tag");
  }
}

void packed_terms_host(std::vector<Interaction> const& terms, std::vector<DevBond>& bonds, std::vector<double>& pool_re,
                       std::vector<double>& pool_im, std::vector<std::uint16_t>& masks) {
  PackedTerms pk = pack_terms(terms);
  bonds = pk.bonds;
  pool_re = pk.pool_re;
  pool_im = pk.pool_im;
  masks = pk.masks;
}

// One column, multi-rank: the all-gather of the Krylov vector runs on its own stream while the
// streaming kernel already handles the elements whose source entries this rank owns; the remote
// class follows once the gather has landed.  (Matrix-free mode: gather, then one kernel.)
void Operator::matvec_sharded(int dtype, void const* x_local, void* y_local, void* xfull, cudaStream_t s) {
  SPED_NVTX("sped: sharded matvec (exchange + class passes)");
  prepare();
  Comm& cm = comm();
  size_t const es = dtype_size(dtype);
  u64 const n_local = dist.n_local, padded = dist.chunk * dist.world;
  unsigned char* mine = static_cast<unsigned char*>(xfull) + (u64)dist.rank * dist.chunk * es;
  if (n_local && x_local != mine)
    CUDA_CHECK(cudaMemcpyAsync(mine, x_local, n_local * es, cudaMemcpyDeviceToDevice, s));
  if (!cm.active()) {
    matmat_device(dtype, 1, xfull, padded, y_local, std::max<u64>(n_local, 1), s);
    return;
  }
  // Which collectives run must not depend on anything rank-local (a rank without rows, or one whose
  // cache did not fit, still has to take part in the same exchange): only on world and environment.
  if (!overlap_enabled()) {
    comm_allgather_inplace(xfull, dist.chunk * es, s);
    matmat_device(dtype, 1, xfull, padded, y_local, std::max<u64>(n_local, 1), s);
    return;
  }
  bool const rounds = exchange_rounds(dist.world, dist.chunk) == 2;
  bool const cached = n_local > 0 && cache_usable() && c_rounds == (rounds ? 2u : 1u);
  // SPED_OVERLAP_TRACE=n: device times of the first n overlapped matvecs
  static int trace_left = [] {
    char const* e = std::getenv("SPED_OVERLAP_TRACE");
    return e && *e ? std::atoi(e) : 0;
  }();
  bool const trace = trace_left > 0;
  cudaEvent_t t[9] = {};
  if (trace)
    for (auto& e : t) CUDA_CHECK(cudaEventCreate(&e));
  auto mark = [&](int k, cudaStream_t st) {
    if (trace) CUDA_CHECK(cudaEventRecord(t[k], st));
  };
  cudaStream_t const g = cm.gather_stream;
  // exchange: one all-gather, or two rounds of peer groups (near ranks first) -- NCCL kernels, or the
  // copy engines pulling the peers' published shards over peer memory (comm.cpp; long shards)
  // (comm_ce_prepare is collective and does something only when the send buffers must grow)
  bool const ce = comm_ce_wanted(dist.chunk * es) && comm_ce_prepare(dist.chunk * es);
  if (ce) comm_ce_publish(x_local, n_local * es, s);
  CUDA_CHECK(cudaEventRecord(cm.ev_ready, s));
  CUDA_CHECK(cudaStreamWaitEvent(g, cm.ev_ready, 0));
  mark(0, g);
  if (ce) comm_ce_barrier(g);
  if (rounds) {
    // (near comes from the world size, not from this rank's cache: a rank without rows or without a
    // cache has to send and receive in the same rounds as everybody else)
    int const near = (int)exchange_near(dist.world, dist.chunk);
    if (ce) comm_ce_pull_round(xfull, dist.chunk * es, 1, near, g);
    else comm_exchange_round(xfull, dist.chunk * es, 1, near, g);
    CUDA_CHECK(cudaEventRecord(cm.ev_round1, g));
    mark(1, g);
    if (ce) comm_ce_pull_round(xfull, dist.chunk * es, near + 1, (int)dist.world - 1, g);
    else comm_exchange_round(xfull, dist.chunk * es, near + 1, (int)dist.world - 1, g);
  } else {
    if (ce) comm_ce_pull_round(xfull, dist.chunk * es, 1, (int)dist.world - 1, g);
    else comm_allgather_inplace(xfull, dist.chunk * es, g);
    mark(1, g);
  }
  if (ce) comm_ce_advance();
  CUDA_CHECK(cudaEventRecord(cm.ev_gathered, g));
  mark(2, g);
  if (!cached) {  // no rows, or matrix-free mode: wait for the whole vector, one kernel
    CUDA_CHECK(cudaStreamWaitEvent(s, cm.ev_gathered, 0));
    if (n_local) matmat_device(dtype, 1, xfull, padded, y_local, n_local, s);
    if (trace) {
      for (auto e : t) cudaEventDestroy(e);
      --trace_left;
    }
    return;
  }
  // product: the pass over class c runs beside the transfer that class c + 1 waits for
  mark(3, s);
  cached_matmat(dtype, 1, xfull, padded, y_local, n_local, s, 0, ~(u64)0, 1, true);
  mark(4, s);
  if (rounds) {
    CUDA_CHECK(cudaStreamWaitEvent(s, cm.ev_round1, 0));
    mark(5, s);
    cached_matmat(dtype, 1, xfull, padded, y_local, n_local, s, 0, ~(u64)0, 2, true);
    mark(6, s);
  }
  CUDA_CHECK(cudaStreamWaitEvent(s, cm.ev_gathered, 0));
  mark(7, s);
  cached_matmat(dtype, 1, xfull, padded, y_local, n_local, s, 0, ~(u64)0, 1 + (int)c_rounds, false);
  mark(8, s);
  if (trace) {
    CUDA_CHECK(cudaStreamSynchronize(s));
    CUDA_CHECK(cudaStreamSynchronize(g));
    auto ms = [&](int a, int b) {
      float v = 0;
      cudaEventElapsedTime(&v, t[a], t[b]);
      return v;
    };
    if (rounds)
      std::fprintf(stderr,
                   "[sped] rank %d overlapped matvec: exchange round 1 %.3f ms, round 2 %.3f ms | local pass %.3f ms, wait %.3f, "
                   "near pass %.3f ms, wait %.3f, far pass %.3f ms | total %.3f ms\n",
                   cm.rank, ms(0, 1), ms(1, 2), ms(3, 4), ms(4, 5), ms(5, 6), ms(6, 7), ms(7, 8), ms(3, 8));
    else
      std::fprintf(stderr,
                   "[sped] rank %d overlapped matvec: gather %.3f ms | local pass %.3f ms, wait %.3f ms, remote pass %.3f ms | "
                   "total %.3f ms\n",
                   cm.rank, ms(0, 2), ms(3, 4), ms(4, 7), ms(7, 8), ms(3, 8));
    for (auto e : t) cudaEventDestroy(e);
    --trace_left;
  }
}

void Operator::count_elements(u64& rows, u64& offdiag) {
  prepare();
  Basis& b = *basis;
  rows = b.n_states;
  if (!counted) {
    DeviceBuffer<unsigned long long> d_count(1);
    CUDA_CHECK(cudaMemset(d_count.ptr, 0, 8));
    MatvecParams p = make_params(*this);
    p.counter = d_count.ptr;
    u64 n_local = dist.n_local;
    if (cache_ready) {
      cached_count(d_count.ptr);  // the cache holds exactly the elements the matrix-free pass found
    } else if (n_local) {
      size_t tsm = terms_smem_bytes(p.terms, false);
      int grid = persistent_grid(n_local, kThreads, 8);
      if (b.trivial()) {
        set_smem(count_kernel<u64, false>, tsm);
        count_kernel<u64, false><<<grid, kThreads, tsm>>>(p, ProgramView<u64>{});
      } else if (b.use32()) {
        size_t psm; bool staged;
        auto prog = program_view32(b, psm, staged);
        set_smem(count_kernel<u32, true>, tsm + psm);
        count_kernel<u32, true><<<grid, kThreads, tsm + psm>>>(p, prog);
      } else {
        size_t psm; bool staged;
        auto prog = program_view64(b, psm, staged);
        set_smem(count_kernel<u64, true>, tsm + psm);
        count_kernel<u64, true><<<grid, kThreads, tsm + psm>>>(p, prog);
      }
      KERNEL_LAUNCHED();
      CUDA_CHECK(cudaGetLastError());
    }
    comm_allreduce_sum_u64(d_count.ptr, 1, nullptr);
    unsigned long long c = 0;
    CUDA_CHECK(cudaMemcpy(&c, d_count.ptr, 8, cudaMemcpyDeviceToHost));
    n_offdiag = c;
    counted = true;
  }
  offdiag = n_offdiag;
}

namespace {

// Conversion between global row order and the replicated [rank][local] layout (device_types.h).
template <class E>
__global__ void __launch_bounds__(kThreads) to_dist_kernel(E const* src, E* dst, RowDist d) {
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < d.n; g += (u64)gridDim.x * blockDim.x)
    dst[dist_global_to_pos(d, g)] = src[g];
}
template <class E>
__global__ void __launch_bounds__(kThreads) from_dist_kernel(E const* src, E* dst, RowDist d) {
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < d.n; g += (u64)gridDim.x * blockDim.x)
    dst[g] = src[dist_global_to_pos(d, g)];
}

void convert_layout(bool to_dist, size_t es, void const* src, void* dst, RowDist const& d, cudaStream_t s) {
  int grid = persistent_grid(std::max<u64>(d.n, 1), kThreads, 8);
  if (es == 4) {
    if (to_dist) to_dist_kernel<u32><<<grid, kThreads, 0, s>>>((u32 const*)src, (u32*)dst, d);
    else from_dist_kernel<u32><<<grid, kThreads, 0, s>>>((u32 const*)src, (u32*)dst, d);
  } else if (es == 8) {
    if (to_dist) to_dist_kernel<u64><<<grid, kThreads, 0, s>>>((u64 const*)src, (u64*)dst, d);
    else from_dist_kernel<u64><<<grid, kThreads, 0, s>>>((u64 const*)src, (u64*)dst, d);
  } else {
    if (to_dist) to_dist_kernel<double2><<<grid, kThreads, 0, s>>>((double2 const*)src, (double2*)dst, d);
    else from_dist_kernel<double2><<<grid, kThreads, 0, s>>>((double2 const*)src, (double2*)dst, d);
  }
  KERNEL_LAUNCHED();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace

// Host-pointer entry (the reference's PRIMME callback shape): stage x to the device, apply, copy
// back.  With several ranks every rank passes the full x and receives the full y.
void Operator::matmat_host(int dtype, u64 size, u64 block, void const* x, u64 xs, void* y, u64 ys) {
  SPED_NVTX("sped: ls_operator_matmat (host buffers)");
  prepare();
  Basis& b = *basis;
  if (size != b.n_states) fail(LS_DIMENSION_MISMATCH, "size does not match the number of representatives");
  if (xs < size || ys < size) fail(LS_INVALID_ARGUMENT, "strides must be at least size");
  if (block == 0 || size == 0) return;
  size_t es = dtype_size(dtype);
  Comm& cm = comm();
  u64 padded = dist.chunk * dist.world;
  // staging buffers are kept between calls (PRIMME calls this once per block per iteration)
  if (cm.active()) {
    // Several ranks (every rank is handed the full x and wants the full y): a rank uploads only the
    // rows it owns -- its blocks of the block-cyclic distribution are a strided 2-D copy -- the shards
    // are exchanged over NVLink inside the sharded matvec, the result shards are all-gathered, and
    // the [rank][local] layout is put back into global row order by the copy engines on the way out
    // (one strided 2-D copy per owner).  PCIe carries N/P entries in and N out per rank.
    if (stage_y.count < 2 * padded * es) stage_y.alloc(2 * padded * es);
    unsigned char* xp = stage_y.ptr;
    unsigned char* yp = stage_y.ptr + padded * es;
    cudaStream_t st = cm.stream;
    u64 const B = (u64)1 << dist.log2b, full = size >> dist.log2b, rem = size & (B - 1);
    auto blocks_of = [&](u32 r) { return full / dist.world + (r < full % dist.world ? 1 : 0); };
    for (u64 c = 0; c < block; ++c) {
      unsigned char const* xc = static_cast<unsigned char const*>(x) + c * xs * es;
      unsigned char* yc = static_cast<unsigned char*>(y) + c * ys * es;
      unsigned char* mine = xp + (u64)dist.rank * dist.chunk * es;
      u64 const nb = blocks_of(dist.rank);
      if (nb)
        CUDA_CHECK(cudaMemcpy2DAsync(mine, B * es, xc + (u64)dist.rank * B * es, (u64)dist.world * B * es, B * es, nb,
                                     cudaMemcpyHostToDevice, st));
      if (rem && dist.rank == full % dist.world)
        CUDA_CHECK(cudaMemcpyAsync(mine + nb * B * es, xc + full * B * es, rem * es, cudaMemcpyHostToDevice, st));
      matvec_sharded(dtype, mine, yp + (u64)dist.rank * dist.chunk * es, xp, st);
      comm_allgather_inplace(yp, dist.chunk * es, st);
      for (u32 r = 0; r < dist.world; ++r) {
        u64 const nbr = blocks_of(r);
        unsigned char const* src = yp + (u64)r * dist.chunk * es;
        if (nbr)
          CUDA_CHECK(cudaMemcpy2DAsync(yc + (u64)r * B * es, (u64)dist.world * B * es, src, B * es, B * es, nbr,
                                       cudaMemcpyDeviceToHost, st));
        if (rem && r == full % dist.world)
          CUDA_CHECK(cudaMemcpyAsync(yc + full * B * es, src + nbr * B * es, rem * es, cudaMemcpyDeviceToHost, st));
      }
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    return;
  }
  if (stage_x.count < size * block * es) stage_x.alloc(size * block * es);
  if (stage_y.count < size * block * es) stage_y.alloc(size * block * es);
  DeviceBuffer<unsigned char>&dx = stage_x, &dy = stage_y;
  // column by column: a 2-D copy would need a pitch of xs * es bytes, which CUDA limits to 2^31 - 1
  for (u64 c = 0; c < block; ++c)
    CUDA_CHECK(cudaMemcpyAsync(dx.ptr + c * size * es, static_cast<unsigned char const*>(x) + c * xs * es, size * es,
                               cudaMemcpyHostToDevice, nullptr));
  // from pageable memory the copy returns once the data is staged, possibly before the last DMA into
  // dx has finished; the kernels below run on non-blocking streams, which the legacy stream does not
  // order -- so wait for it explicitly
  CUDA_CHECK(cudaStreamSynchronize(nullptr));
  constexpr int kPipe = 8;
  u64 const rows_per = ((size + kPipe - 1) / kPipe + 31) & ~(u64)31;
  if (cache_usable() && block == 1 && size >= (u64)1 << 20) {
    // steady state: the streaming kernel runs row chunk by row chunk and the device-to-host copy
    // of a finished chunk overlaps the kernel of the next one
    if (!pipe_compute) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&pipe_compute, cudaStreamNonBlocking));
      CUDA_CHECK(cudaStreamCreateWithFlags(&pipe_copy, cudaStreamNonBlocking));
      for (auto& e : pipe_events) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // all kernels are queued first: a copy into pageable memory blocks the host, and must not
    // keep the next chunk's kernel from being launched
    int chunks = 0;
    for (int c = 0; c < kPipe; ++c) {
      u64 lo = std::min<u64>(size, (u64)c * rows_per), hi = std::min<u64>(size, lo + rows_per);
      if (lo >= hi) break;
      cached_matmat(dtype, 1, dx.ptr, size, dy.ptr, size, pipe_compute, lo, hi);
      CUDA_CHECK(cudaEventRecord(pipe_events[c], pipe_compute));
      chunks = c + 1;
    }
    for (int c = 0; c < chunks; ++c) {
      u64 lo = std::min<u64>(size, (u64)c * rows_per), hi = std::min<u64>(size, lo + rows_per);
      CUDA_CHECK(cudaStreamWaitEvent(pipe_copy, pipe_events[c], 0));
      CUDA_CHECK(cudaMemcpyAsync(static_cast<unsigned char*>(y) + lo * es, dy.ptr + lo * es, (hi - lo) * es,
                                 cudaMemcpyDeviceToHost, pipe_copy));
    }
    CUDA_CHECK(cudaStreamSynchronize(pipe_copy));
    return;
  }
  matmat_device(dtype, block, dx.ptr, size, dy.ptr, size, nullptr);
  CUDA_CHECK(cudaDeviceSynchronize());
  for (u64 c = 0; c < block; ++c)
    CUDA_CHECK(cudaMemcpy(static_cast<unsigned char*>(y) + c * ys * es, dy.ptr + c * size * es, size * es, cudaMemcpyDeviceToHost));
}

// Row-sharded host-pointer entry: this rank's rows in, this rank's rows out (see sped.h).
void Operator::matmat_host_local(int dtype, u64 block, void const* x, u64 xs, void* y, u64 ys) {
  prepare();
  Comm& cm = comm();
  if (!cm.active()) {
    matmat_host(dtype, basis->n_states, block, x, xs, y, ys);
    return;
  }
  if (!dtype_is_complex(dtype) && !is_real()) fail(LS_OPERATOR_IS_COMPLEX, "operator is complex but a real datatype was requested");
  u64 const n_local = dist.n_local, padded = dist.chunk * dist.world;
  if (xs < n_local || ys < n_local) fail(LS_INVALID_ARGUMENT, "strides must be at least the number of local rows");
  if (block == 0) return;
  size_t const es = dtype_size(dtype);
  if (stage_y.count < 2 * padded * es) stage_y.alloc(2 * padded * es);
  unsigned char* xp = stage_y.ptr;
  unsigned char* yl = stage_y.ptr + padded * es;  // local rows of the result
  unsigned char* mine = xp + (u64)dist.rank * dist.chunk * es;
  cudaStream_t st = cm.stream;
  for (u64 c = 0; c < block; ++c) {
    if (n_local)
      CUDA_CHECK(cudaMemcpyAsync(mine, static_cast<unsigned char const*>(x) + c * xs * es, n_local * es, cudaMemcpyHostToDevice, st));
    matvec_sharded(dtype, mine, yl, xp, st);
    if (n_local)
      CUDA_CHECK(cudaMemcpyAsync(static_cast<unsigned char*>(y) + c * ys * es, yl, n_local * es, cudaMemcpyDeviceToHost, st));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
}

template <class T>
static void launch_dot(void const* x, void const* y, u64 n, double2* partial, int grid) {
  dot_partial_kernel<T><<<grid, kThreads>>>((T const*)x, 0, 0, (T const*)y, 0, n, 1u, partial);
}

void Operator::expectation_host(int dtype, u64 size, u64 block, void const* x, u64 xs, cplx* out) {
  SPED_NVTX("sped: ls_operator_expectation");
  prepare();
  Basis& b = *basis;
  if (size != b.n_states) fail(LS_DIMENSION_MISMATCH, "size does not match the number of representatives");
  if (xs < size) fail(LS_INVALID_ARGUMENT, "stride must be at least size");
  for (u64 c = 0; c < block; ++c) out[c] = cplx(0, 0);
  if (block == 0 || size == 0) return;
  size_t es = dtype_size(dtype);
  u64 n_local = dist.n_local;
  u64 padded = dist.chunk * dist.world;
  DeviceBuffer<unsigned char> dx(size * es), xp(padded * es), dy(std::max<u64>(1, n_local) * es);
  int grid = persistent_grid(std::max<u64>(1, n_local), kThreads, 4);
  DeviceBuffer<double2> d_partial((size_t)grid);
  std::vector<double> sums(2 * block, 0.0);
  for (u64 c = 0; c < block; ++c) {
    CUDA_CHECK(cudaMemcpy(dx.ptr, static_cast<unsigned char const*>(x) + c * xs * es, size * es, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(xp.ptr, 0, padded * es));
    convert_layout(true, es, dx.ptr, xp.ptr, dist, nullptr);
    matmat_device(dtype, 1, xp.ptr, padded, dy.ptr, std::max<u64>(1, n_local), nullptr);
    void const* xl = xp.ptr + (u64)dist.rank * dist.chunk * es;  // this rank's rows of x
    switch (dtype) {
      case SPED_F32: launch_dot<float>(xl, dy.ptr, n_local, d_partial.ptr, grid); break;
      case SPED_F64: launch_dot<double>(xl, dy.ptr, n_local, d_partial.ptr, grid); break;
      case SPED_C64: launch_dot<float2>(xl, dy.ptr, n_local, d_partial.ptr, grid); break;
      case SPED_C128: launch_dot<double2>(xl, dy.ptr, n_local, d_partial.ptr, grid); break;
      default: fail(LS_INVALID_DATATYPE, "unknown datatype tag");
    }
    KERNEL_LAUNCHED();
    CUDA_CHECK(cudaGetLastError());
    auto partial = d_partial.download();
    for (int g = 0; g < grid; ++g) {
      sums[2 * c] += partial[g].x;
      sums[2 * c + 1] += partial[g].y;
    }
  }
  if (comm().active()) {
    DeviceBuffer<double> d(2 * block);
    d.upload(sums);
    comm_allreduce_sum_f64(d.ptr, 2 * block, comm().stream);
    CUDA_CHECK(cudaStreamSynchronize(comm().stream));
    sums = d.download();
  }
  for (u64 c = 0; c < block; ++c) out[c] = cplx(sums[2 * c], sums[2 * c + 1]);
}

}  // namespace sped
