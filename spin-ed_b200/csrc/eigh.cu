// eigh.cu -- device-resident block Davidson (GD+k style, no preconditioner) eigensolver (K4).
//
// Replaces the PRIMME call `eigh primmeOptions primmeOperator` of
// /root/reference/src/SpinED.hs:383-404: smallest `n_evals` eigenpairs of a Hermitian operator given
// only its block matvec.  Option semantics follow PRIMME's: max_basis_size bounds the search
// space, max_block_size the number of new directions per outer iteration, min_restart_size the
// number of Ritz vectors kept at a restart (plus the previous-iteration Ritz directions, which is
// what makes a 3-vector basis behave like locally-optimal CG -- the chain_40/42 decks use
// max_primme_basis_size 3/4).  All O(N) work is fused device kernels; the host only solves the
// projected (<= 64 x 64) problem.  Reductions are two-stage with a fixed order: deterministic.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>

#include "device_common.cuh"

namespace sped {

// ---------------------------------------------------------------------------------------------
// small dense Hermitian eigenproblem (cyclic Jacobi), ascending eigenvalues, columns = vectors
// ---------------------------------------------------------------------------------------------
void small_eigh(int m, std::vector<cplx> A /* row-major m x m */, std::vector<double>& evals, std::vector<cplx>& evecs) {
  std::vector<cplx> V((size_t)m * m, cplx(0, 0));
  for (int i = 0; i < m; ++i) V[(size_t)i * m + i] = 1.0;
  auto at = [&](int i, int j) -> cplx& { return A[(size_t)i * m + j]; };
  for (int i = 0; i < m; ++i) {
    at(i, i) = at(i, i).real();
    for (int j = i + 1; j < m; ++j) {
      cplx v = 0.5 * (at(i, j) + std::conj(at(j, i)));
      at(i, j) = v;
      at(j, i) = std::conj(v);
    }
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < m; ++i) {
      diag += std::norm(at(i, i));
      for (int j = i + 1; j < m; ++j) off += 2 * std::norm(at(i, j));
    }
    if (off <= 1e-32 * (diag + off) || off == 0) break;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) {
        double b = std::abs(at(p, q));
        if (b < 1e-300) continue;
        cplx e = at(p, q) / b;
        double a = at(p, p).real(), d = at(q, q).real();
        double tau = (d - a) / (2 * b);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1 + tau * tau));
        double c = 1 / std::sqrt(1 + t * t), s = t * c;
        cplx se = s * e, sce = s * std::conj(e);
        for (int i = 0; i < m; ++i) {
          cplx x = at(i, p), y = at(i, q);
          at(i, p) = x * c - y * sce;
          at(i, q) = x * se + y * c;
          cplx vx = V[(size_t)i * m + p], vy = V[(size_t)i * m + q];
          V[(size_t)i * m + p] = vx * c - vy * sce;
          V[(size_t)i * m + q] = vx * se + vy * c;
        }
        for (int j = 0; j < m; ++j) {
          cplx x = at(p, j), y = at(q, j);
          at(p, j) = c * x - se * y;
          at(q, j) = sce * x + c * y;
        }
        at(p, q) = 0;
        at(q, p) = 0;
        at(p, p) = at(p, p).real();
        at(q, q) = at(q, q).real();
      }
  }
  std::vector<int> order(m);
  for (int i = 0; i < m; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return at(x, x).real() < at(y, y).real(); });
  evals.resize(m);
  evecs.assign((size_t)m * m, cplx(0, 0));
  for (int k = 0; k < m; ++k) {
    evals[k] = at(order[k], order[k]).real();
    for (int i = 0; i < m; ++i) evecs[(size_t)i * m + k] = V[(size_t)i * m + order[k]];
  }
}

namespace {

constexpr int kMaxBasis = 64;
constexpr int kDotChunk = 8;

template <class T> struct VT;
template <> struct VT<float> {
  using Acc = double;
  static constexpr bool cplx = false;
  static __device__ __forceinline__ Acc load(float const* p) { return (double)*p; }
  static __device__ __forceinline__ void store(float* p, Acc v) { *p = (float)v; }
};
template <> struct VT<double> {
  using Acc = double;
  static constexpr bool cplx = false;
  static __device__ __forceinline__ Acc load(double const* p) { return *p; }
  static __device__ __forceinline__ void store(double* p, Acc v) { *p = v; }
};
template <> struct VT<float2> {
  using Acc = double2;
  static constexpr bool cplx = true;
  static __device__ __forceinline__ Acc load(float2 const* p) { float2 v = *p; return make_double2(v.x, v.y); }
  static __device__ __forceinline__ void store(float2* p, Acc v) { *p = make_float2((float)v.x, (float)v.y); }
};
template <> struct VT<double2> {
  using Acc = double2;
  static constexpr bool cplx = true;
  static __device__ __forceinline__ Acc load(double2 const* p) { return *p; }
  static __device__ __forceinline__ void store(double2* p, Acc v) { *p = v; }
};

// complex helpers on double2; real scalars ride in .x
__device__ __forceinline__ double2 to2(double v) { return make_double2(v, 0.0); }
__device__ __forceinline__ double2 to2(double2 v) { return v; }
template <class Acc> __device__ __forceinline__ Acc from2(double2 v);
template <> __device__ __forceinline__ double from2<double>(double2 v) { return v.x; }
template <> __device__ __forceinline__ double2 from2<double2>(double2 v) { return v; }
// acc += conj(a) * b
__device__ __forceinline__ void dot_acc(double2& acc, double a, double b) { acc.x += a * b; }
__device__ __forceinline__ void dot_acc(double2& acc, double2 a, double2 b) {
  acc.x += a.x * b.x + a.y * b.y;
  acc.y += a.x * b.y - a.y * b.x;
}
// a * c  (c complex coefficient as double2)
__device__ __forceinline__ double mulc(double a, double2 c) { return a * c.x; }
__device__ __forceinline__ double2 mulc(double2 a, double2 c) {
  return make_double2(a.x * c.x - a.y * c.y, a.x * c.y + a.y * c.x);
}
__device__ __forceinline__ double addv(double a, double b) { return a + b; }
__device__ __forceinline__ double2 addv(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double subv(double a, double b) { return a - b; }
__device__ __forceinline__ double2 subv(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ void block_reduce_store(double2 v, double2* dst) {
  __shared__ double2 red[kThreads / 32];
  for (int o = 16; o; o >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, o);
    v.y += __shfl_down_sync(0xffffffffu, v.y, o);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double2 s = make_double2(0, 0);
    for (int w = 0; w < kThreads / 32; ++w) {
      s.x += red[w].x;
      s.y += red[w].y;
    }
    *dst = s;
  }
  __syncthreads();
}

// partial[j * grid + block] = sum_rows conj(V_j) w   for j in [0, ncols)
template <class T>
__global__ void __launch_bounds__(kThreads) multi_dot_kernel(T const* V, u64 ld, int ncols, T const* w, u64 n,
                                                             double2* partial) {
  using A = typename VT<T>::Acc;
  for (int j0 = 0; j0 < ncols; j0 += kDotChunk) {
    int nj = min(kDotChunk, ncols - j0);
    double2 acc[kDotChunk];
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j) acc[j] = make_double2(0, 0);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
      A wv = VT<T>::load(w + i);
#pragma unroll
      for (int j = 0; j < kDotChunk; ++j)
        if (j < nj) dot_acc(acc[j], VT<T>::load(V + (u64)(j0 + j) * ld + i), wv);
    }
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j)
      if (j < nj) block_reduce_store(acc[j], partial + (u64)(j0 + j) * gridDim.x + blockIdx.x);
  }
}

// out[j] = sum_b partial[j * nblocks + b]  in fixed order (one thread per j)
__global__ void finish_dot_kernel(double2 const* partial, int nblocks, int ncols, double2* out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  double2 s = make_double2(0, 0);
  for (int b = 0; b < nblocks; ++b) {
    s.x += partial[(u64)j * nblocks + b].x;
    s.y += partial[(u64)j * nblocks + b].y;
  }
  out[j] = s;
}

// w -= sum_j coeff[j] V_j
template <class T>
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(T const* V, u64 ld, int ncols, double2 const* coeff,
                                                              T* w, u64 n) {
  using A = typename VT<T>::Acc;
  __shared__ double2 c[kMaxBasis];
  for (int j = threadIdx.x; j < ncols; j += blockDim.x) c[j] = coeff[j];
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A acc = VT<T>::load(w + i);
    for (int j = 0; j < ncols; ++j) acc = subv(acc, mulc(VT<T>::load(V + (u64)j * ld + i), c[j]));
    VT<T>::store(w + i, acc);
  }
}

// w *= 1 / sqrt(norm2[0].x)   (norm2 on the device)
template <class T>
__global__ void __launch_bounds__(kThreads) normalize_kernel(T* w, u64 n, double2 const* norm2) {
  using A = typename VT<T>::Acc;
  double s = rsqrt(norm2[0].x);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A v = VT<T>::load(w + i);
    VT<T>::store(w + i, mulc(v, make_double2(s, 0)));
  }
}

// out = sum_j (W_j - theta V_j) s[j];  partial[block] = sum |out|^2
template <class T>
__global__ void __launch_bounds__(kThreads) residual_kernel(T const* V, T const* Wm, u64 ld, int m,
                                                            double2 const* s, double theta, T* out, u64 n,
                                                            double2* partial) {
  using A = typename VT<T>::Acc;
  __shared__ double2 c[kMaxBasis];
  for (int j = threadIdx.x; j < m; j += blockDim.x) c[j] = s[j];
  __syncthreads();
  double2 nrm = make_double2(0, 0);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A acc = from2<A>(make_double2(0, 0));
    for (int j = 0; j < m; ++j) {
      A wv = VT<T>::load(Wm + (u64)j * ld + i);
      A vv = VT<T>::load(V + (u64)j * ld + i);
      A d = subv(wv, mulc(vv, make_double2(theta, 0)));
      acc = addv(acc, mulc(d, c[j]));
    }
    // residual norms are measured on what is stored (storage precision)
    VT<T>::store(out + i, acc);
    dot_acc(nrm, acc, acc);
  }
  block_reduce_store(nrm, partial + blockIdx.x);
}

// rows of V (n x m) <- rows * C (m x p), in place; C row-major in global memory
template <class T>
__global__ void __launch_bounds__(kThreads) row_transform_kernel(T* V, u64 ld, int m, int p, double2 const* C, u64 n) {
  using A = typename VT<T>::Acc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* c = reinterpret_cast<double2*>(smem_raw);
  for (int j = threadIdx.x; j < m * p; j += blockDim.x) c[j] = C[j];
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A row[kMaxBasis];
    for (int j = 0; j < m; ++j) row[j] = VT<T>::load(V + (u64)j * ld + i);
    for (int q = 0; q < p; ++q) {
      A acc = from2<A>(make_double2(0, 0));
      for (int j = 0; j < m; ++j) acc = addv(acc, mulc(row[j], c[j * p + q]));
      VT<T>::store(V + (u64)q * ld + i, acc);
    }
  }
}

__device__ __forceinline__ u64 splitmix64(u64 z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// uniform(-1, 1) from splitmix64(seed ^ global_row)  (same data for any rank count)
template <class T>
__global__ void __launch_bounds__(kThreads) random_kernel(T* w, RowDist d, u64 seed) {
  u64 const n = d.n_local;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 h = splitmix64(seed ^ dist_local_to_global(d, i));
    double re = (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    if constexpr (VT<T>::cplx) {
      u64 h2 = splitmix64(h);
      double im = (double)(h2 >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
      VT<T>::store(w + i, make_double2(re, im));
    } else {
      VT<T>::store(w + i, re);
    }
  }
}

double now_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class T>
struct Solver {
  Operator& op;
  int dtype;
  u64 n_global, row0, n;  // local rows [row0, row0 + n)
  u64 ld;                 // leading dimension of V and W (elements)
  int grid;
  cudaStream_t stream;
  DeviceBuffer<T> V, Wm, xfull;
  DeviceBuffer<double2> partial, scal, coeff;
  u64 chunk;              // rows per rank in the replicated vector
  EighStats stats;

  Solver(Operator& o, int dt) : op(o), dtype(dt) {}

  void sync() { CUDA_CHECK(cudaStreamSynchronize(stream)); }

  // out_host[j] = <V_j, w> for j < ncols (all-reduced)
  std::vector<cplx> dots(T const* Vp, int ncols, T const* w, bool to_host, double2* dev_out = nullptr) {
    double2* out = dev_out ? dev_out : scal.ptr;
    multi_dot_kernel<T><<<grid, kThreads, 0, stream>>>(Vp, ld, ncols, w, n, partial.ptr);
    KERNEL_LAUNCHED();
    finish_dot_kernel<<<(ncols + 63) / 64, 64, 0, stream>>>(partial.ptr, grid, ncols, out);
    KERNEL_LAUNCHED();
    comm_allreduce_sum_f64(reinterpret_cast<double*>(out), 2 * (size_t)ncols, stream);
    std::vector<cplx> h;
    if (to_host) {
      std::vector<double2> tmp(ncols);
      CUDA_CHECK(cudaMemcpyAsync(tmp.data(), out, sizeof(double2) * ncols, cudaMemcpyDeviceToHost, stream));
      sync();
      h.resize(ncols);
      for (int j = 0; j < ncols; ++j) h[j] = cplx(tmp[j].x, tmp[j].y);
    }
    return h;
  }

  // w <- (I - V V^H) w, then normalise; returns the norm after projection.  Classical
  // Gram-Schmidt with the DGKS criterion: a unit-length w that keeps more than half of its squared
  // norm after the first projection needs no second pass (residuals of Ritz pairs are orthogonal to
  // the basis up to round-off, so this is the common case and saves a full sweep over V).
  double orthonormalize(int m, T* w, bool unit_norm_input = false) {
    double t0 = now_seconds();
    for (int pass = 0; pass < 2 && m > 0; ++pass) {
      auto c = dots(V.ptr, m, w, unit_norm_input, coeff.ptr);
      multi_axpy_kernel<T><<<grid, kThreads, 0, stream>>>(V.ptr, ld, m, coeff.ptr, w, n);
      KERNEL_LAUNCHED();
      if (unit_norm_input) {
        double removed = 0;
        for (auto const& v : c) removed += std::norm(v);
        if (removed < 0.5) break;
      }
    }
    auto nn = dots(w, 1, w, true);
    double nrm = std::sqrt(std::max(0.0, nn[0].real()));
    if (nrm > 0) {
      normalize_kernel<T><<<grid, kThreads, 0, stream>>>(w, n, scal.ptr);
      KERNEL_LAUNCHED();
    }
    stats.seconds_ortho += now_seconds() - t0;
    return nrm;
  }

  void randomize(T* w, u64 seed) {
    random_kernel<T><<<grid, kThreads, 0, stream>>>(w, op.dist, seed);
    KERNEL_LAUNCHED();
  }

  // Wm[:, j0:j0+nb] = H V[:, j0:j0+nb]
  void apply(int j0, int nb) {
    double t0 = now_seconds();
    Comm& cm = comm();
    if (!cm.active()) {
      op.matmat_device(dtype, nb, V.ptr + (u64)j0 * ld, ld, Wm.ptr + (u64)j0 * ld, ld, stream);
    } else {
      // per column: all-gather of the shard overlapped with the local-source pass (operator.cu)
      for (int c = 0; c < nb; ++c)
        op.matvec_sharded(dtype, V.ptr + (u64)(j0 + c) * ld, Wm.ptr + (u64)(j0 + c) * ld, xfull.ptr, stream);
    }
    sync();
    stats.matvecs += nb;
    stats.seconds_matvec += now_seconds() - t0;
  }

  void transform(T* M, int m, int p, std::vector<cplx> const& C) {
    std::vector<double2> c((size_t)m * p);
    for (size_t i = 0; i < c.size(); ++i) c[i] = make_double2(C[i].real(), C[i].imag());
    if (transform_coeff.count < c.size()) transform_coeff.alloc((size_t)kMaxBasis * kMaxBasis);
    CUDA_CHECK(cudaMemcpyAsync(transform_coeff.ptr, c.data(), c.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    size_t smem = c.size() * sizeof(double2);
    if (smem > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(row_transform_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    row_transform_kernel<T><<<grid, kThreads, smem, stream>>>(M, ld, m, p, transform_coeff.ptr, n);
    KERNEL_LAUNCHED();
    sync();  // `c` is pageable host memory read by the async copy
  }
  DeviceBuffer<double2> transform_coeff;

  int run(u64 k, double eps, int m_max, int b_max, int m_min, double* evals_out, void* evecs_out, double* rnorms_out,
          sped_monitor_fn monitor, void* mctx) {
    double t_start = now_seconds();
    Basis& B = *op.basis;
    Comm& cm = comm();
    n_global = B.n_states;
    op.prepare();
    row0 = 0;
    n = op.dist.n_local;
    chunk = op.dist.chunk;
    if (k == 0 || k > n_global) fail(LS_INVALID_ARGUMENT, "number of eigenpairs must be in 1..dimension");
    // option defaults
    int b = b_max > 0 ? b_max : (int)std::min<u64>(k, 8);
    b = (int)std::min<u64>((u64)b, n_global);
    int mmax = m_max > 0 ? m_max : std::max<int>(4 * b + 2 * (int)k, 24);
    mmax = (int)std::min<u64>((u64)std::min(mmax, kMaxBasis), n_global);
    if (mmax < (int)k + 1 && (u64)mmax < n_global) mmax = (int)std::min<u64>(n_global, k + 1 + (u64)b);
    if (b > mmax) b = mmax;
    int keep = m_min > 0 ? m_min : std::max<int>((int)k, mmax / 3);
    double const mach = (dtype == SPED_F32 || dtype == SPED_C64) ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    double const tol = eps > 0 ? eps : 1e4 * mach;

    stream = cm.active() ? cm.stream : nullptr;
    ld = std::max<u64>(n, 1);
    grid = persistent_grid(std::max<u64>(n, 1), kThreads, 4);
    V.alloc(ld * mmax);
    Wm.alloc(ld * mmax);
    if (cm.active()) xfull.alloc(chunk * cm.world);
    partial.alloc((size_t)grid * kMaxBasis);
    scal.alloc(kMaxBasis);
    coeff.alloc(kMaxBasis);
    if (cm.active()) CUDA_CHECK(cudaMemsetAsync(xfull.ptr, 0, chunk * cm.world * sizeof(T), stream));

    std::vector<cplx> H((size_t)mmax * mmax, cplx(0, 0));  // projected matrix, row-major, leading dim mmax
    auto Hat = [&](int i, int j) -> cplx& { return H[(size_t)i * mmax + j]; };
    int m = 0;
    u64 seed = 0x5EED0002ull;
    auto append_random = [&]() {
      for (int attempt = 0; attempt < 8; ++attempt) {
        randomize(V.ptr + (u64)m * ld, seed);
        seed += 0x1000193ull;
        if (orthonormalize(m, V.ptr + (u64)m * ld) > 1e-8) return;
      }
      fail(SPED_INTERNAL_ERROR, "could not generate a new search direction");
    };
    auto extend_projection = [&](int m_old, int m_new) {
      double t_pr = now_seconds();
      for (int j = m_old; j < m_new; ++j) {
        auto h = dots(V.ptr, m_new, Wm.ptr + (u64)j * ld, true);
        for (int i = 0; i < m_new; ++i) {
          Hat(i, j) = h[i];
          Hat(j, i) = std::conj(h[i]);
        }
        Hat(j, j) = Hat(j, j).real();
      }
      stats.seconds_project += now_seconds() - t_pr;
    };
    int const b0 = std::min<int>(mmax, std::max<int>(b, (int)std::min<u64>(k, (u64)mmax)));
    for (int j = 0; j < b0; ++j) {
      append_random();
      ++m;
    }
    apply(0, m);
    extend_projection(0, m);

    std::vector<double> theta;
    std::vector<cplx> S, S_prev;
    int m_prev = 0, n_prev = 0;
    std::vector<double> rn(k, 0.0), ev(k, 0.0);
    double a_norm = 0;
    int status = SPED_NOT_CONVERGED;
    int const max_outer = 200000;
    for (int it = 0; it < max_outer; ++it) {
      stats.iterations = it + 1;
      std::vector<cplx> Hm((size_t)m * m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Hm[(size_t)i * m + j] = Hat(i, j);
      small_eigh(m, Hm, theta, S);
      for (double t : theta) a_norm = std::max(a_norm, std::abs(t));
      int const kk = (int)std::min<u64>(k, (u64)m);
      // residuals of the wanted pairs; the first `b` unconverged ones become new directions,
      // written straight into the free columns of V when there is room, otherwise after a restart
      std::vector<int> unconverged;
      int n_conv = 0;
      bool have_all = m >= (int)k;
      double t_res = now_seconds();
      for (int i = 0; i < kk; ++i) {
        std::vector<double2> s(m);
        for (int j = 0; j < m; ++j) s[j] = make_double2(S[(size_t)j * m + i].real(), S[(size_t)j * m + i].imag());
        CUDA_CHECK(cudaMemcpyAsync(coeff.ptr, s.data(), sizeof(double2) * m, cudaMemcpyHostToDevice, stream));
        residual_kernel<T><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, coeff.ptr, theta[i], scratch(i), n,
                                                          partial.ptr);
        KERNEL_LAUNCHED();
        finish_dot_kernel<<<1, 64, 0, stream>>>(partial.ptr, grid, 1, scal.ptr);
        KERNEL_LAUNCHED();
        comm_allreduce_sum_f64(reinterpret_cast<double*>(scal.ptr), 2, stream);
        double2 r2;
        CUDA_CHECK(cudaMemcpyAsync(&r2, scal.ptr, sizeof(double2), cudaMemcpyDeviceToHost, stream));
        sync();
        rn[i] = std::sqrt(std::max(0.0, r2.x));
        ev[i] = theta[i];
        if (rn[i] <= tol * a_norm) ++n_conv;
        else unconverged.push_back(i);
      }
      stats.seconds_residual += now_seconds() - t_res;
      if (monitor) {
        sped_eigh_info info{it, m, n_conv, (int)k, stats.matvecs, ev.data(), rn.data(), now_seconds() - t_start};
        if (monitor(&info, mctx) != 0) break;
      }
      if (have_all && n_conv == (int)k) {
        status = LS_SUCCESS;
        break;
      }
      if ((u64)m >= n_global) {
        // the basis spans the whole space: Ritz pairs are exact up to round-off
        status = LS_SUCCESS;
        break;
      }
      int nb = std::min<int>(b, (int)unconverged.size());
      if (!have_all) nb = std::max(nb, 1);
      nb = (int)std::min<u64>((u64)nb, n_global - (u64)m);
      // restart when the new directions do not fit
      if (m + nb > mmax) {
        double t_rs = now_seconds();
        int r = std::min(std::max(keep, (int)std::min<u64>(k, (u64)m)), mmax - nb);
        r = std::max(r, 1);
        int p_room = mmax - nb - r;
        // coefficient block C = [S(:, 0:r) | previous Ritz directions orthogonalised against it]
        std::vector<std::vector<cplx>> cols;
        for (int q = 0; q < r; ++q) {
          std::vector<cplx> c(m);
          for (int j = 0; j < m; ++j) c[j] = S[(size_t)j * m + q];
          cols.push_back(c);
        }
        for (int q = 0; q < n_prev && p_room > 0; ++q) {
          std::vector<cplx> c(m, cplx(0, 0));
          for (int j = 0; j < m_prev; ++j) c[j] = S_prev[(size_t)j * n_prev + q];
          for (int pass = 0; pass < 2; ++pass)
            for (auto const& u : cols) {
              cplx d = 0;
              for (int j = 0; j < m; ++j) d += std::conj(u[j]) * c[j];
              for (int j = 0; j < m; ++j) c[j] -= d * u[j];
            }
          double nn = 0;
          for (auto const& v : c) nn += std::norm(v);
          if (nn < 1e-20) continue;
          for (auto& v : c) v /= std::sqrt(nn);
          cols.push_back(c);
          --p_room;
        }
        int p = (int)cols.size();
        std::vector<cplx> C((size_t)m * p);
        for (int q = 0; q < p; ++q)
          for (int j = 0; j < m; ++j) C[(size_t)j * p + q] = cols[q][j];
        // move the pending residuals out of the way is unnecessary: scratch lives outside V
        transform(V.ptr, m, p, C);
        transform(Wm.ptr, m, p, C);
        std::vector<cplx> Hn((size_t)p * p, cplx(0, 0));
        for (int a = 0; a < p; ++a)
          for (int c2 = 0; c2 < p; ++c2) {
            cplx acc = 0;
            for (int i = 0; i < m; ++i) {
              cplx t = 0;
              for (int j = 0; j < m; ++j) t += Hat(i, j) * cols[c2][j];
              acc += std::conj(cols[a][i]) * t;
            }
            Hn[(size_t)a * p + c2] = acc;
          }
        for (int a = 0; a < p; ++a)
          for (int c2 = 0; c2 < p; ++c2) Hat(a, c2) = Hn[(size_t)a * p + c2];
        // after the transform the first r Ritz vectors are the unit vectors e_0..e_{r-1}
        S.assign((size_t)p * p, cplx(0, 0));
        for (int q = 0; q < p; ++q) S[(size_t)q * p + q] = 1.0;
        m = p;
        ++stats.restarts;
        stats.seconds_restart += now_seconds() - t_rs;
      }
      // remember the current Ritz directions (for the "+k" part of the next restart)
      n_prev = (int)std::min<u64>((u64)std::max(1, std::min(b, (int)k)), (u64)m);
      m_prev = m;
      S_prev.assign((size_t)m * n_prev, cplx(0, 0));
      for (int q = 0; q < n_prev; ++q)
        for (int j = 0; j < m; ++j) S_prev[(size_t)j * n_prev + q] = S[(size_t)j * m + q];
      // expand
      int m_old = m;
      for (int q = 0; q < nb; ++q) {
        T* dst = V.ptr + (u64)m * ld;
        bool ok = false;
        if (q < (int)unconverged.size()) {
          CUDA_CHECK(cudaMemcpyAsync(dst, scratch(unconverged[q]), n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
          // scale to unit length first so the breakdown test is relative
          auto nn = dots(dst, 1, dst, true);
          if (nn[0].real() > 0) {
            normalize_kernel<T><<<grid, kThreads, 0, stream>>>(dst, n, scal.ptr);
            KERNEL_LAUNCHED();
            ok = orthonormalize(m, dst, true) > 1e-7;
          }
        }
        if (!ok) append_random();
        ++m;
      }
      apply(m_old, m - m_old);
      extend_projection(m_old, m);
    }

    // Ritz vectors of the wanted pairs: V <- V S(:, 0:k)
    int const kk = (int)std::min<u64>(k, (u64)m);
    {
      std::vector<cplx> Hm((size_t)m * m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Hm[(size_t)i * m + j] = Hat(i, j);
      small_eigh(m, Hm, theta, S);
      std::vector<cplx> C((size_t)m * kk);
      for (int q = 0; q < kk; ++q)
        for (int j = 0; j < m; ++j) C[(size_t)j * kk + q] = S[(size_t)j * m + q];
      transform(V.ptr, m, kk, C);
    }
    for (u64 i = 0; i < k; ++i) {
      evals_out[i] = i < (u64)kk ? theta[i] : 0.0;
      rnorms_out[i] = i < (u64)kk ? rn[i] : 0.0;
    }
    if (evecs_out) {
      T* host = static_cast<T*>(evecs_out);
      if (!cm.active()) {
        CUDA_CHECK(cudaMemcpy2D(host, n_global * sizeof(T), V.ptr, ld * sizeof(T), n * sizeof(T), kk, cudaMemcpyDeviceToHost));
      } else {
        // all-gather gives the [rank][local] layout; the blocks are put back in global row order
        // on the host (one memcpy per block of 2^log2b rows)
        std::vector<T> tmp(chunk * cm.world);
        RowDist const& d = op.dist;
        u64 const bsz = (u64)1 << d.log2b;
        for (int q = 0; q < kk; ++q) {
          CUDA_CHECK(cudaMemcpyAsync(xfull.ptr + (u64)cm.rank * chunk, V.ptr + (u64)q * ld, n * sizeof(T),
                                     cudaMemcpyDeviceToDevice, stream));
          comm_allgather_inplace(xfull.ptr, chunk * sizeof(T), stream);
          CUDA_CHECK(cudaMemcpyAsync(tmp.data(), xfull.ptr, tmp.size() * sizeof(T), cudaMemcpyDeviceToHost, stream));
          sync();
          for (u64 g0 = 0; g0 < n_global; g0 += bsz) {
            u64 len = std::min(bsz, n_global - g0);
            std::memcpy(host + (u64)q * n_global + g0, tmp.data() + dist_global_to_pos(d, g0), len * sizeof(T));
          }
        }
      }
    }
    stats.seconds_total = now_seconds() - t_start;
    op.last_stats = stats;
    return status;
  }

  // residual scratch: k columns kept in a separate buffer
  DeviceBuffer<T> scratch_buf;
  T* scratch(int i) {
    return scratch_buf.ptr + (u64)i * ld;
  }
};

template <class T>
int run_solver(Operator& op, int dtype, u64 k, double eps, int m_max, int b_max, int m_min, double* evals, void* evecs,
               double* rnorms, sped_monitor_fn monitor, void* ctx) {
  Solver<T> s(op, dtype);
  op.prepare();
  u64 n = op.dist.n_local;
  s.scratch_buf.alloc(std::max<u64>(n, 1) * std::max<u64>(1, std::min<u64>(k, (u64)kMaxBasis)));
  return s.run(k, eps, m_max, b_max, m_min, evals, evecs, rnorms, monitor, ctx);
}

}  // namespace

int eigh(Operator& op, int dtype, u64 n_evals, double eps, int max_basis, int max_block, int min_restart,
         double* evals, void* evecs, double* rnorms, sped_monitor_fn monitor, void* ctx) {
  if (!dtype_is_complex(dtype) && !op.is_real())
    fail(LS_OPERATOR_IS_COMPLEX, "operator is complex but a real datatype was requested");
  if (n_evals > (u64)kMaxBasis / 2) fail(LS_INVALID_ARGUMENT, "at most 32 eigenpairs are supported");
  switch (dtype) {
    case SPED_F32: return run_solver<float>(op, dtype, n_evals, eps, max_basis, max_block, min_restart, evals, evecs, rnorms, monitor, ctx);
    case SPED_F64: return run_solver<double>(op, dtype, n_evals, eps, max_basis, max_block, min_restart, evals, evecs, rnorms, monitor, ctx);
    case SPED_C64: return run_solver<float2>(op, dtype, n_evals, eps, max_basis, max_block, min_restart, evals, evecs, rnorms, monitor, ctx);
    case SPED_C128: return run_solver<double2>(op, dtype, n_evals, eps, max_basis, max_block, min_restart, evals, evecs, rnorms, monitor, ctx);
  }
  fail(LS_INVALID_DATATYPE, "unknown datatype tag");
}

}  // namespace sped
