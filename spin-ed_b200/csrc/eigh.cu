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
#include "eigh_kernels.cuh"

namespace sped {

// ---------------------------------------------------------------------------------------------
// small dense Hermitian eigenproblem (cyclic Jacobi), ascending eigenvalues, columns = vectors
// ---------------------------------------------------------------------------------------------
// Real symmetric case (every deck with real characters): Householder reduction to tridiagonal form
// and the implicit QL iteration -- O(m^3) with a small constant, tens of microseconds for the
// projected problems of the solver, where cyclic Jacobi in complex arithmetic took about a
// millisecond per outer iteration.  a: row-major m x m, overwritten by the eigenvectors (columns);
// d: eigenvalues (unsorted).  Returns false if the QL iteration does not converge.
static bool symmetric_tridiagonal_eigh(int n, std::vector<double>& a, std::vector<double>& d) {
  std::vector<double> e(n, 0.0);
  d.assign(n, 0.0);
  auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
  // Householder reduction (rows n-1 .. 1), transformations accumulated in a
  for (int i = n - 1; i > 0; --i) {
    int const l = i - 1;
    double h = 0, scale = 0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) scale += std::abs(A(i, k));
      if (scale == 0.0) {
        e[i] = A(i, l);
      } else {
        for (int k = 0; k <= l; ++k) {
          A(i, k) /= scale;
          h += A(i, k) * A(i, k);
        }
        double f = A(i, l);
        double g = f >= 0 ? -std::sqrt(h) : std::sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        A(i, l) = f - g;
        f = 0;
        for (int j = 0; j <= l; ++j) {
          A(j, i) = A(i, j) / h;
          g = 0;
          for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
          for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
          e[j] = g / h;
          f += e[j] * A(i, j);
        }
        double const hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = A(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
        }
      }
    } else {
      e[i] = A(i, l);
    }
    d[i] = h;
  }
  d[0] = 0;
  e[0] = 0;
  for (int i = 0; i < n; ++i) {
    int const l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0;
        for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
        for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
      }
    }
    d[i] = A(i, i);
    A(i, i) = 1.0;
    for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
  }
  // implicit QL with Wilkinson shifts on (d, e), rotations applied to the columns of a
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, mm;
    do {
      for (mm = l; mm < n - 1; ++mm) {
        double const dd = std::abs(d[mm]) + std::abs(d[mm + 1]);
        if (std::abs(e[mm]) <= 2.220446049250313e-16 * dd) break;
      }
      if (mm != l) {
        if (iter++ == 80) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? std::abs(r) : -std::abs(r)));
        double sn = 1, cs = 1, p = 0;
        int i;
        for (i = mm - 1; i >= l; --i) {
          double f = sn * e[i];
          double const b = cs * e[i];
          e[i + 1] = r = std::hypot(f, g);
          if (r == 0.0) {
            d[i + 1] -= p;
            e[mm] = 0;
            break;
          }
          sn = f / r;
          cs = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * sn + 2.0 * cs * b;
          p = sn * r;
          d[i + 1] = g + p;
          g = cs * r - b;
          for (int k = 0; k < n; ++k) {
            f = A(k, i + 1);
            A(k, i + 1) = sn * A(k, i) + cs * f;
            A(k, i) = cs * A(k, i) - sn * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[mm] = 0;
      }
    } while (mm != l);
  }
  return true;
}

void small_eigh(int m, std::vector<cplx> A /* row-major m x m */, std::vector<double>& evals, std::vector<cplx>& evecs) {
  {  // real symmetric input: the fast path
    bool real = true;
    for (auto const& v : A) real = real && v.imag() == 0.0;
    if (real && m > 0) {
      std::vector<double> a((size_t)m * m), d;
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) a[(size_t)i * m + j] = 0.5 * (A[(size_t)i * m + j].real() + A[(size_t)j * m + i].real());
      if (symmetric_tridiagonal_eigh(m, a, d)) {
        std::vector<int> order(m);
        for (int i = 0; i < m; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return d[x] < d[y]; });
        evals.resize(m);
        evecs.assign((size_t)m * m, cplx(0, 0));
        for (int k = 0; k < m; ++k) {
          evals[k] = d[order[k]];
          for (int i = 0; i < m; ++i) evecs[(size_t)i * m + k] = a[(size_t)i * m + order[k]];
        }
        return;
      }
    }
  }
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

// out[j] = sum_b partial[j * nblocks + b] in a fixed order: one warp per j, lane l takes the blocks
// l, l + 32, ... and the lanes are combined by a shuffle tree.  (One thread per j walked its 592
// partials one dependent load after the other: 57 us per call, several calls per iteration -- 0.17 s of
// the 0.40 s xxz_triangular_19 solve.)
__global__ void __launch_bounds__(32) finish_dot_kernel(double2 const* partial, int nblocks, int ncols, double2* out) {
  int const j = blockIdx.x;
  if (j >= ncols) return;
  double2 s = make_double2(0, 0);
  for (int b = threadIdx.x; b < nblocks; b += 32) {
    s.x += partial[(u64)j * nblocks + b].x;
    s.y += partial[(u64)j * nblocks + b].y;
  }
  for (int o = 16; o; o >>= 1) {
    s.x += __shfl_down_sync(0xffffffffu, s.x, o);
    s.y += __shfl_down_sync(0xffffffffu, s.y, o);
  }
  if (threadIdx.x == 0) out[j] = s;
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

// ---- block forms: several vectors per pass, so that the basis V is read once for all of them ----
// A kernel with a `gate` returns at once when *gate == 0: the second (DGKS) sweep of the block
// orthogonalisation is queued unconditionally and decides on the device whether it has work.

// partial[(j * NW + c) * gridDim.x + block] = sum_rows conj(V_j) w_c,  j < m, c < NW
template <class T, int NW>
__global__ void __launch_bounds__(kThreads) block_dot_kernel(T const* V, u64 ld, int m, T const* W, u64 ldw, u64 n,
                                                             double2* partial, int const* gate) {
  using A = typename VT<T>::Acc;
  if (gate && *gate == 0) return;
  constexpr int JC = NW == 1 ? 8 : (NW == 2 ? 6 : 4);
  for (int j0 = 0; j0 < m; j0 += JC) {
    double2 acc[JC][NW];
#pragma unroll
    for (int j = 0; j < JC; ++j)
#pragma unroll
      for (int c = 0; c < NW; ++c) acc[j][c] = make_double2(0, 0);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
      A wv[NW];
#pragma unroll
      for (int c = 0; c < NW; ++c) wv[c] = VT<T>::load(W + (u64)c * ldw + i);
#pragma unroll
      for (int j = 0; j < JC; ++j)
        if (j0 + j < m) {
          A const vv = VT<T>::load(V + (u64)(j0 + j) * ld + i);
#pragma unroll
          for (int c = 0; c < NW; ++c) dot_acc(acc[j][c], vv, wv[c]);
        }
    }
#pragma unroll
    for (int j = 0; j < JC; ++j)
      if (j0 + j < m) {
#pragma unroll
        for (int c = 0; c < NW; ++c) block_reduce_store(acc[j][c], partial + (u64)((j0 + j) * NW + c) * gridDim.x + blockIdx.x);
      }
  }
}

// w_c -= sum_j coeff[j * NW + c] V_j,  c < NW
template <class T, int NW>
__global__ void __launch_bounds__(kThreads) block_axpy_kernel(T const* V, u64 ld, int m, double2 const* coeff, T* W, u64 ldw,
                                                              u64 n, int const* gate) {
  using A = typename VT<T>::Acc;
  if (gate && *gate == 0) return;
  __shared__ double2 c[kMaxBasis * NW];
  for (int j = threadIdx.x; j < m * NW; j += blockDim.x) c[j] = coeff[j];
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A acc[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q) acc[q] = VT<T>::load(W + (u64)q * ldw + i);
    for (int j = 0; j < m; ++j) {
      A const vv = VT<T>::load(V + (u64)j * ld + i);
#pragma unroll
      for (int q = 0; q < NW; ++q) acc[q] = subv(acc[q], mulc(vv, c[j * NW + q]));
    }
#pragma unroll
    for (int q = 0; q < NW; ++q) VT<T>::store(W + (u64)q * ldw + i, acc[q]);
  }
}

// w *= 1 / sqrt(norm2[0].x), or 0 when the squared norm is not above `tiny` (a direction that is
// linearly dependent on the basis becomes the zero vector, harmless for every later kernel; the
// host replaces it when it reads the recorded norm).  record[0] = norm2; with `flag` given,
// *flag is set when less than half of the (unit) input survived -- the DGKS criterion for a
// second orthogonalisation sweep.
template <class T>
__global__ void __launch_bounds__(kThreads) scale_kernel(T* w, u64 n, double2 const* norm2, double tiny, double* record, int* flag,
                                                         int const* gate) {
  using A = typename VT<T>::Acc;
  if (gate && *gate == 0) return;
  double const v = norm2[0].x;
  double const s = v > tiny ? rsqrt(v) : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (record) *record = v;
    if (flag && v < 0.5) *flag = 1;
  }
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    VT<T>::store(w + i, mulc(VT<T>::load(w + i), make_double2(s, 0)));
}

// Cholesky factor of the Gram matrix G = W^H W of a group of nw <= 4 vectors (G row-major,
// G[j * nw + c] = <w_j, w_c>), one thread: G = R^H R, and the inverse of the upper-triangular R, so
// that W R^{-1} has orthonormal columns (CholQR; the DGKS flag below makes it CholQR2 when needed).
// record[c] = R_cc^2 = squared norm of w_c after the components along w_0..w_{c-1} are removed;
// a column with R_cc^2 <= tiny is linearly dependent: it becomes the zero vector.  *flag is
// raised when some column kept less than half of its (unit) length.
__global__ void chol_kernel(double2 const* G, int nw, double2* Rinv, double tiny, double* record, int* flag, int const* gate) {
  if (gate && *gate == 0) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double2 R[4][4], X[4][4];
  bool dead[4];
  auto cm = [](double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
  auto cj = [](double2 a) { return make_double2(a.x, -a.y); };
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) R[i][j] = X[i][j] = make_double2(0, 0);
  for (int c = 0; c < nw; ++c) {
    // column c of R: R_jc = (G_jc - sum_{k<j} conj(R_kj) R_kc) / R_jj
    for (int j = 0; j < c; ++j) {
      double2 v = G[j * nw + c];
      for (int k = 0; k < j; ++k) {
        double2 t = cm(cj(R[k][j]), R[k][c]);
        v.x -= t.x;
        v.y -= t.y;
      }
      R[j][c] = dead[j] ? make_double2(0, 0) : make_double2(v.x / R[j][j].x, v.y / R[j][j].x);
    }
    double d = G[c * nw + c].x;
    for (int k = 0; k < c; ++k) d -= R[k][c].x * R[k][c].x + R[k][c].y * R[k][c].y;
    dead[c] = !(d > tiny);
    R[c][c] = make_double2(dead[c] ? 1.0 : sqrt(d), 0);
    if (record) record[c] = d;
    if (flag && d < 0.5) *flag = 1;
  }
  // X = R^{-1} (upper triangular) by back substitution, column by column; dead columns are zero
  for (int c = 0; c < nw; ++c) {
    if (dead[c]) continue;
    X[c][c] = make_double2(1.0 / R[c][c].x, 0);
    for (int i = c - 1; i >= 0; --i) {
      double2 v = make_double2(0, 0);
      for (int k = i + 1; k <= c; ++k) {
        double2 t = cm(R[i][k], X[k][c]);
        v.x -= t.x;
        v.y -= t.y;
      }
      X[i][c] = make_double2(v.x / R[i][i].x, v.y / R[i][i].x);
    }
  }
  for (int i = 0; i < nw; ++i)
    for (int c = 0; c < nw; ++c) Rinv[i * nw + c] = X[i][c];
}

// W <- W X for an nw x nw matrix X (row-major), in place, row by row
template <class T, int NW>
__global__ void __launch_bounds__(kThreads) right_multiply_kernel(T* W, u64 ldw, u64 n, double2 const* X, int const* gate) {
  using A = typename VT<T>::Acc;
  if (gate && *gate == 0) return;
  __shared__ double2 x[NW * NW];
  if (threadIdx.x < NW * NW) x[threadIdx.x] = X[threadIdx.x];
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A row[NW], out[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) row[j] = VT<T>::load(W + (u64)j * ldw + i);
#pragma unroll
    for (int c = 0; c < NW; ++c) {
      out[c] = from2<A>(make_double2(0, 0));
#pragma unroll
      for (int j = 0; j <= c; ++j) out[c] = addv(out[c], mulc(row[j], x[j * NW + c]));
    }
#pragma unroll
    for (int c = 0; c < NW; ++c) VT<T>::store(W + (u64)c * ldw + i, out[c]);
  }
}

// out_q = sum_j (W_j - theta_q V_j) s[j * KQ + q] for q < KQ in ONE pass over V and W;
// partial[q * gridDim.x + block] = sum |out_q|^2
template <class T, int KQ>
__global__ void __launch_bounds__(kThreads) residual_block_kernel(T const* V, T const* Wm, u64 ld, int m, double2 const* s,
                                                                  double const* theta, T* out, u64 ldo, u64 n, double2* partial) {
  using A = typename VT<T>::Acc;
  __shared__ double2 c[kMaxBasis * KQ];
  __shared__ double th[KQ];
  for (int j = threadIdx.x; j < m * KQ; j += blockDim.x) c[j] = s[j];
  if (threadIdx.x < KQ) th[threadIdx.x] = theta[threadIdx.x];
  __syncthreads();
  double2 nrm[KQ];
#pragma unroll
  for (int q = 0; q < KQ; ++q) nrm[q] = make_double2(0, 0);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A acc[KQ];
#pragma unroll
    for (int q = 0; q < KQ; ++q) acc[q] = from2<A>(make_double2(0, 0));
    for (int j = 0; j < m; ++j) {
      A const wv = VT<T>::load(Wm + (u64)j * ld + i);
      A const vv = VT<T>::load(V + (u64)j * ld + i);
#pragma unroll
      for (int q = 0; q < KQ; ++q) acc[q] = addv(acc[q], mulc(subv(wv, mulc(vv, make_double2(th[q], 0))), c[j * KQ + q]));
    }
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
      VT<T>::store(out + (u64)q * ldo + i, acc[q]);
      dot_acc(nrm[q], acc[q], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < KQ; ++q) block_reduce_store(nrm[q], partial + (u64)q * gridDim.x + blockIdx.x);
}

// dst_q = src_q * scale[q]  (new search directions: residuals scaled to unit length)
template <class T>
__global__ void __launch_bounds__(kThreads) copy_scaled_kernel(T const* src, u64 lds, T* dst, u64 ldd, u64 n, double scale) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    VT<T>::store(dst + i, mulc(VT<T>::load(src + i), make_double2(scale, 0)));
  (void)lds;
  (void)ldd;
}

// rows of V (n x m) <- rows * C (m x p), in place; C row-major in global memory.  MAXM bounds m
// at compile time so that the row lives in registers (a dynamically indexed 64-entry array would
// sit in local memory: the restart of a 3-vector basis -- every iteration of the 40/42-spin decks --
// then moves several times the bytes it has to).
template <class T, int MAXM>
__global__ void __launch_bounds__(kThreads) row_transform_kernel(T* V, u64 ld, int m, int p, double2 const* C, u64 n) {
  using A = typename VT<T>::Acc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* c = reinterpret_cast<double2*>(smem_raw);
  for (int j = threadIdx.x; j < m * p; j += blockDim.x) c[j] = C[j];
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A row[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j)
      if (j < m) row[j] = VT<T>::load(V + (u64)j * ld + i);
    for (int q = 0; q < p; ++q) {
      A acc = from2<A>(make_double2(0, 0));
#pragma unroll
      for (int j = 0; j < MAXM; ++j)
        if (j < m) acc = addv(acc, mulc(row[j], c[j * p + q]));
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

// Device time per phase of the solve, measured with CUDA events recorded on the solver's stream and
// resolved once at the end (no host synchronisation inside the iteration for the sake of timing).
struct PhaseTimer {
  enum Phase { kMatvec, kOrtho, kResidual, kRestart, kProject, kOther, kPhases };
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;
  std::vector<int> phase_of;  // phase that ENDS at event i
  size_t used = 0;
  void start(cudaStream_t s) {
    stream = s;
    used = 0;
    phase_of.clear();
    mark(kOther);
  }
  void mark(int phase) {
    if (used == events.size()) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      events.push_back(e);
    }
    CUDA_CHECK(cudaEventRecord(events[used++], stream));
    phase_of.push_back(phase);
  }
  void resolve(double* seconds /* [kPhases] */) {
    CUDA_CHECK(cudaStreamSynchronize(stream));
    for (size_t i = 1; i < used; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, events[i - 1], events[i]);
      seconds[phase_of[i]] += ms * 1e-3;
    }
  }
  ~PhaseTimer() {
    for (auto e : events) cudaEventDestroy(e);
  }
};

template <class T>
struct Solver {
  Operator& op;
  int dtype;
  u64 n_global, n;        // global rows, rows of this rank
  u64 ld;                 // leading dimension of V, W and the residual block (elements)
  int grid;
  cudaStream_t stream;
  // views into the operator's grow-only workspace (see Operator::workspace)
  template <class U>
  struct Ws {
    U* ptr = nullptr;
    size_t count = 0;
    void take(Operator& o, int slot, size_t n) {
      ptr = static_cast<U*>(o.workspace(slot, std::max<size_t>(n, 1) * sizeof(U)));
      count = n;
    }
  };
  Ws<T> V, Wm, xfull, resid;
  Ws<double2> partial, scal, coeff, transform_coeff;
  Ws<double> d_theta, d_norms;
  Ws<int> d_flag;
  u64 chunk;              // rows per rank in the replicated vector
  EighStats stats;
  PhaseTimer timer;
  static constexpr int kGroup = 4;  // vectors handled together by the block kernels

  Solver(Operator& o, int dt) : op(o), dtype(dt) {}

  void sync() { CUDA_CHECK(cudaStreamSynchronize(stream)); }

  // ---- launches (all asynchronous on `stream`) ----
  // out[j * nw + c] = <V_j, w_c>, all-reduced, left on the device
  void block_dots(T const* Vp, int m, T const* W, int nw, double2* out, int const* gate = nullptr) {
    switch (nw) {
      case 1: block_dot_kernel<T, 1><<<grid, kThreads, 0, stream>>>(Vp, ld, m, W, ld, n, partial.ptr, gate); break;
      case 2: block_dot_kernel<T, 2><<<grid, kThreads, 0, stream>>>(Vp, ld, m, W, ld, n, partial.ptr, gate); break;
      case 3: block_dot_kernel<T, 3><<<grid, kThreads, 0, stream>>>(Vp, ld, m, W, ld, n, partial.ptr, gate); break;
      default: block_dot_kernel<T, 4><<<grid, kThreads, 0, stream>>>(Vp, ld, m, W, ld, n, partial.ptr, gate); break;
    }
    KERNEL_LAUNCHED();
    finish_dot_kernel<<<m * nw, 32, 0, stream>>>(partial.ptr, grid, m * nw, out);
    KERNEL_LAUNCHED();
    comm_allreduce_sum_f64(reinterpret_cast<double*>(out), 2 * (size_t)m * nw, stream);
  }
  void block_axpy(T const* Vp, int m, double2 const* c, T* W, int nw, int const* gate = nullptr) {
    switch (nw) {
      case 1: block_axpy_kernel<T, 1><<<grid, kThreads, 0, stream>>>(Vp, ld, m, c, W, ld, n, gate); break;
      case 2: block_axpy_kernel<T, 2><<<grid, kThreads, 0, stream>>>(Vp, ld, m, c, W, ld, n, gate); break;
      case 3: block_axpy_kernel<T, 3><<<grid, kThreads, 0, stream>>>(Vp, ld, m, c, W, ld, n, gate); break;
      default: block_axpy_kernel<T, 4><<<grid, kThreads, 0, stream>>>(Vp, ld, m, c, W, ld, n, gate); break;
    }
    KERNEL_LAUNCHED();
  }
  std::vector<cplx> to_host(double2 const* dev, int count) {
    std::vector<double2> tmp(count);
    CUDA_CHECK(cudaMemcpyAsync(tmp.data(), dev, sizeof(double2) * count, cudaMemcpyDeviceToHost, stream));
    sync();
    std::vector<cplx> h(count);
    for (int j = 0; j < count; ++j) h[j] = cplx(tmp[j].x, tmp[j].y);
    return h;
  }

  // One sweep of the block orthogonalisation of W[:, 0:nw) (unit-length columns on entry of the
  // first sweep): against V[:, 0:m) -- V read once for the whole group -- then inside the group by
  // CholQR; every column ends normalised, its squared norm BEFORE the scaling (after the removal
  // of the components along V and along the earlier columns of the group) recorded in norms_out[c].  first sweep: `raise` is set when some column kept less than half of its
  // length (DGKS: a second sweep is due); second sweep: every kernel is gated on that flag.
  void ortho_sweep(int m, T* W, int nw, double* norms_out, int* raise, int const* gate) {
    if (m > 0) {
      block_dots(V.ptr, m, W, nw, coeff.ptr, gate);
      block_axpy(V.ptr, m, coeff.ptr, W, nw, gate);
    }
    // inside the group: Gram matrix, Cholesky factor on the device, W <- W R^{-1}
    block_dots(W, nw, W, nw, scal.ptr, gate);
    chol_kernel<<<1, 32, 0, stream>>>(scal.ptr, nw, scal.ptr + 16, 1e-24, norms_out, raise, gate);
    KERNEL_LAUNCHED();
    switch (nw) {
      case 1: right_multiply_kernel<T, 1><<<grid, kThreads, 0, stream>>>(W, ld, n, scal.ptr + 16, gate); break;
      case 2: right_multiply_kernel<T, 2><<<grid, kThreads, 0, stream>>>(W, ld, n, scal.ptr + 16, gate); break;
      case 3: right_multiply_kernel<T, 3><<<grid, kThreads, 0, stream>>>(W, ld, n, scal.ptr + 16, gate); break;
      default: right_multiply_kernel<T, 4><<<grid, kThreads, 0, stream>>>(W, ld, n, scal.ptr + 16, gate); break;
    }
    KERNEL_LAUNCHED();
  }

  // New search directions W[:, 0:nw) (columns of V past the first m), unit length on entry:
  // classical Gram-Schmidt by blocks with the DGKS criterion, entirely on the device.  The squared
  // norms that decide whether a direction broke down (linearly dependent on the basis) are left
  // in d_norms[slot .. slot + nw) for the host to read with the next round trip.
  void ortho_block(int m, T* W, int nw, int slot) {
    SPED_NVTX("sped_eigh: block orthogonalisation");
    CUDA_CHECK(cudaMemsetAsync(d_flag.ptr, 0, sizeof(int), stream));
    ortho_sweep(m, W, nw, d_norms.ptr + slot, d_flag.ptr, nullptr);
    ortho_sweep(m, W, nw, nullptr, nullptr, d_flag.ptr);
  }

  // sequential, host-synchronous form (start vectors and replacements of broken-down directions)
  double orthonormalize(int m, T* w) {
    for (int pass = 0; pass < 2 && m > 0; ++pass) {
      block_dots(V.ptr, m, w, 1, coeff.ptr);
      block_axpy(V.ptr, m, coeff.ptr, w, 1);
    }
    block_dots(w, 1, w, 1, scal.ptr);
    double const nn = to_host(scal.ptr, 1)[0].real();
    double const nrm = std::sqrt(std::max(0.0, nn));
    if (nrm > 0) {
      scale_kernel<T><<<grid, kThreads, 0, stream>>>(w, n, scal.ptr, 0.0, nullptr, nullptr, nullptr);
      KERNEL_LAUNCHED();
    }
    return nrm;
  }

  void randomize(T* w, u64 seed) {
    random_kernel<T><<<grid, kThreads, 0, stream>>>(w, op.dist, seed);
    KERNEL_LAUNCHED();
  }

  // Wm[:, j0:j0+nb] = H V[:, j0:j0+nb]   (asynchronous)
  void apply(int j0, int nb) {
    SPED_NVTX("sped_eigh: H V");
    Comm& cm = comm();
    if (!cm.active()) {
      op.matmat_device(dtype, nb, V.ptr + (u64)j0 * ld, ld, Wm.ptr + (u64)j0 * ld, ld, stream);
    } else {
      // per column: exchange of the shards overlapped with the passes over the source classes (operator.cu)
      for (int c = 0; c < nb; ++c)
        op.matvec_sharded(dtype, V.ptr + (u64)(j0 + c) * ld, Wm.ptr + (u64)(j0 + c) * ld, xfull.ptr, stream);
    }
    stats.matvecs += nb;
    timer.mark(PhaseTimer::kMatvec);
  }

  void transform(T* M, int m, int p, std::vector<cplx> const& C) {
    std::vector<double2> c((size_t)m * p);
    for (size_t i = 0; i < c.size(); ++i) c[i] = make_double2(C[i].real(), C[i].imag());
    // (pageable source: the call returns once the data is staged, so `c` may go out of scope)
    CUDA_CHECK(cudaMemcpyAsync(transform_coeff.ptr, c.data(), c.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
    size_t smem = c.size() * sizeof(double2);
    auto launch = [&](auto kernel) {
      if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<grid, kThreads, smem, stream>>>(M, ld, m, p, transform_coeff.ptr, n);
    };
    if (m <= 4) launch(row_transform_kernel<T, 4>);
    else if (m <= 8) launch(row_transform_kernel<T, 8>);
    else if (m <= 16) launch(row_transform_kernel<T, 16>);
    else if (m <= 32) launch(row_transform_kernel<T, 32>);
    else launch(row_transform_kernel<T, kMaxBasis>);
    KERNEL_LAUNCHED();
  }

  int run(u64 k, double eps, int m_max, int b_max, int m_min, double* evals_out, void* evecs_out, double* rnorms_out,
          sped_monitor_fn monitor, void* mctx) {
    SPED_NVTX("sped_eigh");
    double t_start = now_seconds();
    Basis& B = *op.basis;
    Comm& cm = comm();
    n_global = B.n_states;
    op.prepare();
    n = op.dist.n_local;
    chunk = op.dist.chunk;
    if (k == 0 || k > n_global) fail(LS_INVALID_ARGUMENT, "number of eigenpairs must be in 1..dimension");
    // option defaults (PRIMME's meaning of the three sizes); everything is kept within kMaxBasis
    int b = b_max > 0 ? b_max : (int)std::min<u64>(k, 8);
    b = (int)std::min<u64>((u64)std::min(b, kMaxBasis / 2), n_global);
    int mmax = m_max > 0 ? m_max : std::max<int>(4 * b + 2 * (int)k, 24);
    mmax = (int)std::min<u64>((u64)std::min(mmax, kMaxBasis), n_global);
    if (mmax < (int)k + 1 && (u64)mmax < n_global) mmax = (int)std::min<u64>(n_global, k + 1 + (u64)b);
    mmax = std::min(mmax, kMaxBasis);
    if (mmax < (int)k && (u64)mmax < n_global)
      fail(LS_INVALID_ARGUMENT, "max_primme_basis_size is too small for the number of eigenpairs");
    if (b > mmax) b = mmax;
    int keep = m_min > 0 ? m_min : std::max<int>((int)k, mmax / 3);
    double const mach = (dtype == SPED_F32 || dtype == SPED_C64) ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    double const tol = eps > 0 ? eps : 1e4 * mach;

    stream = cm.active() ? cm.stream : nullptr;
    ld = std::max<u64>(n, 1);
    grid = persistent_grid(std::max<u64>(n, 1), kThreads, 4);
    int const kcap = (int)std::min<u64>(k, (u64)kMaxBasis);
    V.take(op, 0, ld * mmax);
    Wm.take(op, 1, ld * mmax);
    // One wanted pair and a basis of at most 8 vectors (the 40/42-spin decks: 3 and 4): when the basis
    // is full, restart and residual are one fused pass (restart_residual_kernel), the residual lands
    // unscaled in the first free column with its dot products against the restarted basis, and one
    // fused axpy + norm pass orthogonalises it; while the basis grows, the residual is written
    // straight into the next column.  No separate residual buffer exists in this mode (one vector
    // of 6.9 GB for chain_40 on one GPU).
    bool const single = k == 1 && mmax <= 8 && mmax >= 2 && (u64)mmax < n_global;
    resid.take(op, 2, single ? 1 : ld * std::max(1, kcap));
    if (cm.active()) xfull.take(op, 3, chunk * cm.world);
    partial.take(op, 4, (size_t)grid * kMaxBasis * kGroup);
    scal.take(op, 5, kMaxBasis);
    coeff.take(op, 6, (size_t)kMaxBasis * kMaxBasis);
    d_theta.take(op, 7, kMaxBasis);
    d_norms.take(op, 8, kMaxBasis);
    d_flag.take(op, 9, 1);
    transform_coeff.take(op, 10, (size_t)kMaxBasis * kMaxBasis);
    if (cm.active()) CUDA_CHECK(cudaMemsetAsync(xfull.ptr, 0, chunk * cm.world * sizeof(T), stream));
    timer.start(stream);

    std::vector<cplx> H((size_t)mmax * mmax, cplx(0, 0));  // projected matrix, row-major, leading dim mmax
    auto Hat = [&](int i, int j) -> cplx& { return H[(size_t)i * mmax + j]; };
    int m = 0;
    u64 seed = 0x5EED0002ull;
    auto append_random = [&](int at) {
      for (int attempt = 0; attempt < 8; ++attempt) {
        randomize(V.ptr + (u64)at * ld, seed);
        seed += 0x1000193ull;
        if (orthonormalize(at, V.ptr + (u64)at * ld) > 1e-8) return;
      }
      fail(SPED_INTERNAL_ERROR, "could not generate a new search direction");
    };
    // H[:, m_old:m_new) = V[:, 0:m_new)^H W[:, m_old:m_new), groups of kGroup columns per pass over V;
    // one device-to-host round trip for the whole block
    auto extend_projection = [&](int m_old, int m_new) {
      SPED_NVTX("sped_eigh: projection V^H H V");
      int off = 0;
      for (int j0 = m_old; j0 < m_new; j0 += kGroup) {
        int const nw = std::min(kGroup, m_new - j0);
        block_dots(V.ptr, m_new, Wm.ptr + (u64)j0 * ld, nw, coeff.ptr + off);
        off += m_new * nw;
      }
      auto h = to_host(coeff.ptr, off);
      off = 0;
      for (int j0 = m_old; j0 < m_new; j0 += kGroup) {
        int const nw = std::min(kGroup, m_new - j0);
        for (int i = 0; i < m_new; ++i)
          for (int c = 0; c < nw; ++c) {
            Hat(i, j0 + c) = h[off + i * nw + c];
            Hat(j0 + c, i) = std::conj(h[off + i * nw + c]);
          }
        off += m_new * nw;
      }
      for (int j = m_old; j < m_new; ++j) Hat(j, j) = Hat(j, j).real();
      timer.mark(PhaseTimer::kProject);
    };
    int const b0 = std::min<int>(mmax, std::max<int>(b, (int)std::min<u64>(k, (u64)mmax)));
    for (int j = 0; j < b0; ++j) {
      append_random(m);
      ++m;
    }
    timer.mark(PhaseTimer::kOrtho);
    apply(0, m);
    extend_projection(0, m);

    std::vector<double> theta;
    std::vector<cplx> S, S_prev;
    int m_prev = 0, n_prev = 0;
    std::vector<double> rn(k, 0.0), ev(k, 0.0);
    double a_norm = 0;
    int status = SPED_NOT_CONVERGED;
    int const max_outer = 200000;
    // Restart coefficients for nb new directions: C = [S(:, 0:r) | previous Ritz directions
    // orthogonalised against it] (the "+k" of GD+k), as columns over the current basis of m vectors.
    auto restart_columns = [&](int nb) {
      int r = std::min(std::max(keep, (int)std::min<u64>(k, (u64)m)), mmax - nb);
      r = std::max(r, 1);
      int p_room = mmax - nb - r;
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
      return cols;
    };
    // bookkeeping of a restart with the columns `cols`: H <- C^H H C, Ritz vectors become unit vectors
    auto restart_projection = [&](std::vector<std::vector<cplx>> const& cols) {
      int const p = (int)cols.size();
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
    };
    auto flatten = [&](std::vector<std::vector<cplx>> const& cols, int rows) {
      int const p = (int)cols.size();
      std::vector<cplx> C((size_t)rows * p);
      for (int q = 0; q < p; ++q)
        for (int j = 0; j < rows; ++j) C[(size_t)j * p + q] = cols[q][j];
      return C;
    };
    for (int it = 0; it < max_outer; ++it) {
      stats.iterations = it + 1;
      std::vector<cplx> Hm((size_t)m * m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Hm[(size_t)i * m + j] = Hat(i, j);
      small_eigh(m, Hm, theta, S);
      for (double t : theta) a_norm = std::max(a_norm, std::abs(t));
      int const kk = (int)std::min<u64>(k, (u64)m);
      // residuals of ALL wanted pairs, kGroup per pass over V and W, one round trip for their norms
      bool const have_all = m >= (int)k;
      bool fused = false;  // this iteration restarted inside the residual pass: r sits in V[:, m], unscaled
      if (single && m == mmax) {
        SPED_NVTX("sped_eigh: restart + residual (fused)");
        auto cols = restart_columns(1);
        int const p = (int)cols.size();
        std::vector<cplx> C = flatten(cols, m);
        std::vector<double2> c(C.size());
        for (size_t i = 0; i < c.size(); ++i) c[i] = make_double2(C[i].real(), C[i].imag());
        CUDA_CHECK(cudaMemcpyAsync(transform_coeff.ptr, c.data(), c.size() * sizeof(double2), cudaMemcpyHostToDevice, stream));
        if (m <= 4) restart_residual_kernel<T, 4><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, p, transform_coeff.ptr, theta[0], n, partial.ptr);
        else restart_residual_kernel<T, 8><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, p, transform_coeff.ptr, theta[0], n, partial.ptr);
        KERNEL_LAUNCHED();
        // scal[0] = |r|^2, scal[1 + q] = <V'_q, r>: left on the device for the orthogonalisation below
        finish_dot_kernel<<<1 + p, 32, 0, stream>>>(partial.ptr, grid, 1 + p, scal.ptr);
        KERNEL_LAUNCHED();
        comm_allreduce_sum_f64(reinterpret_cast<double*>(scal.ptr), 2 * (size_t)(1 + p), stream);
        auto r2 = to_host(scal.ptr, 1);
        rn[0] = std::sqrt(std::max(0.0, r2[0].real()));
        ev[0] = theta[0];
        restart_projection(cols);
        fused = true;
        timer.mark(PhaseTimer::kRestart);
      } else {
        SPED_NVTX("sped_eigh: residuals");
        std::vector<double2> sblock;
        std::vector<double> th;
        int off = 0;
        for (int q0 = 0; q0 < kk; q0 += kGroup) {
          int const kq = std::min(kGroup, kk - q0);
          sblock.assign((size_t)m * kq, make_double2(0, 0));
          th.assign(kq, 0.0);
          for (int q = 0; q < kq; ++q) {
            th[q] = theta[q0 + q];
            for (int j = 0; j < m; ++j) sblock[(size_t)j * kq + q] = make_double2(S[(size_t)j * m + q0 + q].real(), S[(size_t)j * m + q0 + q].imag());
          }
          CUDA_CHECK(cudaMemcpyAsync(coeff.ptr, sblock.data(), sizeof(double2) * sblock.size(), cudaMemcpyHostToDevice, stream));
          CUDA_CHECK(cudaMemcpyAsync(d_theta.ptr, th.data(), sizeof(double) * kq, cudaMemcpyHostToDevice, stream));
          T* out = single ? V.ptr + (u64)m * ld : resid.ptr + (u64)q0 * ld;  // single: m < mmax here, column m is free
          switch (kq) {
            case 1: residual_block_kernel<T, 1><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, coeff.ptr, d_theta.ptr, out, ld, n, partial.ptr); break;
            case 2: residual_block_kernel<T, 2><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, coeff.ptr, d_theta.ptr, out, ld, n, partial.ptr); break;
            case 3: residual_block_kernel<T, 3><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, coeff.ptr, d_theta.ptr, out, ld, n, partial.ptr); break;
            default: residual_block_kernel<T, 4><<<grid, kThreads, 0, stream>>>(V.ptr, Wm.ptr, ld, m, coeff.ptr, d_theta.ptr, out, ld, n, partial.ptr); break;
          }
          KERNEL_LAUNCHED();
          finish_dot_kernel<<<kq, 32, 0, stream>>>(partial.ptr, grid, kq, scal.ptr + off);
          KERNEL_LAUNCHED();
          off += kq;
        }
        comm_allreduce_sum_f64(reinterpret_cast<double*>(scal.ptr), 2 * (size_t)kk, stream);
        auto r2 = to_host(scal.ptr, kk);
        for (int i = 0; i < kk; ++i) {
          rn[i] = std::sqrt(std::max(0.0, r2[i].real()));
          ev[i] = theta[i];
        }
        timer.mark(PhaseTimer::kResidual);
      }
      std::vector<int> unconverged;
      int n_conv = 0;
      for (int i = 0; i < kk; ++i) {
        if (rn[i] <= tol * a_norm) ++n_conv;
        else unconverged.push_back(i);
      }
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
        SPED_NVTX("sped_eigh: restart");
        auto cols = restart_columns(nb);
        std::vector<cplx> C = flatten(cols, m);
        transform(V.ptr, m, (int)cols.size(), C);
        transform(Wm.ptr, m, (int)cols.size(), C);
        restart_projection(cols);
        timer.mark(PhaseTimer::kRestart);
      }
      // remember the current Ritz directions (for the "+k" part of the next restart)
      n_prev = (int)std::min<u64>((u64)std::max(1, std::min(b, (int)k)), (u64)m);
      m_prev = m;
      S_prev.assign((size_t)m * n_prev, cplx(0, 0));
      for (int q = 0; q < n_prev; ++q)
        for (int j = 0; j < m; ++j) S_prev[(size_t)j * n_prev + q] = S[(size_t)j * m + q];
      // expand: the first nb unconverged residuals, scaled to unit length, become columns m .. m+nb
      // of V and are orthonormalised by blocks on the device; a direction without a usable residual
      // (the basis is still smaller than the number of wanted pairs) is a random vector
      int const m_old = m;
      std::vector<int> random_cols;
      bool const fused_ok = fused && nb == 1 && !unconverged.empty() && rn[0] > 0;
      for (int q = 0; q < nb && !fused_ok; ++q) {
        T* dst = V.ptr + (u64)(m_old + q) * ld;
        if (q < (int)unconverged.size() && rn[unconverged[q]] > 0) {
          // (single: the residual already sits in its column and is scaled in place)
          T const* src = (single || fused) ? dst : resid.ptr + (u64)unconverged[q] * ld;
          copy_scaled_kernel<T><<<grid, kThreads, 0, stream>>>(src, ld, dst, ld, n, 1.0 / rn[unconverged[q]]);
          KERNEL_LAUNCHED();
        } else {
          random_cols.push_back(q);
        }
      }
      if (fused_ok) {
        // r (unscaled, in column m_old) minus its components along the restarted basis -- the dot
        // products came out of the fused pass -- and its norm in one pass; then the scaling.  The
        // residual of a Ritz pair is orthogonal to the basis up to rounding, so this first sweep
        // keeps nearly all of it; should it keep less than half, the gated second sweep runs.
        SPED_NVTX("sped_eigh: orthogonalisation (fused)");
        T* w = V.ptr + (u64)m_old * ld;
        CUDA_CHECK(cudaMemsetAsync(d_flag.ptr, 0, sizeof(int), stream));
        axpy_norm_kernel<T><<<grid, kThreads, 0, stream>>>(V.ptr, ld, m_old, scal.ptr + 1, w, n, partial.ptr);
        KERNEL_LAUNCHED();
        finish_dot_kernel<<<1, 32, 0, stream>>>(partial.ptr, grid, 1, scal.ptr + 32);
        KERNEL_LAUNCHED();
        comm_allreduce_sum_f64(reinterpret_cast<double*>(scal.ptr + 32), 2, stream);
        scale_rel_kernel<T><<<grid, kThreads, 0, stream>>>(w, n, scal.ptr + 32, scal.ptr, 1e-24, d_norms.ptr, d_flag.ptr);
        KERNEL_LAUNCHED();
        ortho_sweep(m_old, w, 1, nullptr, nullptr, d_flag.ptr);
      } else if (!random_cols.empty()) {
        // rare: place the residual-based directions first, then the random ones sequentially
        std::vector<int> res_cols;
        for (int q = 0; q < nb; ++q)
          if (std::find(random_cols.begin(), random_cols.end(), q) == random_cols.end()) res_cols.push_back(q);
        int at = m_old;
        for (int q : res_cols) {
          if (m_old + q != at)
            CUDA_CHECK(cudaMemcpyAsync(V.ptr + (u64)at * ld, V.ptr + (u64)(m_old + q) * ld, n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
          ++at;
        }
        int const n_res = (int)res_cols.size();
        for (int g0 = 0; g0 < n_res; g0 += kGroup)
          ortho_block(m_old + g0, V.ptr + (u64)(m_old + g0) * ld, std::min(kGroup, n_res - g0), g0);
        for (int q = n_res; q < nb; ++q) append_random(m_old + q);
      } else {
        for (int g0 = 0; g0 < nb; g0 += kGroup)
          ortho_block(m_old + g0, V.ptr + (u64)(m_old + g0) * ld, std::min(kGroup, nb - g0), g0);
      }
      m = m_old + nb;
      timer.mark(PhaseTimer::kOrtho);
      apply(m_old, nb);
      // breakdown check rides on the projection's round trip: norms recorded by the block sweeps
      std::vector<double> kept(nb, 1.0);
      int const n_checked = nb - (int)random_cols.size();
      if (n_checked > 0)
        CUDA_CHECK(cudaMemcpyAsync(kept.data(), d_norms.ptr, sizeof(double) * n_checked, cudaMemcpyDeviceToHost, stream));
      extend_projection(m_old, m);
      bool redo = false;
      for (int q = 0; q < n_checked; ++q)
        if (!(kept[q] > 1e-24)) {  // linearly dependent direction (stored as zeros): replace it
          append_random(m_old + q);
          redo = true;
        }
      if (redo) {
        timer.mark(PhaseTimer::kOrtho);
        apply(m_old, nb);
        stats.matvecs -= nb;  // counted once
        extend_projection(m_old, m);
      }
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
      // eigenvalues and residual norms of the same (last evaluated) Ritz pairs
      evals_out[i] = i < (u64)kk ? (status == LS_SUCCESS ? theta[i] : ev[i]) : 0.0;
      rnorms_out[i] = i < (u64)kk ? rn[i] : 0.0;
    }
    if (evecs_out) {
      T* host = static_cast<T*>(evecs_out);
      if (!cm.active()) {
        // column by column: a 2-D copy would need a pitch of n_global * sizeof(T) bytes, which CUDA
        // limits to 2^31 - 1 (sectors above ~2.7e8 f64 rows)
        for (int q = 0; q < kk; ++q)
          CUDA_CHECK(cudaMemcpyAsync(host + (u64)q * n_global, V.ptr + (u64)q * ld, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        sync();
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
    timer.mark(PhaseTimer::kOther);
    double phase[PhaseTimer::kPhases] = {0, 0, 0, 0, 0, 0};
    timer.resolve(phase);
    stats.seconds_matvec = phase[PhaseTimer::kMatvec];
    stats.seconds_ortho = phase[PhaseTimer::kOrtho];
    stats.seconds_residual = phase[PhaseTimer::kResidual];
    stats.seconds_restart = phase[PhaseTimer::kRestart];
    stats.seconds_project = phase[PhaseTimer::kProject];
    stats.seconds_total = now_seconds() - t_start;
    op.last_stats = stats;
    return status;
  }
};

template <class T>
int run_solver(Operator& op, int dtype, u64 k, double eps, int m_max, int b_max, int m_min, double* evals, void* evecs,
               double* rnorms, sped_monitor_fn monitor, void* ctx) {
  Solver<T> s(op, dtype);
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
