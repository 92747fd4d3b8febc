// eigh_kernels.cuh -- device code of the eigensolver that is shared with the single-threaded host
// emulation of the test suite (emul.cpp compiles it with the host compiler, the CUDA built-ins
// shimmed): the storage traits and arithmetic helpers, the block reduction, and the fused kernels of
// the single-pair / small-basis iteration (restart + residual, axpy + norm, relative scaling).  The
// other kernels of the solver live in eigh.cu.
#pragma once
#include "device_types.h"

#if !defined(SPED_KERNEL_THREADS)
#define SPED_KERNEL_THREADS 256
#endif
#if !defined(SPED_KERNEL_LINKAGE)
#define SPED_KERNEL_LINKAGE
#endif

namespace sped {
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
  __shared__ double2 red[SPED_KERNEL_THREADS / 32];
  for (int o = 16; o; o >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, o);
    v.y += __shfl_down_sync(0xffffffffu, v.y, o);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double2 s = make_double2(0, 0);
    for (int w = 0; w < SPED_KERNEL_THREADS / 32; ++w) {
      s.x += red[w].x;
      s.y += red[w].y;
    }
    *dst = s;
  }
  __syncthreads();
}

// value as it will be read back from storage (single-precision storage rounds)
template <class T> __device__ __forceinline__ typename VT<T>::Acc stored(typename VT<T>::Acc v) { return v; }
template <> __device__ __forceinline__ double stored<float>(double v) { return (double)(float)v; }
template <> __device__ __forceinline__ double2 stored<float2>(double2 v) { return make_double2((double)(float)v.x, (double)(float)v.y); }

// Restart and residual in ONE pass over the basis (basis full, one wanted pair -- every iteration of
// the 40-spin decks, whose basis holds three vectors): rows of V and W (n x m) <- rows * C (m x p),
// in place, and column p of V <- r = sum_j (W_j - theta V_j) C[j][0], the residual of the first Ritz
// pair (its coefficients are column 0 of C), unscaled: the next search direction.  Also
// partial[b] = sum |r|^2 and partial[(1 + q) * grid + b] = sum conj(V'_q) r for q < p, so that the
// orthogonalisation of r needs no pass of its own for the dot products.  Same arithmetic, in the same
// order, as residual_block_kernel followed by row_transform_kernel on V and on W -- 6 column reads and
// 5 writes for a 3-vector basis instead of 12 and 5.
template <class T, int MAXM>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) restart_residual_kernel(T* V, T* W, u64 ld, int m, int p, double2 const* C, double theta,
                                                                    u64 n, double2* partial) {
  using A = typename VT<T>::Acc;
  __shared__ double2 c[MAXM * MAXM];
  for (int j = threadIdx.x; j < m * p; j += blockDim.x) c[j] = C[j];
  __syncthreads();
  double2 nrm = make_double2(0, 0);
  double2 dots[MAXM - 1];
#pragma unroll
  for (int q = 0; q < MAXM - 1; ++q) dots[q] = make_double2(0, 0);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A v[MAXM], w[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j)
      if (j < m) {
        v[j] = VT<T>::load(V + (u64)j * ld + i);
        w[j] = VT<T>::load(W + (u64)j * ld + i);
      }
    A r = from2<A>(make_double2(0, 0));
#pragma unroll
    for (int j = 0; j < MAXM; ++j)
      if (j < m) r = addv(r, mulc(subv(w[j], mulc(v[j], make_double2(theta, 0))), c[j * p]));
    r = stored<T>(r);
#pragma unroll
    for (int q = 0; q < MAXM - 1; ++q)
      if (q < p) {
        A vq = from2<A>(make_double2(0, 0)), wq = from2<A>(make_double2(0, 0));
#pragma unroll
        for (int j = 0; j < MAXM; ++j)
          if (j < m) {
            vq = addv(vq, mulc(v[j], c[j * p + q]));
            wq = addv(wq, mulc(w[j], c[j * p + q]));
          }
        VT<T>::store(V + (u64)q * ld + i, vq);
        VT<T>::store(W + (u64)q * ld + i, wq);
        dot_acc(dots[q], stored<T>(vq), r);
      }
    VT<T>::store(V + (u64)p * ld + i, r);
    dot_acc(nrm, r, r);
  }
  block_reduce_store(nrm, partial + blockIdx.x);
#pragma unroll
  for (int q = 0; q < MAXM - 1; ++q)
    if (q < p) block_reduce_store(dots[q], partial + (u64)(1 + q) * gridDim.x + blockIdx.x);
}

// w -= sum_j coeff[j] V_j (j < m) and partial[block] = sum |w|^2 of the result, in one pass
template <class T>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) axpy_norm_kernel(T const* V, u64 ld, int m, double2 const* coeff, T* w, u64 n,
                                                             double2* partial) {
  using A = typename VT<T>::Acc;
  __shared__ double2 c[kMaxBasis];
  for (int j = threadIdx.x; j < m; j += blockDim.x) c[j] = coeff[j];
  __syncthreads();
  double2 nrm = make_double2(0, 0);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A acc = VT<T>::load(w + i);
    for (int j = 0; j < m; ++j) acc = subv(acc, mulc(VT<T>::load(V + (u64)j * ld + i), c[j]));
    acc = stored<T>(acc);
    VT<T>::store(w + i, acc);
    dot_acc(nrm, acc, acc);
  }
  block_reduce_store(nrm, partial + blockIdx.x);
}

// w *= 1 / sqrt(norm2), or 0 when norm2 <= tiny * ref2 (linearly dependent on the basis).
// record[0] = norm2 / ref2 = the share of the direction that survived the orthogonalisation; *flag is
// raised when that is less than half (DGKS criterion for a second sweep).
template <class T>
SPED_KERNEL_LINKAGE __global__ void __launch_bounds__(SPED_KERNEL_THREADS) scale_rel_kernel(T* w, u64 n, double2 const* norm2, double2 const* ref2, double tiny,
                                                             double* record, int* flag) {
  double const v = norm2[0].x, ref = ref2[0].x;
  double const kept = ref > 0 ? v / ref : 0.0;
  double const s = kept > tiny ? rsqrt(v) : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *record = kept;
    if (kept < 0.5) *flag = 1;
  }
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    VT<T>::store(w + i, mulc(VT<T>::load(w + i), make_double2(s, 0)));
}

}  // namespace
}  // namespace sped
