// canon.cuh -- canonicalisation functors of the statically compiled kernels (the run-time
// specialised kernels use JitCanon from matvec_kernel.cuh instead).
#pragma once
#include "device_common.cuh"
#include "matvec_kernel.cuh"

namespace sped {

struct TrivialCanon {
  static constexpr bool symmetric = false;
  __device__ __forceinline__ void operator()(u64, u64&, int&) const {}
};

// Interprets the group program staged in shared memory (permprog.h).
template <class W>
struct ProgramCanon {
  static constexpr bool symmetric = true;
  ProgramView<W> P;
  __device__ __forceinline__ void operator()(u64 x, u64& rep, int& phase) const {
    W r;
    u32 step, flipped;
    canonicalize<W>(P, (W)x, r, step, flipped);
    rep = r;
    phase = element_phase<W>(P, step, flipped);
  }
};

}  // namespace sped
