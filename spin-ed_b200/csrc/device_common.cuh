// device_common.cuh -- helpers shared by the statically compiled sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include "internal.h"

namespace sped {

constexpr int kThreads = 256;
constexpr int kSmCount = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of it

inline int persistent_grid(u64 work_items, int threads, int blocks_per_sm) {
  u64 need = (work_items + threads - 1) / threads;
  u64 cap = (u64)kSmCount * blocks_per_sm;
  return (int)std::max<u64>(1, std::min<u64>(need, cap));
}

// Bytes needed to stage a program in shared memory (16-byte aligned sections).
template <class W>
inline size_t program_smem_bytes(u32 n_steps, u32 n_ops) {
  auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
  return up((size_t)n_steps * sizeof(FastStep<W>)) + up((size_t)n_steps * sizeof(PermStep)) +
         up((size_t)n_ops * sizeof(PermOp<W>)) + up((size_t)n_steps * sizeof(std::int32_t));
}

#if defined(__CUDACC__)
// Cooperative copy of the program into shared memory; returns a view addressing the copy (always
// shared memory, so the interpreter's loads are LDS).  Must be called by all threads of the block;
// ends with __syncthreads().
template <class W>
__device__ __forceinline__ ProgramView<W> stage_program(ProgramView<W> g, unsigned char* smem) {
  auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
  FastStep<W>* fast = reinterpret_cast<FastStep<W>*>(smem);
  PermStep* steps = reinterpret_cast<PermStep*>(smem + up((size_t)g.n_steps * sizeof(FastStep<W>)));
  PermOp<W>* ops = reinterpret_cast<PermOp<W>*>(reinterpret_cast<unsigned char*>(steps) +
                                                up((size_t)g.n_steps * sizeof(PermStep)));
  std::int32_t* phase = reinterpret_cast<std::int32_t*>(reinterpret_cast<unsigned char*>(ops) +
                                                        up((size_t)g.n_ops * sizeof(PermOp<W>)));
  for (u32 i = threadIdx.x; i < g.n_steps; i += blockDim.x) {
    fast[i] = g.fast[i];
    steps[i] = g.steps[i];
    phase[i] = g.phase[i];
  }
  for (u32 i = threadIdx.x; i < g.n_ops; i += blockDim.x) ops[i] = g.ops[i];
  __syncthreads();
  ProgramView<W> v = g;
  v.fast = fast;
  v.steps = steps;
  v.ops = ops;
  v.phase = phase;
  return v;
}
#endif

}  // namespace sped
