// Shared host-side helpers for the C-ABI translation units: error string, argument checks.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../include/tic_b200.h"

namespace tic {

void set_error(const char* fmt, ...);

#define TIC_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::tic::set_error(__VA_ARGS__);    \
      return TIC_E_ARG;                 \
    }                                   \
  } while (0)

#define TIC_CHECK_LAUNCH(name)                                                 \
  do {                                                                         \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) {                                                  \
      ::tic::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return TIC_E_LAUNCH;                                                     \
    }                                                                          \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The small-batch step is a chain of ~10 latency-bound kernels.  Every kernel of the library calls pdl_trigger() first
// (its successor in the stream may be launched and run its prologue: barrier init, TMEM alloc, descriptor prefetch) and
// pdl_wait() before it touches global memory (blocks until every predecessor grid has completed and flushed), and is
// launched with the programmatic-stream-serialization attribute.  Under stream capture the edge becomes a programmatic
// dependency of the CUDA graph.  MEASURED (B200, c2 step as one multi-branch CUDA graph): 0.128 ms with PDL edges vs 0.112 ms
// with ordinary edges — waiting CTAs of later kernels take SM slots from the parallel branches — so it is OFF by default;
// TIC_PDL=1 enables it (results are identical either way).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
int set_pdl_override(int mode);

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace tic
