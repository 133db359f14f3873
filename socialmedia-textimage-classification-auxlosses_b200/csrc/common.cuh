// Shared host-side helpers for the C-ABI translation units: error string, argument checks.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdint>
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

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace tic
