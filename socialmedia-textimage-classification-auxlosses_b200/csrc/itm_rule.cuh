// The uniform ITM sampling rule of models/mm_late.py:389-414 on supplied uniforms (integer arithmetic, bit-exact):
// coin < 0.5 -> mismatch (label 0) with a uniform pick among the other B-1 rows; else match (label 1, the row itself).
#pragma once
#include <cuda_runtime.h>

namespace tic {

__device__ __forceinline__ void uniform_rule(const float* u_coin, const float* u_pick, int B, int i, int& label, int& src) {
  label = 1;
  src = i;
  if (B > 1 && u_coin[i] < 0.5f) {
    label = 0;
    int k = static_cast<int>(floorf(__fmul_rn(u_pick[i], static_cast<float>(B - 1))));
    k = min(k, B - 2);
    src = k < i ? k : k + 1;
  }
}

}  // namespace tic
