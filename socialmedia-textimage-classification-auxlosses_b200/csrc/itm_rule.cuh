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

// Bit-reproducible fp32 exp for x <= 0, mirrored op-for-op by oracle/restatement.py:det_exp_f32 — every step is a single
// IEEE-754 round-to-nearest operation (no FMA contraction, no library transcendental).
//
// Written for the FMA / ALU pipes only (this runs once per similarity element inside the tile epilogues, next to one
// ex2.approx per element on the XU pipe): round-to-nearest-even by the 1.5 * 2^23 magic constant (two exact fp32 adds,
// identical to rintf for |v| < 2^22), the integer n read off the magic sum's mantissa, the 2^n scaling by an exponent add.
__device__ __forceinline__ float det_exp(float x) {
  x = fmaxf(x, -80.0f);
  const float t = __fadd_rn(__fmul_rn(x, 1.44269504088896341f), 12582912.0f);
  const float n = __fsub_rn(t, 12582912.0f);                 // == rintf(x * log2e)
  const int ni = __float_as_int(t) - 0x4B400000;             // == (int)n
  float r = __fsub_rn(x, __fmul_rn(n, 0.693359375f));
  r = __fsub_rn(r, __fmul_rn(n, -2.12194440e-4f));
  float p = 1.9875691500e-4f;
  p = __fadd_rn(__fmul_rn(p, r), 1.3981999507e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 8.3334519073e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 4.1665795894e-2f);
  p = __fadd_rn(__fmul_rn(p, r), 1.6666665459e-1f);
  p = __fadd_rn(__fmul_rn(p, r), 5.0000001201e-1f);
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(p, __fmul_rn(r, r)), r), 1.0f);
  return __int_as_float(__float_as_int(y) + (ni << 23));  // exact scaling by 2^n (result stays normal)
}

// Hard-negative sampling weight of one logit against the FIXED reference `ref` >= max S (oracle/restatement.py:
// hard_qweights): q = trunc(det_exp(min(s - ref, 0)) * 2^40).  Integer weights make every prefix sum associative, the fixed
// reference lets the similarity tiles sum them in any order without knowing the row maximum first.
__device__ __forceinline__ unsigned long long hard_qweight(float s, float ref) {
  // trunc(f * 2^40) for the normal fp32 f in (0, 1] by shifting its 24-bit significand: no float->int64 conversion
  // instruction (those share the quarter-rate XU pipe with the exponentials of the softmax statistics)
  const int b = __float_as_int(det_exp(fminf(__fsub_rn(s, ref), 0.0f)));
  const int sh = (b >> 23) - 127 + 17;                                  // f = m * 2^(e-23), m in [2^23, 2^24)  ->  q = m * 2^(e+17)
  const unsigned long long m = static_cast<unsigned long long>((b & 0x7FFFFF) | 0x800000);
  return sh >= 0 ? (m << sh) : (sh > -24 ? (m >> (-sh)) : 0ull);
}
// target = floor(U * total / 2^24), U = trunc(u_pick * 2^24)  (no 128-bit product needed: total < 2^60)
__device__ __forceinline__ unsigned long long hard_target(float u_pick, unsigned long long total) {
  const unsigned long long U = static_cast<unsigned long long>(__fmul_rn(u_pick, 16777216.0f));
  return U * (total >> 24) + ((U * (total & 0xFFFFFFull)) >> 24);
}

// One similarity logit from a tensor-core accumulator value (raw <t_i, v_j>), the two inverse norms and the scale.  The
// forward tiles (when they materialise logits or sum hard-negative weights) and the pick tiles use THIS expression, so
// that a recomputed tile reproduces the logits bit for bit.
__device__ __forceinline__ float itc_logit(float acc, float rinv_t, float rinv_v, float scale) {
  return __fmul_rn(__fmul_rn(__fmul_rn(acc, rinv_t), rinv_v), scale);
}

}  // namespace tic
