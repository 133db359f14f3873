// Fusion-variant kernels that are not plain GEMMs (models/mm_late.py:91-144):
//   attention : the HBM-bound middle of the CLS-row collapse — one streaming pass over x_v per direction, shared by
//               the main and the ITM pass (mm_late.py:98-113,195-210; only ctx[:,0,:] is consumed, :111)
//   aspect-att: tanh-scored 2-way softmax over the (scrambled) pooled pair (mm_late.py:115-131)
//   gmu       : sigmoid gate of the raw concatenation (mm_late.py:133-144)
#include <cstdlib>
#include "common.cuh"
#include "tic_ptx.cuh"

namespace tic {

constexpr int kAttnWarps = 8;
constexpr int kAttnMaxLv = 1024;

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

// One CTA per sample. Each lane owns NV 8-element slices of the E axis (E = 256*NV); each warp walks tokens
// j = warp, warp+8, ... with an online softmax, NPASS query vectors at once so x_v is read from HBM exactly once.
template <int NV, int NPASS>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_pool_fwd_kernel(const __nv_bfloat16* __restrict__ xv, int64_t bstride, int64_t tstride, const float* __restrict__ kq,
                     int64_t ldkq, int B, int Lv, float scale, __nv_bfloat16* __restrict__ xbar_b,
                     __nv_bfloat16* __restrict__ xbar_lo, int64_t ld_xb,
                     float* __restrict__ xbar_f, int64_t ld_xf, float* __restrict__ attn, int64_t ld_attn) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  constexpr int E = NV * 256;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ float sm[];
  float* s_scores = sm;                                   // [NPASS][Lv]
  float* s_ml = s_scores + NPASS * Lv;                    // [NPASS][warps][2]
  float* s_acc = s_ml + NPASS * kAttnWarps * 2;           // [warps][E] (reused per pass)
  float q[NPASS][NV][8], cb[NPASS];
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    const float* kr = kq + static_cast<int64_t>(p * B + b) * ldkq;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) q[p][v][k] = __ldg(kr + v * 256 + lane * 8 + k);
    cb[p] = __ldg(kr + E);  // augmented column: <q0, b_K>
  }
  float m[NPASS], l[NPASS], acc[NPASS][NV][8];
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    m[p] = -INFINITY;
    l[p] = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[p][v][k] = 0.f;
  }
  const __nv_bfloat16* xb = xv + static_cast<int64_t>(b) * bstride;
  for (int j = warp; j < Lv; j += kAttnWarps) {
    float x[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v) load8_bf16(xb + static_cast<int64_t>(j) * tstride + v * 256 + lane * 8, x[v]);
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      float d = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) d = fmaf(q[p][v][k], x[v][k], d);
      d = (warp_sum(d) + cb[p]) * scale;
      if (lane == 0) s_scores[p * Lv + j] = d;
      const float mn = fmaxf(m[p], d);
      const float corr = __expf(m[p] - mn), w = __expf(d - mn);
      l[p] = l[p] * corr + w;
      m[p] = mn;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][v][k] = fmaf(acc[p][v][k], corr, w * x[v][k]);
    }
  }
  // combine the 8 warps' (m, l, acc)
#pragma unroll
  for (int p = 0; p < NPASS; ++p)
    if (lane == 0) { s_ml[(p * kAttnWarps + warp) * 2] = m[p]; s_ml[(p * kAttnWarps + warp) * 2 + 1] = l[p]; }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    float M = -INFINITY;
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, s_ml[(p * kAttnWarps + w) * 2]);
    float L = 0.f;
    for (int w = 0; w < kAttnWarps; ++w) L += s_ml[(p * kAttnWarps + w) * 2 + 1] * __expf(s_ml[(p * kAttnWarps + w) * 2] - M);
    const float mine = (m[p] == -INFINITY) ? 0.f : __expf(m[p] - M) / L;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) s_acc[warp * E + v * 256 + lane * 8 + k] = acc[p][v][k] * mine;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += kAttnWarps * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) s += s_acc[w * E + e];
      const int64_t r = static_cast<int64_t>(p * B + b);
      if (xbar_b) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(s);
        xbar_b[r * ld_xb + e] = hi;
        if (xbar_lo) xbar_lo[r * ld_xb + e] = __float2bfloat16_rn(s - __bfloat162float(hi));
      }
      if (xbar_f) xbar_f[r * ld_xf + e] = s;
    }
    for (int j = threadIdx.x; j < Lv; j += kAttnWarps * 32)
      attn[static_cast<int64_t>(p * B + b) * ld_attn + j] = __expf(s_scores[p * Lv + j] - M) / L;
    __syncthreads();
  }
}

// Backward of the pooling w.r.t. the (augmented) query kq: second streaming pass over x_v.
//   t_j = <dxbar, x_v[j]>,  D = <dxbar, xbar>,  ds_j = a_j (t_j - D) scale,  dkq = sum_j ds_j x_v[j],  dc = sum_j ds_j
template <int NV, int NPASS>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_pool_bwd_kernel(const __nv_bfloat16* __restrict__ xv, int64_t bstride, int64_t tstride, const float* __restrict__ attn,
                     int64_t ld_attn, const float* __restrict__ dxbar, int64_t ld_dxb, const float* __restrict__ xbar_f,
                     int64_t ld_xf, int B, int Lv, float scale, __nv_bfloat16* __restrict__ dkq,
                     __nv_bfloat16* __restrict__ dkq_lo, int64_t ld_dkq) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  constexpr int E = NV * 256;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ float sm[];
  float* s_acc = sm;                       // [warps][E]
  float* s_dc = s_acc + kAttnWarps * E;    // [NPASS][warps]
  float g[NPASS][NV][8], D[NPASS], acc[NPASS][NV][8], dc[NPASS];
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    const int64_t r = static_cast<int64_t>(p * B + b);
    float d = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = v * 256 + lane * 8 + k;
        g[p][v][k] = dxbar[r * ld_dxb + e];
        d = fmaf(g[p][v][k], xbar_f[r * ld_xf + e], d);
        acc[p][v][k] = 0.f;
      }
    D[p] = warp_sum(d);
    dc[p] = 0.f;
  }
  const __nv_bfloat16* xb = xv + static_cast<int64_t>(b) * bstride;
  for (int j = warp; j < Lv; j += kAttnWarps) {
    float x[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v) load8_bf16(xb + static_cast<int64_t>(j) * tstride + v * 256 + lane * 8, x[v]);
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      float t = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) t = fmaf(g[p][v][k], x[v][k], t);
      t = warp_sum(t);
      const float ds = __ldg(attn + static_cast<int64_t>(p * B + b) * ld_attn + j) * (t - D[p]) * scale;
      dc[p] += ds;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][v][k] = fmaf(ds, x[v][k], acc[p][v][k]);
    }
  }
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) s_acc[warp * E + v * 256 + lane * 8 + k] = acc[p][v][k];
    if (lane == 0) s_dc[p * kAttnWarps + warp] = dc[p];
    __syncthreads();
    const int64_t r = static_cast<int64_t>(p * B + b);
    for (int e = threadIdx.x; e < E; e += kAttnWarps * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) s += s_acc[w * E + e];
      const __nv_bfloat16 hi = __float2bfloat16_rn(s);
      dkq[r * ld_dkq + e] = hi;
      if (dkq_lo) dkq_lo[r * ld_dkq + e] = __float2bfloat16_rn(s - __bfloat162float(hi));
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < kAttnWarps; ++w) s += s_dc[p * kAttnWarps + w];
      const __nv_bfloat16 hi = __float2bfloat16_rn(s);
      dkq[r * ld_dkq + E] = hi;  // augmented column: dc
      if (dkq_lo) dkq_lo[r * ld_dkq + E] = __float2bfloat16_rn(s - __bfloat162float(hi));
      for (int e = E + 1; e < E + 8; ++e) {
        dkq[r * ld_dkq + e] = __float2bfloat16_rn(0.f);
        if (dkq_lo) dkq_lo[r * ld_dkq + e] = __float2bfloat16_rn(0.f);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ attention pool, bulk-async pipeline (v2)
// Same arithmetic as the kernels above, but x_v is streamed by a dedicated producer warp with cp.async.bulk (the TMA
// engine's 1-D form) into a 6-stage shared-memory ring, 16 token rows (24 KB) per stage, completion on mbarriers.
// One persistent CTA per SM keeps ~144 KB in flight regardless of register pressure, which is what the HBM system
// needs (the register-staged v1 ran at 12 % warp occupancy and ~25 % of HBM peak).  Requires contiguous token rows.
constexpr int kApRows = 16, kApStages = 6, kApE = 768;
constexpr int kApChunkBytes = kApRows * kApE * 2;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void lds8_bf16(const uint8_t* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

struct ApSmem {
  uint8_t* ring;
  uint32_t ring_u32, full0, empty0;
  float* fl;
};
__device__ __forceinline__ ApSmem ap_carve(uint8_t* raw) {
  ApSmem s;
  const uint32_t base = (smem_u32(raw) + 127u) & ~127u;
  s.ring = raw + (base - smem_u32(raw));
  s.ring_u32 = base;
  s.full0 = base + kApStages * kApChunkBytes;
  s.empty0 = s.full0 + 8 * kApStages;
  s.fl = reinterpret_cast<float*>(s.ring + kApStages * kApChunkBytes + 16 * kApStages);
  return s;
}
__device__ __forceinline__ void ap_producer(const ApSmem& sm, const __nv_bfloat16* xv, int64_t bstride, int B, int Lv) {
  const int nchunks = (Lv + kApRows - 1) / kApRows;
  uint32_t it = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const __nv_bfloat16* xb = xv + static_cast<int64_t>(b) * bstride;
    for (int c = 0; c < nchunks; ++c, ++it) {
      const uint32_t st = it % kApStages, ph = (it / kApStages) & 1u;
      mbar_wait(sm.empty0 + 8 * st, ph ^ 1u);
      const int rows = min(kApRows, Lv - c * kApRows);
      const uint32_t bytes = static_cast<uint32_t>(rows) * kApE * 2;
      mbar_arrive_expect_tx(sm.full0 + 8 * st, bytes);
      bulk_g2s(sm.ring_u32 + st * kApChunkBytes, xb + static_cast<int64_t>(c) * kApRows * kApE, bytes, sm.full0 + 8 * st);
    }
  }
}

template <int NPASS>
__global__ void __launch_bounds__(288, 1)
attn_pool_fwd_v2_kernel(const __nv_bfloat16* __restrict__ xv, int64_t bstride, const float* __restrict__ kq, int64_t ldkq, int B,
                        int Lv, float scale, __nv_bfloat16* __restrict__ xbar_b, __nv_bfloat16* __restrict__ xbar_lo,
                        int64_t ld_xb, float* __restrict__ xbar_f, int64_t ld_xf, float* __restrict__ attn, int64_t ld_attn) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  constexpr int NV = 3, E = kApE;
  extern __shared__ uint8_t ap_raw[];
  const ApSmem sm = ap_carve(ap_raw);
  float* s_scores = sm.fl;                          // [NPASS][Lv]
  float* s_ml = s_scores + NPASS * Lv;              // [NPASS][8][2]
  float* s_acc = s_ml + NPASS * 16;                 // [8][E]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kApStages; ++i) { mbar_init(sm.full0 + 8 * i, 1); mbar_init(sm.empty0 + 8 * i, 8); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) ap_producer(sm, xv, bstride, B, Lv);
    return;
  }
  const int nchunks = (Lv + kApRows - 1) / kApRows;
  uint32_t it = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float q[NPASS][NV][8], cb[NPASS], m[NPASS], l[NPASS], acc[NPASS][NV][8];
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      const float* kr = kq + static_cast<int64_t>(p * B + b) * ldkq;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(kr + v * 256 + lane * 8));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(kr + v * 256 + lane * 8 + 4));
        q[p][v][0] = a.x; q[p][v][1] = a.y; q[p][v][2] = a.z; q[p][v][3] = a.w;
        q[p][v][4] = c4.x; q[p][v][5] = c4.y; q[p][v][6] = c4.z; q[p][v][7] = c4.w;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][v][k] = 0.f;
      }
      cb[p] = __ldg(kr + E);
      m[p] = -INFINITY;
      l[p] = 0.f;
    }
    for (int c = 0; c < nchunks; ++c, ++it) {
      const uint32_t st = it % kApStages, ph = (it / kApStages) & 1u;
      mbar_wait(sm.full0 + 8 * st, ph);
      const uint8_t* chunk = sm.ring + st * kApChunkBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = c * kApRows + warp + 8 * h;
        if (j < Lv) {
          float x[NV][8];
#pragma unroll
          for (int v = 0; v < NV; ++v) lds8_bf16(chunk + (warp + 8 * h) * (E * 2) + (v * 256 + lane * 8) * 2, x[v]);
#pragma unroll
          for (int p = 0; p < NPASS; ++p) {
            float d = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
              for (int k = 0; k < 8; ++k) d = fmaf(q[p][v][k], x[v][k], d);
            d = (warp_sum(d) + cb[p]) * scale;
            if (lane == 0) s_scores[p * Lv + j] = d;
            const float mn = fmaxf(m[p], d);
            const float corr = __expf(m[p] - mn), w = __expf(d - mn);
            l[p] = l[p] * corr + w;
            m[p] = mn;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[p][v][k] = fmaf(acc[p][v][k], corr, w * x[v][k]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty0 + 8 * st);
    }
    // combine the 8 consumer warps (the producer keeps prefetching the next samples meanwhile)
#pragma unroll
    for (int p = 0; p < NPASS; ++p)
      if (lane == 0) { s_ml[(p * 8 + warp) * 2] = m[p]; s_ml[(p * 8 + warp) * 2 + 1] = l[p]; }
    consumer_bar();
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      float M = -INFINITY;
      for (int w = 0; w < 8; ++w) M = fmaxf(M, s_ml[(p * 8 + w) * 2]);
      float L = 0.f;
      for (int w = 0; w < 8; ++w) L += s_ml[(p * 8 + w) * 2 + 1] * __expf(s_ml[(p * 8 + w) * 2] - M);
      const float mine = (m[p] == -INFINITY) ? 0.f : __expf(m[p] - M) / L;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) s_acc[warp * E + v * 256 + lane * 8 + k] = acc[p][v][k] * mine;
      consumer_bar();
      const int64_t r = static_cast<int64_t>(p * B + b);
      for (int e = threadIdx.x; e < E; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_acc[w * E + e];
        if (xbar_b) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(s);
          xbar_b[r * ld_xb + e] = hi;
          if (xbar_lo) xbar_lo[r * ld_xb + e] = __float2bfloat16_rn(s - __bfloat162float(hi));
        }
        if (xbar_f) xbar_f[r * ld_xf + e] = s;
      }
      for (int j = threadIdx.x; j < Lv; j += 256) attn[r * ld_attn + j] = __expf(s_scores[p * Lv + j] - M) / L;
      consumer_bar();
    }
  }
}

template <int NPASS>
__global__ void __launch_bounds__(288, 1)
attn_pool_bwd_v2_kernel(const __nv_bfloat16* __restrict__ xv, int64_t bstride, const float* __restrict__ attn, int64_t ld_attn,
                        const float* __restrict__ dxbar, int64_t ld_dxb, const float* __restrict__ xbar_f, int64_t ld_xf, int B,
                        int Lv, float scale, __nv_bfloat16* __restrict__ dkq, __nv_bfloat16* __restrict__ dkq_lo, int64_t ld_dkq) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  constexpr int NV = 3, E = kApE;
  extern __shared__ uint8_t ap_raw[];
  const ApSmem sm = ap_carve(ap_raw);
  float* s_acc = sm.fl;                 // [8][E]
  float* s_dc = s_acc + 8 * E;          // [NPASS][8]
  float* s_attn = s_dc + NPASS * 8;     // [NPASS][Lv]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kApStages; ++i) { mbar_init(sm.full0 + 8 * i, 1); mbar_init(sm.empty0 + 8 * i, 8); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) ap_producer(sm, xv, bstride, B, Lv);
    return;
  }
  const int nchunks = (Lv + kApRows - 1) / kApRows;
  uint32_t it = 0;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float g[NPASS][NV][8], D[NPASS], acc[NPASS][NV][8], dc[NPASS];
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      const int64_t r = static_cast<int64_t>(p * B + b);
      for (int j = threadIdx.x; j < Lv; j += 256) s_attn[p * Lv + j] = __ldg(attn + r * ld_attn + j);
      float d = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float* gp = dxbar + r * ld_dxb + v * 256 + lane * 8;
        const float* xp = xbar_f + r * ld_xf + v * 256 + lane * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(gp)), a2 = __ldg(reinterpret_cast<const float4*>(gp + 4));
        const float4 c = __ldg(reinterpret_cast<const float4*>(xp)), c2 = __ldg(reinterpret_cast<const float4*>(xp + 4));
        g[p][v][0] = a.x; g[p][v][1] = a.y; g[p][v][2] = a.z; g[p][v][3] = a.w;
        g[p][v][4] = a2.x; g[p][v][5] = a2.y; g[p][v][6] = a2.z; g[p][v][7] = a2.w;
        d += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w + a2.x * c2.x + a2.y * c2.y + a2.z * c2.z + a2.w * c2.w;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][v][k] = 0.f;
      }
      D[p] = warp_sum(d);
      dc[p] = 0.f;
    }
    consumer_bar();   // s_attn visible
    for (int c = 0; c < nchunks; ++c, ++it) {
      const uint32_t st = it % kApStages, ph = (it / kApStages) & 1u;
      mbar_wait(sm.full0 + 8 * st, ph);
      const uint8_t* chunk = sm.ring + st * kApChunkBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = c * kApRows + warp + 8 * h;
        if (j < Lv) {
          float x[NV][8];
#pragma unroll
          for (int v = 0; v < NV; ++v) lds8_bf16(chunk + (warp + 8 * h) * (E * 2) + (v * 256 + lane * 8) * 2, x[v]);
#pragma unroll
          for (int p = 0; p < NPASS; ++p) {
            float t = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
              for (int k = 0; k < 8; ++k) t = fmaf(g[p][v][k], x[v][k], t);
            t = warp_sum(t);
            const float ds = s_attn[p * Lv + j] * (t - D[p]) * scale;
            dc[p] += ds;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[p][v][k] = fmaf(ds, x[v][k], acc[p][v][k]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty0 + 8 * st);
    }
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) s_acc[warp * E + v * 256 + lane * 8 + k] = acc[p][v][k];
      if (lane == 0) s_dc[p * 8 + warp] = dc[p];
      consumer_bar();
      const int64_t r = static_cast<int64_t>(p * B + b);
      for (int e = threadIdx.x; e < E; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_acc[w * E + e];
        const __nv_bfloat16 hi = __float2bfloat16_rn(s);
        dkq[r * ld_dkq + e] = hi;
        if (dkq_lo) dkq_lo[r * ld_dkq + e] = __float2bfloat16_rn(s - __bfloat162float(hi));
      }
      if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_dc[p * 8 + w];
        const __nv_bfloat16 hi = __float2bfloat16_rn(s);
        dkq[r * ld_dkq + E] = hi;
        if (dkq_lo) dkq_lo[r * ld_dkq + E] = __float2bfloat16_rn(s - __bfloat162float(hi));
        for (int e = E + 1; e < E + 8; ++e) {
          dkq[r * ld_dkq + e] = __float2bfloat16_rn(0.f);
          if (dkq_lo) dkq_lo[r * ld_dkq + e] = __float2bfloat16_rn(0.f);
        }
      }
      consumer_bar();
    }
  }
}

static int ap_grid(int B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return B < sms ? B : sms;
}

// ------------------------------------------------------------------ aspect-att
__device__ __forceinline__ const __nv_bfloat16* aspect_row(const __nv_bfloat16* t, int64_t ldt, const __nv_bfloat16* v, int64_t ldv,
                                                           int B, int f) {
  return f < B ? t + static_cast<int64_t>(f) * ldt : v + static_cast<int64_t>(f - B) * ldv;
}
// warp per sample i; pairs flat rows 2i, 2i+1 of [t_pool ; v_pool] (stack -> reshape, mm_late.py:120-121)
__global__ void aspect_fwd_kernel(const __nv_bfloat16* __restrict__ t, int64_t ldt, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                                  int B, int E, const float* __restrict__ w_a, const float* __restrict__ b_a,
                                  float* __restrict__ out, int64_t ldo, float* __restrict__ alpha) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= B) return;
  const __nv_bfloat16* r0 = aspect_row(t, ldt, v, ldv, B, 2 * i);
  const __nv_bfloat16* r1 = aspect_row(t, ldt, v, ldv, B, 2 * i + 1);
  float d0 = 0.f, d1 = 0.f;
  for (int k = lane; k < E; k += 32) {
    const float w = w_a[k];
    d0 = fmaf(w, __bfloat162float(r0[k]), d0);
    d1 = fmaf(w, __bfloat162float(r1[k]), d1);
  }
  const float e0 = tanhf(warp_sum(d0) + b_a[0]), e1 = tanhf(warp_sum(d1) + b_a[0]);
  const float mx = fmaxf(e0, e1);
  const float p0 = expf(e0 - mx), p1 = expf(e1 - mx);
  const float a0 = p0 / (p0 + p1), a1 = p1 / (p0 + p1);
  if (lane == 0) { alpha[2 * i] = a0; alpha[2 * i + 1] = a1; }
  for (int k = lane; k < E; k += 32)
    out[static_cast<int64_t>(i) * ldo + k] = fmaxf(a0 * __bfloat162float(r0[k]) + a1 * __bfloat162float(r1[k]), 0.f);
}
// dout already includes everything downstream; the relu mask is re-derived from `out`.
__global__ void aspect_bwd_kernel(const __nv_bfloat16* __restrict__ t, int64_t ldt, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                                  int B, int E, const float* __restrict__ w_a, const float* __restrict__ b_a,
                                  const float* __restrict__ out, int64_t ldo, const float* __restrict__ alpha,
                                  const float* __restrict__ dout, int64_t lddo, float* __restrict__ dt, int64_t lddt,
                                  float* __restrict__ dw_a, float* __restrict__ db_a) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= B) return;
  const int f0 = 2 * i, f1 = 2 * i + 1;
  const __nv_bfloat16* r0 = aspect_row(t, ldt, v, ldv, B, f0);
  const __nv_bfloat16* r1 = aspect_row(t, ldt, v, ldv, B, f1);
  const float a0 = alpha[2 * i], a1 = alpha[2 * i + 1];
  // dalpha_k = <g, V_k>, g = dout * (out > 0)
  float da0 = 0.f, da1 = 0.f, d0 = 0.f, d1 = 0.f;
  for (int k = lane; k < E; k += 32) {
    const float g = out[static_cast<int64_t>(i) * ldo + k] > 0.f ? dout[static_cast<int64_t>(i) * lddo + k] : 0.f;
    const float x0 = __bfloat162float(r0[k]), x1 = __bfloat162float(r1[k]);
    da0 = fmaf(g, x0, da0);
    da1 = fmaf(g, x1, da1);
    d0 = fmaf(w_a[k], x0, d0);
    d1 = fmaf(w_a[k], x1, d1);
  }
  da0 = warp_sum(da0); da1 = warp_sum(da1);
  const float e0 = tanhf(warp_sum(d0) + b_a[0]), e1 = tanhf(warp_sum(d1) + b_a[0]);
  const float dot = a0 * da0 + a1 * da1;
  const float de0 = a0 * (da0 - dot) * (1.f - e0 * e0), de1 = a1 * (da1 - dot) * (1.f - e1 * e1);  // through softmax, tanh
  for (int k = lane; k < E; k += 32) {
    const float g = out[static_cast<int64_t>(i) * ldo + k] > 0.f ? dout[static_cast<int64_t>(i) * lddo + k] : 0.f;
    const float x0 = __bfloat162float(r0[k]), x1 = __bfloat162float(r1[k]);
    const float w = w_a[k];
    if (f0 < B) dt[static_cast<int64_t>(f0) * lddt + k] = a0 * g + de0 * w;   // flat rows >= B are the frozen vision pools
    if (f1 < B) dt[static_cast<int64_t>(f1) * lddt + k] = a1 * g + de1 * w;
    atomicAdd(dw_a + k, de0 * x0 + de1 * x1);
  }
  if (lane == 0) atomicAdd(db_a, de0 + de1);
}

// ------------------------------------------------------------------ gmu gate
__global__ void gmu_gate_fwd_kernel(const __nv_bfloat16* __restrict__ X, int64_t ldx, const float* __restrict__ tp,
                                    const float* __restrict__ vp, int64_t ldp, int B, int E2, __nv_bfloat16* __restrict__ G,
                                    __nv_bfloat16* __restrict__ G_lo, int64_t ldg) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * E2) return;
  const int64_t r = idx / E2, c = idx % E2;
  const float z = 1.f / (1.f + expf(-__bfloat162float(X[r * ldx + c])));
  const float gv = z * tp[r * ldp + c] + (1.f - z) * vp[r * ldp + c];
  const __nv_bfloat16 hi = __float2bfloat16_rn(gv);
  G[r * ldg + c] = hi;
  if (G_lo) G_lo[r * ldg + c] = __float2bfloat16_rn(gv - __bfloat162float(hi));
}
__global__ void gmu_gate_bwd_kernel(const __nv_bfloat16* __restrict__ X, int64_t ldx, const float* __restrict__ tp,
                                    const float* __restrict__ vp, int64_t ldp, const float* __restrict__ dG, int64_t lddg, int B,
                                    int E2, __nv_bfloat16* __restrict__ dtp, __nv_bfloat16* __restrict__ dvp,
                                    __nv_bfloat16* __restrict__ dtp_lo, __nv_bfloat16* __restrict__ dvp_lo, int64_t lddp,
                                    float* __restrict__ dX, int64_t lddx) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * E2) return;
  const int64_t r = idx / E2, c = idx % E2;
  const float z = 1.f / (1.f + expf(-__bfloat162float(X[r * ldx + c])));
  const float g = dG[r * lddg + c];
  const float a = g * z, bq = g * (1.f - z);
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(bq);
  dtp[r * lddp + c] = ah;
  dvp[r * lddp + c] = bh;
  if (dtp_lo) dtp_lo[r * lddp + c] = __float2bfloat16_rn(a - __bfloat162float(ah));
  if (dvp_lo) dvp_lo[r * lddp + c] = __float2bfloat16_rn(bq - __bfloat162float(bh));
  if (dX) dX[r * lddx + c] = g * (tp[r * ldp + c] - vp[r * ldp + c]) * z * (1.f - z);
}

// tensor-core form of the two kernels above (attn_mma.cu); TIC_ATTN_V3=0 keeps the SIMT pipeline (A/B measurement switch)
int attn_pool_fwd_mma(const void* xv, int64_t bstride, const float* kq, int64_t ldkq, int B, int npass, int Lv, float scale,
                      void* xbar_b, void* xbar_lo, int64_t ld_xb, float* xbar_f, int64_t ld_xf, float* attn, int64_t ld_attn,
                      cudaStream_t st);
int attn_pool_bwd_mma(const void* xv, int64_t bstride, const float* attn, int64_t ld_attn, const float* dxbar, int64_t ld_dxb,
                      const float* xbar_f, int64_t ld_xf, int B, int npass, int Lv, float scale, void* dkq, void* dkq_lo,
                      int64_t ld_dkq, cudaStream_t st);
static bool attn_v3() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_ATTN_V3"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

}  // namespace tic

using namespace tic;

extern "C" {

int tic_attn_pool_fwd(const void* xv, int64_t xv_batch_stride, int64_t xv_tok_stride, const void* kq, int64_t ldkq, int B,
                      int npass, int Lv, int E, float scale, void* xbar_bf16, void* xbar_bf16_lo, int64_t ld_xb, float* xbar_f32,
                      int64_t ld_xf, float* attn, int64_t ld_attn, void* stream) {
  TIC_CHECK_ARG(xv && kq && attn && B > 0 && Lv > 0 && Lv <= kAttnMaxLv, "tic_attn_pool_fwd: bad arguments");
  TIC_CHECK_ARG(E == 768 && (npass == 1 || npass == 2), "tic_attn_pool_fwd: E must be 768 (models/config.py:82-84), npass 1|2");
  TIC_CHECK_ARG((xv_batch_stride & 7) == 0 && (xv_tok_stride & 7) == 0 && ldkq >= E + 1 && aligned16(xv),
                "tic_attn_pool_fwd: 16-byte aligned x_v rows / augmented kq (ld >= E+1) required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (xv_tok_stride == E && (xv_batch_stride & 7) == 0 && (ldkq & 3) == 0 && (reinterpret_cast<uintptr_t>(kq) & 15) == 0) {
    if (attn_v3() && (ld_xb & 1) == 0 && (ld_xf & 1) == 0 && ((reinterpret_cast<uintptr_t>(xbar_f32) | reinterpret_cast<uintptr_t>(xbar_bf16) |
                                                               reinterpret_cast<uintptr_t>(xbar_bf16_lo)) & 7) == 0) {
      int rc = attn_pool_fwd_mma(xv, xv_batch_stride, static_cast<const float*>(kq), ldkq, B, npass, Lv, scale, xbar_bf16, xbar_bf16_lo,
                                 ld_xb, xbar_f32, ld_xf, attn, ld_attn, st);
      if (rc) set_error("tic_attn_pool_fwd: tensor-core kernel launch failed (%d)", rc);
      return rc;
    }
    const size_t sm2 = 128 + kApStages * kApChunkBytes + 16 * kApStages + sizeof(float) * (npass * Lv + npass * 16 + 8 * E);
    auto xb2 = static_cast<__nv_bfloat16*>(xbar_bf16);
    auto xl2 = static_cast<__nv_bfloat16*>(xbar_bf16_lo);
    if (npass == 1) {
      auto k = attn_pool_fwd_v2_kernel<1>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      launch_k(k, dim3(ap_grid(B)), dim3(288), sm2, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, static_cast<const float*>(kq), ldkq,
                                      B, Lv, scale, xb2, xl2, ld_xb, xbar_f32, ld_xf, attn, ld_attn);
    } else {
      auto k = attn_pool_fwd_v2_kernel<2>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      launch_k(k, dim3(ap_grid(B)), dim3(288), sm2, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, static_cast<const float*>(kq), ldkq,
                                      B, Lv, scale, xb2, xl2, ld_xb, xbar_f32, ld_xf, attn, ld_attn);
    }
    TIC_CHECK_LAUNCH("tic_attn_pool_fwd");
    return TIC_OK;
  }
  const size_t smem = sizeof(float) * (npass * Lv + npass * kAttnWarps * 2 + kAttnWarps * E);
  auto xb = static_cast<__nv_bfloat16*>(xbar_bf16);
  auto xl = static_cast<__nv_bfloat16*>(xbar_bf16_lo);
  auto kqf = static_cast<const float*>(kq);
  if (npass == 1) {
    auto k = attn_pool_fwd_kernel<3, 1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    launch_k(k, dim3(B), dim3(kAttnWarps * 32), smem, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, xv_tok_stride,
                                        kqf, ldkq, B, Lv, scale, xb, xl, ld_xb, xbar_f32, ld_xf, attn, ld_attn);
  } else {
    auto k = attn_pool_fwd_kernel<3, 2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    launch_k(k, dim3(B), dim3(kAttnWarps * 32), smem, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, xv_tok_stride,
                                        kqf, ldkq, B, Lv, scale, xb, xl, ld_xb, xbar_f32, ld_xf, attn, ld_attn);
  }
  TIC_CHECK_LAUNCH("tic_attn_pool_fwd");
  return TIC_OK;
}

int tic_attn_pool_bwd(const void* xv, int64_t xv_batch_stride, int64_t xv_tok_stride, const float* attn, int64_t ld_attn,
                      const float* dxbar, int64_t ld_dxb, const float* xbar_f32, int64_t ld_xf, int B, int npass, int Lv, int E,
                      float scale, void* dkq_bf16, void* dkq_bf16_lo, int64_t ld_dkq, void* stream) {
  TIC_CHECK_ARG(xv && attn && dxbar && xbar_f32 && dkq_bf16 && B > 0 && Lv > 0, "tic_attn_pool_bwd: bad arguments");
  TIC_CHECK_ARG(E == 768 && (npass == 1 || npass == 2) && ld_dkq >= E + 8, "tic_attn_pool_bwd: E must be 768, npass 1|2, ld_dkq >= E+8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (xv_tok_stride == E && (xv_batch_stride & 7) == 0 && (ld_dxb & 3) == 0 && (ld_xf & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(dxbar) | reinterpret_cast<uintptr_t>(xbar_f32)) & 15) == 0) {
    if (attn_v3() && (ld_dkq & 1) == 0 && ((reinterpret_cast<uintptr_t>(dkq_bf16) | reinterpret_cast<uintptr_t>(dkq_bf16_lo)) & 3) == 0) {
      int rc = attn_pool_bwd_mma(xv, xv_batch_stride, attn, ld_attn, dxbar, ld_dxb, xbar_f32, ld_xf, B, npass, Lv, scale, dkq_bf16,
                                 dkq_bf16_lo, ld_dkq, st);
      if (rc) set_error("tic_attn_pool_bwd: tensor-core kernel launch failed (%d)", rc);
      return rc;
    }
    const size_t sm2 = 128 + kApStages * kApChunkBytes + 16 * kApStages + sizeof(float) * (8 * E + npass * 8 + npass * Lv);
    if (npass == 1) {
      auto k = attn_pool_bwd_v2_kernel<1>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      launch_k(k, dim3(ap_grid(B)), dim3(288), sm2, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, attn, ld_attn, dxbar, ld_dxb, xbar_f32,
                                      ld_xf, B, Lv, scale, static_cast<__nv_bfloat16*>(dkq_bf16),
                                      static_cast<__nv_bfloat16*>(dkq_bf16_lo), ld_dkq);
    } else {
      auto k = attn_pool_bwd_v2_kernel<2>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      launch_k(k, dim3(ap_grid(B)), dim3(288), sm2, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, attn, ld_attn, dxbar, ld_dxb, xbar_f32,
                                      ld_xf, B, Lv, scale, static_cast<__nv_bfloat16*>(dkq_bf16),
                                      static_cast<__nv_bfloat16*>(dkq_bf16_lo), ld_dkq);
    }
    TIC_CHECK_LAUNCH("tic_attn_pool_bwd");
    return TIC_OK;
  }
  const size_t smem = sizeof(float) * (kAttnWarps * E + npass * kAttnWarps);
  if (npass == 1) {
    auto k = attn_pool_bwd_kernel<3, 1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    launch_k(k, dim3(B), dim3(kAttnWarps * 32), smem, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, xv_tok_stride, attn, ld_attn,
                                        dxbar, ld_dxb, xbar_f32, ld_xf, B, Lv, scale, static_cast<__nv_bfloat16*>(dkq_bf16),
                                        static_cast<__nv_bfloat16*>(dkq_bf16_lo), ld_dkq);
  } else {
    auto k = attn_pool_bwd_kernel<3, 2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    launch_k(k, dim3(B), dim3(kAttnWarps * 32), smem, st, static_cast<const __nv_bfloat16*>(xv), xv_batch_stride, xv_tok_stride, attn, ld_attn,
                                        dxbar, ld_dxb, xbar_f32, ld_xf, B, Lv, scale, static_cast<__nv_bfloat16*>(dkq_bf16),
                                        static_cast<__nv_bfloat16*>(dkq_bf16_lo), ld_dkq);
  }
  TIC_CHECK_LAUNCH("tic_attn_pool_bwd");
  return TIC_OK;
}

int tic_aspect_fwd(const void* t_pool, int64_t ldt, const void* v_pool, int64_t ldv, int B, int E, const float* w_a,
                   const float* b_a, float* out, int64_t ldo, float* alpha, void* stream) {
  TIC_CHECK_ARG(t_pool && v_pool && w_a && b_a && out && alpha && B > 0 && E > 0, "tic_aspect_fwd: bad arguments");
  launch_k(aspect_fwd_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(t_pool), ldt, static_cast<const __nv_bfloat16*>(v_pool), ldv, B, E, w_a, b_a, out, ldo, alpha);
  TIC_CHECK_LAUNCH("tic_aspect_fwd");
  return TIC_OK;
}

int tic_aspect_bwd(const void* t_pool, int64_t ldt, const void* v_pool, int64_t ldv, int B, int E, const float* w_a,
                   const float* b_a, const float* out, int64_t ldo, const float* alpha, const float* dout, int64_t lddo,
                   float* dt_pool, int64_t lddt, float* dw_a, float* db_a, void* stream) {
  TIC_CHECK_ARG(t_pool && v_pool && w_a && b_a && out && alpha && dout && dt_pool && dw_a && db_a && B > 0,
                "tic_aspect_bwd: bad arguments");
  launch_k(aspect_bwd_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(t_pool), ldt, static_cast<const __nv_bfloat16*>(v_pool), ldv, B, E, w_a, b_a, out, ldo,
      alpha, dout, lddo, dt_pool, lddt, dw_a, db_a);
  TIC_CHECK_LAUNCH("tic_aspect_bwd");
  return TIC_OK;
}

int tic_gmu_gate_fwd(const void* Xcat, int64_t ldx, const float* tp, const float* vp, int64_t ldp, int B, int E2, void* G_bf16,
                     void* G_bf16_lo, int64_t ldg, void* stream) {
  TIC_CHECK_ARG(Xcat && tp && vp && G_bf16 && B > 0 && E2 > 0, "tic_gmu_gate_fwd: bad arguments");
  const int64_t n = static_cast<int64_t>(B) * E2;
  launch_k(gmu_gate_fwd_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(Xcat), ldx, tp, vp, ldp, B, E2, static_cast<__nv_bfloat16*>(G_bf16),
      static_cast<__nv_bfloat16*>(G_bf16_lo), ldg);
  TIC_CHECK_LAUNCH("tic_gmu_gate_fwd");
  return TIC_OK;
}

int tic_gmu_gate_bwd(const void* Xcat, int64_t ldx, const float* tp, const float* vp, int64_t ldp, const float* dG, int64_t lddg,
                     int B, int E2, void* dtp_bf16, void* dvp_bf16, void* dtp_lo, void* dvp_lo, int64_t lddp, float* dXcat_gate,
                     int64_t lddx, void* stream) {
  TIC_CHECK_ARG(Xcat && tp && vp && dG && dtp_bf16 && dvp_bf16 && B > 0 && E2 > 0, "tic_gmu_gate_bwd: bad arguments");
  const int64_t n = static_cast<int64_t>(B) * E2;
  launch_k(gmu_gate_bwd_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(Xcat), ldx, tp, vp, ldp, dG, lddg, B, E2, static_cast<__nv_bfloat16*>(dtp_bf16),
      static_cast<__nv_bfloat16*>(dvp_bf16), static_cast<__nv_bfloat16*>(dtp_lo), static_cast<__nv_bfloat16*>(dvp_lo), lddp,
      dXcat_gate, lddx);
  TIC_CHECK_LAUNCH("tic_gmu_gate_bwd");
  return TIC_OK;
}

}  // extern "C"
