// Per-sample (HBM/latency-bound) pieces of the late-fusion head: CLS-row packing for linear_fusion, the classifier and
// ITM heads with their losses (forward + backward in one pass), bias/weight-gradient column reductions, casts, loss mix.
// Reference: models/mm_late.py:92-96,160-193 (head), :473-487 (mix); run_mm_late.py:85,97 (loss constructors).
#include "common.cuh"
#include "tic_ptx.cuh"
#include "itm_rule.cuh"

namespace tic {

constexpr int kMaxClasses = 8;
constexpr int kBlkC = 4;   // the block-per-row heads kernel keeps the class weights in registers: tasks have 2..4 classes (config.py:18-48)

// ------------------------------------------------------------------ pack / unpack
// One warp per output row; 16-byte copies. Row r < B: [xt[r] | xv[r]];  row B + i: [xt[src[i]] | xv[i]].
__global__ void pack_cls_pairs_kernel(const __nv_bfloat16* __restrict__ xt, int64_t xt_stride,
                                      const __nv_bfloat16* __restrict__ xv, int64_t xv_stride, int B, int E,
                                      const int32_t* __restrict__ src, __nv_bfloat16* __restrict__ X, int64_t ldx, int rows,
                                      const float* __restrict__ u_coin, const float* __restrict__ u_pick) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int i = r < B ? r : r - B;
  int it = r;
  if (r >= B) {
    if (u_coin != nullptr) {   // uniform ITM rule evaluated in place: the pack does not wait for the sampler kernel
      int lbl;
      uniform_rule(u_coin, u_pick, B, i, lbl, it);
    } else {
      it = src[i];
    }
  }
  const __nv_bfloat16* a = xt + static_cast<int64_t>(it) * xt_stride;
  const __nv_bfloat16* b = xv + static_cast<int64_t>(i) * xv_stride;
  __nv_bfloat16* d = X + static_cast<int64_t>(r) * ldx;
  const bool vec = (E & 7) == 0 && (xt_stride & 7) == 0 && (xv_stride & 7) == 0 && (ldx & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(xt) | reinterpret_cast<uintptr_t>(xv) | reinterpret_cast<uintptr_t>(X)) & 15) == 0;
  if (vec) {
    for (int c = lane * 8; c < E; c += 256) {
      *reinterpret_cast<uint4*>(d + c) = __ldg(reinterpret_cast<const uint4*>(a + c));
      if (xv) *reinterpret_cast<uint4*>(d + E + c) = __ldg(reinterpret_cast<const uint4*>(b + c));
    }
  } else {
    for (int c = lane; c < E; c += 32) { d[c] = a[c]; if (xv) d[E + c] = b[c]; }
  }
}

// dxt[target(r), :] += dX[r, :E] (+ dX2[r, :])  for every packed row r: target = r (main rows) or src[r - B] (ITM rows).
// dxt is zero on entry (it lives in the plan's zero block), so the main and the scattered part are ONE launch of fp32 atomics.
__global__ void unpack_accum_kernel(const float* __restrict__ dX, int64_t ldd, const float* __restrict__ dX2, int64_t ldd2, int B,
                                    int E, int rows, const int32_t* __restrict__ src, float* __restrict__ dxt, int64_t ldo) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int E4 = E >> 2;
  if (idx >= static_cast<int64_t>(rows) * E4) return;
  const int r = static_cast<int>(idx / E4), c = static_cast<int>(idx % E4) * 4;
  const int t = r < B ? r : src[r - B];
  float4 v = *reinterpret_cast<const float4*>(dX + static_cast<int64_t>(r) * ldd + c);
  if (dX2) {
    const float4 w = *reinterpret_cast<const float4*>(dX2 + static_cast<int64_t>(r) * ldd2 + c);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  float* d = dxt + static_cast<int64_t>(t) * ldo + c;
  atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
}

// ------------------------------------------------------------------ heads: logits, losses, dlogits, dH
// One warp per row of H. rows [0,B) feed linear_cls (+ dropout keep mask), rows [B,2B) feed linear_tim.
//
// Pairwise form of concat fusion (Pt != nullptr): linear_fusion([x_t | x_v]) = x_t W_f[:, :E]^T + (x_v W_f[:, E:]^T + b), so the
// two halves are projected ONCE per sample (Pt = text half, Pv = image half + bias, fp32 [B, E]) and the fused vector of the
// ITM pair (text of sample src[i], image of sample i; mm_late.py:170-181) is relu(Pt[src[i]] + Pv[i]) — no packed [2B, 2E]
// operand, half the fusion FLOPs.  The kernel then builds its row of H itself (and stores it: mm_features / weight gradients).
__global__ void heads_rows_kernel(float* __restrict__ H, int64_t ldh, int B, int E, int C, int has_tim,
                                  const float* __restrict__ W_cls, const float* __restrict__ b_cls,
                                  const float* __restrict__ W_tim, const float* __restrict__ b_tim,
                                  const float* __restrict__ y_soft, const float* __restrict__ class_w,
                                  const int64_t* __restrict__ lbl_tim, const uint8_t* __restrict__ keep, float keep_scale,
                                  float c_cls, float c_tim, float* __restrict__ logits_cls, float* __restrict__ logits_tim,
                                  float* __restrict__ losses, float* __restrict__ dlogits /* [rows, kMaxClasses] */,
                                  const float* __restrict__ dz_ext /* upstream dL/dlogits or nullptr (fused losses) */,
                                  __nv_bfloat16* __restrict__ dHb, __nv_bfloat16* __restrict__ dHb_lo, int64_t ld_dhb,
                                  float* __restrict__ dHf, int64_t ld_dhf, int relu_mask, const float* __restrict__ Pt,
                                  const float* __restrict__ Pv, int64_t ldp, const int32_t* __restrict__ src) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int rows = has_tim ? 2 * B : B;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + wib;
  __shared__ float sl[2][32];
  float loss_c = 0.f, loss_t = 0.f;
  if (r < rows) {
    const bool is_cls = r < B;
    const int i = is_cls ? r : r - B;
    const int nc = is_cls ? C : 2;
    const float* W = is_cls ? W_cls : W_tim;
    const float* bias = is_cls ? b_cls : b_tim;
    float* h = H + static_cast<int64_t>(r) * ldh;
    const uint8_t* kp = (is_cls && keep) ? keep + static_cast<int64_t>(i) * E : nullptr;
    float z[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) z[c] = 0.f;
    // 16-byte path: every lane owns 4 consecutive features per 128-wide chunk (6 independent iterations at E = 768)
    const bool vec = (E & 3) == 0 && (ldh & 3) == 0 && (ld_dhb & 3) == 0 && (ld_dhf & 3) == 0 && (ldp & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(W_cls) | reinterpret_cast<uintptr_t>(W_tim) |
                       reinterpret_cast<uintptr_t>(dHf) | reinterpret_cast<uintptr_t>(keep) | reinterpret_cast<uintptr_t>(Pt) |
                       reinterpret_cast<uintptr_t>(Pv)) & 15) == 0 &&
                     ((reinterpret_cast<uintptr_t>(dHb) | reinterpret_cast<uintptr_t>(dHb_lo)) & 7) == 0;
    if (Pt != nullptr) {   // pairwise form: this row of H = relu(Pt[text row] + Pv[image row]); every lane re-reads only its own stores
      const float* pt = Pt + static_cast<int64_t>(is_cls ? i : src[i]) * ldp;
      const float* pv = Pv + static_cast<int64_t>(i) * ldp;
      if (vec) {
#pragma unroll 2
        for (int k = lane * 4; k < E; k += 128) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(pt + k)), b = __ldg(reinterpret_cast<const float4*>(pv + k));
          *reinterpret_cast<float4*>(h + k) = make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f),
                                                          fmaxf(a.w + b.w, 0.f));
        }
      } else {
        for (int k = lane; k < E; k += 32) h[k] = fmaxf(__ldg(pt + k) + __ldg(pv + k), 0.f);
      }
    }
    if (vec) {
#pragma unroll 2
      for (int k = lane * 4; k < E; k += 128) {
        float4 hv = *reinterpret_cast<const float4*>(h + k);
        if (kp) {
          const uchar4 m = *reinterpret_cast<const uchar4*>(kp + k);
          hv.x = m.x ? hv.x * keep_scale : 0.f; hv.y = m.y ? hv.y * keep_scale : 0.f;
          hv.z = m.z ? hv.z * keep_scale : 0.f; hv.w = m.w ? hv.w * keep_scale : 0.f;
        }
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c < nc) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + c * E + k));
            z[c] = fmaf(hv.x, w.x, fmaf(hv.y, w.y, fmaf(hv.z, w.z, fmaf(hv.w, w.w, z[c]))));
          }
      }
    } else {
    for (int k = lane; k < E; k += 32) {
      float hv = h[k];
      if (kp) hv = kp[k] ? hv * keep_scale : 0.f;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < nc) z[c] = fmaf(hv, __ldg(W + c * E + k), z[c]);
    }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < nc) { z[c] = warp_sum(z[c]) + bias[c]; mx = fmaxf(mx, z[c]); }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < nc) se += expf(z[c] - mx);
    const float lse = mx + logf(se);
    float dz[kMaxClasses];
    if (is_cls) {
      float wy = 0.f, l = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < nc) {
          const float w = class_w ? class_w[c] : 1.f;
          const float y = y_soft[static_cast<int64_t>(i) * C + c];
          wy += w * y;
          l -= w * y * (z[c] - lse);
        }
      loss_c = l / B;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < nc) {
          const float w = class_w ? class_w[c] : 1.f;
          const float y = y_soft[static_cast<int64_t>(i) * C + c];
          dz[c] = c_cls / B * (expf(z[c] - lse) * wy - w * y);
          if (lane == 0) logits_cls[static_cast<int64_t>(i) * C + c] = z[c];
        }
    } else {
      const int y = static_cast<int>(lbl_tim[i]);
      loss_t = -((y == 0 ? z[0] : z[1]) - lse) / B;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        dz[c] = c_tim / B * (expf(z[c] - lse) - (c == y ? 1.f : 0.f));
        if (lane == 0) logits_tim[static_cast<int64_t>(i) * 2 + c] = z[c];
      }
    }
    if (dz_ext != nullptr) {  // autograd mode: the caller owns the loss; use its gradient w.r.t. the logits
      loss_c = 0.f;
      loss_t = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) dz[c] = c < nc ? __ldg(dz_ext + static_cast<int64_t>(r) * kMaxClasses + c) : 0.f;
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) dlogits[static_cast<int64_t>(r) * kMaxClasses + c] = c < nc ? dz[c] : 0.f;
    }
    if ((dHb || dHf) && vec) {
#pragma unroll 2
      for (int k = lane * 4; k < E; k += 128) {
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c < nc) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + c * E + k));
            g[0] = fmaf(dz[c], w.x, g[0]); g[1] = fmaf(dz[c], w.y, g[1]);
            g[2] = fmaf(dz[c], w.z, g[2]); g[3] = fmaf(dz[c], w.w, g[3]);
          }
        if (kp) {
          const uchar4 m = *reinterpret_cast<const uchar4*>(kp + k);
          g[0] = m.x ? g[0] * keep_scale : 0.f; g[1] = m.y ? g[1] * keep_scale : 0.f;
          g[2] = m.z ? g[2] * keep_scale : 0.f; g[3] = m.w ? g[3] * keep_scale : 0.f;
        }
        if (relu_mask) {
          const float4 hv = *reinterpret_cast<const float4*>(h + k);
          if (!(hv.x > 0.f)) g[0] = 0.f;
          if (!(hv.y > 0.f)) g[1] = 0.f;
          if (!(hv.z > 0.f)) g[2] = 0.f;
          if (!(hv.w > 0.f)) g[3] = 0.f;
        }
        if (dHf) *reinterpret_cast<float4*>(dHf + static_cast<int64_t>(r) * ld_dhf + k) = make_float4(g[0], g[1], g[2], g[3]);
        if (dHb) {
          uint2 u;
          u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]);
          *reinterpret_cast<uint2*>(dHb + static_cast<int64_t>(r) * ld_dhb + k) = u;
          if (dHb_lo) {
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] -= __bfloat162float(__float2bfloat16_rn(g[j]));
            u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]);
            *reinterpret_cast<uint2*>(dHb_lo + static_cast<int64_t>(r) * ld_dhb + k) = u;
          }
        }
      }
    } else if (dHb || dHf) {
      for (int k = lane; k < E; k += 32) {
        float g = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c < nc) g = fmaf(dz[c], __ldg(W + c * E + k), g);
        if (kp) g = kp[k] ? g * keep_scale : 0.f;
        if (relu_mask && !(h[k] > 0.f)) g = 0.f;
        if (dHb) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(g);
          dHb[static_cast<int64_t>(r) * ld_dhb + k] = hi;
          if (dHb_lo) dHb_lo[static_cast<int64_t>(r) * ld_dhb + k] = __float2bfloat16_rn(g - __bfloat162float(hi));
        }
        if (dHf) dHf[static_cast<int64_t>(r) * ld_dhf + k] = g;
      }
    }
  }
  if (lane == 0) { sl[0][wib] = loss_c; sl[1][wib] = loss_t; }
  __syncthreads();
  if (wib == 0) {
    const int nw = blockDim.x >> 5;
    float a = lane < nw ? sl[0][lane] : 0.f, b = lane < nw ? sl[1][lane] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      if (a != 0.f) atomicAdd(losses + 0, a);
      if (b != 0.f) atomicAdd(losses + 1, b);
    }
  }
}

// Latency-oriented form of heads_rows_kernel for 16-byte-aligned rows with E <= 4096: ONE BLOCK PER ROW, every thread owns
// 4 consecutive features.  All global loads of a row (the two projected halves or H, the nc weight rows, the dropout mask)
// are issued at once and everything — the weights too — stays in registers between the logits and the dH pass, so a row
// costs one memory round trip + one block reduction instead of the warp-per-row kernel's chain of dependent passes
// (CUPTI timeline of the c2 step: 20 us for 512 rows, the longest kernel of the fusion chain).
__global__ void __launch_bounds__(768, 1)
heads_rows_block_kernel(float* __restrict__ H, int64_t ldh, int B, int E, int C, int has_tim, const float* __restrict__ W_cls,
                        const float* __restrict__ b_cls, const float* __restrict__ W_tim, const float* __restrict__ b_tim,
                        const float* __restrict__ y_soft, const float* __restrict__ class_w, const int64_t* __restrict__ lbl_tim,
                        const uint8_t* __restrict__ keep, float keep_scale, float c_cls, float c_tim,
                        float* __restrict__ logits_cls, float* __restrict__ logits_tim, float* __restrict__ losses,
                        float* __restrict__ dlogits, const float* __restrict__ dz_ext, __nv_bfloat16* __restrict__ dHb,
                        __nv_bfloat16* __restrict__ dHb_lo, int64_t ld_dhb, float* __restrict__ dHf, int64_t ld_dhf, int relu_mask,
                        const float* __restrict__ Pt, const float* __restrict__ Pv, int64_t ldp, const int32_t* __restrict__ src, int tpr) {
  pdl_trigger();
  pdl_wait();
  // tpr threads per row (a multiple of 32), rpb rows side by side in a block; the grid is capped (see the launch) so that the
  // kernel occupies only part of the machine: its many small blocks otherwise land on EVERY SM and their registers keep the
  // 8-CTA cluster of the ITC chain (51 K registers per CTA) from being scheduled until they retire (CUPTI timeline: 6 us).
  const int rows = has_tim ? 2 * B : B;
  const int rpb = blockDim.x / tpr, sub = threadIdx.x / tpr;
  const int tid = threadIdx.x - sub * tpr, lane = tid & 31, wib = tid >> 5, nw = tpr >> 5;
  __shared__ float sz[10][8][kBlkC];
  for (int base = blockIdx.x * rpb; base < rows; base += gridDim.x * rpb) {
  const int r = min(base + sub, rows - 1);
  const bool rvalid = base + sub < rows;
  const bool is_cls = r < B;
  const int i = is_cls ? r : r - B;
  const int nc = is_cls ? C : 2;
  const float* W = is_cls ? W_cls : W_tim;
  const float* bias = is_cls ? b_cls : b_tim;
  const int k = tid * 8;               // 8 consecutive features per thread (two 16-byte groups): E = 768 -> 96 threads per row
  const bool act0 = k < E && rvalid, act1 = k + 4 < E && rvalid;
  // ---- every load of this row, issued back to back
  float4 hv[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  float4 w[2][kBlkC];
  uchar4 m[2] = {make_uchar4(1, 1, 1, 1), make_uchar4(1, 1, 1, 1)};
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int kv = k + 4 * v;
    if (v == 0 ? act0 : act1) {
      if (Pt != nullptr) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(Pt + static_cast<int64_t>(is_cls ? i : src[i]) * ldp + kv));
        const float4 b = __ldg(reinterpret_cast<const float4*>(Pv + static_cast<int64_t>(i) * ldp + kv));
        hv[v] = make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
      } else {
        hv[v] = *reinterpret_cast<const float4*>(H + static_cast<int64_t>(r) * ldh + kv);
      }
      if (is_cls && keep) m[v] = *reinterpret_cast<const uchar4*>(keep + static_cast<int64_t>(i) * E + kv);
#pragma unroll
      for (int c = 0; c < kBlkC; ++c)
        if (c < nc) w[v][c] = __ldg(reinterpret_cast<const float4*>(W + c * E + kv));
    }
  }
  float yv[kBlkC], cw[kBlkC];
  int ylab = 0;
  if (is_cls) {
#pragma unroll
    for (int c = 0; c < kBlkC; ++c)
      if (c < nc) { yv[c] = __ldg(y_soft + static_cast<int64_t>(i) * C + c); cw[c] = class_w ? __ldg(class_w + c) : 1.f; }
  } else {
    ylab = static_cast<int>(lbl_tim[i]);
  }
  float z[kBlkC];
#pragma unroll
  for (int c = 0; c < kBlkC; ++c) z[c] = 0.f;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const bool act = v == 0 ? act0 : act1;
    if (act && Pt != nullptr) *reinterpret_cast<float4*>(H + static_cast<int64_t>(r) * ldh + k + 4 * v) = hv[v];   // mm_features / weight gradients
    // ---- logits: per-thread partial dot products
    const float4 hd = make_float4(m[v].x ? hv[v].x * keep_scale : 0.f, m[v].y ? hv[v].y * keep_scale : 0.f,
                                  m[v].z ? hv[v].z * keep_scale : 0.f, m[v].w ? hv[v].w * keep_scale : 0.f);
    const float4 hx = (is_cls && keep) ? hd : hv[v];
#pragma unroll
    for (int c = 0; c < kBlkC; ++c)
      if (c < nc && act) z[c] = fmaf(hx.x, w[v][c].x, fmaf(hx.y, w[v][c].y, fmaf(hx.z, w[v][c].z, fmaf(hx.w, w[v][c].w, z[c]))));
  }
#pragma unroll
  for (int c = 0; c < kBlkC; ++c)
    if (c < nc) z[c] = warp_sum(z[c]);
  __syncthreads();      // (the previous row group has read its sums)
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kBlkC; ++c) sz[sub][wib][c] = c < nc ? z[c] : 0.f;
  }
  __syncthreads();
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kBlkC; ++c)
    if (c < nc) {
      float t = bias[c];
      for (int q = 0; q < nw; ++q) t += sz[sub][q][c];     // fixed order: every thread computes the same logits
      z[c] = t;
      mx = fmaxf(mx, t);
    }
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < kBlkC; ++c)
    if (c < nc) se += expf(z[c] - mx);
  const float lse = mx + logf(se);
  float dz[kBlkC];
  float loss_c = 0.f, loss_t = 0.f;
  if (is_cls) {
    float wy = 0.f, l = 0.f;
#pragma unroll
    for (int c = 0; c < kBlkC; ++c)
      if (c < nc) { wy += cw[c] * yv[c]; l -= cw[c] * yv[c] * (z[c] - lse); }
    loss_c = l / B;
#pragma unroll
    for (int c = 0; c < kBlkC; ++c)
      if (c < nc) {
        dz[c] = c_cls / B * (expf(z[c] - lse) * wy - cw[c] * yv[c]);
        if (tid == 0 && rvalid) logits_cls[static_cast<int64_t>(i) * C + c] = z[c];
      }
  } else {
    loss_t = -((ylab == 0 ? z[0] : z[1]) - lse) / B;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      dz[c] = c_tim / B * (expf(z[c] - lse) - (c == ylab ? 1.f : 0.f));
      if (tid == 0 && rvalid) logits_tim[static_cast<int64_t>(i) * 2 + c] = z[c];
    }
  }
  if (dz_ext != nullptr) {  // autograd mode: the caller owns the loss; use its gradient w.r.t. the logits
    loss_c = 0.f;
    loss_t = 0.f;
#pragma unroll
    for (int c = 0; c < kBlkC; ++c) dz[c] = c < nc ? __ldg(dz_ext + static_cast<int64_t>(r) * kMaxClasses + c) : 0.f;
  }
  if (tid == 0 && rvalid) {
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) dlogits[static_cast<int64_t>(r) * kMaxClasses + c] = (c < kBlkC && c < nc) ? dz[c < kBlkC ? c : 0] : 0.f;
    if (loss_c != 0.f) atomicAdd(losses + 0, loss_c);
    if (loss_t != 0.f) atomicAdd(losses + 1, loss_t);
  }
  // ---- dH for this thread's 8 features, from the weights still in registers
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int kv = k + 4 * v;
    if ((v == 0 ? act0 : act1) && (dHb || dHf)) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < kBlkC; ++c)
        if (c < nc) {
          g[0] = fmaf(dz[c], w[v][c].x, g[0]); g[1] = fmaf(dz[c], w[v][c].y, g[1]);
          g[2] = fmaf(dz[c], w[v][c].z, g[2]); g[3] = fmaf(dz[c], w[v][c].w, g[3]);
        }
      if (is_cls && keep) {
        g[0] = m[v].x ? g[0] * keep_scale : 0.f; g[1] = m[v].y ? g[1] * keep_scale : 0.f;
        g[2] = m[v].z ? g[2] * keep_scale : 0.f; g[3] = m[v].w ? g[3] * keep_scale : 0.f;
      }
      if (relu_mask) {
        if (!(hv[v].x > 0.f)) g[0] = 0.f;
        if (!(hv[v].y > 0.f)) g[1] = 0.f;
        if (!(hv[v].z > 0.f)) g[2] = 0.f;
        if (!(hv[v].w > 0.f)) g[3] = 0.f;
      }
      if (dHf) *reinterpret_cast<float4*>(dHf + static_cast<int64_t>(r) * ld_dhf + kv) = make_float4(g[0], g[1], g[2], g[3]);
      if (dHb) {
        uint2 u;
        u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]);
        *reinterpret_cast<uint2*>(dHb + static_cast<int64_t>(r) * ld_dhb + kv) = u;
        if (dHb_lo) {
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] -= __bfloat162float(__float2bfloat16_rn(g[j]));
          u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]);
          *reinterpret_cast<uint2*>(dHb_lo + static_cast<int64_t>(r) * ld_dhb + kv) = u;
        }
      }
    }
  }
  }   // row groups
}

// Backward of the pairwise form: gradients w.r.t. the two projected halves from dH [2B, E] (bf16 hi + lo):
//   dPv[i] = dH[i] + dH[B+i]                                  (image half of sample i feeds its main row and its ITM row)
//   dPt[k] = dH[k] + sum over {i : src[i] == k} of dH[B+i]    (text half of sample k feeds its main row and every ITM row that drew it)
// One warp per sample; the inverse of the gather is found by scanning src (ballot over 32 candidates per step): no atomics,
// fixed summation order, nothing to zero.  Outputs are bf16 (hi, lo) pairs: A operands of the dX / dW GEMMs.
__global__ void __launch_bounds__(256) fusion_pair_grad_kernel(const __nv_bfloat16* __restrict__ dH, const __nv_bfloat16* __restrict__ dH_lo,
                                                               int64_t ld, int B, int E, int has_tim, const int32_t* __restrict__ src,
                                                               __nv_bfloat16* __restrict__ dPt, __nv_bfloat16* __restrict__ dPt_lo,
                                                               __nv_bfloat16* __restrict__ dPv, __nv_bfloat16* __restrict__ dPv_lo,
                                                               int64_t ldo) {
  pdl_trigger();
  pdl_wait();
  constexpr int MAXC = 4;                         // E <= 1024: 8 consecutive features per lane per 256-wide chunk
  constexpr int kSrcChunk = 2048;                 // src is staged through shared memory, shared by the block's 8 samples
  __shared__ int32_t ssrc[kSrcChunk];
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const bool live = k < B;
  auto ld8 = [&](const __nv_bfloat16* base, const __nv_bfloat16* base_lo, int row, int c, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(row) * ld + c));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (base_lo != nullptr) v = __ldg(reinterpret_cast<const uint4*>(base_lo + static_cast<int64_t>(row) * ld + c));
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
    const __nv_bfloat162* hl = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 t = __bfloat1622float2(hh[q]), t2 = __bfloat1622float2(hl[q]);
      f[2 * q] = t.x + t2.x; f[2 * q + 1] = t.y + t2.y;
    }
  };
  auto st8 = [&](__nv_bfloat16* hi, __nv_bfloat16* lo, int row, int c, float (&g)[8]) {
    uint4 u;
    u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]); u.z = pack_bf16x2(g[4], g[5]); u.w = pack_bf16x2(g[6], g[7]);
    *reinterpret_cast<uint4*>(hi + static_cast<int64_t>(row) * ldo + c) = u;
    if (lo != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] -= __bfloat162float(__float2bfloat16_rn(g[j]));
      u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]); u.z = pack_bf16x2(g[4], g[5]); u.w = pack_bf16x2(g[6], g[7]);
      *reinterpret_cast<uint4*>(lo + static_cast<int64_t>(row) * ldo + c) = u;
    }
  };
  // the first chunk of src and this sample's own two rows are requested together (one memory round trip)
  if (has_tim)
    for (int j = threadIdx.x; j < min(B, kSrcChunk); j += blockDim.x) ssrc[j] = __ldg(src + j);
  float at[MAXC][8], av[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int col = c * 256 + lane * 8;
    if (live && col < E) {
      ld8(dH, dH_lo, k, col, at[c]);
      if (has_tim) ld8(dH, dH_lo, B + k, col, av[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) av[c][j] = has_tim ? av[c][j] + at[c][j] : at[c][j];
  if (has_tim) {
    for (int base = 0; base < B; base += kSrcChunk) {
      if (base > 0) {
        __syncthreads();
        for (int j = threadIdx.x; j < min(B - base, kSrcChunk); j += blockDim.x) ssrc[j] = __ldg(src + base + j);
      }
      __syncthreads();
      const int n = min(B - base, kSrcChunk);
      for (int j0 = 0; j0 < n && live; j0 += 32) {
        const int j = j0 + lane;
        unsigned hit = __ballot_sync(0xffffffffu, j < n && ssrc[j] == k);
        while (hit != 0) {                       // ascending j: a fixed summation order
          const int jj = base + j0 + __ffs(hit) - 1;
          hit &= hit - 1;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            const int col = c * 256 + lane * 8;
            if (col < E) {
              float t[8];
              ld8(dH, dH_lo, B + jj, col, t);
#pragma unroll
              for (int q = 0; q < 8; ++q) at[c][q] += t[q];
            }
          }
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < E) {
      st8(dPt, dPt_lo, k, col, at[c]);
      st8(dPv, dPv_lo, k, col, av[c]);
    }
  }
}

// dW[c,k] += sum_rows dlogits[r,c] * hd[r,k], db[c] += sum_rows dlogits[r,c]; thread per k, rows split over blockIdx.y,
// blockIdx.z = head (0: linear_cls on rows [0,B), 1: linear_tim on rows [B,2B)); 4 rows of loads in flight per thread.
__global__ void heads_wgrad_kernel(const float* __restrict__ H, int64_t ldh, int B, int E, int C,
                                   const float* __restrict__ dlogits, const uint8_t* __restrict__ keep, float keep_scale,
                                   float* __restrict__ dW_cls, float* __restrict__ db_cls, float* __restrict__ dW_tim,
                                   float* __restrict__ db_tim, int rows_per_blk) {
  pdl_trigger();
  pdl_wait();
  const int head = blockIdx.z;
  const int row0 = head ? B : 0, nc = head ? 2 : C;
  float* dW = head ? dW_tim : dW_cls;
  float* db = head ? db_tim : db_cls;
  const uint8_t* kp = head ? nullptr : keep;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int rb = blockIdx.y * rows_per_blk;
  const int re = min(rb + rows_per_blk, B);
  float acc[kMaxClasses], accb[kMaxClasses];
#pragma unroll
  for (int c = 0; c < kMaxClasses; ++c) { acc[c] = 0.f; accb[c] = 0.f; }
  if (k < E) {
    for (int r = rb; r < re; r += 4) {
      float hv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = min(r + u, re - 1);
        hv[u] = (r + u < re) ? H[static_cast<int64_t>(row0 + rr) * ldh + k] : 0.f;
        if (kp) hv[u] = kp[static_cast<int64_t>(rr) * E + k] ? hv[u] * keep_scale : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u < re) {
          const float* dl = dlogits + static_cast<int64_t>(row0 + r + u) * kMaxClasses;
#pragma unroll
          for (int c = 0; c < kMaxClasses; ++c)
            if (c < nc) {
              const float d = __ldg(dl + c);
              acc[c] = fmaf(d, hv[u], acc[c]);
              accb[c] += d;
            }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < nc) {
        atomicAdd(dW + c * E + k, acc[c]);
        if (k == 0) atomicAdd(db + c, accb[c]);
      }
  }
}

// out[n] += sum_m X[m,n]  (bf16 in) — bias gradients of the fusion / projection linears.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ X_lo, int64_t ldx, int rows,
                                   int cols, float* __restrict__ out, int rows_per_blk) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= cols) return;
  const int rb = blockIdx.y * rows_per_blk, re = min(rb + rows_per_blk, rows);
  float s = 0.f;
  for (int r = rb; r < re; ++r) s += __bfloat162float(X[static_cast<int64_t>(r) * ldx + n]);
  if (X_lo != nullptr)   // residual twin of a split-precision operand: one launch sums the pair
    for (int r = rb; r < re; ++r) s += __bfloat162float(X_lo[static_cast<int64_t>(r) * ldx + n]);
  atomicAdd(out + n, s);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, int64_t lds, __nv_bfloat16* __restrict__ d, int64_t ldd,
                                     int rows, int cols) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(rows) * cols) return;
  const int64_t r = idx / cols, c = idx % cols;
  d[r * ldd + c] = __float2bfloat16_rn(s[r * lds + c]);
}
// Multi-tensor weight refresh: up to kMaxRefresh fp32 master matrices -> their bf16 working copies in ONE launch
// (blockIdx.y = matrix; 4 elements per thread when the row geometry allows 16-byte loads), plus scale = exp(logit_scale)
// by thread 0.  This is the first node of a captured training step: the optimiser updates the fp32 parameters in place,
// the next replay of the same graph sees them (and the trainable logit_scale, models/mm_late.py:59-69) through this kernel.
constexpr int kMaxRefresh = 8;
struct RefreshDesc {
  const float* src;
  __nv_bfloat16* dst;
  int64_t lds, ldd;
  int rows, cols;
};
struct RefreshArgs {
  RefreshDesc d[kMaxRefresh];
  int n;
  const float* logit_scale;   // optional device scalar (dual_encoder.logit_scale)
  float* scale_out;           // [1] exp(logit_scale), clamped to (0, 40]
  uint32_t* status;           // [1] sticky: 1 if logit_scale left the supported range (exp > 40 or not finite)
  float* zero[2];             // optional: up to two fp32 ranges set to zero (the step's loss sums / small accumulators), so that
  int nzero[2];               // the first node of a captured step is this kernel and not a memset in front of it
};
__global__ void __launch_bounds__(256) refresh_weights_kernel(RefreshArgs a) {
  pdl_trigger();
  pdl_wait();
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && a.logit_scale != nullptr) {
    float sc = expf(*a.logit_scale);
    if (!(sc > 0.f) || !(sc <= 40.f)) {   // fixed-shift softmax statistics need exp(-2*scale) to stay normal in fp32
      if (a.status) *a.status = 1u;
      sc = 40.f;
    }
    *a.scale_out = sc;
  }
  {
    const int gtid = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x, gthreads = gridDim.x * gridDim.y * blockDim.x;
#pragma unroll
    for (int z = 0; z < 2; ++z)
      for (int i = gtid; i < a.nzero[z]; i += gthreads) a.zero[z][i] = 0.f;
  }
  if (static_cast<int>(blockIdx.y) >= a.n) return;
  const RefreshDesc d = a.d[blockIdx.y];
  const bool vec = (d.cols & 3) == 0 && (d.lds & 3) == 0 && (d.ldd & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(d.dst) & 7) == 0);
  if (vec) {
    // this kernel is the root of the captured step and reads cold HBM: 4 independent 16-byte loads per thread and round
    const int64_t n4 = static_cast<int64_t>(d.rows) * (d.cols >> 2), c4 = d.cols >> 2;
    const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * nthr) {
      float4 v[4];
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int64_t i = i0 + u4 * nthr;
        if (i < n4) {
          const int64_t r = i / c4, c = (i - r * c4) << 2;
          v[u4] = __ldg(reinterpret_cast<const float4*>(d.src + r * d.lds + c));
        }
      }
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const int64_t i = i0 + u4 * nthr;
        if (i < n4) {
          const int64_t r = i / c4, c = (i - r * c4) << 2;
          uint2 u;
          u.x = pack_bf16x2(v[u4].x, v[u4].y);
          u.y = pack_bf16x2(v[u4].z, v[u4].w);
          *reinterpret_cast<uint2*>(d.dst + r * d.ldd + c) = u;
        }
      }
    }
  } else {
    const int64_t n = static_cast<int64_t>(d.rows) * d.cols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / d.cols, c = i - r * d.cols;
      d.dst[r * d.ldd + c] = __float2bfloat16_rn(d.src[r * d.lds + c]);
    }
  }
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s, int64_t lds, float* __restrict__ d, int64_t ldd,
                                     int rows, int cols) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(rows) * cols) return;
  const int64_t r = idx / cols, c = idx % cols;
  d[r * ldd + c] = __bfloat162float(s[r * lds + c]);
}

__global__ void loss_mix_kernel(const float* losses, const float* itc_sums, int n_global, float bi, float bm, int use_itc,
                                int use_itm, float* out) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const float l_cls = losses[0];
  const float l_itm = use_itm ? losses[1] : 0.f;
  const float l_itc = use_itc ? 0.5f * (itc_sums[0] + itc_sums[1]) / n_global : 0.f;
  float w = 1.f;
  if (use_itc) w -= bi;
  if (use_itm) w -= bm;
  if (use_itc && use_itm) w = 1.f - (bi + bm);
  out[0] = w * l_cls + (use_itc ? bi * l_itc : 0.f) + (use_itm ? bm * l_itm : 0.f);
  out[1] = l_cls;
  out[2] = l_itc;
  out[3] = l_itm;
}

}  // namespace tic

using namespace tic;

extern "C" {

int tic_pack_cls_pairs(const void* xt, int64_t xt_stride, const void* xv, int64_t xv_stride, int B, int E,
                       const int32_t* src_idx, void* Xcat, int64_t ldx, const float* u_coin, const float* u_pick, void* stream) {
  TIC_CHECK_ARG(xt && Xcat && B > 0 && E > 0 && ldx >= 2 * E, "tic_pack_cls_pairs: bad arguments");
  TIC_CHECK_ARG((u_coin == nullptr) == (u_pick == nullptr), "tic_pack_cls_pairs: u_coin and u_pick go together");
  const int rows = (src_idx || u_coin) ? 2 * B : B;
  launch_k(pack_cls_pairs_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      static_cast<const __nv_bfloat16*>(xt), xt_stride, static_cast<const __nv_bfloat16*>(xv), xv_stride, B, E, src_idx,
      static_cast<__nv_bfloat16*>(Xcat), ldx, rows, u_coin, u_pick);
  TIC_CHECK_LAUNCH("tic_pack_cls_pairs");
  return TIC_OK;
}

int tic_unpack_cls_grad(const float* dXcat, int64_t ldd, const float* dX2, int64_t ldd2, int B, int E, const int32_t* src_idx,
                        float* dxt, int64_t ld_dxt, void* stream) {
  TIC_CHECK_ARG(dXcat && dxt && B > 0 && E > 0, "tic_unpack_cls_grad: bad arguments");
  TIC_CHECK_ARG((E & 3) == 0 && (ldd & 3) == 0 && (ldd2 & 3) == 0 && aligned16(dXcat) && aligned16(dX2),
                "tic_unpack_cls_grad: E and the leading dimensions must be multiples of 4 (16-byte rows)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = src_idx ? 2 * B : B;
  const int64_t n = static_cast<int64_t>(rows) * (E / 4);
  launch_k(unpack_accum_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, st, dXcat, ldd, dX2, ldd2, B, E, rows, src_idx,
           dxt, ld_dxt);
  TIC_CHECK_LAUNCH("tic_unpack_cls_grad");
  return TIC_OK;
}

int tic_fusion_pair_grad(const void* dH, const void* dH_lo, int64_t ld, int B, int E, int has_tim, const int32_t* src_idx, void* dPt,
                         void* dPt_lo, void* dPv, void* dPv_lo, int64_t ldo, void* stream) {
  TIC_CHECK_ARG(dH && dPt && dPv && B > 0 && E > 0 && (!has_tim || src_idx), "tic_fusion_pair_grad: bad arguments");
  TIC_CHECK_ARG((E & 7) == 0 && E <= 1024 && (ld & 7) == 0 && (ldo & 7) == 0 && aligned16(dH) && aligned16(dH_lo) && aligned16(dPt) &&
                    aligned16(dPt_lo) && aligned16(dPv) && aligned16(dPv_lo),
                "tic_fusion_pair_grad: E must be a multiple of 8 (<= 1024), rows 16-byte aligned");
  launch_k(fusion_pair_grad_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const __nv_bfloat16*>(dH), static_cast<const __nv_bfloat16*>(dH_lo), ld, B, E, has_tim, src_idx,
           static_cast<__nv_bfloat16*>(dPt), static_cast<__nv_bfloat16*>(dPt_lo), static_cast<__nv_bfloat16*>(dPv),
           static_cast<__nv_bfloat16*>(dPv_lo), ldo);
  TIC_CHECK_LAUNCH("tic_fusion_pair_grad");
  return TIC_OK;
}

int tic_heads_fwd_bwd(float* H, int64_t ldh, int B, int E, int C, int has_tim, const float* W_cls, const float* b_cls,
                      const float* W_tim, const float* b_tim, const float* y_soft, const float* class_w,
                      const int64_t* lbl_tim, const uint8_t* keep, float keep_scale, float c_cls, float c_tim,
                      float* logits_cls, float* logits_tim, float* losses, void* dH_bf16, void* dH_bf16_lo, int64_t ld_dhb, float* dH_f32,
                      int64_t ld_dhf, float* dW_cls, float* db_cls, float* dW_tim, float* db_tim, int relu_mask, float* ws, const float* dlogits_ext,
                      const float* Pt, const float* Pv, int64_t ldp, const int32_t* src_idx, void* stream) {
  TIC_CHECK_ARG(H && W_cls && b_cls && y_soft && logits_cls && losses && ws, "tic_heads_fwd_bwd: null pointer");
  TIC_CHECK_ARG((Pt == nullptr) == (Pv == nullptr) && (!Pt || !has_tim || src_idx), "tic_heads_fwd_bwd: pairwise form needs Pt, Pv (and src_idx with ITM)");
  TIC_CHECK_ARG(B > 0 && E > 0 && C >= 1 && C <= kMaxClasses, "tic_heads_fwd_bwd: need 1 <= C <= %d", kMaxClasses);
  TIC_CHECK_ARG(!has_tim || (W_tim && b_tim && lbl_tim && logits_tim), "tic_heads_fwd_bwd: ITM head pointers missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = has_tim ? 2 * B : B;
  float* dlogits = ws;
  const bool blk = C <= kBlkC && (E & 3) == 0 && E <= 2048 && (ldh & 3) == 0 && (ld_dhb & 3) == 0 && (ld_dhf & 3) == 0 && (ldp & 3) == 0 &&
                   aligned16(H) && aligned16(W_cls) && aligned16(W_tim) && aligned16(dH_f32) && aligned16(keep) && aligned16(Pt) &&
                   aligned16(Pv) && ((reinterpret_cast<uintptr_t>(dH_bf16) | reinterpret_cast<uintptr_t>(dH_bf16_lo)) & 7) == 0;
  if (blk) {
    const int tpr = ceil_div(ceil_div(E, 8), 32) * 32;    // 8 features per thread
    int rpb = 768 / tpr;
    if (rpb > 8) rpb = 8;
    if (rpb < 1) rpb = 1;
    int grid = ceil_div(rows, rpb);
    if (grid > 64) grid = 64;       // part of the machine only (see the kernel comment); c2: 64 blocks of 8 rows, ONE round
    launch_k(heads_rows_block_kernel, dim3(grid), dim3(tpr * rpb), 0, st, H, ldh, B, E, C, has_tim, W_cls, b_cls, W_tim, b_tim, y_soft,
             class_w, lbl_tim, keep, keep_scale, c_cls, c_tim, logits_cls, logits_tim, losses, dlogits, dlogits_ext,
             static_cast<__nv_bfloat16*>(dH_bf16), static_cast<__nv_bfloat16*>(dH_bf16_lo), ld_dhb, dH_f32, ld_dhf, relu_mask, Pt, Pv,
             ldp, src_idx, tpr);
  } else
  launch_k(heads_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, st, H, ldh, B, E, C, has_tim, W_cls, b_cls, W_tim, b_tim, y_soft, class_w,
                                                        lbl_tim, keep, keep_scale, c_cls, c_tim, logits_cls, logits_tim, losses,
                                                        dlogits, dlogits_ext, static_cast<__nv_bfloat16*>(dH_bf16), static_cast<__nv_bfloat16*>(dH_bf16_lo), ld_dhb,
                                                        dH_f32, ld_dhf,
                                                        relu_mask, Pt, Pv, ldp, src_idx);
  if (dW_cls && db_cls) {
    const int rpb = max(8, ceil_div(B, 128));
    const int heads = (has_tim && dW_tim && db_tim) ? 2 : 1;
    dim3 grid(ceil_div(E, 128), ceil_div(B, rpb), heads);
    launch_k(heads_wgrad_kernel, grid, dim3(128), 0, st, H, ldh, B, E, C, dlogits, keep, keep_scale, dW_cls, db_cls, dW_tim, db_tim,
             rpb);
  }
  TIC_CHECK_LAUNCH("tic_heads_fwd_bwd");
  return TIC_OK;
}

int tic_heads_wgrad(const float* H, int64_t ldh, int B, int E, int C, int has_tim, const float* dlogits, const uint8_t* keep,
                    float keep_scale, float* dW_cls, float* db_cls, float* dW_tim, float* db_tim, void* stream) {
  TIC_CHECK_ARG(H && dlogits && dW_cls && db_cls && B > 0 && E > 0 && C >= 1 && C <= kMaxClasses, "tic_heads_wgrad: bad arguments");
  TIC_CHECK_ARG(!has_tim || (dW_tim && db_tim), "tic_heads_wgrad: ITM head pointers missing");
  const int rpb = max(8, ceil_div(B, 128));
  dim3 grid(ceil_div(E, 128), ceil_div(B, rpb), has_tim ? 2 : 1);
  launch_k(heads_wgrad_kernel, grid, dim3(128), 0, static_cast<cudaStream_t>(stream), H, ldh, B, E, C, dlogits, keep, keep_scale,
           dW_cls, db_cls, dW_tim, db_tim, rpb);
  TIC_CHECK_LAUNCH("tic_heads_wgrad");
  return TIC_OK;
}

int tic_colsum_bf16_pair(const void* X, const void* X_lo, int64_t ldx, int rows, int cols, float* out, void* stream) {
  TIC_CHECK_ARG(X && out && rows > 0 && cols > 0, "tic_colsum_bf16: bad arguments");
  const int rpb = max(8, ceil_div(rows, 128));
  dim3 grid(ceil_div(cols, 128), ceil_div(rows, rpb));
  launch_k(colsum_bf16_kernel, dim3(grid), dim3(128), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(X),
           static_cast<const __nv_bfloat16*>(X_lo), ldx, rows, cols, out, rpb);
  TIC_CHECK_LAUNCH("tic_colsum_bf16");
  return TIC_OK;
}
int tic_colsum_bf16(const void* X, int64_t ldx, int rows, int cols, float* out, void* stream) {
  return tic_colsum_bf16_pair(X, nullptr, ldx, rows, cols, out, stream);
}

int tic_cast_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, void* stream) {
  TIC_CHECK_ARG(src && dst && rows > 0 && cols > 0, "tic_cast_f32_to_bf16: bad arguments");
  const int64_t n = static_cast<int64_t>(rows) * cols;
  launch_k(cast_f32_bf16_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      src, lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  TIC_CHECK_LAUNCH("tic_cast_f32_to_bf16");
  return TIC_OK;
}
int tic_cast_bf16_to_f32(const void* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, void* stream) {
  TIC_CHECK_ARG(src && dst && rows > 0 && cols > 0, "tic_cast_bf16_to_f32: bad arguments");
  const int64_t n = static_cast<int64_t>(rows) * cols;
  launch_k(cast_bf16_f32_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(src), lds, dst, ldd, rows, cols);
  TIC_CHECK_LAUNCH("tic_cast_bf16_to_f32");
  return TIC_OK;
}

int tic_refresh_weights(int n, const float* const* src_host, void* const* dst_host, const int64_t* lds_host, const int64_t* ldd_host,
                        const int* rows_host, const int* cols_host, const float* logit_scale, float* scale_out, uint32_t* status,
                        float* zero0, int nzero0, float* zero1, int nzero1, void* stream) {
  TIC_CHECK_ARG(n >= 0 && n <= kMaxRefresh, "tic_refresh_weights: at most %d matrices per call", kMaxRefresh);
  TIC_CHECK_ARG(n > 0 || logit_scale || nzero0 > 0 || nzero1 > 0, "tic_refresh_weights: nothing to do");
  TIC_CHECK_ARG(nzero0 >= 0 && nzero1 >= 0 && (nzero0 == 0 || zero0) && (nzero1 == 0 || zero1), "tic_refresh_weights: bad zero ranges");
  TIC_CHECK_ARG(!logit_scale || scale_out, "tic_refresh_weights: scale_out is NULL");
  RefreshArgs a{};
  a.n = n;
  a.logit_scale = logit_scale;
  a.scale_out = scale_out;
  a.status = status;
  a.zero[0] = zero0; a.nzero[0] = nzero0;
  a.zero[1] = zero1; a.nzero[1] = nzero1;
  int64_t most = 1;
  for (int i = 0; i < n; ++i) {
    TIC_CHECK_ARG(src_host[i] && dst_host[i] && rows_host[i] > 0 && cols_host[i] > 0, "tic_refresh_weights: bad matrix %d", i);
    a.d[i] = RefreshDesc{src_host[i], static_cast<__nv_bfloat16*>(dst_host[i]), lds_host[i], ldd_host[i], rows_host[i], cols_host[i]};
    const int64_t e = static_cast<int64_t>(rows_host[i]) * cols_host[i];
    if (e > most) most = e;
  }
  // 4 vectors per thread (one round of 4 independent loads) for the largest matrix: ~1150 blocks for the c2 weights = one
  // resident wave; smaller matrices leave their surplus blocks idle (they exit at once)
  int gx = static_cast<int>((most / 4 + 1023) / 1024);
  if (gx < 1) gx = 1;
  if (gx > 592) gx = 592;
  launch_k(refresh_weights_kernel, dim3(gx, n > 0 ? n : 1), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
  TIC_CHECK_LAUNCH("tic_refresh_weights");
  return TIC_OK;
}

int tic_loss_mix(const float* losses, const float* itc_sums, int n_global, float beta_itc, float beta_itm, int use_itc,
                 int use_itm, float* out, void* stream) {
  TIC_CHECK_ARG(losses && out && (!use_itc || itc_sums), "tic_loss_mix: bad arguments");
  launch_k(loss_mix_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), losses, itc_sums, n_global, beta_itc, beta_itm, use_itc,
                                                                  use_itm, out);
  TIC_CHECK_LAUNCH("tic_loss_mix");
  return TIC_OK;
}

}  // extern "C"
