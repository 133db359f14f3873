// Image-text contrastive (ITC) path: similarity tiles on tcgen05/TMEM with the bidirectional softmax
// cross-entropy fused into the epilogue (logits are never written unless asked for), and its backward.
//
// Reference arithmetic: HF VisionTextDualEncoderModel.forward :268-273 (normalise, exp(logit_scale) * T V^T),
// in-tree statement models/mm_early.py:96-103; loss models/utils.py:225-231.
//
// The L2 normalisation is folded into the tile epilogue:  S_ij = scale * rinv_t[i] * rinv_v[j] * <t_i, v_j>,
// so the MMA consumes the raw bf16 embeddings (no extra rounding of normalised copies).
// Softmax statistics use the fixed shift `shift >= max S` (= scale, because |cos| <= 1), which turns both the row and
// the column log-sum-exp into plain sums: per-tile partial sums are written out and reduced in a fixed order
// (deterministic; across ranks it is a SUM all-reduce).
#include <cstdlib>
#include "common.cuh"
#include "tic_umma.cuh"
#include "itm_rule.cuh"

namespace tic {

constexpr int kItcBN = 256;        // wide tiles: best tensor-pipe efficiency at large batch
constexpr int kItcBNSmall = 64;    // narrow tiles below kItcSmallN columns: 4x more CTAs for the latency-bound small batch
constexpr int kItcSmallN = 2048;
inline int itc_bn(int n_global) { return n_global <= kItcSmallN ? kItcBNSmall : kItcBN; }
// cluster split-K factor of the narrow-tile (small batch) similarity kernels; not combined with segment-consuming mode
inline int itc_small_kc(int m_local, int n_global, int P, const void* T_lo, const void* V_lo, const SegOrder* sop) {
  if (sop != nullptr) return 1;
  const int tiles = ceil_div(m_local, kBM) * ceil_div(n_global, kItcBNSmall);
  const int total_kb = ceil_div(P, kBK) * (1 + (T_lo ? 1 : 0) + (V_lo ? 1 : 0));
  return pick_cluster_k(tiles, total_kb);
}
// TIC_ITC_MULTICAST=0 disables the 2-CTA TMA-multicast variant (A/B measurement switch, not a fallback).
inline bool itc_multicast() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_ITC_MULTICAST"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// TIC_ITC_TMA_STORE=0 keeps the direct 16-byte stores of the gradient operand (A/B measurement switch).
inline bool itc_tma_store() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TIC_ITC_TMA_STORE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
constexpr int kItcEpiWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Sum the 32 per-lane values of each of 32 columns across the warp: returns, in lane l, sum over lanes of v[l].
// Recursive-halving butterfly: 31 shuffles for 32 columns.
__device__ __forceinline__ float warp_col_sums(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 16];
    const float keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 8];
    const float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 4];
    const float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? v[i] : v[i + 2];
    const float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// ------------------------------------------------------------------ forward epilogue
struct ItcFwdEpi {
  struct Params {
    float* rinv_t;       // inputs, or OUTPUTS when the sum-of-squares partials below are given
    float* rinv_v;
    const float* ss_t_part;   // optional [n_ss_t][M] / [n_ss_v][N] row sum-of-squares partials from tic_gemm_bf16_rowss:
    const float* ss_v_part;   // the L2 normalisation (HF :268-269) then costs no kernel of its own
    int n_ss_t, n_ss_v;
    float scale_log2e;   // scale * log2(e)
    float shift_log2e;   // shift * log2(e)
    float scale;
    float shift;
    float* row_part;     // [n_tiles * nparts][M]
    float* col_part;     // [m_tiles][N]
    float* diag;         // [M]
    float* logits;       // optional [M, ld_logits]
    int64_t ld_logits;
    int row_offset;
    const float* scale_dev;     // optional DEVICE scalar exp(logit_scale): overrides scale and shift (a trainable logit_scale
                                // reaches a captured step by pointer, never as a launch-time constant)
    unsigned long long* qpart;  // optional [ceil(N/32)][M]: integer sums of the hard-negative sampling weights per 32-column chunk
  };
  static constexpr int scratch_bytes(int bn) { return 6 * bn * 4 + 256; }   // 2 x rinv_v + 4 x partial column sums
  // column norms of this tile -> shared memory, this thread's row norm -> cx.pre[0]; runs while the MMAs of the tile are in flight
  template <int BN>
  __device__ static void prefetch(const Params& p, EpiCtx& cx) {
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    const bool valid_row = row < cx.M;
    float* sb = reinterpret_cast<float*>(cx.scratch) + (cx.iter & 1) * BN;   // rinv_v of this tile (double-buffered)
    for (int j = cx.epi_tid; j < BN; j += cx.epi_threads) {
      float b = 0.f;
      if (cx.n0 + j < cx.N) {
        if (p.ss_v_part != nullptr) {
          float ssum = 0.f;
          for (int q = 0; q < p.n_ss_v; ++q) ssum += __ldg(p.ss_v_part + static_cast<int64_t>(q) * cx.N + cx.n0 + j);
          b = 1.0f / sqrtf(ssum);   // no epsilon (HF :268-269)
          if (cx.m_blk == 0) p.rinv_v[cx.n0 + j] = b;
        } else {
          b = p.rinv_v[cx.n0 + j];
        }
      }
      sb[j] = b;
    }
    epi_bar_sync(cx.epi_threads);
    float rt = 0.f;
    if (valid_row) {
      if (p.ss_t_part != nullptr) {
        float ssum = 0.f;
        for (int q = 0; q < p.n_ss_t; ++q) ssum += __ldg(p.ss_t_part + static_cast<int64_t>(q) * cx.M + row);
        rt = 1.0f / sqrtf(ssum);
        if (cx.n_blk == 0 && cx.part == 0) p.rinv_t[row] = rt;
      } else {
        rt = p.rinv_t[row];
      }
    }
    cx.pre[0] = rt;
    float sc = p.scale, sh = p.shift;
    if (p.scale_dev != nullptr) { sc = __ldg(p.scale_dev); sh = sc; }
    cx.pre[1] = sc;
    cx.pre[2] = sh;
  }
  template <int BN>
  __device__ static void tile(const Params& p, const EpiCtx& cx) {
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    const bool valid_row = row < cx.M;
    const float* sb = reinterpret_cast<const float*>(cx.scratch) + (cx.iter & 1) * BN;   // rinv_v of this tile (prefetch)
    float* scol = reinterpret_cast<float*>(cx.scratch) + 2 * BN;                         // [4 quads][BN] partial column sums
    const float rt = cx.pre[0];
    const float sc = cx.pre[1], sh = cx.pre[2];                      // scale, shift (host constants or the device scalar)
    const float scale_log2e = sc * kLog2e, shift_log2e = sh * kLog2e;
    const float rt2 = rt * scale_log2e;
    const float negshift = valid_row ? -shift_log2e : -INFINITY;
    const int gcol = p.row_offset + row;  // column holding this row's positive
    const int cols_per_part = BN / cx.nparts;
    float rowsum = 0.f;
    const uint32_t trow = cx.tmem_acc + (static_cast<uint32_t>(cx.quad * 32) << 16);
    const int nchunk = cols_per_part / 32;
    uint32_t vn[32];
    if (cx.n0 + cx.part * cols_per_part < cx.N) tmem_ld_32x32(trow + cx.part * cols_per_part, vn);
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
      const int cl = cx.part * cols_per_part + c * 32;
      const int col0 = cx.n0 + cl;
      if (col0 >= cx.N) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = vn[j];
      if (c + 1 < nchunk && col0 + 32 < cx.N) tmem_ld_32x32(trow + cl + 32, vn);   // prefetch the next 32 columns
      float e[32];
      if (col0 + 32 <= cx.N && p.logits == nullptr && p.qpart != nullptr) {
        // hard-negative fast path (interior tile, logits not materialised): the statistics as in the plain fast path plus the
        // integer sampling weight of every logit, formed on the FMA / ALU pipes only (itm_rule.cuh) beside the one ex2 per
        // element; the logit itself is itc_logit() — the expression the pick tiles repeat bit for bit.
        unsigned long long qs = 0;
        float dv = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float cosv = __fmul_rn(__fmul_rn(__uint_as_float(v[j4 + q]), rt), bb[q]);
            const float sij = __fmul_rn(cosv, sc);
            e[j4 + q] = ex2_approx(fmaf(cosv, scale_log2e, negshift));
            const unsigned long long qw = hard_qweight(sij, sh);
            if (col0 + j4 + q == gcol) dv = sij; else qs += qw;
          }
        }
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) { r0 += e[j]; r1 += e[j + 1]; r2 += e[j + 2]; r3 += e[j + 3]; }
        rowsum += (r0 + r1) + (r2 + r3);
        if (valid_row) {
          if (gcol >= col0 && gcol < col0 + 32) p.diag[row] = dv;
          p.qpart[static_cast<int64_t>(col0 >> 5) * cx.M + row] = qs;
        }
      } else if (col0 + 32 <= cx.N && p.logits == nullptr) {
        // fast path (interior tile, logits not materialised): 1 FMUL + 1 FFMA + 1 MUFU + 1 FADD per element.
        // invalid rows carry negshift = -inf, so their exp is exactly 0 and they drop out of the column sums.
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j4);
          e[j4 + 0] = ex2_approx(fmaf(__uint_as_float(v[j4 + 0]) * b4.x, rt2, negshift));
          e[j4 + 1] = ex2_approx(fmaf(__uint_as_float(v[j4 + 1]) * b4.y, rt2, negshift));
          e[j4 + 2] = ex2_approx(fmaf(__uint_as_float(v[j4 + 2]) * b4.z, rt2, negshift));
          e[j4 + 3] = ex2_approx(fmaf(__uint_as_float(v[j4 + 3]) * b4.w, rt2, negshift));
        }
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) { r0 += e[j]; r1 += e[j + 1]; r2 += e[j + 2]; r3 += e[j + 3]; }
        rowsum += (r0 + r1) + (r2 + r3);
        if (gcol >= col0 && gcol < col0 + 32 && valid_row) {   // the positive of this row lives in this chunk (rare)
          float dv = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j == gcol) dv = itc_logit(__uint_as_float(v[j]), rt, sb[cl + j], sc);
          p.diag[row] = dv;
        }
      } else {
        // generic path: edge tiles (column tail), materialised logits (drop-in API) and the hard-negative weight sums.
        // The logit is formed by itc_logit() — the expression the pick tiles repeat bit for bit.
        float dsel = 0.f;
        unsigned long long qsum = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float cosv = __fmul_rn(__fmul_rn(__uint_as_float(v[j]), rt), sb[cl + j]);   // cosine similarity
          const float sij = __fmul_rn(cosv, sc);                                            // == itc_logit(...)
          const bool ok = valid_row && (col0 + j < cx.N);
          e[j] = ok ? exp2f(fmaf(cosv, scale_log2e, -shift_log2e)) : 0.f;
          rowsum += e[j];
          if (col0 + j == gcol) dsel = sij;
          else if (p.qpart != nullptr && ok) qsum += hard_qweight(sij, sh);
          v[j] = __float_as_uint(sij);
        }
        if (valid_row && gcol >= col0 && gcol < col0 + 32) p.diag[row] = dsel;
        if (p.qpart != nullptr && valid_row) p.qpart[static_cast<int64_t>(col0 >> 5) * cx.M + row] = qsum;
        if (p.logits != nullptr && valid_row) {
          float* d = p.logits + static_cast<int64_t>(row) * p.ld_logits + col0;
          if (col0 + 32 <= cx.N && (p.ld_logits & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(d + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < cx.N) d[j] = __uint_as_float(v[j]);
          }
        }
      }
      if (p.col_part != nullptr) {   // symmetric multi-GPU mode computes the column statistics as rows of the swapped block
        const float cs = warp_col_sums(e);
        scol[cx.quad * BN + cl + lane] = cs;
      }
    }
    if (valid_row) p.row_part[static_cast<int64_t>(cx.n_blk * cx.nparts + cx.part) * cx.M + row] = rowsum;
    if (p.col_part == nullptr) return;
    epi_bar_sync(cx.epi_threads);
    for (int j = cx.epi_tid; j < BN; j += cx.epi_threads)
      if (cx.n0 + j < cx.N)
        p.col_part[static_cast<int64_t>(cx.m_blk) * cx.N + cx.n0 + j] =
            (scol[j] + scol[BN + j]) + (scol[2 * BN + j] + scol[3 * BN + j]);
  }
};

// ------------------------------------------------------------------ hard-negative pick epilogue
// Third step of the tile-stream hard-negative sampler (SURVEY.md a-5', BASELINE north_star "fused hard-negative
// multinomial-sampling"): the forward tiles summed the integer sampling weights per (row, column part), tic_itm_hard_locate
// turned each row's uniform into (part, residual target); these tiles recompute S and walk the ONE located part of every
// mismatch row: src = first column whose running weight sum exceeds the residual.  S never exists in HBM; what the sampler
// reads instead is 8 bytes per (row, part).  Recomputed accumulators are bit-identical (same operands, same UMMA shapes,
// same k order), and the logit / weight expressions are the shared itc_logit() / hard_qweight().
struct ItcPickEpi {
  struct Params {
    const float* rinv_t;
    const float* rinv_v;
    float scale, shift;
    const float* scale_dev;
    int row_offset;
    const int32_t* loc_part;               // [M] 32-column chunk holding the row's target, -1 = nothing to pick
    const unsigned long long* loc_res;     // [M] target minus the weight of the parts before it
    int32_t* src_idx;                      // [M] out (only located rows are written)
  };
  static constexpr int scratch_bytes(int bn) { return 2 * bn * 4 + 256; }
  template <int BN>
  __device__ static void prefetch(const Params& p, EpiCtx& cx) {
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    float* sb = reinterpret_cast<float*>(cx.scratch) + (cx.iter & 1) * BN;
    for (int j = cx.epi_tid; j < BN; j += cx.epi_threads) sb[j] = (cx.n0 + j < cx.N) ? __ldg(p.rinv_v + cx.n0 + j) : 0.f;
    epi_bar_sync(cx.epi_threads);
    const bool valid_row = row < cx.M;
    cx.pre[0] = valid_row ? __ldg(p.rinv_t + row) : 0.f;
    float sc = p.scale, sh = p.shift;
    if (p.scale_dev != nullptr) { sc = __ldg(p.scale_dev); sh = sc; }
    cx.pre[1] = sc;
    cx.pre[2] = sh;
    const int lp = valid_row ? __ldg(p.loc_part + row) : -1;
    cx.pre[3] = __int_as_float(lp);
    const unsigned long long res = lp >= 0 ? __ldg(p.loc_res + row) : 0ull;
    cx.pre[4] = __uint_as_float(static_cast<uint32_t>(res));
    cx.pre[5] = __uint_as_float(static_cast<uint32_t>(res >> 32));
  }
  template <int BN>
  __device__ static void tile(const Params& p, const EpiCtx& cx) {
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    const int cols_per_part = BN / cx.nparts;
    const int lp = __float_as_int(cx.pre[3]);              // 32-column chunk holding this row's target (-1: none)
    const float* sb = reinterpret_cast<const float*>(cx.scratch) + (cx.iter & 1) * BN;
    const float rt = cx.pre[0], sc = cx.pre[1], sh = cx.pre[2];
    const unsigned long long res = static_cast<unsigned long long>(__float_as_uint(cx.pre[4])) |
                                   (static_cast<unsigned long long>(__float_as_uint(cx.pre[5])) << 32);
    const int gcol = p.row_offset + row;
    const uint32_t trow = cx.tmem_acc + (static_cast<uint32_t>(cx.quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < cols_per_part / 32; ++c) {
      const int cl = cx.part * cols_per_part + c * 32;
      const int col0 = cx.n0 + cl;
      if (col0 >= cx.N) break;  // warp-uniform
      const bool mine = lp == (col0 >> 5);
      if (!__any_sync(0xffffffffu, mine)) continue;        // warp-uniform: no row of this warp has its target in this chunk
      uint32_t v[32];
      tmem_ld_32x32(trow + cl, v);
      tmem_ld_wait();
      if (mine) {
        unsigned long long q[32];                          // 32 independent weights first (ILP), then the short integer scan
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          q[j] = (col < cx.N && col != gcol) ? hard_qweight(itc_logit(__uint_as_float(v[j]), rt, sb[cl + j], sc), sh) : 0ull;
        }
        unsigned long long run = 0;
        int found = -1;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          run += q[j];
          if (found < 0 && run > res) found = col0 + j;
        }
        if (found >= 0) p.src_idx[row] = found;
      }
    }
  }
};

// ------------------------------------------------------------------ backward epilogue: gradient operands
struct ItcBwdEpi {
  struct Params {
    const float* rinv_t;
    const float* rinv_v;
    const float* lse_row;
    const float* lse_col;
    float scale_log2e;
    float gscale;        // g / (2B)
    __nv_bfloat16* GA;   // [M, ld_ga]   Gp * rinv_v[j]
    int64_t ld_ga;
    __nv_bfloat16* GBT;  // "GB" [M, ld_ga]  Gp * rinv_t[i], row-major like GA (read MN-major by the image-side GEMM); NULL = single-operand modes
    int64_t ld_gbt;
    __nv_bfloat16* GA_lo;   // optional bf16 residuals (split precision for small batches)
    __nv_bfloat16* GBT_lo;
    // inline statistics (small batch): lse straight from the forward partials, so the lse kernel leaves the critical path
    const float* row_part;  // [n_row_parts][M] or nullptr -> lse_row
    const float* col_part;  // [n_col_parts][N] or nullptr -> lse_col
    int n_row_parts, n_col_parts;
    float shift;
    float shift_log2e;   // common offset of the factored form below (keeps 2^(lr-off), 2^(off-lc) in fp32 range)
    int factored;        // fast path: ONE ex2 per element (host enables it for scale <= 20)
    int tma_store;       // GA leaves through shared memory + TMA 32x32 tile stores (full lines) instead of 16-byte pieces
    const float* scale_dev;   // optional DEVICE scalar exp(logit_scale): overrides scale_log2e / shift / shift_log2e / factored
    alignas(64) CUtensorMap tmap_ga;
  };
  // narrow tiles never use the TMA-store staging area (that path needs the single-operand large-batch mode)
  static constexpr int scratch_bytes(int bn) { return bn == 64 ? 6 * bn * 4 + 256 : kEpiScratchBytes; }
  // per-column vectors of this tile -> shared memory, this thread's row norm / row lse -> cx.pre[0..1]; runs while the MMAs of
  // the tile are in flight (small batch: lse comes from up to 64 forward partials per row and column — a long load chain).
  // COH: the statistics were written by OTHER CTAs of this very kernel (fused forward+backward, below): they must not be
  // read through the non-coherent path (ld.global.nc is only defined for data that is read-only during the kernel).
  template <int BN, bool COH = false>
  __device__ static void prefetch(const Params& p, EpiCtx& cx) {
    const bool coh = COH || cx.coherent;
    auto ldf = [coh](const float* q) { return coh ? __ldcg(q) : __ldg(q); };
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    const bool valid_row = row < cx.M;
    float* sb = reinterpret_cast<float*>(cx.scratch) + (cx.iter & 1) * 3 * BN;  // rinv_v | -lse_col*log2e | gscale*rinv_v
    float* sl = sb + BN;
    float* sg = sb + 2 * BN;
    float scale_log2e = p.scale_log2e, shift = p.shift, shift_log2e = p.shift_log2e;
    int factored = p.factored;
    if (p.scale_dev != nullptr) {
      const float sc = __ldg(p.scale_dev);
      scale_log2e = sc * kLog2e; shift = sc; shift_log2e = scale_log2e;
      factored = p.factored && sc <= 20.f;
    }
    for (int j = cx.epi_tid; j < BN; j += cx.epi_threads) {
      const bool ok = cx.n0 + j < cx.N;
      const float b = ok ? ldf(p.rinv_v + cx.n0 + j) : 0.f;
      sb[j] = b;
      float lc = 0.f;
      if (ok) {
        if (p.col_part != nullptr) {
          float cs = 0.f;
          for (int q = 0; q < p.n_col_parts; ++q) cs += ldf(p.col_part + static_cast<int64_t>(q) * cx.N + cx.n0 + j);
          lc = shift + logf(cs);
        } else {
          lc = ldf(p.lse_col + cx.n0 + j);
        }
      }
      // factored fast path:  e^{S-lse_row} + e^{S-lse_col} = e1 * (1 + Er_i * Ec_j),  e1 = 2^(s2 - lr),
      //   Er_i = 2^(lr - off), Ec_j = 2^(off - lc)  ->  GA = e1 * (g_j + Er_i * (Ec_j g_j)),  g_j = gscale * rinv_v[j]
      sl[j] = factored ? (ok ? exp2f(shift_log2e - lc * kLog2e) * b * p.gscale : 0.f) : lc * kLog2e;
      sg[j] = b * p.gscale;
    }
    epi_bar_sync(cx.epi_threads);
    const float rt = valid_row ? ldf(p.rinv_t + row) : 0.f;
    float lr = 0.f;
    if (valid_row) {
      if (p.row_part != nullptr) {
        float rs = 0.f;
        for (int q = 0; q < p.n_row_parts; ++q) rs += ldf(p.row_part + static_cast<int64_t>(q) * cx.M + row);
        lr = (shift + logf(rs)) * kLog2e;
      } else {
        lr = ldf(p.lse_row + row) * kLog2e;
      }
    }
    cx.pre[0] = rt;
    cx.pre[1] = lr;
    cx.pre[2] = scale_log2e;
    cx.pre[3] = shift_log2e;
    cx.pre[4] = __int_as_float(factored);
  }
  template <int BN>
  __device__ static void tile(const Params& p, const EpiCtx& cx) {
    const int lane = threadIdx.x & 31;
    const int row = cx.m0 + cx.quad * 32 + lane;
    const bool valid_row = row < cx.M;
    const float* sb = reinterpret_cast<const float*>(cx.scratch) + (cx.iter & 1) * 3 * BN;  // filled by prefetch
    const float* sl = sb + BN;
    const float* sg = sb + 2 * BN;
    const float rt = cx.pre[0], lr = cx.pre[1];
    const float scale_log2e = cx.pre[2], shift_log2e = cx.pre[3];
    const bool factored = __float_as_int(cx.pre[4]) != 0;
    const float rt2 = rt * scale_log2e;
    const int cols_per_part = BN / cx.nparts;
    const bool vec_ok = (p.ld_ga & 7) == 0;
    const uint32_t trow = cx.tmem_acc + (static_cast<uint32_t>(cx.quad * 32) << 16);
    const int nchunk = cols_per_part / 32;
    uint32_t vn[32];
    if (cx.n0 + cx.part * cols_per_part < cx.N) tmem_ld_32x32(trow + cx.part * cols_per_part, vn);
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
      const int cl = cx.part * cols_per_part + c * 32;
      const int col0 = cx.n0 + cl;
      if (col0 >= cx.N) break;
      uint32_t v[32];
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = vn[j];
      if (c + 1 < nchunk && col0 + 32 < cx.N) tmem_ld_32x32(trow + cl + 32, vn);   // prefetch the next 32 columns
      float ga[32];
      if (col0 + 32 <= cx.N && p.GBT == nullptr && factored) {
        // fast path (interior tile, no transposed operand): FMUL, FFMA, MUFU, FFMA, FMUL per element — the XU pipe (16 ex2
        // per clock per SM) is what bounds this epilogue, so the second exponential is replaced by the factored form.
        const float er = valid_row ? exp2f(lr - shift_log2e) : 0.f;
        const float nlr = valid_row ? -lr : -INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j4);
          const float4 c4 = *reinterpret_cast<const float4*>(sl + cl + j4);
          const float4 g4 = *reinterpret_cast<const float4*>(sg + cl + j4);
          ga[j4 + 0] = ex2_approx(fmaf(__uint_as_float(v[j4 + 0]) * b4.x, rt2, nlr)) * fmaf(er, c4.x, g4.x);
          ga[j4 + 1] = ex2_approx(fmaf(__uint_as_float(v[j4 + 1]) * b4.y, rt2, nlr)) * fmaf(er, c4.y, g4.y);
          ga[j4 + 2] = ex2_approx(fmaf(__uint_as_float(v[j4 + 2]) * b4.z, rt2, nlr)) * fmaf(er, c4.z, g4.z);
          ga[j4 + 3] = ex2_approx(fmaf(__uint_as_float(v[j4 + 3]) * b4.w, rt2, nlr)) * fmaf(er, c4.w, g4.w);
        }
      } else if (col0 + 32 <= cx.N && p.GBT == nullptr) {
        // two-exponential form of the same fast path (scale > 20: the factored product could leave the fp32 range)
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j4);
          const float4 l4 = *reinterpret_cast<const float4*>(sl + cl + j4);
          const float4 g4 = *reinterpret_cast<const float4*>(sg + cl + j4);
          const float s0 = __uint_as_float(v[j4 + 0]) * b4.x * rt2, s1 = __uint_as_float(v[j4 + 1]) * b4.y * rt2;
          const float s2_ = __uint_as_float(v[j4 + 2]) * b4.z * rt2, s3 = __uint_as_float(v[j4 + 3]) * b4.w * rt2;
          ga[j4 + 0] = (ex2_approx(s0 - lr) + ex2_approx(s0 - l4.x)) * g4.x;
          ga[j4 + 1] = (ex2_approx(s1 - lr) + ex2_approx(s1 - l4.y)) * g4.y;
          ga[j4 + 2] = (ex2_approx(s2_ - lr) + ex2_approx(s2_ - l4.z)) * g4.z;
          ga[j4 + 3] = (ex2_approx(s3 - lr) + ex2_approx(s3 - l4.w)) * g4.w;
        }
      } else if (p.GBT == nullptr) {
        // edge tile of the single-operand modes
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s2 = __uint_as_float(v[j]) * rt * sb[cl + j] * scale_log2e;  // S * log2e
          if (factored) ga[j] = exp2f(s2 - lr) * fmaf(exp2f(lr - shift_log2e), sl[cl + j], sg[cl + j]);   // sl = Ec_j * g_j, sg = g_j
          else ga[j] = p.gscale * (exp2f(s2 - lr) + exp2f(s2 - sl[cl + j])) * sb[cl + j];
        }
      } else {
        // Two-operand (small batch, split precision) mode: GA = Gp * rinv_v[j] and GB = Gp * rinv_t[i], BOTH row-major with the
        // layout of S — the image-side GEMM reads GB MN-major, so no transposed operand is written (the transposed form cost
        // 64 scattered 2-byte stores per thread and chunk, and was the longest part of this epilogue at B = 256).
        // ga[] holds Gp here; the two scalings happen at the stores.
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j4);
          const float4 l4 = *reinterpret_cast<const float4*>(sl + cl + j4);
          const float s0 = __uint_as_float(v[j4 + 0]) * b4.x * rt2, s1 = __uint_as_float(v[j4 + 1]) * b4.y * rt2;
          const float s2_ = __uint_as_float(v[j4 + 2]) * b4.z * rt2, s3 = __uint_as_float(v[j4 + 3]) * b4.w * rt2;
          ga[j4 + 0] = p.gscale * (ex2_approx(s0 - lr) + ex2_approx(s0 - l4.x));
          ga[j4 + 1] = p.gscale * (ex2_approx(s1 - lr) + ex2_approx(s1 - l4.y));
          ga[j4 + 2] = p.gscale * (ex2_approx(s2_ - lr) + ex2_approx(s2_ - l4.z));
          ga[j4 + 3] = p.gscale * (ex2_approx(s3 - lr) + ex2_approx(s3 - l4.w));
        }
        if (cx.stage != nullptr && col0 + 32 <= cx.N && vec_ok) {
          // Staged stores: a thread owns one ROW of the 32x32 chunk, so its direct 16-byte stores touch 32 different lines per
          // instruction, each sector half-written — with a cold / dirty L2 these partial-sector writes took 11 of the kernel's
          // 24 us (phase stamps, scripts/kernel_times.py).  Through a swizzled 2 KB staging tile per warp every store
          // instruction writes 8 rows x 64 contiguous bytes (full sectors).  Swizzle as in the TMA-store path: 16-byte chunk q
          // of row r sits at q ^ ((r >> 1) & 3): conflict-free on both sides.
          const int wslot = (threadIdx.x >> 5) - 2;
          uint8_t* stg = cx.stage + wslot * 2048;
          const uint32_t rowb = smem_u32(stg) + lane * 64, sw = (lane >> 1) & 3;
          // ONE rolled loop over the four outputs (hi / residual x column- / row-scaled): the cold instruction fetch of this
          // epilogue is on the step's critical chain (scripts/fused_in_step.py), so its code is kept small.
#pragma unroll 1
          for (int which = 0; which < 4; ++which) {
            __nv_bfloat16* dst = which == 0 ? p.GA : (which == 1 ? p.GA_lo : (which == 2 ? p.GBT : p.GBT_lo));
            if (dst == nullptr) continue;
            const bool by_col = which < 2, residual = which & 1;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float x[8];
#pragma unroll
              for (int e8 = 0; e8 < 8; ++e8) {
                x[e8] = ga[8 * q + e8] * (by_col ? sb[cl + 8 * q + e8] : rt);
                if (residual) x[e8] -= __bfloat162float(__float2bfloat16_rn(x[e8]));
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((q ^ sw) << 4)), "r"(pack_bf16x2(x[0], x[1])),
                           "r"(pack_bf16x2(x[2], x[3])), "r"(pack_bf16x2(x[4], x[5])), "r"(pack_bf16x2(x[6], x[7])) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + (lane >> 2), c = lane & 3;
              const uint4 u = *reinterpret_cast<const uint4*>(stg + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
              const int grow = cx.m0 + cx.quad * 32 + r;
              if (grow < cx.M) *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(grow) * p.ld_ga + col0 + c * 8) = u;
            }
          }
        } else if (valid_row) {
          const bool full = col0 + 32 <= cx.N && vec_ok;
          auto put = [&](__nv_bfloat16* hi, __nv_bfloat16* lo, bool by_col) {
            __nv_bfloat16* d = hi + static_cast<int64_t>(row) * p.ld_ga + col0;
            __nv_bfloat16* dl = lo ? lo + static_cast<int64_t>(row) * p.ld_ga + col0 : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float x[8], r8[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                x[q] = ga[j + q] * (by_col ? sb[cl + j + q] : rt);
                r8[q] = x[q] - __bfloat162float(__float2bfloat16_rn(x[q]));
              }
              if (full) {
                uint4 u;
                u.x = pack_bf16x2(x[0], x[1]); u.y = pack_bf16x2(x[2], x[3]); u.z = pack_bf16x2(x[4], x[5]); u.w = pack_bf16x2(x[6], x[7]);
                *reinterpret_cast<uint4*>(d + j) = u;
                if (dl) {
                  u.x = pack_bf16x2(r8[0], r8[1]); u.y = pack_bf16x2(r8[2], r8[3]); u.z = pack_bf16x2(r8[4], r8[5]); u.w = pack_bf16x2(r8[6], r8[7]);
                  *reinterpret_cast<uint4*>(dl + j) = u;
                }
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (col0 + j + q < cx.N) {
                    d[j + q] = __float2bfloat16_rn(x[q]);
                    if (dl) dl[j + q] = __float2bfloat16_rn(r8[q]);
                  }
              }
            }
          };
          put(p.GA, p.GA_lo, true);
          put(p.GBT, p.GBT_lo, false);
        }
        continue;
      }
      if (p.tma_store && col0 + 32 <= cx.N) {
        // Each lane's 32 bf16 (64 B) go to its row of this warp's 32x32 staging tile (64-byte swizzle: 16-byte chunk q of row r
        // sits at chunk q ^ ((r >> 1) & 3), bank-conflict free), then ONE bulk tensor store writes full 64-byte row segments;
        // rows past M are clipped by the TMA unit.  The direct form (16-byte pieces at a 2N-byte stride per lane) made every
        // warp store touch 32 half-filled sectors and its back-pressure was the top stall of this kernel.
        const int wslot = (threadIdx.x >> 5) - 2;
        const uint32_t stg = smem_u32(cx.scratch) + kEpiStageOff + wslot * 2048;
        if (lane == 0) tma_store_wait_read<0>();   // the previous tile store of this warp has read its staging tile
        __syncwarp();
        const uint32_t rowb = stg + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t a = rowb + ((q ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(ga[8 * q], ga[8 * q + 1])),
                       "r"(pack_bf16x2(ga[8 * q + 2], ga[8 * q + 3])), "r"(pack_bf16x2(ga[8 * q + 4], ga[8 * q + 5])),
                       "r"(pack_bf16x2(ga[8 * q + 6], ga[8 * q + 7])) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tmap_ga, stg, col0, cx.m0 + cx.quad * 32);
          tma_store_commit();
        }
      } else if (valid_row) {
        __nv_bfloat16* d = p.GA + static_cast<int64_t>(row) * p.ld_ga + col0;
        if (col0 + 32 <= cx.N && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16x2(ga[j], ga[j + 1]);
            u.y = pack_bf16x2(ga[j + 2], ga[j + 3]);
            u.z = pack_bf16x2(ga[j + 4], ga[j + 5]);
            u.w = pack_bf16x2(ga[j + 6], ga[j + 7]);
            *reinterpret_cast<uint4*>(d + j) = u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < cx.N) d[j] = __float2bfloat16_rn(ga[j]);
        }
        if (p.GA_lo != nullptr) {
          __nv_bfloat16* dl = p.GA_lo + static_cast<int64_t>(row) * p.ld_ga + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) ga[j] -= __bfloat162float(__float2bfloat16_rn(ga[j]));
          if (col0 + 32 <= cx.N && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16x2(ga[j], ga[j + 1]);
              u.y = pack_bf16x2(ga[j + 2], ga[j + 3]);
              u.z = pack_bf16x2(ga[j + 4], ga[j + 5]);
              u.w = pack_bf16x2(ga[j + 6], ga[j + 7]);
              *reinterpret_cast<uint4*>(dl + j) = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < cx.N) dl[j] = __float2bfloat16_rn(ga[j]);
          }
        }
      }
    }
  }
};

// ------------------------------------------------------------------ fused forward + backward tiles (small batch)
// The small-batch step is a chain of latency-bound launches; forward tiles and gradient-operand tiles are two of them and
// each runs the SAME k-loop over the same operands.  When the whole similarity matrix is at most 8 tiles of 128 x 64
// (B <= 256 on one GPU) one thread-block CLUSTER computes it: every CTA keeps its S tile in TMEM, runs the forward
// epilogue (row / column partial sums, positives, inverse norms, optional logits / hard-negative weight sums), the
// cluster meets at ONE barrier (release/acquire at cluster scope: the partial sums every CTA wrote to global memory are
// visible), and the backward epilogue derives lse_row / lse_col from those partials and emits the gradient operands from
// the accumulator that is still in TMEM.  One launch and one k-loop less on the critical chain of the step.
// Measurement switch (tic_debug_set_trace): when a device buffer is registered, the fused kernel stamps %globaltimer at its
// phase boundaries into trace[blockIdx.x * 16 + i] — where the 27 us of this 8-CTA kernel go cannot be seen from outside.
__device__ unsigned long long* g_tic_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int i) {
  unsigned long long* t = g_tic_trace;
  if (t != nullptr) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    t[blockIdx.x * 16 + i] = now;
  }
}
constexpr int kFusedSmallStages = 4;   // 4 x 24 KB ring + 3 KB scratch: the cluster's CTAs co-reside with other small launches
constexpr int kFusedSmallSmem = kFusedSmallStages * UmmaCfg<kItcBNSmall>::kStageBytes + 1024 + 256 + 2 * (6 * kItcBNSmall * 4 + 256);
template <int BN>
__global__ void __launch_bounds__(64 + 32 * kItcEpiWarps, 1)
itc_fused_small_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo,
                       const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b_lo, int split, int M,
                       int N, int K, const __grid_constant__ ItcFwdEpi::Params fp, const __grid_constant__ ItcBwdEpi::Params bp) {
  using Cfg = UmmaCfg<BN>;
  constexpr int STAGES = kFusedSmallStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
  auto full_bar = [&](int st) { return bar_base + 8u * st; };
  auto empty_bar = [&](int st) { return bar_base + 8u * (STAGES + st); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* scratch = smem_gen + STAGES * Cfg::kStageBytes + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (N + BN - 1) / BN;
  const int m_blk = blockIdx.x / n_tiles, n_blk = blockIdx.x % n_tiles;   // one tile per CTA, the grid is one cluster
  const int num_kb = (K + kBK - 1) / kBK;
  const int nseg = 1 + (split & 1) + ((split >> 1) & 1);
  const int total_kb = num_kb * nseg;
  if (threadIdx.x == 64) trace_stamp(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (split & 1) tma_prefetch_desc(&tmap_a_lo);
    if (split & 2) tma_prefetch_desc(&tmap_b_lo);
    for (int st = 0; st < STAGES; ++st) { mbar_init(full_bar(st), 1); mbar_init(empty_bar(st), 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  pdl_wait();
  if (threadIdx.x == 64) trace_stamp(1);
  const int m0 = m_blk * kBM, n0 = n_blk * BN;
  EpiCtx cx;
  if (warp == 0) {
    if (lane == 0) {     // ---- TMA producer (same segment order as umma_gemm_kernel: hi*hi, lo*hi, hi*lo)
      int st = 0;
      uint32_t ph = 0;
      for (int kt = 0; kt < total_kb; ++kt) {
        mbar_wait(empty_bar(st), ph ^ 1u);
        const int seg = kt / num_kb, kb = kt - seg * num_kb;
        const bool a_lo = (seg == 1) && (split & 1);
        const bool b_lo = (seg >= 1) && !a_lo;
        const uint32_t sa = smem_base + st * Cfg::kStageBytes;
        mbar_arrive_expect_tx(full_bar(st), Cfg::kStageBytes);
        tma_load_2d(sa, a_lo ? &tmap_a_lo : &tmap_a, full_bar(st), kb * kBK, m0);
        tma_load_2d(sa + kABytes, b_lo ? &tmap_b_lo : &tmap_b, full_bar(st), kb * kBK, n0);
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
      trace_stamp(2);
      pdl_trigger();
    }
  } else if (warp == 1) {
    if (lane == 0) {     // ---- MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, false, false);
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + st * Cfg::kStageBytes;
        const uint64_t da = umma_smem_desc_sw128(sa, 16u, 1024u);
        const uint64_t db = umma_smem_desc_sw128(sa + kABytes, 16u, 1024u);
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k)
          umma_bf16(tmem_base, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(st));
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
      umma_commit(tfull_bar);
    }
  } else {               // ---- epilogue, phase 1: forward statistics
    cx.M = M; cx.N = N;
    cx.quad = warp & 3;
    cx.part = (warp - 2) >> 2;
    cx.nparts = kItcEpiWarps / 4;
    cx.epi_tid = threadIdx.x - 64;
    cx.epi_threads = 32 * kItcEpiWarps;
    cx.scratch = scratch;
    cx.iter = 0;
    cx.ks = 0; cx.ksplit = 1;
    cx.m_blk = m_blk; cx.n_blk = n_blk; cx.m0 = m0; cx.n0 = n0;
    cx.tmem_acc = tmem_base;
    ItcFwdEpi::template prefetch<BN>(fp, cx);
    if (threadIdx.x == 64) trace_stamp(8);
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(3);
    ItcFwdEpi::template tile<BN>(fp, cx);
    tc_fence_before();
    if (threadIdx.x == 64) trace_stamp(4);
  }
  __syncwarp();
  cluster_sync_all();    // every CTA's partial sums / inverse norms are in global memory and visible cluster-wide
  if (warp >= 2) {       // ---- epilogue, phase 2: gradient operands from the accumulator still in TMEM
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(5);
    cx.iter = 1;         // the other half of the double-buffered epilogue scratch
    cx.stage = smem_gen; // the operand ring is idle (every MMA has completed): 2 KB of staging per epilogue warp
    ItcBwdEpi::template prefetch<BN, true>(bp, cx);
    if (threadIdx.x == 64) trace_stamp(6);
    ItcBwdEpi::template tile<BN>(bp, cx);
    tma_store_wait<0>();
    if (threadIdx.x == 64) trace_stamp(7);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ small HBM-bound kernels
// rinv[i] = 1/||X[i,:]||  — one warp per row, 16-byte loads when aligned.
__global__ void row_rnorm_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ X_lo, int64_t ldx, int rows,
                                 int cols, float* __restrict__ rinv, __nv_bfloat16* __restrict__ Xhat, int64_t ldh) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const __nv_bfloat16* x = X + static_cast<int64_t>(warp) * ldx;
  float ss = 0.f;
  if (X_lo != nullptr) {
    const __nv_bfloat16* xl = X_lo + static_cast<int64_t>(warp) * ldx;
    for (int c = lane; c < cols; c += 32) {
      const float f = __bfloat162float(x[c]) + __bfloat162float(xl[c]);
      ss = fmaf(f, f, ss);
    }
  } else if ((ldx & 7) == 0 && (cols & 7) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    for (int c = lane * 8; c < cols; c += 256) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + c));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      const float f = __bfloat162float(x[c]);
      ss = fmaf(f, f, ss);
    }
  }
  ss = warp_sum(ss);
  const float ri = 1.0f / sqrtf(ss);  // no epsilon (HF :268-269)
  if (lane == 0) rinv[warp] = ri;
  if (Xhat != nullptr) {               // normalised bf16 copy: B operand of the image-side gradient GEMM at large batch
    __nv_bfloat16* h = Xhat + static_cast<int64_t>(warp) * ldh;
    for (int c = lane; c < cols; c += 32) h[c] = __float2bfloat16_rn(ri * __bfloat162float(x[c]));
  }
}

__global__ void reduce_parts_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ out) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[static_cast<int64_t>(p) * n + j];
  out[j] = s;
}

// Multi-block: fixed-order sum of the per-tile partials -> lse vectors (thread t owns row t and column t).
__global__ void itc_lse_kernel(const float* __restrict__ row_part, int nrp, const float* __restrict__ col_part, int ncp, int M,
                               int N, float shift, const float* __restrict__ scale_dev, float* __restrict__ lse_row,
                               float* __restrict__ lse_col) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  if (scale_dev != nullptr) shift = __ldg(scale_dev);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N) {
    float s = 0.f;
    for (int p = 0; p < ncp; ++p) s += col_part[static_cast<int64_t>(p) * N + t];
    lse_col[t] = shift + logf(s);
  }
  if (t < M) {
    float s = 0.f;
    for (int p = 0; p < nrp; ++p) s += row_part[static_cast<int64_t>(p) * M + t];
    lse_row[t] = shift + logf(s);
  }
}
// Single block, fixed-order tree: the two loss sums (deterministic).
__global__ void itc_loss_kernel(const float* __restrict__ lse_row, const float* __restrict__ lse_col,
                                const float* __restrict__ diag, int M, int row_offset, float* __restrict__ loss_sums) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  __shared__ float sr[32], sc[32];
  float ar = 0.f, ac = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float d = diag[i];
    ar += lse_row[i] - d;
    ac += lse_col[row_offset + i] - d;
  }
  ar = warp_sum(ar);
  ac = warp_sum(ac);
  if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = ar; sc[threadIdx.x >> 5] = ac; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    ar = threadIdx.x < nw ? sr[threadIdx.x] : 0.f;
    ac = threadIdx.x < nw ? sc[threadIdx.x] : 0.f;
    ar = warp_sum(ar);
    ac = warp_sum(ac);
    if (threadIdx.x == 0) { loss_sums[0] += ar; loss_sums[1] += ac; }
  }
}

// Symmetric (peer-memory) mode: both softmax directions are ROW statistics — of the row block S[rows_r, :] and of the
// swapped block S^T[cols_r, :].  blockIdx.y selects the direction; lse = shift + log(sum of the per-tile partials);
// the loss terms sum_i (lse_i - diag_i) are reduced in a fixed order (block partials, last block adds them up).
// Optional PUSH of the two lse vectors (multi-GPU symmetric form): every value is also stored into the remote peers' gathered
// vectors (slot of this rank), and the last block of the launch releases the peers' flags = step[1] — the exchange costs no
// kernel of its own (it was a 6-8 us launch between the forward and the gradient-operand tiles).
constexpr int kLseRowsPerBlock = 32;   // itc_lse_rows_kernel: 256 threads = 32 rows x 8 lanes
struct LsePush {
  uint8_t* base[8];       // every rank's peer-mapped block (world == 0: no push)
  int world, rank;
  int64_t off_a, off_b;   // byte offsets of the gathered vectors (element 0 of rank 0's slot) inside every block
  int64_t flag_off;       // uint32[world] flag words inside every block
  const uint32_t* step;   // {completed steps, epoch in flight} (tic_peer_signal)
};
__global__ void itc_lse_rows_kernel(const float* __restrict__ part_a, const float* __restrict__ part_b, int nparts, int m,
                                    const float* __restrict__ diag, float shift, const float* __restrict__ scale_dev,
                                    float* __restrict__ lse_a, float* __restrict__ lse_b, float* __restrict__ loss_sums,
                                    float* __restrict__ ws, LsePush px) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  if (scale_dev != nullptr) shift = __ldg(scale_dev);
  const int dir = blockIdx.y;
  const float* part = dir == 0 ? part_a : part_b;
  float* lse = dir == 0 ? lse_a : lse_b;
  float* blk_part = ws + 2 + dir * gridDim.x;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws) + dir;
  // 8 lanes per row (32 rows per block): the partials of a row are loaded by 8 threads at once — one thread per row walked
  // up to 64 partials through a serial add chain, ~12 us at 4 GPUs for 2 x 256 rows on the step's critical chain
  // (profiles/r02_timeline_c2_g4.txt)
  const int sub = threadIdx.x & 7;
  const int i = blockIdx.x * kLseRowsPerBlock + (threadIdx.x >> 3);
  float term = 0.f;
  {
    float s = 0.f;
    if (i < m)
      for (int p = sub; p < nparts; p += 8) s += part[static_cast<int64_t>(p) * m + i];
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (i < m) {
      const float l = shift + logf(s);
      if (sub == 0) {
        lse[i] = l;
        term = l - diag[i];
      } else if (sub < px.world) {          // lane k of the row's group: the remote copy in peer rank + k
        int q = px.rank + sub;
        if (q >= px.world) q -= px.world;
        reinterpret_cast<float*>(px.base[q] + (dir == 0 ? px.off_a : px.off_b))[static_cast<int64_t>(px.rank) * m + i] = l;
      }
    }
  }
  __shared__ float sw[32];
  __shared__ bool last, pub;
  if (threadIdx.x == 0) pub = false;
  term = warp_sum(term);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sw[w];
    blk_part[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (last) {
      __threadfence();
      float tot = 0.f;
      for (int b = 0; b < static_cast<int>(gridDim.x); ++b) tot += reinterpret_cast<volatile float*>(blk_part)[b];
      loss_sums[dir] += tot;
      *ticket = 0u;
    }
    if (px.world > 1) {      // the last block of BOTH directions publishes the flags (release at system scope, cumulative)
      unsigned int* all = reinterpret_cast<unsigned int*>(ws) + 2 + 2 * gridDim.x;
      __threadfence();
      pub = atomicAdd(all, 1u) == 2 * gridDim.x - 1;
      if (pub) {
        *all = 0u;
        __threadfence();
      }
    }
  }
  if (px.world > 1) {
    __syncthreads();
    if (pub && threadIdx.x >= 1 && static_cast<int>(threadIdx.x) < px.world) {   // one releasing thread per peer (csrc/peer.cu)
      const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(px.step + 1);
      int q = px.rank + threadIdx.x;
      if (q >= px.world) q -= px.world;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(px.base[q] + px.flag_off) + px.rank), "r"(epoch) : "memory");
    }
  }
}

// One warp per row: diagonal term, normalise-backward, dlogit_scale contribution.
__global__ void itc_grad_finalize_kernel(const float* __restrict__ acc, int64_t ld_acc, const __nv_bfloat16* __restrict__ X,
                                         const __nv_bfloat16* __restrict__ X_lo, int64_t ldx, const float* __restrict__ rinv,
                                         const __nv_bfloat16* __restrict__ Xo, const __nv_bfloat16* __restrict__ Xo_lo,
                                         int64_t ldxo, const float* __restrict__ rinv_o, int rows, int P, float scale,
                                         float diag_coef, float* __restrict__ dXf, int64_t ld_df,
                                         __nv_bfloat16* __restrict__ dXb, __nv_bfloat16* __restrict__ dXb_lo, int64_t ld_db,
                                         float* __restrict__ r_sum, int acc_div_rinv, const float* __restrict__ scale_dev) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  const int warp_in_blk = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp_in_blk;
  __shared__ float sblk[32];
  float r = 0.f;
  if (row < rows) {
    const float ri = rinv[row];
    const float co = (Xo != nullptr && diag_coef != 0.f) ? diag_coef * rinv_o[row] : 0.f;
    const float asc = acc_div_rinv ? 1.0f / ri : 1.0f;   // accumulator carries an extra factor rinv[row] (GA-shared mode)
    const float* a = acc + static_cast<int64_t>(row) * ld_acc;
    const __nv_bfloat16* x = X + static_cast<int64_t>(row) * ldx;
    const __nv_bfloat16* xo = Xo ? Xo + static_cast<int64_t>(row) * ldxo : nullptr;
    const __nv_bfloat16* xl = X_lo ? X_lo + static_cast<int64_t>(row) * ldx : nullptr;
    const __nv_bfloat16* xol = (Xo && Xo_lo) ? Xo_lo + static_cast<int64_t>(row) * ldxo : nullptr;
    auto xval = [&](int k) { return __bfloat162float(x[k]) + (xl ? __bfloat162float(xl[k]) : 0.f); };
    auto xoval = [&](int k) { return __bfloat162float(xo[k]) + (xol ? __bfloat162float(xol[k]) : 0.f); };
    for (int k = lane; k < P; k += 32) {
      const float dxh = scale * (a[k] * asc - (xo ? co * xoval(k) : 0.f));
      r = fmaf(ri * xval(k), dxh, r);
    }
    r = warp_sum(r);
    for (int k = lane; k < P; k += 32) {
      const float dxh = scale * (a[k] * asc - (xo ? co * xoval(k) : 0.f));
      const float xh = ri * xval(k);
      const float g = ri * (dxh - xh * r);
      if (dXf) dXf[static_cast<int64_t>(row) * ld_df + k] = g;
      if (dXb) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(g);
        dXb[static_cast<int64_t>(row) * ld_db + k] = hi;
        if (dXb_lo) dXb_lo[static_cast<int64_t>(row) * ld_db + k] = __float2bfloat16_rn(g - __bfloat162float(hi));
      }
    }
  }
  if (r_sum != nullptr) {
    if (lane == 0) sblk[warp_in_blk] = (row < rows) ? r : 0.f;
    __syncthreads();
    if (warp_in_blk == 0) {
      float t = lane < (blockDim.x >> 5) ? sblk[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) atomicAdd(r_sum, t);
    }
  }
}

// Vector form of the kernel above (P % 8 == 0, P <= 1024, 16-byte aligned rows): every lane owns 8 consecutive elements
// of each 256-wide chunk, everything is read once with 16-byte loads and stays in registers between the two passes.
__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ void add8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  float t[8];
  ld8_bf16(p, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] += t[k];
}
__global__ void __launch_bounds__(256)
itc_grad_finalize_vec_kernel(const float* __restrict__ acc, int64_t ld_acc, const __nv_bfloat16* __restrict__ X,
                             const __nv_bfloat16* __restrict__ X_lo, int64_t ldx, const float* __restrict__ rinv,
                             const __nv_bfloat16* __restrict__ Xo, const __nv_bfloat16* __restrict__ Xo_lo, int64_t ldxo,
                             const float* __restrict__ rinv_o, int rows, int P, float scale, float diag_coef,
                             float* __restrict__ dXf, int64_t ld_df, __nv_bfloat16* __restrict__ dXb,
                             __nv_bfloat16* __restrict__ dXb_lo, int64_t ld_db, float* __restrict__ r_sum, int acc_div_rinv,
                             const float* __restrict__ scale_dev) {
  pdl_trigger();
  pdl_wait();
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  constexpr int MAXC = 4;
  const int warp_in_blk = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp_in_blk;
  __shared__ float sblk[32];
  float r = 0.f;
  if (row < rows) {
    const float ri = rinv[row];
    const bool has_o = Xo != nullptr && diag_coef != 0.f;
    const float co = has_o ? diag_coef * rinv_o[row] : 0.f;
    const float asc = acc_div_rinv ? 1.0f / ri : 1.0f;
    float xh[MAXC][8], dxh[MAXC][8];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = c * 256 + lane * 8;
      if (k < P) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(acc + static_cast<int64_t>(row) * ld_acc + k));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(acc + static_cast<int64_t>(row) * ld_acc + k + 4));
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        ld8_bf16(X + static_cast<int64_t>(row) * ldx + k, xh[c]);
        if (X_lo) add8_bf16(X_lo + static_cast<int64_t>(row) * ldx + k, xh[c]);
        float xo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (has_o) {
          ld8_bf16(Xo + static_cast<int64_t>(row) * ldxo + k, xo);
          if (Xo_lo) add8_bf16(Xo_lo + static_cast<int64_t>(row) * ldxo + k, xo);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dxh[c][j] = scale * (a[j] * asc - co * xo[j]);
          xh[c][j] *= ri;
          r = fmaf(xh[c][j], dxh[c][j], r);
        }
      }
    }
    r = warp_sum(r);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = c * 256 + lane * 8;
      if (k < P) {
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = ri * (dxh[c][j] - xh[c][j] * r);
        if (dXf) {
          float* d = dXf + static_cast<int64_t>(row) * ld_df + k;
          *reinterpret_cast<float4*>(d) = make_float4(g[0], g[1], g[2], g[3]);
          *reinterpret_cast<float4*>(d + 4) = make_float4(g[4], g[5], g[6], g[7]);
        }
        if (dXb) {
          uint4 u;
          u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]); u.z = pack_bf16x2(g[4], g[5]); u.w = pack_bf16x2(g[6], g[7]);
          *reinterpret_cast<uint4*>(dXb + static_cast<int64_t>(row) * ld_db + k) = u;
          if (dXb_lo) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] -= __bfloat162float(__float2bfloat16_rn(g[j]));
            u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]); u.z = pack_bf16x2(g[4], g[5]); u.w = pack_bf16x2(g[6], g[7]);
            *reinterpret_cast<uint4*>(dXb_lo + static_cast<int64_t>(row) * ld_db + k) = u;
          }
        }
      }
    }
  }
  if (r_sum != nullptr) {
    if (lane == 0) sblk[warp_in_blk] = (row < rows) ? r : 0.f;
    __syncthreads();
    if (warp_in_blk == 0) {
      float t = lane < (blockDim.x >> 5) ? sblk[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) atomicAdd(r_sum, t);
    }
  }
}

// Gradient operands from a MATERIALISED dL/dS (autograd path at drop-in batch sizes):
//   GA[i,j] = dS[i,j] * rinv_v[j],  GB[i,j] = dS[i,j] * rinv_t[i]   (+ bf16 residuals), both row-major like dS.
__global__ void itc_ds_operands_kernel(const float* __restrict__ dS, int64_t ldds, int M, int N, const float* __restrict__ rinv_t,
                                       const float* __restrict__ rinv_v, __nv_bfloat16* __restrict__ GA,
                                       __nv_bfloat16* __restrict__ GA_lo, int64_t ld_ga, __nv_bfloat16* __restrict__ GB,
                                       __nv_bfloat16* __restrict__ GB_lo) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M || j >= N) return;
  const float v = dS[static_cast<int64_t>(i) * ldds + j];
  const float ga = v * rinv_v[j];
  const __nv_bfloat16 hi = __float2bfloat16_rn(ga);
  GA[static_cast<int64_t>(i) * ld_ga + j] = hi;
  if (GA_lo) GA_lo[static_cast<int64_t>(i) * ld_ga + j] = __float2bfloat16_rn(ga - __bfloat162float(hi));
  if (GB == nullptr) return;
  const float gb = v * rinv_t[i];
  const __nv_bfloat16 hb = __float2bfloat16_rn(gb);
  GB[static_cast<int64_t>(i) * ld_ga + j] = hb;
  if (GB_lo) GB_lo[static_cast<int64_t>(i) * ld_ga + j] = __float2bfloat16_rn(gb - __bfloat162float(hb));
}

}  // namespace tic

using namespace tic;

extern "C" {

int tic_itc_q_parts(int n_global) { return ceil_div(n_global, 32); }
int tic_itc_row_parts(int n_global) { return ceil_div(n_global, itc_bn(n_global)) * (kItcEpiWarps / 4); }
int tic_itc_col_parts(int m_local) { return ceil_div(m_local, kBM); }

int tic_row_rnorm_bf16(const void* X, const void* X_lo, int64_t ldx, int rows, int cols, float* rinv, void* Xhat, int64_t ldh,
                       void* stream) {
  TIC_CHECK_ARG(X && rinv && rows > 0 && cols > 0, "tic_row_rnorm_bf16: bad arguments");
  const int wpb = 8;
  launch_k(row_rnorm_kernel, dim3(ceil_div(rows, wpb)), dim3(wpb * 32), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(X_lo), ldx, rows, cols, rinv,
      static_cast<__nv_bfloat16*>(Xhat), ldh);
  TIC_CHECK_LAUNCH("tic_row_rnorm_bf16");
  return TIC_OK;
}

int tic_itc_fwd(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv, float* rinv_t,
                float* rinv_v,
                int m_local, int n_global, int P, int row_offset, float scale, float shift, float* row_part,
                float* col_part, float* diag, float* logits_out, int64_t ld_logits, const float* ss_t_part, int n_ss_t,
                const float* ss_v_part, int n_ss_v, const uint32_t* seg_ready, const uint32_t* seg_epoch, int seg_cols,
                int my_seg, const float* scale_dev, void* qpart, void* stream) {
  TIC_CHECK_ARG(T && V && rinv_t && rinv_v && row_part && diag, "tic_itc_fwd: null pointer");
  TIC_CHECK_ARG(m_local > 0 && n_global > 0 && P > 0, "tic_itc_fwd: empty problem");
  TIC_CHECK_ARG(row_offset >= 0 && row_offset + m_local <= n_global, "tic_itc_fwd: row block outside the global batch");
  if (scale_dev == nullptr && (!(scale > 0.f) || scale > 40.f || shift < scale)) {
    set_error("tic_itc_fwd: scale=%g shift=%g outside the supported range (0 < scale <= 40, shift >= scale)", scale, shift);
    return TIC_E_RANGE;
  }
  TIC_CHECK_ARG((!ss_t_part || n_ss_t > 0) && (!ss_v_part || n_ss_v > 0), "tic_itc_fwd: empty sum-of-squares partial list");
  ItcFwdEpi::Params ep{rinv_t, rinv_v, ss_t_part, ss_v_part, n_ss_t, n_ss_v, scale * kLog2e, shift * kLog2e, scale, shift,
                       row_part, col_part, diag, logits_out, ld_logits, row_offset, scale_dev,
                       static_cast<unsigned long long*>(qpart)};
  const bool mc = itc_multicast() && itc_bn(n_global) == kItcBN && !T_lo && !V_lo && m_local >= 2 * kBM;
  // V arriving segment by segment (peer pull running beside this kernel): tiles go segment-major, local segment first
  SegOrder so{nullptr, nullptr, 0, 0, 0, 0};
  const SegOrder* sop = nullptr;
  if (seg_ready != nullptr) {
    const int remote = seg_cols < 0 ? 1 : 0;        // negative seg_cols: the ready words are written by the peers (push form)
    if (remote) seg_cols = -seg_cols;
    TIC_CHECK_ARG(seg_epoch && seg_cols > 0 && n_global % seg_cols == 0 && my_seg >= 0 && my_seg < n_global / seg_cols,
                  "tic_itc_fwd: bad segment description (seg_cols=%d my_seg=%d n_global=%d)", seg_cols, my_seg, n_global);
    const int bn = itc_bn(n_global);
    so = SegOrder{seg_ready, seg_epoch, seg_cols % bn == 0 ? seg_cols / bn : 0, my_seg, n_global / seg_cols, remote};
    sop = &so;
  }
  int rc = mc ? launch_umma_gemm_cluster2<kItcBN, false, false, kItcEpiWarps, ItcFwdEpi>(T, ldt, V, ldv, m_local, n_global, P, ep,
                                                                                        static_cast<cudaStream_t>(stream), sop)
           : itc_bn(n_global) == kItcBN
               ? launch_umma_gemm<kItcBN, false, false, kItcEpiWarps, ItcFwdEpi>(T, T_lo, ldt, V, V_lo, ldv, m_local, n_global,
                                                                                P, ep, static_cast<cudaStream_t>(stream), 1, 0, sop)
               : (qpart == nullptr && itc_small_kc(m_local, n_global, P, T_lo, V_lo, sop) == 4)   // (pick tiles recompute without KC)
                   ? launch_umma_gemm_kc<kItcBNSmall, false, false, kItcEpiWarps, ItcFwdEpi, 4>(
                         T, T_lo, ldt, V, V_lo, ldv, m_local, n_global, P, ep, static_cast<cudaStream_t>(stream))
               : (qpart == nullptr && itc_small_kc(m_local, n_global, P, T_lo, V_lo, sop) == 2)
                   ? launch_umma_gemm_kc<kItcBNSmall, false, false, kItcEpiWarps, ItcFwdEpi, 2>(
                         T, T_lo, ldt, V, V_lo, ldv, m_local, n_global, P, ep, static_cast<cudaStream_t>(stream))
                   : launch_umma_gemm<kItcBNSmall, false, false, kItcEpiWarps, ItcFwdEpi>(T, T_lo, ldt, V, V_lo, ldv, m_local,
                                                                                         n_global, P, ep,
                                                                                         static_cast<cudaStream_t>(stream), 1, 0, sop);
  if (rc == -3) { set_error("tic_itc_fwd: cudaFuncSetAttribute failed"); return TIC_E_ATTR; }
  if (rc == -4) { set_error("tic_itc_fwd: launch failed"); return TIC_E_LAUNCH; }
  return rc;
}

int tic_itc_pick(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv, const float* rinv_t,
                 const float* rinv_v, int m_local, int n_global, int P, int row_offset, float scale, float shift,
                 const float* scale_dev, const int32_t* loc_part, const void* loc_res, int32_t* src_idx, void* stream) {
  TIC_CHECK_ARG(T && V && rinv_t && rinv_v && loc_part && loc_res && src_idx, "tic_itc_pick: null pointer");
  TIC_CHECK_ARG(m_local > 0 && n_global > 0 && P > 0, "tic_itc_pick: empty problem");
  ItcPickEpi::Params ep{rinv_t, rinv_v, scale, shift, scale_dev, row_offset, loc_part,
                        static_cast<const unsigned long long*>(loc_res), src_idx};
  // the SAME tile shapes as tic_itc_fwd picks for this problem: the recomputed accumulators must be bit-identical
  const bool mc = itc_multicast() && itc_bn(n_global) == kItcBN && !T_lo && !V_lo && m_local >= 2 * kBM;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = mc ? launch_umma_gemm_cluster2<kItcBN, false, false, kItcEpiWarps, ItcPickEpi>(T, ldt, V, ldv, m_local, n_global, P, ep, st)
           : itc_bn(n_global) == kItcBN
               ? launch_umma_gemm<kItcBN, false, false, kItcEpiWarps, ItcPickEpi>(T, T_lo, ldt, V, V_lo, ldv, m_local, n_global, P, ep, st, 1)
               : launch_umma_gemm<kItcBNSmall, false, false, kItcEpiWarps, ItcPickEpi>(T, T_lo, ldt, V, V_lo, ldv, m_local, n_global, P, ep,
                                                                                     st, 1);
  if (rc == -3) { set_error("tic_itc_pick: cudaFuncSetAttribute failed"); return TIC_E_ATTR; }
  if (rc == -4) { set_error("tic_itc_pick: launch failed"); return TIC_E_LAUNCH; }
  return rc;
}

int tic_reduce_parts(const float* part, int nparts, int n, float* out, void* stream) {
  TIC_CHECK_ARG(part && out && nparts > 0 && n > 0, "tic_reduce_parts: bad arguments");
  launch_k(reduce_parts_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), part, nparts, n, out);
  TIC_CHECK_LAUNCH("tic_reduce_parts");
  return TIC_OK;
}

int tic_itc_lse_loss(const float* row_part, int n_row_parts, const float* col_part, int n_col_parts, const float* diag,
                     int m_local, int n_global, int row_offset, float shift, float* lse_row, float* lse_col,
                     float* loss_sums, const float* scale_dev, void* stream) {
  TIC_CHECK_ARG(row_part && col_part && diag && lse_row && lse_col && loss_sums && n_row_parts > 0 && n_col_parts > 0,
                "tic_itc_lse_loss: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int mx = m_local > n_global ? m_local : n_global;
  launch_k(itc_lse_kernel, dim3(ceil_div(mx, 256)), dim3(256), 0, st, row_part, n_row_parts, col_part, n_col_parts, m_local, n_global, shift,
           scale_dev, lse_row, lse_col);
  launch_k(itc_loss_kernel, dim3(1), dim3(1024), 0, st, lse_row, lse_col, diag, m_local, row_offset, loss_sums);
  TIC_CHECK_LAUNCH("tic_itc_lse_loss");
  return TIC_OK;
}

int64_t tic_itc_lse_rows_workspace_bytes(int m) { return static_cast<int64_t>(3 + 2 * ceil_div(m, kLseRowsPerBlock)) * 4; }

int tic_itc_lse_rows(const float* part_a, const float* part_b, int n_parts, int m, const float* diag, float shift, float* lse_a,
                     float* lse_b, float* loss_sums, void* workspace, const float* scale_dev, void* stream) {
  TIC_CHECK_ARG(part_a && part_b && diag && lse_a && lse_b && loss_sums && workspace && n_parts > 0 && m > 0,
                "tic_itc_lse_rows: bad arguments");
  dim3 grid(ceil_div(m, kLseRowsPerBlock), 2);
  launch_k(itc_lse_rows_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), part_a, part_b, n_parts, m, diag, shift,
           scale_dev, lse_a, lse_b, loss_sums, static_cast<float*>(workspace), LsePush{});
  TIC_CHECK_LAUNCH("tic_itc_lse_rows");
  return TIC_OK;
}

int tic_itc_lse_rows_push(const float* part_a, const float* part_b, int n_parts, int m, const float* diag, float shift, float* lse_a,
                          float* lse_b, float* loss_sums, void* workspace, const float* scale_dev, void* const* bases_host, int world,
                          int rank, int64_t off_a, int64_t off_b, int64_t flag_off, const uint32_t* step, void* stream) {
  TIC_CHECK_ARG(part_a && part_b && diag && lse_a && lse_b && loss_sums && workspace && n_parts > 0 && m > 0,
                "tic_itc_lse_rows_push: bad arguments");
  TIC_CHECK_ARG(bases_host && step && world >= 1 && world <= 8 && rank >= 0 && rank < world && (off_a & 3) == 0 && (off_b & 3) == 0 &&
                    (flag_off & 3) == 0,
                "tic_itc_lse_rows_push: bad peer description");
  LsePush px{};
  for (int p = 0; p < world; ++p) {
    TIC_CHECK_ARG(bases_host[p] != nullptr, "tic_itc_lse_rows_push: rank %d has no mapped block", p);
    px.base[p] = static_cast<uint8_t*>(bases_host[p]);
  }
  px.world = world; px.rank = rank; px.off_a = off_a; px.off_b = off_b; px.flag_off = flag_off; px.step = step;
  dim3 grid(ceil_div(m, kLseRowsPerBlock), 2);
  launch_k(itc_lse_rows_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), part_a, part_b, n_parts, m, diag, shift,
           scale_dev, lse_a, lse_b, loss_sums, static_cast<float*>(workspace), px);
  TIC_CHECK_LAUNCH("tic_itc_lse_rows_push");
  return TIC_OK;
}

int tic_itc_bwd_g(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv,
                  const float* rinv_t, const float* rinv_v,
                  const float* lse_row, const float* lse_col, int m_local, int n_global, int P, float scale, float gscale,
                  void* GA, int64_t ld_ga, void* GBT, int64_t ld_gbt, void* GA_lo, void* GBT_lo, const float* row_part,
                  int n_row_parts, const float* col_part, int n_col_parts, float shift, const float* scale_dev,
                  const uint32_t* seg_ready, const uint32_t* seg_epoch, int seg_cols, int my_seg, void* stream) {
  TIC_CHECK_ARG(T && V && rinv_t && rinv_v && GA, "tic_itc_bwd_g: null pointer");
  SegOrder so{nullptr, nullptr, 0, 0, 0, 0};
  const SegOrder* sop = nullptr;
  if (seg_ready != nullptr) {       // gathered lse vectors landing segment by segment (push form: always remote)
    const int cols = seg_cols < 0 ? -seg_cols : seg_cols;
    TIC_CHECK_ARG(seg_epoch && cols > 0 && n_global % cols == 0 && my_seg >= 0 && my_seg < n_global / cols,
                  "tic_itc_bwd_g: bad segment description (seg_cols=%d my_seg=%d n_global=%d)", seg_cols, my_seg, n_global);
    const int bn = itc_bn(n_global);
    so = SegOrder{seg_ready, seg_epoch, cols % bn == 0 ? cols / bn : 0, my_seg, n_global / cols, 1};
    sop = &so;
  }
  TIC_CHECK_ARG((lse_row || (row_part && n_row_parts > 0)) && (lse_col || (col_part && n_col_parts > 0)),
                "tic_itc_bwd_g: need lse_row/lse_col or the forward partials");
  TIC_CHECK_ARG(m_local > 0 && n_global > 0 && P > 0, "tic_itc_bwd_g: empty problem");
  ItcBwdEpi::Params ep{rinv_t, rinv_v, lse_row, lse_col, scale * kLog2e, gscale, static_cast<__nv_bfloat16*>(GA), ld_ga,
                       static_cast<__nv_bfloat16*>(GBT), ld_gbt, static_cast<__nv_bfloat16*>(GA_lo),
                       static_cast<__nv_bfloat16*>(GBT_lo), row_part, col_part, n_row_parts, n_col_parts, shift,
                       scale * kLog2e, ((scale_dev != nullptr || scale <= 20.f) && GBT == nullptr) ? 1 : 0, 0, scale_dev, {}};
  if (GA_lo == nullptr && GBT == nullptr && (ld_ga & 7) == 0 && aligned16(GA) && itc_tma_store()) {
    int trc = make_tmap_bf16_store32(&ep.tmap_ga, GA, static_cast<uint64_t>(n_global), static_cast<uint64_t>(m_local),
                                     static_cast<uint64_t>(ld_ga));
    if (trc) return trc;
    ep.tma_store = 1;
  }
  const bool mc = itc_multicast() && itc_bn(n_global) == kItcBN && !T_lo && !V_lo && m_local >= 2 * kBM;
  int rc = mc ? launch_umma_gemm_cluster2<kItcBN, false, false, kItcEpiWarps, ItcBwdEpi>(T, ldt, V, ldv, m_local, n_global, P, ep,
                                                                                        static_cast<cudaStream_t>(stream), sop)
           : itc_bn(n_global) == kItcBN
               ? launch_umma_gemm<kItcBN, false, false, kItcEpiWarps, ItcBwdEpi>(T, T_lo, ldt, V, V_lo, ldv, m_local, n_global,
                                                                                P, ep, static_cast<cudaStream_t>(stream), 1, 0, sop)
               : itc_small_kc(m_local, n_global, P, T_lo, V_lo, sop) == 4
                   ? launch_umma_gemm_kc<kItcBNSmall, false, false, kItcEpiWarps, ItcBwdEpi, 4>(
                         T, T_lo, ldt, V, V_lo, ldv, m_local, n_global, P, ep, static_cast<cudaStream_t>(stream))
               : itc_small_kc(m_local, n_global, P, T_lo, V_lo, sop) == 2
                   ? launch_umma_gemm_kc<kItcBNSmall, false, false, kItcEpiWarps, ItcBwdEpi, 2>(
                         T, T_lo, ldt, V, V_lo, ldv, m_local, n_global, P, ep, static_cast<cudaStream_t>(stream))
                   : launch_umma_gemm<kItcBNSmall, false, false, kItcEpiWarps, ItcBwdEpi>(T, T_lo, ldt, V, V_lo, ldv, m_local,
                                                                                         n_global, P, ep,
                                                                                         static_cast<cudaStream_t>(stream), 1, 0, sop);
  if (rc == -3) { set_error("tic_itc_bwd_g: cudaFuncSetAttribute failed"); return TIC_E_ATTR; }
  if (rc == -4) { set_error("tic_itc_bwd_g: launch failed"); return TIC_E_LAUNCH; }
  return rc;
}

int tic_debug_set_trace(void* device_u64_buffer) {
  unsigned long long* p = static_cast<unsigned long long*>(device_u64_buffer);
  if (cudaMemcpyToSymbol(g_tic_trace, &p, sizeof(p)) != cudaSuccess) { set_error("tic_debug_set_trace: cudaMemcpyToSymbol failed"); return TIC_E_CUDA; }
  return TIC_OK;
}

int tic_itc_fused_small_ok(int m_local, int n_global) {
  return itc_bn(n_global) == kItcBNSmall && ceil_div(m_local, kBM) * ceil_div(n_global, kItcBNSmall) <= 8 ? 1 : 0;
}

int tic_itc_fwd_bwd_small(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv, float* rinv_t,
                          float* rinv_v, int m_local, int n_global, int P, int row_offset, float scale, float* row_part,
                          float* col_part, float* diag, float* logits_out, int64_t ld_logits, const float* ss_t_part, int n_ss_t,
                          const float* ss_v_part, int n_ss_v, const float* scale_dev, void* qpart, float gscale, void* GA,
                          int64_t ld_ga, void* GBT, int64_t ld_gbt, void* GA_lo, void* GBT_lo, void* stream) {
  TIC_CHECK_ARG(T && V && rinv_t && rinv_v && row_part && col_part && diag && GA, "tic_itc_fwd_bwd_small: null pointer");
  TIC_CHECK_ARG(m_local > 0 && n_global > 0 && P > 0 && row_offset >= 0 && row_offset + m_local <= n_global,
                "tic_itc_fwd_bwd_small: bad problem");
  TIC_CHECK_ARG(tic_itc_fused_small_ok(m_local, n_global), "tic_itc_fwd_bwd_small: more than 8 tiles of 128x64 (m=%d n=%d)", m_local,
                n_global);
  if (scale_dev == nullptr && (!(scale > 0.f) || scale > 40.f)) {
    set_error("tic_itc_fwd_bwd_small: scale=%g outside the supported range (0 < scale <= 40)", scale);
    return TIC_E_RANGE;
  }
  TIC_CHECK_ARG((!ss_t_part || n_ss_t > 0) && (!ss_v_part || n_ss_v > 0), "tic_itc_fwd_bwd_small: empty sum-of-squares partial list");
  constexpr int BN = kItcBNSmall;
  using Cfg = UmmaCfg<BN>;
  const int m_tiles = ceil_div(m_local, kBM), n_tiles = ceil_div(n_global, BN);
  ItcFwdEpi::Params fp{rinv_t, rinv_v, ss_t_part, ss_v_part, n_ss_t, n_ss_v, scale * kLog2e, scale * kLog2e, scale, scale,
                       row_part, col_part, diag, logits_out, ld_logits, row_offset, scale_dev,
                       static_cast<unsigned long long*>(qpart)};
  // backward statistics straight from the forward partials of this launch (inline lse)
  ItcBwdEpi::Params bp{rinv_t, rinv_v, nullptr, nullptr, scale * kLog2e, gscale, static_cast<__nv_bfloat16*>(GA), ld_ga,
                       static_cast<__nv_bfloat16*>(GBT), ld_gbt, static_cast<__nv_bfloat16*>(GA_lo),
                       static_cast<__nv_bfloat16*>(GBT_lo), row_part, col_part, n_tiles * (kItcEpiWarps / 4), m_tiles, scale,
                       scale * kLog2e, ((scale_dev != nullptr || scale <= 20.f) && GBT == nullptr) ? 1 : 0, 0, scale_dev, {}};
  CUtensorMap ta, tb, ta_lo, tb_lo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&ta, T, (uint64_t)P, (uint64_t)m_local, (uint64_t)ldt, kBK, kBM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tb, V, (uint64_t)P, (uint64_t)n_global, (uint64_t)ldv, kBK, BN))) return rc;
  if ((rc = make_tmap_bf16_2d(&ta_lo, T_lo ? T_lo : T, (uint64_t)P, (uint64_t)m_local, (uint64_t)ldt, kBK, kBM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tb_lo, V_lo ? V_lo : V, (uint64_t)P, (uint64_t)n_global, (uint64_t)ldv, kBK, BN))) return rc;
  const int split = (T_lo ? 1 : 0) | (V_lo ? 2 : 0);
  auto kern = itc_fused_small_kernel<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmallSmem) != cudaSuccess) {
      set_error("tic_itc_fwd_bwd_small: cudaFuncSetAttribute failed");
      return TIC_E_ATTR;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m_tiles * n_tiles);
  cfg.blockDim = dim3(64 + 32 * kItcEpiWarps);
  cfg.dynamicSmemBytes = kFusedSmallSmem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = m_tiles * n_tiles;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, ta_lo, tb, tb_lo, split, m_local, n_global, P, fp, bp);
  if (e != cudaSuccess) {
    set_error("tic_itc_fwd_bwd_small: launch failed: %s", cudaGetErrorString(e));
    return TIC_E_LAUNCH;
  }
  return TIC_OK;
}

int tic_itc_grad_finalize(const float* acc, int64_t ld_acc, const void* X, const void* X_lo, int64_t ldx, const float* rinv,
                          const void* Xo, const void* Xo_lo, int64_t ldxo, const float* rinv_o, int rows, int P, float scale, float diag_coef, float* dX_f32,
                          int64_t ld_df, void* dX_bf16, void* dX_bf16_lo, int64_t ld_db, float* r_sum, int acc_div_rinv,
                          const float* scale_dev, void* stream) {
  TIC_CHECK_ARG(acc && X && rinv && rows > 0 && P > 0, "tic_itc_grad_finalize: bad arguments");
  TIC_CHECK_ARG(dX_f32 || dX_bf16, "tic_itc_grad_finalize: no output requested");
  const bool vec = (P & 7) == 0 && P <= 1024 && (ld_acc & 3) == 0 && (ldx & 7) == 0 && (ldxo & 7) == 0 && (ld_df & 3) == 0 &&
                   (ld_db & 7) == 0 && aligned16(acc) && aligned16(X) && aligned16(X_lo) && aligned16(Xo) && aligned16(Xo_lo) &&
                   aligned16(dX_f32) && aligned16(dX_bf16) && aligned16(dX_bf16_lo);
  const int wpb = vec ? 4 : 8;   // small blocks: the small-batch step has few rows, spread them over the SMs
  auto kern = vec ? itc_grad_finalize_vec_kernel : itc_grad_finalize_kernel;
  launch_k(kern, dim3(ceil_div(rows, wpb)), dim3(wpb * 32), 0, static_cast<cudaStream_t>(stream),
      acc, ld_acc, static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(X_lo), ldx, rinv,
      static_cast<const __nv_bfloat16*>(Xo), static_cast<const __nv_bfloat16*>(Xo_lo), ldxo, rinv_o,
      rows, P, scale, diag_coef, dX_f32, ld_df, static_cast<__nv_bfloat16*>(dX_bf16), static_cast<__nv_bfloat16*>(dX_bf16_lo), ld_db,
      r_sum, acc_div_rinv, scale_dev);
  TIC_CHECK_LAUNCH("tic_itc_grad_finalize");
  return TIC_OK;
}

int tic_itc_ds_operands(const float* dS, int64_t ldds, int m_local, int n_global, const float* rinv_t, const float* rinv_v,
                        void* GA, void* GA_lo, int64_t ld_ga, void* GBT, void* GBT_lo, int64_t ld_gbt, void* stream) {
  TIC_CHECK_ARG(dS && rinv_t && rinv_v && GA && m_local > 0 && n_global > 0, "tic_itc_ds_operands: bad arguments");
  dim3 grid(ceil_div(n_global, 32), ceil_div(m_local, 8)), block(32, 8);
  (void)ld_gbt;   // GB shares GA's layout
  launch_k(itc_ds_operands_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream),
      dS, ldds, m_local, n_global, rinv_t, rinv_v, static_cast<__nv_bfloat16*>(GA), static_cast<__nv_bfloat16*>(GA_lo), ld_ga,
      static_cast<__nv_bfloat16*>(GBT), static_cast<__nv_bfloat16*>(GBT_lo));
  TIC_CHECK_LAUNCH("tic_itc_ds_operands");
  return TIC_OK;
}

}  // extern "C"
