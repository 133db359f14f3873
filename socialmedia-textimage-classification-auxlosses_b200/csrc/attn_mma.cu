// Attention-pool kernels, tensor-core form (v3) — the HBM-bound middle of the `attention` fusion after the CLS-row collapse
// (models/mm_late.py:98-113,195-210; only ctx[:,0,:] is consumed, :111).  Same arithmetic and interface as the v2
// kernels in fusion.cu; what changes is where the FLOPs go:
//
//   per sample, x_v [Lv, 768] bf16 streams ONCE through a 6-stage shared-memory ring (bulk async copies, one per token
//   row, rows padded by 16 bytes so ldmatrix is bank-conflict free), 16 tokens per stage;
//   phase A   scores   S[16 tok x 8]   = X[16 x 768] * Qm[768 x 8]        mma.sync m16n8k16, K split over the 8 warps
//             Qm columns = (q_0 hi, q_0 lo, q_1 hi, q_1 lo, 0...) : the fp32 queries as bf16 (hi, lo) pairs
//   softmax   online (running max / sum per query), 32 scores per chunk, done redundantly by every warp
//   phase B   pooled   Acc[8 x 768]   += Wm[16 x 16 tok] * X[16 tok x 768]  each warp owns 96 feature columns
//             Wm rows    = (w_0 hi, w_0 lo, w_1 hi, w_1 lo, 0...) : the softmax weights as bf16 (hi, lo) pairs
//
// v2 spent ~150 SIMT instructions per token row per warp in dependent FMA/shuffle chains and reached 40-50 % of the HBM
// roofline; here a 16-token chunk costs each warp 12 ldmatrix + 18 mma + a few dozen scalar instructions.
// The backward kernel is the same pipeline with q := dL/dxbar and weights ds_j = attn_j (g.x_j - g.xbar) * scale.
#include <cstdlib>
#include "common.cuh"
#include "tic_ptx.cuh"

namespace tic {

constexpr int kA3E = 768, kA3Rows = 16, kA3Warps = 8;
// Two CTAs per SM with a 3-stage ring each: the same 144 KB in flight per SM as one 6-stage CTA, but twice the warps, and the
// per-chunk latency chain (partials -> shuffles -> exp -> weights) of one CTA overlaps the other CTA's tensor work.
constexpr int kA3Stages = 3, kA3CtasPerSm = 2;
constexpr int kA3RowBytes = kA3E * 2 + 16;                 // padded smem row: 8 consecutive rows hit 8 distinct 16-byte banks
constexpr int kA3ChunkBytes = kA3Rows * kA3RowBytes;       // 24832
constexpr int kA3KSlice = kA3E / kA3Warps;                 // 96 = 6 k-steps of 16 (phase A) = 12 column groups of 8 (phase B)

__device__ __forceinline__ void a3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void a3_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (hi, lo) bf16 split of two fp32 values by TRUNCATION: hi = upper 16 bits (one PRMT packs both), lo = rn_bf16(x - hi).
// hi + lo carries ~16 mantissa bits, exactly like the round-to-nearest split, without the slow F2F conversions.
__device__ __forceinline__ uint32_t split_pack(float x0, float x1, int lo) {
  const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
  if (!lo) return __byte_perm(u0, u1, 0x7632);
  return pack_bf16x2(x0 - __uint_as_float(u0 & 0xffff0000u), x1 - __uint_as_float(u1 & 0xffff0000u));
}

struct A3Smem {
  uint32_t ring, full0, empty0, pbar;
  uint8_t* ring_ptr;
  float* part;      // [2][8 warps][16 tok][8]
  float* vec;       // [NPASS][Lv]  scores (forward) / attention weights (backward)
};
__device__ __forceinline__ A3Smem a3_carve(uint8_t* raw) {
  A3Smem s;
  const uint32_t base = (smem_u32(raw) + 127u) & ~127u;
  uint8_t* p = raw + (base - smem_u32(raw));
  s.ring = base;
  s.ring_ptr = p;
  s.full0 = base + kA3Stages * kA3ChunkBytes;
  s.empty0 = s.full0 + 8 * kA3Stages;
  s.pbar = s.empty0 + 8 * kA3Stages;      // "partial scores of a chunk are written": one arrival per consumer warp
  s.part = reinterpret_cast<float*>(p + kA3Stages * kA3ChunkBytes + 128);
  s.vec = s.part + 2 * kA3Warps * kA3Rows * 8;
  return s;
}
static size_t a3_smem_bytes(int npass, int Lv) {
  return 128 + static_cast<size_t>(kA3Stages) * kA3ChunkBytes + 128 + sizeof(float) * (2 * kA3Warps * kA3Rows * 8 + npass * Lv);
}

// BWD == false: kq = augmented queries [NPASS*B, ldkq] fp32 (column E = bias term), outputs xbar / attn.
// BWD == true : kq = dL/dxbar [NPASS*B, ldkq] fp32, `attn` and `xbar_f` are inputs, outputs dkq (+ augmented column dc).
template <int NPASS, bool BWD, bool DRY = false>
__global__ void __launch_bounds__(288, kA3CtasPerSm)
attn_pool_mma_kernel(const __nv_bfloat16* __restrict__ xv, int64_t bstride, const float* __restrict__ kq, int64_t ldkq, int B, int Lv,
                     float scale, __nv_bfloat16* __restrict__ out_b, __nv_bfloat16* __restrict__ out_lo, int64_t ld_ob,
                     float* __restrict__ xbar_f, int64_t ld_xf, float* __restrict__ attn, int64_t ld_attn) {
  pdl_trigger();
  pdl_wait();
  constexpr int E = kA3E;
  extern __shared__ uint8_t a3_raw[];
  const A3Smem sm = a3_carve(a3_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the ring starts zeroed: rows past the end of a sample are never copied, and 0 x (stale NaN pattern) would poison the MMAs
  for (int i = threadIdx.x; i < kA3Stages * kA3ChunkBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sm.ring_ptr)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kA3Stages; ++i) { mbar_init(sm.full0 + 8 * i, 1); mbar_init(sm.empty0 + 8 * i, kA3Warps); }
    mbar_init(sm.pbar, kA3Warps);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  const int nchunks = (Lv + kA3Rows - 1) / kA3Rows;

  if (warp == kA3Warps) {
    // ===================== producer warp: one bulk copy per token row =====================
    uint32_t it = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      const __nv_bfloat16* xb = xv + static_cast<int64_t>(b) * bstride;
      for (int c = 0; c < nchunks; ++c, ++it) {
        const uint32_t st = it % kA3Stages, ph = (it / kA3Stages) & 1u;
        const int rows = min(kA3Rows, Lv - c * kA3Rows);
        if (lane == 0) {
          mbar_wait(sm.empty0 + 8 * st, ph ^ 1u);
          mbar_arrive_expect_tx(sm.full0 + 8 * st, static_cast<uint32_t>(rows) * E * 2);
        }
        __syncwarp();
        if (lane < rows)
          a3_bulk_g2s(sm.ring + st * kA3ChunkBytes + lane * kA3RowBytes, xb + static_cast<int64_t>(c * kA3Rows + lane) * E, E * 2,
                      sm.full0 + 8 * st);
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  const int g = lane >> 2, t = lane & 3;               // mma fragment coordinates
  const int pl = (NPASS == 2) ? (lane >> 4) : 0;       // the query this lane does the softmax bookkeeping for
  const int jl = lane & 15;                            // ... and the token of the chunk
  const int ka = warp * kA3KSlice;                     // this warp's K slice (phase A) = feature columns (phase B)
  uint32_t it = 0;
  int buf = 0;   // 3 partial-score buffers: a warp may run one chunk ahead of the slowest reader (split arrive / wait below)
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    // ---- per-sample setup: B fragments of Qm for my K slice (column n = g: query n>>1, hi/lo part n&1)
    uint32_t qb[6][2];
    {
      const int n = g;
      const bool live = n < 2 * NPASS;
      const float* kr = kq + static_cast<int64_t>((live ? (n >> 1) : 0) * B + b) * ldkq + ka + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) {
        const float2 f0 = __ldg(reinterpret_cast<const float2*>(kr + 16 * ks));
        const float2 f8 = __ldg(reinterpret_cast<const float2*>(kr + 16 * ks + 8));
        qb[ks][0] = live ? split_pack(f0.x, f0.y, n & 1) : 0u;
        qb[ks][1] = live ? split_pack(f8.x, f8.y, n & 1) : 0u;
      }
    }
    float cb = 0.f, Dp = 0.f, m_l = -INFINITY, l_l = 0.f, dc_l = 0.f;
    if (!BWD) {
      cb = __ldg(kq + static_cast<int64_t>(pl * B + b) * ldkq + E);
    } else {
      // D_p = <g_p, xbar_p>; attention weights of this sample -> smem
      const float* gp = kq + static_cast<int64_t>(pl * B + b) * ldkq;
      const float* xp = xbar_f + static_cast<int64_t>(pl * B + b) * ld_xf;
      float d = 0.f;
      for (int k = jl * 4; k < E; k += 64) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(gp + k)), c4 = __ldg(reinterpret_cast<const float4*>(xp + k));
        d += a.x * c4.x + a.y * c4.y + a.z * c4.z + a.w * c4.w;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      Dp = d;
      for (int p = 0; p < NPASS; ++p)
        for (int j = threadIdx.x; j < Lv; j += 256) sm.vec[p * Lv + j] = __ldg(attn + static_cast<int64_t>(p * B + b) * ld_attn + j);
      a3_bar();
    }
    float acc[12][2];
#pragma unroll
    for (int n = 0; n < 12; ++n) { acc[n][0] = 0.f; acc[n][1] = 0.f; }

    // Software pipeline over the 16-token chunks: the phase-A MMAs of chunk c+1 are issued BEFORE the softmax bookkeeping and
    // phase B of chunk c, so the shuffle/exp latency chain of one chunk overlaps the tensor work of the next (one named
    // barrier per chunk publishes the partial scores).
    auto phase_a = [&](uint32_t chunk, float (&sc)[4]) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t a_addr = chunk + (lane & 15) * kA3RowBytes + (ka + (lane >> 4) * 8) * 2;
#pragma unroll
      for (int ks = 0; ks < 6; ks += 2) {   // two independent accumulator chains
        uint32_t a[4], a2[4];
        ldsm_x4(a_addr + ks * 32, a);
        ldsm_x4(a_addr + (ks + 1) * 32, a2);
        mma_bf16(s0, a, qb[ks][0], qb[ks][1]);
        mma_bf16(s1, a2, qb[ks + 1][0], qb[ks + 1][1]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) sc[q] = s0[q] + s1[q];
    };
    // partial scores in smem: [buf][warp][query][token][hi, lo] — only the 2*NPASS live columns, consecutive lanes touch
    // consecutive 8-byte words on both the write and the read side (bank-conflict free)
    auto write_part = [&](int bf, const float (&sc)[4]) {
      if (t < NPASS) {
        float* part = sm.part + ((bf * kA3Warps + warp) * NPASS + t) * (kA3Rows * 2);
        *reinterpret_cast<float2*>(part + g * 2) = make_float2(sc[0], sc[1]);
        *reinterpret_cast<float2*>(part + (g + 8) * 2) = make_float2(sc[2], sc[3]);
      }
    };
    if (DRY) {   // measurement aid (TIC_ATTN_V3_DRY=1): the copy pipeline alone, no arithmetic, results undefined
      for (int c = 0; c < nchunks; ++c, ++it) {
        const uint32_t st = it % kA3Stages, ph = (it / kA3Stages) & 1u;
        mbar_wait(sm.full0 + 8 * st, ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.empty0 + 8 * st);
      }
      continue;
    }
    {
      const uint32_t st = it % kA3Stages, ph = (it / kA3Stages) & 1u;
      mbar_wait(sm.full0 + 8 * st, ph);
      float sc[4];
      phase_a(sm.ring + st * kA3ChunkBytes, sc);
      write_part(buf, sc);
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.pbar);
    }
    for (int c = 0; c < nchunks; ++c, ++it) {
      const uint32_t st = it % kA3Stages;
      const uint32_t chunk = sm.ring + st * kA3ChunkBytes;
      // ---- phase A of the NEXT chunk (independent of everything below)
      float sc_next[4] = {0.f, 0.f, 0.f, 0.f};
      const bool has_next = c + 1 < nchunks;
      if (has_next) {
        const uint32_t st1 = (it + 1) % kA3Stages, ph1 = ((it + 1) / kA3Stages) & 1u;
        mbar_wait(sm.full0 + 8 * st1, ph1);
        phase_a(sm.ring + st1 * kA3ChunkBytes, sc_next);
      }
      // Split barrier on the partial scores: wait for chunk `it` (every warp arrived for it one chunk ago), THEN arrive for
      // chunk it+1 — in this order a warp's arrival can never be counted into a phase an earlier chunk still needs.
      mbar_wait(sm.pbar, it & 1u);
      if (has_next) {
        write_part(buf == 2 ? 0 : buf + 1, sc_next);
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.pbar);
      }
      // ---- scores / weights of (token jl, query pl) of THIS chunk, identically in every warp
      const int jj = c * kA3Rows + jl;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kA3Warps; ++w) {
        const float2 v = *reinterpret_cast<const float2*>(sm.part + ((buf * kA3Warps + w) * NPASS + pl) * (kA3Rows * 2) + jl * 2);
        s += v.x + v.y;
      }
      float wgt, corr = 1.f;
      if (!BWD) {
        s = (jj < Lv) ? (s + cb) * scale : -INFINITY;
        if (warp == 0 && jj < Lv && (NPASS == 2 || lane < 16)) sm.vec[pl * Lv + jj] = s;
        float mx = s;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float m_new = fmaxf(m_l, mx);
        corr = __expf(m_l - m_new);            // first chunk: exp(-inf) = 0
        wgt = (jj < Lv) ? __expf(s - m_new) : 0.f;
        float ws = wgt;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
        l_l = l_l * corr + ws;
        m_l = m_new;
      } else {
        wgt = (jj < Lv) ? sm.vec[pl * Lv + jj] * (s - Dp) * scale : 0.f;
        dc_l += wgt;
      }
      // ---- phase B: Acc[rows (w_p hi, w_p lo)] += Wm * X over my 96 feature columns
      uint32_t wa[4];
      {
        const int pa = (g >> 1) & (NPASS - 1), part_lo = g & 1;
        const bool live = g < 2 * NPASS;
        const int src = pa * 16 + 2 * t;
        const float w0 = __shfl_sync(0xffffffffu, wgt, src), w1 = __shfl_sync(0xffffffffu, wgt, src + 1);
        const float w8 = __shfl_sync(0xffffffffu, wgt, src + 8), w9 = __shfl_sync(0xffffffffu, wgt, src + 9);
        wa[0] = live ? split_pack(w0, w1, part_lo) : 0u;
        wa[2] = live ? split_pack(w8, w9, part_lo) : 0u;
        wa[1] = 0u;
        wa[3] = 0u;
        if (!BWD) {
          const float cg = __shfl_sync(0xffffffffu, corr, pa * 16);
#pragma unroll
          for (int n = 0; n < 12; ++n) { acc[n][0] *= cg; acc[n][1] *= cg; }
        }
      }
      {
        const uint32_t b_addr = chunk + ((lane & 7) + ((lane >> 3) & 1) * 8) * kA3RowBytes + (ka + (lane >> 4) * 8) * 2;
#pragma unroll
        for (int np = 0; np < 6; ++np) {
          uint32_t bx[4];
          ldsm_x4_trans(b_addr + np * 32, bx);
          float d0[4] = {acc[2 * np][0], acc[2 * np][1], 0.f, 0.f};
          float d1[4] = {acc[2 * np + 1][0], acc[2 * np + 1][1], 0.f, 0.f};
          mma_bf16(d0, wa, bx[0], bx[1]);
          mma_bf16(d1, wa, bx[2], bx[3]);
          acc[2 * np][0] = d0[0]; acc[2 * np][1] = d0[1];
          acc[2 * np + 1][0] = d1[0]; acc[2 * np + 1][1] = d1[1];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty0 + 8 * st);
      buf = buf == 2 ? 0 : buf + 1;
    }

    // ---- per-sample epilogue: hi + lo rows, normalise, write out
    {
      const int p = (g >> 1) & (NPASS - 1);
      const float lp = __shfl_sync(0xffffffffu, l_l, p * 16);
      const float inv = BWD ? 1.f : 1.f / lp;
      const int64_t r = static_cast<int64_t>(p) * B + b;
#pragma unroll
      for (int n = 0; n < 12; ++n) {
        const float v0 = (acc[n][0] + __shfl_xor_sync(0xffffffffu, acc[n][0], 4)) * inv;
        const float v1 = (acc[n][1] + __shfl_xor_sync(0xffffffffu, acc[n][1], 4)) * inv;
        if (g < 2 * NPASS && (g & 1) == 0) {
          const int col = ka + 8 * n + 2 * t;
          if (!BWD && xbar_f) *reinterpret_cast<float2*>(xbar_f + r * ld_xf + col) = make_float2(v0, v1);
          if (out_b) {
            *reinterpret_cast<uint32_t*>(out_b + r * ld_ob + col) = pack_bf16x2(v0, v1);
            if (out_lo)
              *reinterpret_cast<uint32_t*>(out_lo + r * ld_ob + col) =
                  pack_bf16x2(v0 - __bfloat162float(__float2bfloat16_rn(v0)), v1 - __bfloat162float(__float2bfloat16_rn(v1)));
          }
        }
      }
    }
    if (!BWD) {
      a3_bar();   // every score of this sample is in smem (written by warp 0)
      const float M0 = __shfl_sync(0xffffffffu, m_l, 0), L0 = __shfl_sync(0xffffffffu, l_l, 0);
      const float M1 = __shfl_sync(0xffffffffu, m_l, 16), L1 = __shfl_sync(0xffffffffu, l_l, 16);
      for (int j = threadIdx.x; j < Lv; j += 256) {
        attn[static_cast<int64_t>(b) * ld_attn + j] = __expf(sm.vec[j] - M0) / L0;
        if (NPASS == 2) attn[static_cast<int64_t>(B + b) * ld_attn + j] = __expf(sm.vec[Lv + j] - M1) / L1;
      }
      a3_bar();   // ... and has been consumed before the next sample overwrites it
    } else {
      // augmented column E: dc_p = sum_j ds_j ; columns E+1..E+7 are the zero padding of the [W_K | b_K] operand
      float dcs = dc_l;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dcs += __shfl_xor_sync(0xffffffffu, dcs, o);
      if (warp == 0 && jl < 8 && (NPASS == 2 || lane < 16)) {
        const int64_t r = static_cast<int64_t>(pl) * B + b;
        const float v = jl == 0 ? dcs : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        out_b[r * ld_ob + E + jl] = hi;
        if (out_lo) out_lo[r * ld_ob + E + jl] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
      a3_bar();   // sm.vec (attention weights) is reloaded for the next sample
    }
  }
}

static int a3_grid(int B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return B < kA3CtasPerSm * sms ? B : kA3CtasPerSm * sms;
}

template <int NPASS, bool BWD>
static int a3_launch(const void* xv, int64_t bstride, const float* kq, int64_t ldkq, int B, int Lv, float scale, void* out_b,
                     void* out_lo, int64_t ld_ob, float* xbar_f, int64_t ld_xf, float* attn, int64_t ld_attn, cudaStream_t st) {
  static int dry = -1;
  if (dry < 0) { const char* e = getenv("TIC_ATTN_V3_DRY"); dry = (e && e[0] == '1') ? 1 : 0; }
  auto k = dry ? attn_pool_mma_kernel<NPASS, BWD, true> : attn_pool_mma_kernel<NPASS, BWD, false>;
  const size_t smem = a3_smem_bytes(NPASS, Lv);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return TIC_E_ATTR;
  launch_k(k, dim3(a3_grid(B)), dim3(32 * (kA3Warps + 1)), smem, st, static_cast<const __nv_bfloat16*>(xv), bstride, kq, ldkq, B, Lv,
           scale, static_cast<__nv_bfloat16*>(out_b), static_cast<__nv_bfloat16*>(out_lo), ld_ob, xbar_f, ld_xf, attn, ld_attn);
  return cudaGetLastError() == cudaSuccess ? TIC_OK : TIC_E_LAUNCH;
}

// Entry points used by tic_attn_pool_fwd / tic_attn_pool_bwd (fusion.cu) when the layout allows the bulk-copy pipeline.
int attn_pool_fwd_mma(const void* xv, int64_t bstride, const float* kq, int64_t ldkq, int B, int npass, int Lv, float scale,
                      void* xbar_b, void* xbar_lo, int64_t ld_xb, float* xbar_f, int64_t ld_xf, float* attn, int64_t ld_attn,
                      cudaStream_t st) {
  return npass == 1 ? a3_launch<1, false>(xv, bstride, kq, ldkq, B, Lv, scale, xbar_b, xbar_lo, ld_xb, xbar_f, ld_xf, attn, ld_attn, st)
                    : a3_launch<2, false>(xv, bstride, kq, ldkq, B, Lv, scale, xbar_b, xbar_lo, ld_xb, xbar_f, ld_xf, attn, ld_attn, st);
}
int attn_pool_bwd_mma(const void* xv, int64_t bstride, const float* attn, int64_t ld_attn, const float* dxbar, int64_t ld_dxb,
                      const float* xbar_f, int64_t ld_xf, int B, int npass, int Lv, float scale, void* dkq, void* dkq_lo,
                      int64_t ld_dkq, cudaStream_t st) {
  float* xf = const_cast<float*>(xbar_f);
  float* at = const_cast<float*>(attn);
  return npass == 1 ? a3_launch<1, true>(xv, bstride, dxbar, ld_dxb, B, Lv, scale, dkq, dkq_lo, ld_dkq, xf, ld_xf, at, ld_attn, st)
                    : a3_launch<2, true>(xv, bstride, dxbar, ld_dxb, B, Lv, scale, dkq, dkq_lo, ld_dkq, xf, ld_xf, at, ld_attn, st);
}

}  // namespace tic
