// clip_loss on a MATERIALISED similarity matrix (models/utils.py:225-231):  (CE(S, I) + CE(S^T, I)) / 2.
// This is the drop-in for `utils.clip_loss(similarity)` when the caller hands over an arbitrary [B,B] fp32 matrix
// (the fused ITC path in itc.cu never materialises S).  HBM-bound: the forward reads every element of S exactly
// once — each 64x128 tile lives in the registers of 8 warps and feeds both the row and the column softmax statistics with
// one exponential per element — and the backward is one read of S plus one write of dS.
#include "common.cuh"
#include "tic_ptx.cuh"

namespace tic {

constexpr int kCeCols = 128;  // columns per strip
constexpr int kCeRows = 64;   // rows per tile (two half tiles of 32)
constexpr int kCeLseThreads = 64;   // combine kernel: small blocks, so that 2 x B outputs spread over the SMs

__host__ __device__ inline int ce_nstrips(int B) { return (B + kCeCols - 1) / kCeCols; }
__host__ __device__ inline int ce_nseg(int B) {
  const int tiles = (B + kCeRows - 1) / kCeRows;
  int want = (3 * 148) / ce_nstrips(B);      // 3 resident blocks per SM: one full wave, no tail
  if (want < 1) want = 1;
  return want < tiles ? want : tiles;
}

__device__ __forceinline__ void ms_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { m = mn; s = 0.f; return; }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

// grid (nstrips, nseg); block 256 = 8 warps.  Row partials: rp[strip][row] = (reference, sum of exp(x - reference)) over the
// strip's 128 columns.  Column partials: cp[seg][col] = the same over the segment's rows.
//
// Instruction-bound, not latency-bound: 16.7 M elements at B = 4096 leave ~20 issue slots per element at HBM speed.  The
// round-1 form (10 shuffles per 512 bytes) and the first round-2 form (shared-memory tile + two branchy online-softmax walks,
// 2 exponentials and ~22 instructions per element: 70 us = 15 % of HBM) were both issue-bound.  This form spends ~8:
//   * warp w owns rows w, w+8, .., w+56 of each 64 x 128 tile; a lane owns 4 consecutive columns (one 16-byte load per row);
//   * ONE exponential per element against a warp-wide running reference (the maximum the warp has seen so far in its
//     columns): row sums and column sums are plain additions of the same e = 2^((x - ref) log2 e);
//   * column sums accumulate in registers across all tiles of the block (rescaled when the reference grows);
//   * the 8 row sums of a tile are reduced across the 32 lanes with a transposing butterfly (9 shuffles for 8 rows).
// Exactness on arbitrary input: a row (or column) whose own maximum lies more than ~69 below the shared reference would
// underflow; such sums (< 1e-30) are recomputed exactly against their own maximum (warp-cooperative for rows, a serial
// re-read of the column segment for columns) — never taken for logits of bounded spread, but the kernel is a drop-in for
// utils.clip_loss on ANY matrix (-inf padding included).
constexpr float kCeTiny = 1e-30f;
constexpr float kCeLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ce_ex2(float x) {     // ex2.approx.ftz: 2^(-inf) = +0, relative error ~2^-22
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 3)
ce_bidir_fwd_kernel(const float* __restrict__ S, int64_t lds, int B, float2* __restrict__ rp, float2* __restrict__ cp,
                    float* __restrict__ diag, unsigned int* __restrict__ ticket) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  __shared__ float sm_s[8][kCeCols];
  __shared__ float sm_m[8];
  const int strip = blockIdx.x, seg = blockIdx.y, nseg = gridDim.y;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (strip == 0 && seg == 0 && t == 0) *ticket = 0u;      // the lse kernel behind this one counts its blocks on it
  const int c0 = strip * kCeCols, c4 = lane * 4;
  const int tiles = (B + kCeRows - 1) / kCeRows;
  // half tiles of 32 rows: warp w owns rows w, w+8, w+16, w+24 of each
  const int h_begin = 2 * static_cast<int>(static_cast<int64_t>(tiles) * seg / nseg);
  const int h_end = 2 * static_cast<int>(static_cast<int64_t>(tiles) * (seg + 1) / nseg);
  const bool vec = (lds & 3) == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0 && c0 + kCeCols <= B;
  float mref = -INFINITY;                      // warp-uniform running reference
  float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;   // column sums of exp(x - mref), columns c0 + c4 .. + 3, this warp's rows
  const bool b4 = lane & 16, b3 = lane & 8;
  const int my_i = (b4 ? 2 : 0) + (b3 ? 1 : 0);       // the row (of the 4 in a half tile) whose sum the butterfly leaves here

  auto load4 = [&](float4 (&f)[4], int h) {
    const int r0 = h * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + w + 8 * i;
      f[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (r < B) {
        const float* p = S + static_cast<int64_t>(r) * lds + c0 + c4;
        if (vec) {
          f[i] = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (c0 + c4 + 0 < B) f[i].x = __ldg(p + 0);
          if (c0 + c4 + 1 < B) f[i].y = __ldg(p + 1);
          if (c0 + c4 + 2 < B) f[i].z = __ldg(p + 2);
          if (c0 + c4 + 3 < B) f[i].w = __ldg(p + 3);
        }
      }
    }
  };
  auto consume = [&](const float4 (&f)[4], int h) {
    const int r0 = h * 32;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mx = fmaxf(fmaxf(mx, fmaxf(f[i].x, f[i].y)), fmaxf(f[i].z, f[i].w));
      const int r = r0 + w + 8 * i, dc = r - (c0 + c4);      // the diagonal element of this row, if it sits in this float4
      if (r < B && dc >= 0 && dc < 4) diag[r] = dc == 0 ? f[i].x : (dc == 1 ? f[i].y : (dc == 2 ? f[i].z : f[i].w));
    }
    // The reference only has to keep exp(x - ref) inside the fp32 range, not to BE the maximum: it is raised when some
    // element exceeds it by more than 40 (one vote per half tile; the 5-shuffle warp maximum runs a few times per block).
    if (__any_sync(0xffffffffu, mx > mref + 40.f)) {
      mx = warp_max(mx);
      if (mref != -INFINITY) {
        const float sc = __expf(mref - mx);
        cs0 *= sc; cs1 *= sc; cs2 *= sc; cs3 *= sc;
      }
      mref = mx;
    }
    const float mm = mref == -INFINITY ? 0.f : mref;
    float rs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float e0 = ce_ex2((f[i].x - mm) * kCeLog2e), e1 = ce_ex2((f[i].y - mm) * kCeLog2e);
      const float e2 = ce_ex2((f[i].z - mm) * kCeLog2e), e3 = ce_ex2((f[i].w - mm) * kCeLog2e);
      cs0 += e0; cs1 += e1; cs2 += e2; cs3 += e3;
      rs[i] = (e0 + e1) + (e2 + e3);
    }
    // transposing butterfly: 4 row sums x 32 lanes -> lane holds the full sum of row my_i (2 + 1 + 3 shuffles)
    const float a0 = (b4 ? rs[2] : rs[0]) + __shfl_xor_sync(0xffffffffu, b4 ? rs[0] : rs[2], 16);
    const float a1 = (b4 ? rs[3] : rs[1]) + __shfl_xor_sync(0xffffffffu, b4 ? rs[1] : rs[3], 16);
    float d = (b3 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, b3 ? a0 : a1, 8);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    const int r = r0 + w + 8 * my_i;
    const bool writer = (lane & 7) == 0 && r < B;
    const bool tiny = !(d >= kCeTiny);         // underflow against the shared reference (or an all -inf row)
    if (writer && !tiny) rp[static_cast<int64_t>(strip) * B + r] = make_float2(mref, d);
    const unsigned fl = __ballot_sync(0xffffffffu, tiny && writer);
    if (fl != 0u) {                            // rare: exact partial of the flagged rows against their own maximum
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int holder = ((i >> 1) & 1) * 16 + (i & 1) * 8;
        if ((fl >> holder) & 1u) {             // warp-uniform
          const float m = warp_max(fmaxf(fmaxf(f[i].x, f[i].y), fmaxf(f[i].z, f[i].w)));
          float sx = 0.f;
          if (m != -INFINITY) sx = (__expf(f[i].x - m) + __expf(f[i].y - m)) + (__expf(f[i].z - m) + __expf(f[i].w - m));
          sx = warp_sum(sx);
          if (lane == 0) rp[static_cast<int64_t>(strip) * B + r0 + w + 8 * i] = make_float2(m, sx);
        }
      }
    }
  };

  // software pipeline over half tiles, depth 2: the loads of the next TWO half tiles (8 x 16 bytes per thread) are in flight
  // while this one is consumed — 96 KB outstanding per SM with 3 resident blocks.  Measured equal to depth 1 x 4 blocks (64 KB
  // outstanding): 24.3 vs 24.1 us for the call at B = 4096, so bytes in flight are not what holds the statistics kernel at
  // 3.5 TB/s; the 512-byte row pieces per warp (16 KB apart) are the remaining suspect.
  float4 f0[4], f1[4], f2[4];
  if (h_begin < h_end) load4(f0, h_begin);
  if (h_begin + 1 < h_end) load4(f1, h_begin + 1);
  for (int h = h_begin; h < h_end; h += 3) {
    if (h + 2 < h_end) load4(f2, h + 2);
    consume(f0, h);
    if (h + 3 < h_end) load4(f0, h + 3);
    if (h + 1 < h_end) consume(f1, h + 1);
    if (h + 4 < h_end) load4(f1, h + 4);
    if (h + 2 < h_end) consume(f2, h + 2);
  }
  // ---- column partials of the block: combine the 8 warps' (reference, sums)
  sm_s[w][c4 + 0] = cs0; sm_s[w][c4 + 1] = cs1; sm_s[w][c4 + 2] = cs2; sm_s[w][c4 + 3] = cs3;
  if (lane == 0) sm_m[w] = mref;
  __syncthreads();
  if (t < kCeCols && c0 + t < B) {
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, sm_m[k]);
    float sum = 0.f;
    if (m != -INFINITY) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (sm_m[k] != -INFINITY) sum += sm_s[k][t] * __expf(sm_m[k] - m);
    }
    if (!(sum >= kCeTiny)) {                   // rare: exact online softmax over the segment's rows of this column
      const int r_begin = h_begin * 32, r_end = min(h_end * 32, B);
      m = -INFINITY;
      for (int r = r_begin; r < r_end; ++r) m = fmaxf(m, __ldg(S + static_cast<int64_t>(r) * lds + c0 + t));
      sum = 0.f;
      if (m != -INFINITY)
        for (int r = r_begin; r < r_end; ++r) sum += __expf(__ldg(S + static_cast<int64_t>(r) * lds + c0 + t) - m);
    }
    cp[static_cast<int64_t>(seg) * B + c0 + t] = make_float2(m, sum);
  }
}

// Combine the partials into the lse vectors, many blocks; blockIdx.y = direction.  Block partials of the loss go to `bp`; the
// LAST block to finish (ticket, zeroed by the forward kernel) adds them in index order -> loss (deterministic, no third launch).
__global__ void __launch_bounds__(kCeLseThreads)
ce_bidir_lse_kernel(const float2* __restrict__ rp, int nstrips, const float2* __restrict__ cp, int nseg, const float* __restrict__ diag,
                    int B, float* __restrict__ lse_row, float* __restrict__ lse_col, float* __restrict__ bp,
                    unsigned int* __restrict__ ticket, float* __restrict__ loss) {
  pdl_trigger();
  pdl_wait();
  const int dir = blockIdx.y;
  const float2* part = dir == 0 ? rp : cp;
  const int np = dir == 0 ? nstrips : nseg;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float term = 0.f;
  if (i < B) {
    // two passes over the (few, L2-resident) partials instead of a serial chain of combines: all loads of a pass are
    // independent, and the second pass costs one exponential per partial
    float m = -INFINITY;
#pragma unroll 8
    for (int k = 0; k < np; ++k) m = fmaxf(m, part[static_cast<int64_t>(k) * B + i].x);
    float s = 0.f;
    if (m != -INFINITY) {
#pragma unroll 8
      for (int k = 0; k < np; ++k) {
        const float2 p = part[static_cast<int64_t>(k) * B + i];
        s += p.y * __expf(fmaxf(p.x - m, -INFINITY));      // (-inf) - m = -inf -> 0; never NaN (m is finite here)
      }
    }
    const float l = m + logf(s);
    (dir == 0 ? lse_row : lse_col)[i] = l;
    term = l - diag[i];
  }
  __shared__ float sw[kCeLseThreads / 32];
  __shared__ bool last;
  term = warp_sum(term);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < kCeLseThreads / 32; ++w) a += sw[w];
    bp[dir * gridDim.x + blockIdx.x] = a;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    const int n = gridDim.x * gridDim.y;
    float acc = 0.f;
    for (int k = threadIdx.x; k < n; k += 32) acc += reinterpret_cast<const volatile float*>(bp)[k];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) loss[0] = acc / (2.0f * B);
  }
}

__global__ void ce_bidir_bwd_kernel(const float* __restrict__ S, int64_t lds, int B, const float* __restrict__ lse_row,
                                    const float* __restrict__ lse_col, const float* __restrict__ grad_loss,
                                    float* __restrict__ dS, int64_t ldds) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int r = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= B) return;
  const float g = grad_loss[0] / (2.0f * B);
  const float lr = lse_row[r];
  const float* p = S + static_cast<int64_t>(r) * lds + c;
  float* d = dS + static_cast<int64_t>(r) * ldds + c;
  const bool vec = (lds & 3) == 0 && (ldds & 3) == 0 && c + 3 < B &&
                   ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(dS)) & 15) == 0;
  float x[4], o[4];
  if (vec) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p));
    x[0] = f.x; x[1] = f.y; x[2] = f.z; x[3] = f.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = (c + k < B) ? p[k] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cc = c + k;
    const float lc = cc < B ? __ldg(lse_col + cc) : 0.f;
    o[k] = g * (__expf(x[k] - lr) + __expf(x[k] - lc) - (cc == r ? 2.f : 0.f));
  }
  if (vec) {
    *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c + k < B) d[k] = o[k];
  }
}

}  // namespace tic

using namespace tic;

extern "C" {

int64_t tic_ce_bidir_workspace_bytes(int B) {
  if (B <= 0) return 0;
  return static_cast<int64_t>(sizeof(float2)) * B * (ce_nstrips(B) + ce_nseg(B)) + sizeof(float) * B +
         sizeof(float) * 2 * ((B + kCeLseThreads - 1) / kCeLseThreads) + 64;
}

int tic_ce_bidir_fwd(const float* S, int64_t lds, int B, float* lse_row, float* lse_col, float* loss, void* workspace,
                     void* stream) {
  TIC_CHECK_ARG(S && lse_row && lse_col && loss && workspace && B > 0 && lds >= B, "tic_ce_bidir_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ns = ce_nstrips(B), ng = ce_nseg(B);
  float2* rp = static_cast<float2*>(workspace);
  float2* cp = rp + static_cast<int64_t>(ns) * B;
  float* diag = reinterpret_cast<float*>(cp + static_cast<int64_t>(ng) * B);
  float* bp = diag + B;
  const int nb = ceil_div(B, kCeLseThreads);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(bp + 2 * nb);      // inside the 64-byte tail of the workspace
  launch_k(ce_bidir_fwd_kernel, dim3(dim3(ns, ng)), dim3(256), 0, st, S, lds, B, rp, cp, diag, ticket);
  launch_k(ce_bidir_lse_kernel, dim3(nb, 2), dim3(kCeLseThreads), 0, st, rp, ns, cp, ng, diag, B, lse_row, lse_col, bp, ticket, loss);
  TIC_CHECK_LAUNCH("tic_ce_bidir_fwd");
  return TIC_OK;
}

int tic_ce_bidir_bwd(const float* S, int64_t lds, int B, const float* lse_row, const float* lse_col, const float* grad_loss,
                     float* dS, int64_t ldds, void* stream) {
  TIC_CHECK_ARG(S && lse_row && lse_col && grad_loss && dS && B > 0, "tic_ce_bidir_bwd: bad arguments");
  dim3 grid(ceil_div(ceil_div(B, 4), 256), B);
  launch_k(ce_bidir_bwd_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), S, lds, B, lse_row, lse_col, grad_loss, dS, ldds);
  TIC_CHECK_LAUNCH("tic_ce_bidir_bwd");
  return TIC_OK;
}

}  // extern "C"
