// clip_loss on a MATERIALISED similarity matrix (models/utils.py:225-231):  (CE(S, I) + CE(S^T, I)) / 2.
// This is the drop-in for `utils.clip_loss(similarity)` when the caller hands over an arbitrary [B,B] fp32 matrix
// (the fused ITC path in itc.cu never materialises S).  HBM-bound: the forward reads every element of S exactly
// once — each 64x128 tile is staged in shared memory and feeds both the row and the column online-softmax — and the
// backward is one read of S plus one write of dS.
#include "common.cuh"
#include "tic_ptx.cuh"

namespace tic {

constexpr int kCeCols = 128;  // columns per strip
constexpr int kCeRows = 64;   // rows per smem tile

__host__ __device__ inline int ce_nstrips(int B) { return (B + kCeCols - 1) / kCeCols; }
__host__ __device__ inline int ce_nseg(int B) {
  const int tiles = (B + kCeRows - 1) / kCeRows;
  int want = (4 * 148 + ce_nstrips(B) - 1) / ce_nstrips(B);
  if (want < 1) want = 1;
  return want < tiles ? want : tiles;
}

__device__ __forceinline__ void ms_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { m = mn; s = 0.f; return; }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

// grid (nstrips, nseg); block 256. Row partials: rp[strip][row] = (max, sumexp) over the strip's 128 columns.
// Column partials: cp[seg][col] = (max, sumexp) over the segment's rows.
// Bandwidth-oriented: every thread issues its 8 16-byte loads of a 64x128 tile back to back (32 KB in flight per block),
// the tile is staged in shared memory with a 129-float row pitch (conflict-free for both walks), then 128 threads walk
// COLUMNS (64 rows each) and 128 threads walk ROW halves (64 columns each) with a one-pass online softmax — each element of S
// is read from HBM once and from shared memory twice.  (The round-1 form reduced every row with 10 warp shuffles per 512
// bytes loaded and finished in a single 1024-thread block: 99 us at B = 4096 = 10 % of HBM.)
__device__ __forceinline__ void ms_push(float& m, float& s, float x) {
  if (x <= m) {
    s += __expf(fmaxf(x - m, -INFINITY));   // (-inf) - (-inf) = NaN -> fmaxf picks -inf -> adds 0 (padding / masked logits)
  } else {          // new maximum (rare after the first few elements): rescale the running sum
    s = s * __expf(m - x) + 1.f;
    m = x;
  }
}

__global__ void __launch_bounds__(256)
ce_bidir_fwd_kernel(const float* __restrict__ S, int64_t lds, int B, float2* __restrict__ rp, float2* __restrict__ cp,
                    float* __restrict__ diag) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  constexpr int kPitch = kCeCols + 1;
  __shared__ float tile[kCeRows * kPitch];
  const int strip = blockIdx.x, seg = blockIdx.y, nseg = gridDim.y;
  const int t = threadIdx.x;
  const int c0 = strip * kCeCols;
  const int tiles = (B + kCeRows - 1) / kCeRows;
  const int t_begin = static_cast<int>(static_cast<int64_t>(tiles) * seg / nseg);
  const int t_end = static_cast<int>(static_cast<int64_t>(tiles) * (seg + 1) / nseg);
  const bool vec = (lds & 3) == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0 && c0 + kCeCols <= B;
  float cm = -INFINITY, cs = 0.f;      // column walker state (threads 0..127: column c0 + t)
  for (int tl = t_begin; tl < t_end; ++tl) {
    const int r0 = tl * kCeRows;
    // ---- global -> registers: 8 independent 16-byte loads per thread
    float4 f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = t + 256 * i, rr = q >> 5, c4 = (q & 31) * 4;
      const int r = r0 + rr;
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
    __syncthreads();          // the previous tile has been consumed
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = t + 256 * i, rr = q >> 5, c4 = (q & 31) * 4;
      float* d = tile + rr * kPitch + c4;
      d[0] = f[i].x; d[1] = f[i].y; d[2] = f[i].z; d[3] = f[i].w;
      const int r = r0 + rr, dc = r - (c0 + c4);      // the diagonal element of this row, if it sits in this float4
      if (r < B && dc >= 0 && dc < 4) diag[r] = dc == 0 ? f[i].x : (dc == 1 ? f[i].y : (dc == 2 ? f[i].z : f[i].w));
    }
    __syncthreads();
    if (t < kCeCols) {
      // ---- column walk: 64 rows of column t
#pragma unroll 8
      for (int rr = 0; rr < kCeRows; ++rr) ms_push(cm, cs, tile[rr * kPitch + t]);
    } else {
      // ---- row walk: thread pair (2 x 64 columns) per row
      const int rr = (t - kCeCols) >> 1, h = t & 1;
      float m = -INFINITY, s = 0.f;
      const float* row = tile + rr * kPitch + h * 64;
#pragma unroll 8
      for (int c = 0; c < 64; ++c) ms_push(m, s, row[c]);
      const float m2 = __shfl_xor_sync(0xffffffffu, m, 1), s2 = __shfl_xor_sync(0xffffffffu, s, 1);
      ms_combine(m, s, m2, s2);
      if (h == 0 && r0 + rr < B) rp[static_cast<int64_t>(strip) * B + r0 + rr] = make_float2(m, s);
    }
  }
  if (t < kCeCols && c0 + t < B) cp[static_cast<int64_t>(seg) * B + c0 + t] = make_float2(cm, cs);
}

// Combine the partials into the lse vectors, many blocks; blockIdx.y = direction.  Block partials of the loss go to `bp`.
__global__ void __launch_bounds__(256)
ce_bidir_lse_kernel(const float2* __restrict__ rp, int nstrips, const float2* __restrict__ cp, int nseg, const float* __restrict__ diag,
                    int B, float* __restrict__ lse_row, float* __restrict__ lse_col, float* __restrict__ bp) {
  pdl_trigger();
  pdl_wait();
  const int dir = blockIdx.y;
  const float2* part = dir == 0 ? rp : cp;
  const int np = dir == 0 ? nstrips : nseg;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float term = 0.f;
  if (i < B) {
    float m = -INFINITY, s = 0.f;
    for (int k = 0; k < np; ++k) { const float2 p = part[static_cast<int64_t>(k) * B + i]; ms_combine(m, s, p.x, p.y); }
    const float l = m + logf(s);
    (dir == 0 ? lse_row : lse_col)[i] = l;
    term = l - diag[i];
  }
  __shared__ float sw[8];
  term = warp_sum(term);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sw[w];
    bp[dir * gridDim.x + blockIdx.x] = a;
  }
}

// Single small block: fixed-order sum of the block partials -> loss (deterministic).
__global__ void ce_bidir_loss_kernel(const float* __restrict__ bp, int n, int B, float* __restrict__ loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sred[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += bp[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
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
         sizeof(float) * 2 * ((B + 255) / 256) + 64;
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
  const int nb = ceil_div(B, 256);
  launch_k(ce_bidir_fwd_kernel, dim3(dim3(ns, ng)), dim3(256), 0, st, S, lds, B, rp, cp, diag);
  launch_k(ce_bidir_lse_kernel, dim3(nb, 2), dim3(256), 0, st, rp, ns, cp, ng, diag, B, lse_row, lse_col, bp);
  launch_k(ce_bidir_loss_kernel, dim3(1), dim3(256), 0, st, bp, 2 * nb, B, loss);
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
