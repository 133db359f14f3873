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
  int want = (2 * 148 + ce_nstrips(B) - 1) / ce_nstrips(B);
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
__global__ void __launch_bounds__(256)
ce_bidir_fwd_kernel(const float* __restrict__ S, int64_t lds, int B, float2* __restrict__ rp, float2* __restrict__ cp,
                    float* __restrict__ diag) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  __shared__ float tile[kCeRows][kCeCols + 4];
  __shared__ float2 ccomb[kCeCols];
  const int strip = blockIdx.x, seg = blockIdx.y, nseg = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = strip * kCeCols;
  const int tiles = (B + kCeRows - 1) / kCeRows;
  const int t_begin = static_cast<int>(static_cast<int64_t>(tiles) * seg / nseg);
  const int t_end = static_cast<int>(static_cast<int64_t>(tiles) * (seg + 1) / nseg);
  const int ccol = threadIdx.x & (kCeCols - 1), chalf = threadIdx.x >> 7;  // 2 threads per column, 32 rows each
  float cm = -INFINITY, cs = 0.f;
  const bool vec = (lds & 3) == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0;
  for (int t = t_begin; t < t_end; ++t) {
    const int r0 = t * kCeRows;
    for (int rr = warp; rr < kCeRows; rr += 8) {
      const int r = r0 + rr;
      float x[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      const int c = c0 + lane * 4;
      if (r < B) {
        const float* p = S + static_cast<int64_t>(r) * lds + c;
        if (vec && c + 3 < B) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(p));
          x[0] = f.x; x[1] = f.y; x[2] = f.z; x[3] = f.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c + k < B) x[k] = __ldg(p + k);
        }
        if (r >= c && r < c + 4) diag[r] = x[r - c];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) tile[rr][lane * 4 + k] = x[k];
      float m = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
      m = warp_max(m);
      float s = 0.f;
      if (m != -INFINITY) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s += __expf(x[k] - m);
      }
      s = warp_sum(s);
      if (lane == 0 && r < B) rp[static_cast<int64_t>(strip) * B + r] = make_float2(m, s);
    }
    __syncthreads();
    {
      float m = -INFINITY;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) m = fmaxf(m, tile[chalf * 32 + rr][ccol]);
      if (m != -INFINITY) {
        float s = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) s += __expf(tile[chalf * 32 + rr][ccol] - m);
        ms_combine(cm, cs, m, s);
      }
    }
    __syncthreads();
  }
  if (chalf == 1) ccomb[ccol] = make_float2(cm, cs);
  __syncthreads();
  if (chalf == 0) {
    ms_combine(cm, cs, ccomb[ccol].x, ccomb[ccol].y);
    if (c0 + ccol < B) cp[static_cast<int64_t>(seg) * B + c0 + ccol] = make_float2(cm, cs);
  }
}

// Single block: combine the partials into lse vectors and the loss.
__global__ void ce_bidir_finalize_kernel(const float2* __restrict__ rp, int nstrips, const float2* __restrict__ cp, int nseg,
                                         const float* __restrict__ diag, int B, float* __restrict__ lse_row,
                                         float* __restrict__ lse_col, float* __restrict__ loss) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  __shared__ float sred[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float m = -INFINITY, s = 0.f;
    for (int k = 0; k < nstrips; ++k) { const float2 p = rp[static_cast<int64_t>(k) * B + i]; ms_combine(m, s, p.x, p.y); }
    const float lr = m + logf(s);
    m = -INFINITY; s = 0.f;
    for (int k = 0; k < nseg; ++k) { const float2 p = cp[static_cast<int64_t>(k) * B + i]; ms_combine(m, s, p.x, p.y); }
    const float lc = m + logf(s);
    lse_row[i] = lr;
    lse_col[i] = lc;
    acc += (lr - diag[i]) + (lc - diag[i]);
  }
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
  return static_cast<int64_t>(sizeof(float2)) * B * (ce_nstrips(B) + ce_nseg(B)) + sizeof(float) * B + 64;
}

int tic_ce_bidir_fwd(const float* S, int64_t lds, int B, float* lse_row, float* lse_col, float* loss, void* workspace,
                     void* stream) {
  TIC_CHECK_ARG(S && lse_row && lse_col && loss && workspace && B > 0 && lds >= B, "tic_ce_bidir_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ns = ce_nstrips(B), ng = ce_nseg(B);
  float2* rp = static_cast<float2*>(workspace);
  float2* cp = rp + static_cast<int64_t>(ns) * B;
  float* diag = reinterpret_cast<float*>(cp + static_cast<int64_t>(ng) * B);
  launch_k(ce_bidir_fwd_kernel, dim3(dim3(ns, ng)), dim3(256), 0, st, S, lds, B, rp, cp, diag);
  launch_k(ce_bidir_finalize_kernel, dim3(1), dim3(1024), 0, st, rp, ns, cp, ng, diag, B, lse_row, lse_col, loss);
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
