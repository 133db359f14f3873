// Image-text matching (ITM) negative sampling + pair gather — the device replacement of the host loop in
// models/mm_late.py:389-414 (prepare_itm_inputs): one coin + one pick per row, then row copies of ids / mask.
//
//  uniform mode (reference behaviour): integer arithmetic on the supplied uniforms, bit-exact by construction.
//  hard mode (extension, ALBEF-style; spec = oracle/restatement.py:itm_sample_hard): multinomial over
//     w_j = exp(S_ij - max_j S_ij), j != i, by inverse CDF on 2^30 fixed-point weights; integer prefix sums are
//     associative, and det_exp() uses only single IEEE-754 fp32 operations, so any reduction order reproduces the
//     oracle bit for bit.
#include "common.cuh"
#include "tic_ptx.cuh"
#include "itm_rule.cuh"

namespace tic {

__global__ void itm_sample_uniform_kernel(const float* __restrict__ u_coin, const float* __restrict__ u_pick, int B,
                                          int64_t* __restrict__ labels, int32_t* __restrict__ src_idx) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int l, s;
  uniform_rule(u_coin, u_pick, B, i, l, s);
  labels[i] = l;
  src_idx[i] = s;
}

// One 256-thread block per row; thread t owns the contiguous column chunk [t*chunk, (t+1)*chunk).  Materialised-S form of the
// hard-negative sampler (drop-in batch sizes, where logits_per_text exists anyway); the fused step uses the tile-stream
// form instead (tic_itc_fwd(qpart) -> tic_itm_hard_locate -> tic_itc_pick), which never needs S in memory.
__global__ void __launch_bounds__(256) itm_sample_hard_kernel(const float* __restrict__ u_coin, const float* __restrict__ u_pick,
                                                              int B, const float* __restrict__ S, int64_t lds, float ref,
                                                              const float* __restrict__ ref_dev,
                                                              int64_t* __restrict__ labels, int32_t* __restrict__ src_idx) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  if (ref_dev != nullptr) ref = __ldg(ref_dev);
  const int i = blockIdx.x, t = threadIdx.x;
  int label, src;
  uniform_rule(u_coin, u_pick, B, i, label, src);
  if (label == 1) {  // block-uniform
    if (t == 0) { labels[i] = 1; src_idx[i] = i; }
    return;
  }
  __shared__ unsigned long long ssum[256];
  __shared__ int sres;
  const float* row = S + static_cast<int64_t>(i) * lds;
  if (t == 0) sres = -1;
  const int chunk = (B + 255) / 256;
  const int j0 = t * chunk, j1 = min(j0 + chunk, B);
  unsigned long long local = 0;
  for (int j = j0; j < j1; ++j)
    if (j != i) local += hard_qweight(row[j], ref);
  ssum[t] = local;
  __syncthreads();
  // exclusive prefix of the 256 chunk sums (integer => order independent); 256 adds per thread is negligible
  unsigned long long excl = 0, total = 0;
  for (int k = 0; k < 256; ++k) {
    const unsigned long long s = ssum[k];
    if (k < t) excl += s;
    total += s;
  }
  if (total != 0) {
    const unsigned long long target = hard_target(u_pick[i], total);
    if (local != 0 && target >= excl && target < excl + local) {
      unsigned long long c = excl;
      for (int j = j0; j < j1; ++j) {
        if (j != i) c += hard_qweight(row[j], ref);
        if (c > target) { sres = j; break; }
      }
    }
  }
  __syncthreads();
  if (t == 0) {
    labels[i] = 0;
    src_idx[i] = (total != 0 && sres >= 0) ? sres : src;
  }
}

// Tile-stream form, middle step: the similarity tiles (tic_itc_fwd with qpart) left, for every row, the integer weight sum
// of each column part; one warp per row turns them into (part holding the target, residual target inside that part).
// Also writes the row's label and its default source (itself, or the uniform pick when every weight is zero);
// tic_itc_pick then overwrites src_idx of the located rows.
__global__ void __launch_bounds__(256) itm_hard_locate_kernel(const float* __restrict__ u_coin, const float* __restrict__ u_pick,
                                                              int m_local, int n_global, int row_offset,
                                                              const unsigned long long* __restrict__ qpart, int n_parts,
                                                              int64_t* __restrict__ labels, int32_t* __restrict__ src_idx,
                                                              int32_t* __restrict__ loc_part,
                                                              unsigned long long* __restrict__ loc_res) {
  pdl_trigger();
  pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= m_local) return;
  int label = 1, src = row_offset + r;
  if (n_global > 1 && u_coin[r] < 0.5f) {     // the uniform rule with this row's GLOBAL index (itm_rule.cuh: uniform_rule)
    label = 0;
    int k = static_cast<int>(floorf(__fmul_rn(u_pick[r], static_cast<float>(n_global - 1))));
    k = min(k, n_global - 2);
    src = k < row_offset + r ? k : k + 1;
  }
  int part = -1;
  unsigned long long res = 0;
  if (label == 0) {
    unsigned long long tot = 0;
    for (int p = lane; p < n_parts; p += 32) tot += qpart[static_cast<int64_t>(p) * m_local + r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (tot != 0) {
      const unsigned long long target = hard_target(u_pick[r], tot);
      unsigned long long run = 0;      // weights of the parts before the current group of 32
      for (int p0 = 0; p0 < n_parts && part < 0; p0 += 32) {
        const int p = p0 + lane;
        const unsigned long long q = p < n_parts ? qpart[static_cast<int64_t>(p) * m_local + r] : 0ull;
        unsigned long long inc = q;    // inclusive scan over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += up;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, run + inc > target);
        if (hit != 0) {
          const int l = __ffs(hit) - 1;
          part = p0 + l;
          const unsigned long long before = run + __shfl_sync(0xffffffffu, inc - q, l);
          res = target - before;
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
  }
  if (lane == 0) {
    labels[r] = label;
    src_idx[r] = src;
    loc_part[r] = part;
    loc_res[r] = res;
  }
}

// dst[r,:] = src[idx[r],:] for row_bytes bytes; one warp per row, 16-byte words when everything is aligned.
// blockIdx.y selects one of up to two (src, dst) pairs so ids and mask move in one launch.
// When `u_coin` is given the uniform sampling rule is evaluated in-kernel (single-launch sample + gather).
__global__ void gather_rows_kernel(const uint8_t* __restrict__ s0, uint8_t* __restrict__ d0, const uint8_t* __restrict__ s1,
                                   uint8_t* __restrict__ d1, int64_t spitch, int64_t dpitch, int64_t row_bytes,
                                   const int32_t* __restrict__ idx, int rows, const float* __restrict__ u_coin,
                                   const float* __restrict__ u_pick, int64_t* __restrict__ labels,
                                   int32_t* __restrict__ src_out) {
  pdl_trigger();   // let the next kernel of the chain start its prologue (programmatic dependent launch)
  pdl_wait();      // ... while this one waits here for its own predecessors' writes
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  int sr;
  if (u_coin != nullptr) {
    int l;
    uniform_rule(u_coin, u_pick, rows, r, l, sr);
    if (blockIdx.y == 0 && lane == 0) { labels[r] = l; src_out[r] = sr; }
  } else {
    sr = idx[r];
  }
  const uint8_t* s = (blockIdx.y == 0 ? s0 : s1) + static_cast<int64_t>(sr) * spitch;
  uint8_t* d = (blockIdx.y == 0 ? d0 : d1) + static_cast<int64_t>(r) * dpitch;
  const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | static_cast<uintptr_t>(row_bytes)) & 15) == 0;
  if (vec) {
    for (int64_t b = lane * 16; b < row_bytes; b += 512) *reinterpret_cast<uint4*>(d + b) = __ldg(reinterpret_cast<const uint4*>(s + b));
  } else {
    for (int64_t b = lane; b < row_bytes; b += 32) d[b] = s[b];
  }
}

}  // namespace tic

using namespace tic;

extern "C" {

int tic_itm_sample(const float* u_coin, const float* u_pick, int B, int mode, const float* S, int64_t lds, float hard_ref,
                   const float* hard_ref_dev, int64_t* labels, int32_t* src_idx, void* stream) {
  TIC_CHECK_ARG(u_coin && u_pick && labels && src_idx && B > 0, "tic_itm_sample: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == TIC_ITM_UNIFORM) {
    launch_k(itm_sample_uniform_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, st, u_coin, u_pick, B, labels, src_idx);
  } else if (mode == TIC_ITM_HARD) {
    TIC_CHECK_ARG(S != nullptr && lds >= B, "tic_itm_sample: hard mode needs the similarity matrix");
    launch_k(itm_sample_hard_kernel, dim3(B), dim3(256), 0, st, u_coin, u_pick, B, S, lds, hard_ref, hard_ref_dev, labels, src_idx);
  } else {
    set_error("tic_itm_sample: unknown mode %d", mode);
    return TIC_E_ARG;
  }
  TIC_CHECK_LAUNCH("tic_itm_sample");
  return TIC_OK;
}

int tic_gather_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes, int64_t row_bytes,
                    const int32_t* src_idx, int rows, void* stream) {
  TIC_CHECK_ARG(src && dst && src_idx && rows > 0 && row_bytes > 0, "tic_gather_rows: bad arguments");
  dim3 grid(ceil_div(rows, 8), 1);
  launch_k(gather_rows_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), nullptr, nullptr, src_pitch_bytes, dst_pitch_bytes,
      row_bytes, src_idx, rows, nullptr, nullptr, nullptr, nullptr);
  TIC_CHECK_LAUNCH("tic_gather_rows");
  return TIC_OK;
}

int tic_itm_sample_gather(const float* u_coin, const float* u_pick, int B, int mode, const float* S, int64_t lds,
                          float hard_ref, const float* hard_ref_dev, const void* ids, const void* mask, int64_t row_bytes, void* tim_ids, void* tim_mask,
                          int64_t* labels, int32_t* src_idx, void* stream) {
  TIC_CHECK_ARG(u_coin && u_pick && ids && mask && tim_ids && tim_mask && labels && src_idx && B > 0 && row_bytes > 0,
                "tic_itm_sample_gather: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(ceil_div(B, 8), 2);
  if (mode == TIC_ITM_UNIFORM) {
    launch_k(gather_rows_kernel, dim3(grid), dim3(256), 0, st, static_cast<const uint8_t*>(ids), static_cast<uint8_t*>(tim_ids),
                                             static_cast<const uint8_t*>(mask), static_cast<uint8_t*>(tim_mask), row_bytes,
                                             row_bytes, row_bytes, nullptr, B, u_coin, u_pick, labels, src_idx);
  } else {
    int rc = tic_itm_sample(u_coin, u_pick, B, mode, S, lds, hard_ref, hard_ref_dev, labels, src_idx, stream);
    if (rc) return rc;
    launch_k(gather_rows_kernel, dim3(grid), dim3(256), 0, st, static_cast<const uint8_t*>(ids), static_cast<uint8_t*>(tim_ids),
                                             static_cast<const uint8_t*>(mask), static_cast<uint8_t*>(tim_mask), row_bytes,
                                             row_bytes, row_bytes, src_idx, B, nullptr, nullptr, nullptr, nullptr);
  }
  TIC_CHECK_LAUNCH("tic_itm_sample_gather");
  return TIC_OK;
}

int tic_itm_hard_locate(const float* u_coin, const float* u_pick, int m_local, int n_global, int row_offset, const void* qpart,
                        int n_parts, int64_t* labels, int32_t* src_idx, int32_t* loc_part, void* loc_res, void* stream) {
  TIC_CHECK_ARG(u_coin && u_pick && qpart && labels && src_idx && loc_part && loc_res && m_local > 0 && n_parts > 0 &&
                    row_offset >= 0 && row_offset + m_local <= n_global,
                "tic_itm_hard_locate: bad arguments");
  launch_k(itm_hard_locate_kernel, dim3(ceil_div(m_local, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), u_coin, u_pick,
           m_local, n_global, row_offset, static_cast<const unsigned long long*>(qpart), n_parts, labels, src_idx, loc_part,
           static_cast<unsigned long long*>(loc_res));
  TIC_CHECK_LAUNCH("tic_itm_hard_locate");
  return TIC_OK;
}

}  // extern "C"
