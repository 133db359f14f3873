// Evaluation bookkeeping on the device — SURVEY.md §8 row f-4.
//
// The reference's eval loop (models/mm_late.py:594-612) synchronises once per batch (`loss.item()`, `.cpu().numpy()`), grows
// Python lists of 0-d tensors (`predictions += pred`) and hands them to six torchmetrics objects
// (models/utils.py:294-325).  Here one launch per batch appends predictions / targets to preallocated device arrays and
// updates a C x C confusion matrix, the per-batch accuracy sum and the loss sum; one launch at the end of the epoch turns the
// confusion matrix into the six scores.  Nothing synchronises until the caller reads the results.
//
// Integer work (argmax, counts) is exact.  Both kernels are HBM/latency-trivial: B*(C*4 + 16) bytes per batch.
#include "common.cuh"

namespace tic {

constexpr int EVAL_MAX_C_SMEM = 16;   // confusion matrix kept in shared memory up to 16 classes

// pred = first index of the row maximum (torch.argmax's tie rule; softmax is monotone, mm_late.py:597-600),
// target = argmax of the float one-hot label row (mm_late.py:601) or the given integer label.
__global__ void __launch_bounds__(256) eval_accumulate_kernel(const float* __restrict__ logits, int64_t ldl,
                                                              const float* __restrict__ y_soft, int64_t ldy,
                                                              const int64_t* __restrict__ y_int, int B, int C,
                                                              const float* __restrict__ batch_loss, int64_t* __restrict__ preds,
                                                              int64_t* __restrict__ targets, unsigned long long* __restrict__ conf,
                                                              unsigned long long* __restrict__ counts, float* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  __shared__ unsigned int hist[EVAL_MAX_C_SMEM * EVAL_MAX_C_SMEM];
  __shared__ unsigned int correct_s;
  const bool use_smem = C <= EVAL_MAX_C_SMEM;
  if (use_smem)
    for (int k = threadIdx.x; k < C * C; k += blockDim.x) hist[k] = 0;
  if (threadIdx.x == 0) correct_s = 0;
  __syncthreads();
  unsigned int my_correct = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const float* row = logits + static_cast<int64_t>(i) * ldl;
    int p = 0;
    float best = row[0];
    for (int c = 1; c < C; ++c) {
      const float v = row[c];
      if (v > best) { best = v; p = c; }     // strict: the first maximum wins; NaN never wins (torch: NaN wins — not produced here)
    }
    int t;
    if (y_int != nullptr) {
      t = static_cast<int>(y_int[i]);
    } else {
      const float* yr = y_soft + static_cast<int64_t>(i) * ldy;
      t = 0;
      float yb = yr[0];
      for (int c = 1; c < C; ++c) {
        const float v = yr[c];
        if (v > yb) { yb = v; t = c; }
      }
    }
    preds[i] = p;
    targets[i] = t;
    my_correct += (p == t);
    if (t >= 0 && t < C) {
      if (use_smem) atomicAdd(&hist[t * C + p], 1u);
      else atomicAdd(&conf[static_cast<int64_t>(t) * C + p], 1ull);
    }
  }
  // block reduction of the correct count: warp shuffle, then one shared atomic per warp
  for (int o = 16; o > 0; o >>= 1) my_correct += __shfl_xor_sync(0xffffffffu, my_correct, o);
  if ((threadIdx.x & 31) == 0 && my_correct) atomicAdd(&correct_s, my_correct);
  __syncthreads();
  if (use_smem)
    for (int k = threadIdx.x; k < C * C; k += blockDim.x)
      if (hist[k]) atomicAdd(&conf[k], static_cast<unsigned long long>(hist[k]));
  if (threadIdx.x == 0) {
    if (correct_s) atomicAdd(&counts[0], static_cast<unsigned long long>(correct_s));
    if (blockIdx.x == 0) {
      atomicAdd(&counts[1], static_cast<unsigned long long>(B));
      atomicAdd(&counts[2], 1ull);                                 // batches seen
      if (batch_loss != nullptr) atomicAdd(&sums[0], batch_loss[0]);   // mean of per-batch losses (mm_late.py:594,615)
    }
  }
}

// The reference averages PER-BATCH accuracies (mm_late.py:607-608,616), so a short last batch weighs as much as a full one:
// sums[1] += 100 * correct_in_this_batch / B.  Needs this batch's own correct count, hence a second tiny kernel fed by a
// per-batch counter (counts[3], reset here).
__global__ void eval_batch_acc_kernel(unsigned long long* __restrict__ counts, float* __restrict__ sums, int B) {
  pdl_trigger();
  pdl_wait();
  const unsigned long long total = counts[0];
  const unsigned long long before = counts[3];
  sums[1] += 100.0f * static_cast<float>(total - before) / static_cast<float>(B);
  counts[3] = total;
}

// One warp; lane c owns class c (strided when C > 32).  torchmetrics 0.11 multiclass semantics (functional/classification/
// precision_recall.py, f_beta.py: _safe_divide -> 0 where the denominator is 0; _adjust_weights_safe_divide: `weighted`
// weighs by support tp+fn, `macro` weighs 1 per class except classes with tp+fp+fn == 0, which are dropped).
__global__ void metrics_from_confusion_kernel(const unsigned long long* __restrict__ conf, int C, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // f1_w, f1_m, p_w, p_m, r_w, r_m, sum of supports, number of live classes
  for (int c = lane; c < C; c += 32) {
    unsigned long long tp = conf[static_cast<int64_t>(c) * C + c], row = 0, col = 0;
    for (int k = 0; k < C; ++k) {
      row += conf[static_cast<int64_t>(c) * C + k];   // support of class c (targets == c)
      col += conf[static_cast<int64_t>(k) * C + c];   // predictions == c
    }
    const unsigned long long fn = row - tp, fp = col - tp;
    const double prec = (tp + fp) ? static_cast<double>(tp) / static_cast<double>(tp + fp) : 0.0;
    const double rec = (tp + fn) ? static_cast<double>(tp) / static_cast<double>(tp + fn) : 0.0;
    const double f1 = (2 * tp + fp + fn) ? 2.0 * static_cast<double>(tp) / static_cast<double>(2 * tp + fp + fn) : 0.0;
    const double wsup = static_cast<double>(row);
    const double live = (tp + fp + fn) ? 1.0 : 0.0;
    acc[0] += wsup * f1;  acc[1] += live * f1;
    acc[2] += wsup * prec; acc[3] += live * prec;
    acc[4] += wsup * rec; acc[5] += live * rec;
    acc[6] += wsup;       acc[7] += live;
  }
  for (int k = 0; k < 8; ++k)
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (lane == 0) {
    const double sw = acc[6], sl = acc[7];
    out[0] = static_cast<float>(sw > 0 ? acc[0] / sw : 0.0);
    out[1] = static_cast<float>(sl > 0 ? acc[1] / sl : 0.0);
    out[2] = static_cast<float>(sw > 0 ? acc[2] / sw : 0.0);
    out[3] = static_cast<float>(sl > 0 ? acc[3] / sl : 0.0);
    out[4] = static_cast<float>(sw > 0 ? acc[4] / sw : 0.0);
    out[5] = static_cast<float>(sl > 0 ? acc[5] / sl : 0.0);
  }
}

}  // namespace tic

using namespace tic;

extern "C" {

int tic_eval_state_words(int C) { return C > 0 ? C * C + 4 : -1; }

int tic_eval_accumulate(const float* logits, int64_t ldl, const float* y_soft, int64_t ldy, const int64_t* y_int, int B, int C,
                        const float* batch_loss, int64_t* preds_out, int64_t* targets_out, void* state, float* sums,
                        void* stream) {
  TIC_CHECK_ARG(logits && preds_out && targets_out && state && sums, "tic_eval_accumulate: null pointer");
  TIC_CHECK_ARG((y_soft != nullptr) != (y_int != nullptr), "tic_eval_accumulate: exactly one of y_soft / y_int must be given");
  TIC_CHECK_ARG(B > 0 && C > 0 && C <= 4096 && ldl >= C && (y_soft == nullptr || ldy >= C), "tic_eval_accumulate: bad shape B=%d C=%d", B, C);
  unsigned long long* conf = static_cast<unsigned long long*>(state);
  unsigned long long* counts = conf + static_cast<int64_t>(C) * C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = min(ceil_div(B, 256), 148 * 4);
  launch_k(eval_accumulate_kernel, dim3(grid), dim3(256), 0, st, logits, ldl, y_soft, ldy, y_int, B, C, batch_loss, preds_out,
           targets_out, conf, counts, sums);
  TIC_CHECK_LAUNCH("tic_eval_accumulate");
  launch_k(eval_batch_acc_kernel, dim3(1), dim3(1), 0, st, counts, sums, B);
  TIC_CHECK_LAUNCH("tic_eval_accumulate (batch accuracy)");
  return TIC_OK;
}

int tic_metrics_from_confusion(const void* state, int C, float* out6, void* stream) {
  TIC_CHECK_ARG(state && out6 && C > 0, "tic_metrics_from_confusion: bad arguments");
  launch_k(metrics_from_confusion_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream),
           static_cast<const unsigned long long*>(state), C, out6);
  TIC_CHECK_LAUNCH("tic_metrics_from_confusion");
  return TIC_OK;
}

}  // extern "C"
