/*
 * tic_b200.h — C ABI of libtic_b200.so: the B200 (sm_100a) implementation of the late-fusion head and
 * image-text auxiliary-loss path of danaesavi/SocialMedia-TextImage-Classification-AuxLosses.
 *
 * The reference has no FFI: its boundary for this path is the Python surface of models/mm_late.py,
 * models/utils.py:225-231 (clip_loss) and models/run_mm_late.py (flags).  Each entry point below names the
 * reference lines it replaces; the Python mirror (package `tic_b200`) binds them with ctypes, and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; the caller owns every buffer,
 *     outputs and workspaces are caller-allocated (sizes documented per function);
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     (so calls are CUDA-graph capturable) and the library keeps no global state besides an error string;
 *   - return value: 0 = OK, negative = error (see TIC_E_*); functions never throw across the ABI;
 *   - thread-safety: re-entrant per stream; the error string is thread-local.
 */
#ifndef TIC_B200_H
#define TIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIC_OK 0
#define TIC_E_ARG (-1)        /* bad shape / alignment / null pointer */
#define TIC_E_TMAP (-2)       /* cuTensorMapEncodeTiled failed */
#define TIC_E_ATTR (-3)       /* cudaFuncSetAttribute failed */
#define TIC_E_LAUNCH (-4)     /* kernel launch failed */
#define TIC_E_RANGE (-5)      /* value outside the supported numeric range */
#define TIC_E_CUDA (-6)       /* other CUDA runtime error */

#define TIC_ITM_UNIFORM 0 /* reference behaviour, mm_late.py:389-414 */
#define TIC_ITM_HARD 1    /* similarity-weighted extension (not in the reference) */

const char* tic_last_error_string(void);
int tic_version(void);
int tic_sm_count(void);
/* Programmatic dependent launch for the launches this THREAD makes next: 1 = every kernel is launched with the
 * programmatic-stream-serialization attribute (its prologue may overlap the tail of its predecessor in the stream; all
 * kernels of the library call griddepcontrol.launch_dependents / .wait), 0 = ordinary stream order, -1 = follow the TIC_PDL
 * environment variable (default off).  Returns the previous mode.  Lets a caller turn it on for the kernels of a
 * latency-critical chain only: on the multi-branch small-batch step, early-launched CTAs of side-branch kernels (one CTA
 * per SM) otherwise take SMs from the chains. */
int tic_set_pdl(int mode);

/* ------------------------------------------------------------------------------------------------ GEMM
 * D[m,n] = alpha * sum_k A(m,k) * B(n,k) (+ bias[n]) (relu)   bf16 operands, fp32 accumulate (tcgen05 / TMEM).
 *   A(m,k) = a_mn_major ? A[k*lda + m] : A[m*lda + k];   B(n,k) = b_mn_major ? B[k*ldb + n] : B[n*ldb + k].
 * Replaces the ATen/cuBLAS GEMM call sites of the head: nn.Linear forward/backward for text_projection,
 * visual_projection (HF VisionTextDualEncoderModel.forward, built at mm_late.py:59-61), linear_fusion
 * (mm_late.py:81,95,112,143), fc_Q/fc_K/fc_V collapsed products (mm_late.py:105), linear_gmu_* (mm_late.py:88-89).
 * lda/ldb must be multiples of 8 elements and A/B 16-byte aligned (TMA). d_dtype: 0 = fp32, 1 = bf16. */
int tic_gemm_bf16(const void* A, const void* A_lo, int64_t lda, int a_mn_major, const void* B, const void* B_lo, int64_t ldb,
                  int b_mn_major, void* D, void* D_lo, int64_t ldd, int d_dtype, int M, int N, int K, float alpha,
                  const float* bias, int relu, int accumulate, void* stream);
/* Split precision: an operand that was produced on the device (a gradient or an intermediate activation) may be
 * passed as a bf16 (hi, lo) pair with identical layout (A_lo / B_lo, NULL = plain bf16); the kernel then runs the extra
 * K-segment(s) D += A_lo*B (+ A*B_lo) into the same TMEM accumulator, so the rounding of that operand drops from 2^-9 to
 * ~2^-17.  D_lo (bf16 output only) receives the residual x - bf16(x) so the result can itself be consumed as a pair.
 * accumulate = 1 (fp32 output, no ReLU): D += result with fp32 atomics (D must hold its initial value, e.g. zeros); this
 * also lets the library split K across CTAs so that long-K weight-gradient GEMMs with few output tiles fill all SMs. */
/* Same GEMM (no accumulate) that additionally writes row_ss_part[t][m] = sum over the columns of 64-wide column tile t of
 * D[m, :]^2 (fp32, before any rounding to bf16), t < tic_gemm_rowss_parts(N): the L2-norm statistics of the projected
 * embeddings (HF :268-269) for tic_itc_fwd, so the normalisation needs no kernel of its own on the small-batch path. */
int tic_gemm_rowss_parts(int N);
int tic_gemm_bf16_rowss(const void* A, const void* A_lo, int64_t lda, int a_mn_major, const void* B, const void* B_lo,
                        int64_t ldb, int b_mn_major, void* D, void* D_lo, int64_t ldd, int d_dtype, int M, int N, int K,
                        float alpha, const float* bias, int relu, float* row_ss_part, void* stream);
/* The launch shape tic_gemm_bf16 picks for a problem (no launch): tile width, fp32-atomic split-K factor (accumulate only) and
 * CLUSTER split-K factor (1, 2 or 4): the CTAs of a thread-block cluster each take a K-slice of the same output tile and the
 * partial tiles are folded through distributed shared memory before the ordinary epilogue — used for the small problems of
 * the latency-bound small-batch step, where a CTA's k-loop rate, not the machine, bounds the kernel.  n_split_operands =
 * number of operands passed as (hi, lo) pairs.  Cluster split-K is opt-in: TIC_CLUSTER_K=<CTA budget> enables it (measured: its
 * fixed cost only pays from K >= ~2560 on, beyond the GEMMs of this path; see profiles/r01_cluster_splitk.md). */
int tic_gemm_plan(int M, int N, int K, int n_split_operands, int accumulate, int* tile_n, int* ksplit, int* cluster_k);
/* TEST-ONLY: reference-quality SIMT fp32-accumulate GEMM with the same semantics.  Called by csrc/selftest.cu and the GPU
 * tests as the in-library cross-check of the tcgen05 path; no product code path calls it. */
int tic_gemm_bf16_simt(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* D,
                       int64_t ldd, int d_dtype, int M, int N, int K, float alpha, const float* bias, int relu,
                       void* stream);

/* ------------------------------------------------------------------------------------------------ L2 norm
 * rinv[i] = 1 / ||X[i,:]||_2 (no epsilon, as HF modeling_vision_text_dual_encoder.py:268-269 and
 * mm_early.py:98-99).  X bf16 [rows, cols]. */
int tic_row_rnorm_bf16(const void* X, const void* X_lo /* optional residual: norm of hi+lo */, int64_t ldx, int rows, int cols,
                       float* rinv, void* Xhat /* optional normalised bf16 copy [rows, ldh] */, int64_t ldh, void* stream);

/* ------------------------------------------------------------------------------------------------ ITC (fused)
 * Logits S[i,j] = scale * rinv_t[i] * rinv_v[j] * <T[i,:], V[j,:]>  (HF :272-273; mm_early.py:101-102), never
 * materialised unless `logits_out` is given.  T is this rank's row block [m_local, P]; V is the gathered
 * [n_global, P]; `row_offset` is the global column index of local row 0 (diagonal = positives).
 *
 * tic_itc_fwd writes per-tile partial sums of exp(S - shift):
 *     row_part [tic_itc_row_parts(n_global)][m_local],  col_part [tic_itc_col_parts(m_local)][n_global]
 * T_lo / V_lo (optional) are bf16 residuals of embeddings produced on the device (the projection GEMM): the tiles then run
 * the two extra K-segments T_lo*V and T*V_lo.
 * and diag[i] = S[i, row_offset+i].  `shift` must satisfy shift >= max S; scale (=exp(logit_scale)) works since
 * |cos| <= 1; scale > 40 is rejected with TIC_E_RANGE (fp32 underflow of exp(-2*scale)). */
int tic_itc_row_parts(int n_global);
int tic_itc_col_parts(int m_local);
int tic_itc_fwd(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv,
                float* rinv_t, float* rinv_v,
                int m_local, int n_global, int P, int row_offset, float scale, float shift, float* row_part,
                float* col_part /* NULL: skip the column statistics (symmetric multi-GPU mode) */, float* diag,
                float* logits_out, int64_t ld_logits, const float* ss_t_part, int n_ss_t, const float* ss_v_part, int n_ss_v,
                const uint32_t* seg_ready, const uint32_t* seg_epoch, int seg_cols, int my_seg,
                const float* scale_dev /* optional DEVICE scalar exp(logit_scale): overrides scale and shift */,
                void* qpart /* optional uint64 [tic_itc_q_parts(n_global)][m_local]: hard-negative weight sums, see below */,
                void* stream);
/* Trainable temperature: logit_scale is a parameter of the reference model (mm_late.py:59-69 keeps it trainable; HF :272
 * applies exp()).  Every ITC entry point takes `scale_dev`, a DEVICE pointer to exp(logit_scale) (written by
 * tic_refresh_weights at the head of the step); when it is non-NULL it replaces the host `scale` / `shift` arguments, so
 * a step captured into a CUDA graph follows the optimiser's updates of logit_scale without re-capture.
 * Hard-negative sampling without a materialised S (a-5', tile-stream form; spec oracle/restatement.py:itm_sample_hard with
 * ref = shift): with `qpart` the forward tiles also write, per row and per column part (part p = columns [32p, 32p+32);
 * tic_itc_q_parts(n_global) parts), the integer sum of q_ij = trunc(det_exp(min(S_ij - shift, 0)) * 2^40), j != positive.
 * tic_itm_hard_locate turns these sums and the row's uniform into (part, residual target); tic_itc_pick recomputes the
 * tiles (bit-identical accumulators: same operands, shapes and k order) and walks the located part:
 * src = first column whose running weight exceeds the residual.  Traffic: 8 bytes per 32 logits instead of 128. */
int tic_itc_q_parts(int n_global);
int tic_itc_pick(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv, const float* rinv_t,
                 const float* rinv_v, int m_local, int n_global, int P, int row_offset, float scale, float shift,
                 const float* scale_dev, const int32_t* loc_part, const void* loc_res /* uint64 [m_local] */,
                 int32_t* src_idx /* only located rows are written */, void* stream);
/* Consuming a gathered V as it lands (multi-GPU): seg_ready (NULL = V is complete) points at n_global/seg_cols device
 * words; segment p (columns [p*seg_cols, (p+1)*seg_cols), written by a tic_peer_pull running beside this kernel) may be
 * read once seg_ready[p] >= *seg_epoch.  Tiles are visited segment-major starting with my_seg (always ready); if seg_cols is
 * not a multiple of the tile width the kernel waits for every segment before its first load instead. */
/* Fused normalisation (small batches): when ss_t_part [n_ss_t][m_local] / ss_v_part [n_ss_v][n_global] — the per-tile row
 * sums of squares written by tic_gemm_bf16_rowss while it projected the embeddings — are given, the tiles derive
 * rinv = 1/sqrt(sum of partials) themselves and WRITE rinv_t / rinv_v for the kernels that follow; no norm kernel runs. */
/* out[j] = sum_p part[p][j]  (deterministic fixed-order reduction of the partials above). */
int tic_reduce_parts(const float* part, int nparts, int n, float* out, void* stream);
/* Sums the partials in a fixed order (deterministic), then lse = shift + log(sum);
 *   loss_sums[0] += sum_i (lse_row[i] - diag[i]),  loss_sums[1] += sum_i (lse_col[row_offset+i] - diag[i])
 * utils.py:225-231: clip_loss = (loss_sums[0]/B + loss_sums[1]/B) / 2 with B = n_global.  Multi-GPU: reduce the column
 * partials with tic_reduce_parts, all-reduce(SUM) the result across ranks, then call this with n_col_parts = 1. */
int tic_itc_lse_loss(const float* row_part, int n_row_parts, const float* col_part, int n_col_parts, const float* diag,
                     int m_local, int n_global, int row_offset, float shift, float* lse_row, float* lse_col,
                     float* loss_sums, const float* scale_dev /* optional: shift = *scale_dev */, void* stream);
/* Symmetric (peer-memory) multi-GPU mode: rank r runs tic_itc_fwd twice — on its row block S[rows_r, :] and, with the
 * operands swapped, on S^T[cols_r, :] (col_part = NULL in both) — so BOTH softmax directions are complete row statistics on
 * the rank that owns them and no cross-rank reduction exists.  This call turns the two sets of row partials into
 *   lse_a[i] (= lse_row of my text rows), lse_b[i] (= lse_col of my image columns),
 *   loss_sums[0] += sum_i (lse_a[i]-diag[i]),  loss_sums[1] += sum_i (lse_b[i]-diag[i])   (fixed-order, deterministic).
 * workspace: tic_itc_lse_rows_workspace_bytes(m) bytes, zero-initialised once (self-resetting). utils.py:225-231. */
int64_t tic_itc_lse_rows_workspace_bytes(int m);
int tic_itc_lse_rows(const float* part_a, const float* part_b, int n_parts, int m, const float* diag, float shift, float* lse_a,
                     float* lse_b, float* loss_sums, void* workspace, const float* scale_dev, void* stream);
/* Same, and the PUSH of the two vectors rides on the launch (see tic_peer_push): lse_a / lse_b are this rank's slots of its
 * own gathered vectors; every value is also stored into the remote ranks' blocks at off_a / off_b + (rank * m + i) * 4, and
 * the last block releases flags[rank] = step[1] at flag_off of every remote block.  bases_host: HOST array of world pointers. */
int tic_itc_lse_rows_push(const float* part_a, const float* part_b, int n_parts, int m, const float* diag, float shift,
                          float* lse_a, float* lse_b, float* loss_sums, void* workspace, const float* scale_dev,
                          void* const* bases_host, int world, int rank, int64_t off_a, int64_t off_b, int64_t flag_off,
                          const uint32_t* step, void* stream);
/* Recompute S tiles and emit the bf16 gradient operands (g = dLoss/d(clip_loss), B = n_global):
 *   Gp[i,j] = g/(2B) * (exp(S-lse_row[i]) + exp(S-lse_col[j]))      (the -I/B diagonal is applied in fp32 later)
 *   GA [m_local, ld_ga ] row-major:  Gp[i,j] * rinv_v[j]            (A operand of dT = GA * V)
 *   GB ("GBT" argument) [m_local, ld_ga] row-major: Gp[i,j] * rinv_t[i] at [i,j] — the SAME orientation and leading dimension
 *       as GA; dV = GB^T * T reads it MN-major (a_mn_major = 1), so no transposed operand is ever written (ld_gbt is ignored)
 * GBT may be NULL ("GA-shared" mode, used from 4096 columns on): the image-side gradient is then computed as
 *   dV_acc'[j,:] = sum_i GA[i,j] * That[i,:]  (GA read MN-major, That = normalised bf16 text embeddings from
 *   tic_row_rnorm_bf16), which equals rinv_v[j] * dV_acc[j,:]; tic_itc_grad_finalize(acc_div_rinv=1) divides it out. */
int tic_itc_bwd_g(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv,
                  const float* rinv_t, const float* rinv_v,
                  const float* lse_row, const float* lse_col, int m_local, int n_global, int P, float scale,
                  float gscale, void* GA, int64_t ld_ga, void* GBT, int64_t ld_gbt, void* GA_lo /* optional residuals */,
                  void* GBT_lo, const float* row_part, int n_row_parts, const float* col_part, int n_col_parts, float shift,
                  const float* scale_dev, const uint32_t* seg_ready, const uint32_t* seg_epoch, int seg_cols, int my_seg,
                  void* stream);
/* seg_ready (optional, as in tic_itc_fwd): the gathered lse_col / rinv_v vectors arrive segment by segment (tic_peer_push);
 * the tiles of a column segment start once its flag has reached *seg_epoch.  In both kernels a NEGATIVE seg_cols marks flags
 * that are written by the peers themselves (push form): the launch then keeps the whole machine (no pull kernel beside it). */
/* Fused forward + backward tiles for the small-batch step (at most 8 tiles of 128 x 64: tic_itc_fused_small_ok): ONE launch
 * = tic_itc_fwd (same outputs: partials, diag, inverse norms, optional logits / qpart) followed, after a cluster-wide
 * barrier, by tic_itc_bwd_g with the inline statistics — the S tiles stay in TMEM in between, the second k-loop and one
 * launch leave the critical chain of the step (HF :268-273 + utils.py:225-231 forward, and the gradient operands). */
int tic_itc_fused_small_ok(int m_local, int n_global);
/* Measurement switch: register a device buffer (uint64[16 * CTAs], NULL = off) in which tic_itc_fwd_bwd_small stamps
 * %globaltimer at its phase boundaries (scripts/kernel_times.py). */
int tic_debug_set_trace(void* device_u64_buffer);
int tic_itc_fwd_bwd_small(const void* T, const void* T_lo, int64_t ldt, const void* V, const void* V_lo, int64_t ldv,
                          float* rinv_t, float* rinv_v, int m_local, int n_global, int P, int row_offset, float scale,
                          float* row_part, float* col_part, float* diag, float* logits_out, int64_t ld_logits,
                          const float* ss_t_part, int n_ss_t, const float* ss_v_part, int n_ss_v, const float* scale_dev,
                          void* qpart, float gscale, void* GA, int64_t ld_ga, void* GBT, int64_t ld_gbt, void* GA_lo,
                          void* GBT_lo, void* stream);
/* Inline statistics (small batches): when row_part / col_part (the partials written by tic_itc_fwd) are given, the kernel
 * derives lse_row / lse_col = shift + log(sum of partials) itself (same expression as tic_itc_lse_loss) and the matching
 * lse pointer may be NULL — tic_itc_lse_loss then only produces the loss and runs beside the backward, not before it. */
/* Autograd path (materialised logits, drop-in batch sizes): the same operands from an upstream dL/dS [m_local, n_global]:
 *   GA[i,j] = dS[i,j]*rinv_v[j],  GB[i,j] = dS[i,j]*rinv_t[i]  (+ optional bf16 residuals), layouts as above. Use diag_coef = 0 afterwards. */
int tic_itc_ds_operands(const float* dS, int64_t ldds, int m_local, int n_global, const float* rinv_t, const float* rinv_v,
                        void* GA, void* GA_lo, int64_t ld_ga, void* GBT, void* GBT_lo, int64_t ld_gbt, void* stream);
/* Normalise-backward + diagonal term, one warp per row (HF :268-269 backward):
 *   dxh = scale*acc[i,:] - diag_coef * scale * rinv_o[i] * Xo[i,:]   (diag_coef = g/B, 0 if row has no local positive)
 *   r = <xh, dxh>, xh = rinv[i]*X[i,:];   dX[i,:] = rinv[i] * (dxh - xh * r);   dscale_part[block] += r (for dlogit_scale)
 * acc fp32 [rows, P] is the raw GEMM output; X the embedding being differentiated, Xo the other modality's
 * row with the same global index (may be NULL when diag_coef == 0). dX written as fp32 and/or bf16 (either may be NULL). */
int tic_itc_grad_finalize(const float* acc, int64_t ld_acc, const void* X, const void* X_lo, int64_t ldx, const float* rinv,
                          const void* Xo, const void* Xo_lo, int64_t ldxo, const float* rinv_o, int rows, int P, float scale,
                          float diag_coef, float* dX_f32, int64_t ld_df, void* dX_bf16, void* dX_bf16_lo, int64_t ld_db,
                          float* r_sum /* [1], atomically accumulated */,
                          int acc_div_rinv /* 1: acc rows carry an extra factor rinv[row] (GA-shared mode, see tic_itc_bwd_g) */,
                          const float* scale_dev, void* stream);

/* ------------------------------------------------------------------------------------------------ clip_loss on a given matrix
 * utils.py:225-231 on a materialised similarity S [B,B] fp32: loss = (CE(S, I) + CE(S^T, I)) / 2.
 * fwd reads S once; lse_row/lse_col [B] are saved for backward; bwd writes dS = g*(softmax_row+softmax_col-2I)/(2B).
 * workspace: tic_ce_bidir_workspace_bytes(B). */
int64_t tic_ce_bidir_workspace_bytes(int B);
int tic_ce_bidir_fwd(const float* S, int64_t lds, int B, float* lse_row, float* lse_col, float* loss, void* workspace,
                     void* stream);
int tic_ce_bidir_bwd(const float* S, int64_t lds, int B, const float* lse_row, const float* lse_col,
                     const float* grad_loss /* device scalar */, float* dS, int64_t ldds, void* stream);

/* ------------------------------------------------------------------------------------------------ ITM sampling + gather
 * mm_late.py:389-414 (prepare_itm_inputs).  Row i is swapped iff u_coin[i] < 0.5 (label 0), else kept (label 1);
 * B == 1 keeps everything.  Uniform mode: k = min(floor(u_pick[i]*(B-1)), B-2), src = k < i ? k : k+1.
 * Hard mode (extension named by BASELINE.json only): src ~ Multinomial(w), w[j] = exp(S[i,j] - ref) for j != i with a FIXED
 * reference ref >= max S (hard_ref, or *hard_ref_dev when given: the step passes exp(logit_scale)), by inverse CDF on
 * fixed-point (2^40) weights with a bit-reproducible exp (oracle/restatement.py: det_exp_f32, itm_sample_hard).  This
 * entry point is the materialised-S form (S = logits_per_text of the drop-in API); the fused step uses the tile-stream
 * form (tic_itc_fwd(qpart) -> tic_itm_hard_locate -> tic_itc_pick), which computes the same indices without S in memory.
 * Gathers nrowsets row-major byte matrices: dst_k[i,:] = src_k[src[i],:]  (ids, mask; any row_bytes).
 * labels int64 [B], src_idx int32 [B]. */
int tic_itm_sample(const float* u_coin, const float* u_pick, int B, int mode, const float* S, int64_t lds, float hard_ref,
                   const float* hard_ref_dev, int64_t* labels, int32_t* src_idx, void* stream);
/* Middle step of the tile-stream form: per row, label (as above, with the row's global index row_offset + i), default
 * source (the row itself, or the uniform pick when every weight is zero), the part p with
 * sum_{p' < p} qpart[p'][i] <= target < sum_{p' <= p} qpart[p'][i] and the residual target - sum_{p' < p} (uint64);
 * loc_part = -1 for rows with nothing to pick.  target = floor(trunc(u_pick * 2^24) * total / 2^24). */
int tic_itm_hard_locate(const float* u_coin, const float* u_pick, int m_local, int n_global, int row_offset, const void* qpart,
                        int n_parts, int64_t* labels, int32_t* src_idx, int32_t* loc_part, void* loc_res, void* stream);
int tic_gather_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes, int64_t row_bytes,
                    const int32_t* src_idx, int rows, void* stream);
/* Fused: sample + gather of ids and mask in one launch (the mm_late.py:396-409 loop). */
int tic_itm_sample_gather(const float* u_coin, const float* u_pick, int B, int mode, const float* S, int64_t lds, float hard_ref,
                          const float* hard_ref_dev, const void* ids, const void* mask, int64_t row_bytes, void* tim_ids,
                          void* tim_mask, int64_t* labels, int32_t* src_idx, void* stream);

/* ------------------------------------------------------------------------------------------------ fusion heads
 * Pack the CLS rows for linear_fusion (torch.cat at mm_late.py:94,111,141):
 *   Xcat[i,      :] = [ xt[i*xt_stride : +E]        | xv[i*xv_stride : +E] ]      i < B   (main pass)
 *   Xcat[B + i,  :] = [ xt[src[i]*xt_stride : +E]   | xv[i*xv_stride : +E] ]      if src != NULL (ITM pass, :170-181)
 * xt/xv bf16 with row strides in elements (so x_t[:,0,:] of a [B,L,E] tensor is xt_stride = L*E). */
/* xv may be NULL: then only the text half [:, :E] is written (attention fusion fills the other half by GEMM).
 * u_coin / u_pick (both or neither): evaluate the uniform ITM rule (mm_late.py:396-409, as tic_itm_sample) in place
 * instead of reading src_idx, so the pack does not have to wait for the sampler kernel (src_idx is then ignored). */
int tic_pack_cls_pairs(const void* xt, int64_t xt_stride, const void* xv, int64_t xv_stride, int B, int E,
                       const int32_t* src_idx, void* Xcat, int64_t ldx, const float* u_coin, const float* u_pick,
                       void* stream);
/* Gradient of the pack w.r.t. xt (vision is frozen, mm_late.py:67-69):
 *   dxt[i,:] += dXcat[i,:E] + sum_{k: src[k]==i} dXcat[B+k,:E]   (fp32 atomics, ONE launch: dxt must be zero on entry;
 *   E, ldd, ldd2 multiples of 4). */
int tic_unpack_cls_grad(const float* dXcat, int64_t ldd, const float* dX2 /* optional second addend, same layout */,
                        int64_t ldd2, int B, int E, const int32_t* src_idx, float* dxt, int64_t ld_dxt, void* stream);

/* Classifier + ITM heads with their losses, forward and backward in one launch (mm_late.py:163-164,182;
 * losses run_mm_late.py:85,97 and the mix mm_late.py:473-487):
 *   rows [0,B):   logits_cls = (H*keep_scale) W_cls^T + b_cls;  L_cls = -(1/B) sum_i sum_c w_c y_ic log softmax_ic
 *   rows [B,2B):  logits_tim = H W_tim^T + b_tim;                L_tim = mean CE(logits_tim, lbl_tim)    (if has_tim)
 * H fp32 [B or 2B, E] is the post-ReLU fusion output; dH (bf16 and/or fp32) receives c_cls*dL_cls/dH resp.
 * c_tim*dL_tim/dH, already multiplied by the ReLU mask (H > 0).  dW/db are accumulated (+=) in fp32 with atomics
 * into zero-initialised buffers.  losses[0] += L_cls, losses[1] += L_tim (unweighted). keep (uint8 dropout mask
 * [B,E], may be NULL) and keep_scale = 1/(1-p) reproduce nn.Dropout on the classifier branch only. */
int tic_heads_fwd_bwd(float* H, int64_t ldh, int B, int E, int C, int has_tim, const float* W_cls,
                      const float* b_cls, const float* W_tim, const float* b_tim, const float* y_soft,
                      const float* class_w, const int64_t* lbl_tim, const uint8_t* keep, float keep_scale, float c_cls,
                      float c_tim, float* logits_cls, float* logits_tim, float* losses, void* dH_bf16, void* dH_bf16_lo, int64_t ld_dhb,
                      float* dH_f32 /* optional */, int64_t ld_dhf, float* dW_cls, float* db_cls, float* dW_tim, float* db_tim, int relu_mask,
                      float* ws /* [rows * 8] floats scratch (dlogits) */,
                      const float* dlogits_ext /* NULL = fused losses; else upstream dL/dlogits [rows, 8] (autograd mode) */,
                      const float* Pt, const float* Pv, int64_t ldp, const int32_t* src_idx /* pairwise form, see below */,
                      void* stream);
/* Pairwise form of concat fusion (mm_late.py:92-96 with the ITM pairs of :170-181): linear_fusion([x_t | x_v]) =
 * x_t W_f[:, :E]^T + (x_v W_f[:, E:]^T + b_f).  With Pt = text halves and Pv = image halves (+ bias), fp32 [B, ldp], the
 * kernel above builds H itself — H[i] = relu(Pt[i] + Pv[i]), H[B+i] = relu(Pt[src_idx[i]] + Pv[i]) — and STORES it (H is an
 * output then): every sample is projected once, no packed [2B, 2E] operand exists, the fusion GEMM does half the FLOPs.
 * tic_fusion_pair_grad is its backward: from dH [2B, E] (bf16 hi + optional lo)
 *   dPv[i] = dH[i] + dH[B+i],   dPt[k] = dH[k] + sum_{i: src_idx[i] == k} dH[B+i]      (bf16 hi + optional lo, [B, ldo])
 * without atomics (the inverse of the gather is found by scanning src_idx; fixed summation order). */
int tic_fusion_pair_grad(const void* dH, const void* dH_lo, int64_t ld, int B, int E, int has_tim, const int32_t* src_idx,
                         void* dPt, void* dPt_lo, void* dPv, void* dPv_lo, int64_t ldo, void* stream);

/* The parameter-gradient half of tic_heads_fwd_bwd alone (dW/db of linear_cls and linear_tim from the dlogits that call left
 * in `ws`): call tic_heads_fwd_bwd with dW_* = NULL and this on another stream so that it runs beside the input-gradient
 * chain instead of inside it.  Accumulates (+=) with fp32 atomics into zero-initialised buffers. */
int tic_heads_wgrad(const float* H, int64_t ldh, int B, int E, int C, int has_tim, const float* dlogits, const uint8_t* keep,
                    float keep_scale, float* dW_cls, float* db_cls, float* dW_tim, float* db_tim, void* stream);

/* attention fusion, CLS-row collapse of mm_late.py:98-113,195-210 (exact algebra):
 *   q0 = fc_Q(x_t[:,0]);  kq = W_K^T q0;  c = <q0,b_K>;  s_j = (<kq, x_v[j]> + c) * E^-1/2;  a = softmax_j(s)
 *   xbar = sum_j a_j x_v[j];  ctx0 = W_V xbar + b_V.
 * These two kernels are the HBM-bound middle: ONE streaming read of x_v [B, Lv, E] (bf16) per direction, shared by
 * the main pass and the ITM pass (npass = 2: kq rows [0,B) are the main queries, rows [B,2B) the ITM queries of the
 * same images).  kq is the augmented product q0 * [W_K | b_K] (fp32, ld >= E+1): column E carries c.
 * Outputs: xbar (bf16 for the W_V GEMM and fp32 for backward) [npass*B, E]; attn [npass*B, Lv] fp32. E must be 768. */
int tic_attn_pool_fwd(const void* xv, int64_t xv_batch_stride, int64_t xv_tok_stride, const void* kq, int64_t ldkq, int B,
                      int npass, int Lv, int E, float scale, void* xbar_bf16, void* xbar_bf16_lo, int64_t ld_xb,
                      float* xbar_f32, int64_t ld_xf, float* attn, int64_t ld_attn, void* stream);
/* backward w.r.t. the augmented kq (x_v is frozen, mm_late.py:67-69):
 *   t_j = <dxbar, x_v[j]>, D = <dxbar, xbar>, ds_j = a_j (t_j - D) scale, dkq = sum_j ds_j x_v[j], dkq[E] = sum_j ds_j.
 * dkq is written as bf16 [npass*B, ld_dkq >= E+8] (columns E+1.. zeroed) — directly the operand of the W_K GEMMs. */
int tic_attn_pool_bwd(const void* xv, int64_t xv_batch_stride, int64_t xv_tok_stride, const float* attn, int64_t ld_attn,
                      const float* dxbar, int64_t ld_dxb, const float* xbar_f32, int64_t ld_xf, int B, int npass, int Lv,
                      int E, float scale, void* dkq_bf16, void* dkq_bf16_lo, int64_t ld_dkq, void* stream);

/* aspect-att fusion (mm_late.py:115-131) including the reference's stack->reshape row scrambling:
 * sample i pairs flat rows 2i, 2i+1 of [t_pool; v_pool].  out = relu(sum_k alpha_k V_k), alpha = softmax_k tanh(w.V_k + b). */
int tic_aspect_fwd(const void* t_pool, int64_t ldt, const void* v_pool, int64_t ldv, int B, int E, const float* w_a,
                   const float* b_a, float* out, int64_t ldo, float* alpha /* [B,2] saved */, void* stream);
int tic_aspect_bwd(const void* t_pool, int64_t ldt, const void* v_pool, int64_t ldv, int B, int E, const float* w_a,
                   const float* b_a, const float* out, int64_t ldo, const float* alpha, const float* dout, int64_t lddo,
                   float* dt_pool, int64_t lddt, float* dw_a, float* db_a, void* stream);

/* gmu gate (mm_late.py:133-144): G[i,:] = z*tp + (1-z)*vp, z = sigmoid([xt_cls | xv_cls]) (no learned gate).
 * tp/vp fp32 [B,2E] are linear_gmu_t/v outputs; Xcat bf16 [B,2E] from tic_pack_cls_pairs. */
int tic_gmu_gate_fwd(const void* Xcat, int64_t ldx, const float* tp, const float* vp, int64_t ldp, int B, int E2,
                     void* G_bf16, void* G_bf16_lo, int64_t ldg, void* stream);
int tic_gmu_gate_bwd(const void* Xcat, int64_t ldx, const float* tp, const float* vp, int64_t ldp, const float* dG,
                     int64_t lddg, int B, int E2, void* dtp_bf16, void* dvp_bf16, void* dtp_lo, void* dvp_lo, int64_t lddp,
                     float* dXcat_gate, int64_t lddx, void* stream);

/* ------------------------------------------------------------------------------------------------ eval bookkeeping
 * the per-batch tail of MMLate_Model.eval / compute_predictions (mm_late.py:594-612, 661-690) and
 * utils.compute_metrics (utils.py:294-325) without a host synchronisation per batch.
 * tic_eval_accumulate, one call per batch of B rows:
 *   pred[i]   = first index of the maximum of logits[i, :C]   (torch.argmax(softmax(output), dim=1), mm_late.py:597-600)
 *   target[i] = first index of the maximum of y_soft[i, :C]   (float one-hot labels, mm_late.py:601)  or  y_int[i]
 *   preds_out[i], targets_out[i] (int64, the caller passes the write position inside its epoch-long arrays),
 *   state: uint64 words [C*C confusion matrix, row = target, col = pred | correct | rows | batches | scratch]
 *          (tic_eval_state_words(C) words, zero at the start of an epoch),
 *   sums (fp32[2], zero at the start of an epoch): sums[0] += batch_loss[0] (optional, device scalar: the reference
 *   averages per-batch losses, :594,615), sums[1] += 100 * correct_in_batch / B (per-batch accuracy, :607-608,616).
 * tic_metrics_from_confusion: out6 = f1_weighted, f1_macro, precision_weighted, precision_macro, recall_weighted,
 *   recall_macro with torchmetrics-0.11 multiclass semantics (0 where a denominator is 0; `macro` skips classes with
 *   tp+fp+fn == 0).  torchmetrics is a third-party dependency pinned in timrel-env.yml:120 and absent here: parity
 *   against it is unpinned; the restatement (oracle/restatement.py:metrics_from_confusion) is checked against sklearn. */
int tic_eval_state_words(int C);
int tic_eval_accumulate(const float* logits, int64_t ldl, const float* y_soft, int64_t ldy, const int64_t* y_int, int B, int C,
                        const float* batch_loss, int64_t* preds_out, int64_t* targets_out, void* state, float* sums,
                        void* stream);
int tic_metrics_from_confusion(const void* state, int C, float* out6, void* stream);

/* ------------------------------------------------------------------------------------------------ utilities */
/* Head of a captured TRAINING step: refresh the bf16 working copies of up to 8 fp32 master weight matrices (the optimiser
 * updates the masters in place between replays; models/utils.py:280-292 selects them) and scale_out[0] = exp(*logit_scale)
 * (HF :272; values outside (0, 40] are clamped to 40 and *status = 1, sticky) in ONE launch.  The *_host arrays are HOST
 * arrays of n entries; matrix i is [rows, cols] fp32 with leading dimension lds[i], written as bf16 with ldd[i].
 * zero0 / zero1: up to two fp32 ranges the same launch sets to zero (the step's loss sums and small gradient accumulators),
 * so that no memset node sits in front of the step's first kernel. */
int tic_refresh_weights(int n, const float* const* src_host, void* const* dst_host, const int64_t* lds_host,
                        const int64_t* ldd_host, const int* rows_host, const int* cols_host, const float* logit_scale,
                        float* scale_out, uint32_t* status, float* zero0, int nzero0, float* zero1, int nzero1, void* stream);
int tic_cast_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, void* stream);
int tic_cast_bf16_to_f32(const void* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, void* stream);
/* dst[r, c] (+)= column sums etc. are done by GEMMs; bias gradient: db[n] = sum_m dY[m,n] (bf16 in, fp32 out). */
int tic_colsum_bf16(const void* X, int64_t ldx, int rows, int cols, float* out, void* stream);
/* Same for a split-precision (hi, lo) pair of identical layout in ONE launch (X_lo may be NULL). */
int tic_colsum_bf16_pair(const void* X, const void* X_lo, int64_t ldx, int rows, int cols, float* out, void* stream);
/* out[0] = (1-bi-bm)*losses[0] + bi*0.5*(itc[0]+itc[1])/B + bm*losses[1]  (mm_late.py:473-487);
 * out[1..3] = L_cls, L_itc, L_itm.  out has 4 floats. */
int tic_loss_mix(const float* losses, const float* itc_sums, int n_global, float beta_itc, float beta_itm,
                 int use_itc, int use_itm, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------ peer memory
 * Multi-GPU exchange of the row-sharded ITC step (the reference is single-device, mm_late.py:30) WITHOUT a
 * collective library on the data path: every rank owns one block of device memory that all peers map through CUDA IPC,
 * and loads the other ranks' embeddings / softmax statistics over NVLink with its own kernel.
 *   tic_peer_alloc   cudaMalloc + zero a block (returns the device pointer in *out_ptr)
 *   tic_peer_export  64-byte IPC handle of a block (host buffer of tic_peer_handle_bytes() bytes)
 *   tic_peer_open    map a peer's block from its handle (enables peer access lazily); tic_peer_close unmaps it
 *   tic_peer_exchange  ONE kernel: system-scope release/acquire barrier across the `world` ranks (flags at byte offset
 *       flag_off of every block: uint32[8], zero-initialised), then, for each of the nseg byte ranges
 *       [src_off[s], src_off[s]+bytes[s]) of EVERY rank's block p (its own included), a pull into the local buffer
 *       dst[s] + p * dst_stride[s].  Offsets, sizes and destinations are multiples of 16 bytes.  `ctr` = 2 zero-initialised
 *       uint32 in local device memory (epoch + ticket).  All ranks must issue the same sequence of exchanges; nothing
 *       synchronises with the host, so the call is CUDA-graph capturable.  A peer that never arrives traps after 20 s.
 * The *_host arguments are HOST arrays (world pointers / nseg values). */
int tic_peer_handle_bytes(void);
int tic_peer_alloc(int64_t bytes, void** out_ptr);
int tic_peer_free(void* ptr);
int tic_peer_export(const void* ptr, void* handle_out_host);
int tic_peer_open(const void* handle_host, void** out_ptr);
int tic_peer_close(void* ptr);
int tic_peer_exchange(void* const* bases_host, int world, int rank, int64_t flag_off, uint32_t* ctr, int nseg,
                      const int64_t* src_off_host, const int64_t* bytes_host, void* const* dst_host,
                      const int64_t* dst_stride_host, void* stream);
/* tic_peer_pull: the pull half alone (the ranks were ordered by the preceding tic_peer_exchange on the same stream or an
 * ancestor of it): peer by peer starting with `rank`, and after peer p's ranges have landed ready[p] = ctr[0] (the epoch of
 * that exchange) is published with release semantics — the consumer side of tic_itc_fwd(seg_ready = ready, seg_epoch = ctr)
 * may run concurrently on another stream.  ready / tickets: `world` zero-initialised uint32 each, local device memory.
 * max_blocks bounds the grid (0 = 148 blocks of 128 threads, 64 registers each) so the pull lives beside a persistent GEMM. */
/* PUSH form (small global batches, where the exchange latency is on the critical chain).  Every rank keeps its share of a
 * gathered buffer IN PLACE (the producing kernel writes this rank's slot of the rank's own gathered buffer); tic_peer_push
 * stores the nseg local ranges [src[s], +bytes[s]) into every REMOTE rank's block at dst_off[s] + rank * dst_stride[s] and
 * then publishes flags[rank] = step[1] (uint32[world] at flag_off of every remote block) with release semantics at system
 * scope.  No rank waits in this kernel.  The consumer is tic_itc_fwd / tic_itc_bwd_g with seg_ready = the LOCAL flag words,
 * seg_epoch = step + 1 and a negative seg_cols: the TMA producer polls a column segment's flag right before its first load
 * from it, visits the local segment first and never waits for it — so the consumers need no stream dependency on the push
 * and run beside it.
 *   step (2 x uint32, local, initialised {0, 1}): step[0] = completed steps, step[1] = epoch of the step in flight; advanced
 *   by tic_peer_signal at the tail of a step, which also publishes flags[rank] = step[0] at its own flag_off.
 *   wait_off >= 0: before storing into the peers, wait until every peer's word in the local uint32[world] at wait_off has
 *   reached step[1] - 1: it has finished reading what the previous step delivered.   ticket: 1 zeroed uint32 per push. */
int tic_peer_push(void* const* bases_host, int world, int rank, int64_t flag_off, int64_t wait_off, const uint32_t* step,
                  uint32_t* ticket, int nseg, void* const* src_host, const int64_t* bytes_host, const int64_t* dst_off_host,
                  const int64_t* dst_stride_host, void* stream);
int tic_peer_signal(void* const* bases_host, int world, int rank, int64_t flag_off, uint32_t* step, void* stream);
int tic_peer_pull(void* const* bases_host, int world, int rank, const uint32_t* ctr, uint32_t* ready, uint32_t* tickets, int nseg,
                  const int64_t* src_off_host, const int64_t* bytes_host, void* const* dst_host, const int64_t* dst_stride_host,
                  int max_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIC_B200_H */
