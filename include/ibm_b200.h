/*
 * ibm_b200.h — C ABI of libibm_b200.so, the sm_100a (B200) implementation of the
 * InferBiomechanics data-parallel hot path.
 *
 * The reference (jbejjani2022/InferBiomechanics) is 100 % Python and has no FFI layer; its hot
 * path dispatches ATen ops from the Python sites cited per entry point below (paths relative to
 * the reference root).  Each function here replaces the group of ATen launches at the cited
 * site.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with h_ (host);
 *   - the caller owns every buffer; the library allocates nothing persistent except a small
 *     per-device scratch for cross-block reductions (ibm_workspace_bytes / caller-provided);
 *   - all functions are asynchronous on `stream` (a cudaStream_t passed as void*) and re-entrant
 *     per stream; none synchronises the device;
 *   - return value: 0 on success, non-zero IBM_E_* on failure; the message is available through
 *     ibm_last_error().  There is NO CPU fallback and NO other-architecture dispatch: on a
 *     device that is not compute capability 10.x every compute entry point returns IBM_E_ARCH;
 *   - bf16 matrices are row-major with a leading dimension `ld` (elements) that must be a
 *     multiple of 8 (16-byte rows, TMA-legal); base pointers must be 16-byte aligned;
 *   - "rows30" layout: one row per (window, frame) holding the 30 output channels
 *     [CoP 6 | force 6 | torque 6 | wrench 12]  (src/models/Groundlink.py:151-156).
 */
#ifndef IBM_B200_H_
#define IBM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBM_OK 0
#define IBM_E_ARG 1      /* invalid argument (shape, alignment, enum) */
#define IBM_E_ARCH 2     /* device is not sm_100 */
#define IBM_E_CUDA 3     /* CUDA runtime / driver error */
#define IBM_E_UNSUPPORTED 4

/* activation ids (src/models/FeedForwardRegressionBaseline.py:7-11; Groundlink.py:46; TransformerBaseline.py:16) */
#define IBM_ACT_NONE 0
#define IBM_ACT_RELU 1
#define IBM_ACT_SIGMOID 2
#define IBM_ACT_TANH 3
#define IBM_ACT_ELU 4
#define IBM_ACT_SILU 5

/* element types */
#define IBM_F32 0
#define IBM_BF16 1

/* ---- library ----------------------------------------------------------------------------- */
int ibm_version(void);
/* copies the calling thread's last error message; returns its length */
size_t ibm_last_error(char* buf, size_t cap);
/* 0 iff `device` is compute capability 10.x */
int ibm_device_check(int device);
/* Walk order of the big row-streaming kernels (GEMM tiles, LayerNorm rows, attention windows).  0 = ascending (default,
 * or IBM_WALK_ORDER unset); 1 = consecutive big launches alternate ascending / descending, so that a consumer starts on the
 * rows its producer wrote last (still in L2); 2 = every launch alternates, whatever its size (tests).  Results do not depend
 * on it (split-K sums are taken in a different order).  Process-wide host state: set it once, from the
 * thread that launches (the reference drives one device from one Python thread per rank, src/cli/train.py:100-102). */
int ibm_set_walk_order(int32_t mode);
/* bytes of zero-initialised device scratch the reduction kernels need (loss, layernorm bwd) */
size_t ibm_workspace_bytes(void);

/* ---- window batcher  (src/data/AddBiomechanicsDataset.py:121-139, 161-285;
 *                       src/models/FeedForwardRegressionBaseline.py:97-108) ---------------------- */

/* Candidate-window validity, Dataset.py:131-139: for candidate c of trial j (start ws),
 * valid[c] = !any(missing[ws : ws+window_size : stride]).  cand_trial/cand_start are int32[n]. */
int ibm_window_valid_mask(const uint8_t* missing, const int64_t* trial_frame_base,
                          const int32_t* cand_trial, const int32_t* cand_start, int64_t n_cand,
                          int32_t window_size, int32_t stride, uint8_t* valid, void* stream);

/* Gather F frames (stride `stride`) of each selected window from the frame store into packed
 * model inputs.  frames: fp32 [total_frames, frame_ld] whose first C columns are the model's
 * per-frame concat (FeedForward…py:97-108 order).  win_row0[i] = absolute frame index of the
 * window's first frame.  Any of out_f32 / out_bf16 may be NULL.
 *   out_f32 : [n_win, F, C] contiguous (bit-exact copy)
 *   out_bf16: element (i, f, c) at  out_bf16[(i*F + f)*bf16_frame_stride + i*bf16_win_extra + bf16_col0 + c]
 *             (round-to-nearest-even); use frame_stride=C, win_extra=ld-F*C for the FeedForward
 *             row-per-window layout, frame_stride=ld, win_extra=0 for row-per-frame layouts. */
int ibm_pack_windows(const float* frames, int64_t frame_ld, int32_t C, const int64_t* win_row0,
                     int64_t n_win, int32_t F, int32_t stride, float* out_f32, void* out_bf16,
                     int64_t bf16_frame_stride, int64_t bf16_win_extra, int64_t bf16_col0, void* stream);

/* Concatenate n_src (<= 10) fp32 tensors [n_rows, width_k] (contiguous) along the channel axis in
 * the given order (the reference's torch.concat at FeedForward…py:97-108 / Groundlink.py:122-133)
 * into packed rows; outputs addressed exactly as in ibm_pack_windows (row r = window*F + frame).
 * h_src: host array of device pointers; h_widths: host int32[n_src]. */
int ibm_pack_inputs(const void* const* h_src, const int32_t* h_widths, int32_t n_src, int64_t n_rows,
                    int32_t F, float* out_f32, void* out_bf16, int64_t bf16_frame_stride,
                    int64_t bf16_win_extra, int64_t bf16_col0, void* stream);

/* TransformerBaseline input pack (src/models/TransformerBaseline.py:108-126): the reference concatenates n_src
 * channel-major tensors (B, C_i, T) on dim 1, transposes to (B, T, C) and appends the learned temporal embedding
 * emb[t, 0:E] (expand + cat, not add).  Here: fp32 sources -> bf16 rows [B*T, ld], row b*T+t = [src_0[b,:,t] | … |
 * emb[t,:] | 0…], ld %% 8 == 0, ld >= sum C_i + E.  n_src <= 8.  Bit-exact RNE bf16 of the fp32 values. */
int ibm_pack_channel_major(const void* const* srcs, const int32_t* channels, int32_t n_src, int64_t B, int32_t T,
                           const float* emb, int32_t E, void* out_bf16, int64_t ld, void* stream);

/* Pre-packed analysis stream: rows converted once on the host to frame-major bf16 [n_rows = B*T, ld_src] (C channels per frame)
 * -> the TransformerBaseline's input rows [n_rows, ld] = [C channels | E temporal-embedding columns of frame r % T | zeros],
 * bit-identical to ibm_pack_channel_major on the fp32 (B, C, T) tensors (TransformerBaseline.py:108-126); v_out (may be NULL):
 * bf16 [n_rows, 8] <- columns [v_col0, v_col0+3) | zeros (the CoM accelerations blended by SimpleAttention, …:135-137). */
int ibm_expand_rows_bf16(const void* src_bf16, int64_t ld_src, int32_t C, int64_t n_rows, int32_t T, const float* emb,
                         int32_t E, void* out_bf16, int64_t ld, void* v_out_bf16, int32_t v_col0, void* stream);

/* Label rows (Dataset.py:216-261): raw first-pass per-frame [cop 3nb | force 3nb | torque 3nb |
 * wrench 6nb] in the SUBJECT's body order → rows30 in the DATASET's body order, force/torque/
 * wrench divided by the subject mass (IEEE fp32 division), absent bodies → 0.
 * contact_idx: int32 [n_win, nb]; mass: fp32 [n_win].  last_frame_only!=0 keeps only frame F-1. */
int ibm_pack_labels(const float* raw, int64_t raw_ld, int32_t nb, const int64_t* win_row0,
                    const int32_t* contact_idx, const float* mass, int64_t n_win, int32_t F,
                    int32_t stride, int32_t last_frame_only, float* out_rows, int64_t out_ld, void* stream);

/* ---- regression loss  (src/loss/RegressionLossEvaluator.py:160-221 step 1, 230-263 step 2.2) ----- */

/* Quantity order everywhere: 0 CoP, 1 force, 2 torque, 3 wrench (6,6,6,12 channels).
 * h_out / h_lab: host arrays of 4 device pointers (fp32); h_out_strides / h_lab_strides: host
 * int64[8] = {stride_b, stride_f} per quantity in elements (channel stride is 1).
 * h_weights: host fp32[30] = multiplicity of each component in the args.predict_* lists
 * (…py:217-220; a repeated index counts twice, an absent one 0).
 * result: fp32[40] = loss, then the 30 per-component means in quantity order cop[6], force[6],
 * moment[6], wrench[12], then reports{force, moment, cop, wrench_moment, wrench, com_acc}
 * (last frame only, …py:119-158), 3 pad.
 * workspace: ibm_workspace_bytes() of zeroed device memory (left zeroed on return). */
int ibm_regression_loss_fwd(const void* const* h_out, const int64_t* h_out_strides,
                            const void* const* h_lab, const int64_t* h_lab_strides,
                            int64_t B, int64_t F, const float* h_weights, float threshold,
                            float* result, void* workspace, void* stream);

/* d loss / d outputs = upstream * 2 w_c m^2 (o-l) / (B*F)   (SURVEY §9.1).
 * upstream: device fp32 scalar (NULL ⇒ 1).  h_grad: host array of 4 device pointers of
 * grad_dtype (IBM_F32 | IBM_BF16) with h_grad_strides like h_out_strides. */
int ibm_regression_loss_bwd(const void* const* h_out, const int64_t* h_out_strides,
                            const void* const* h_lab, const int64_t* h_lab_strides,
                            int64_t B, int64_t F, const float* h_weights, float threshold,
                            const float* upstream, void* const* h_grad, const int64_t* h_grad_strides,
                            int32_t grad_dtype, void* stream);

/* ---- the loss evaluator's four static helpers with their general contract -------------------------
 * (src/loss/RegressionLossEvaluator.py:73-158; exercised by test/loss/test_RegressionLossEvaluator.py)
 * All tensors are fp32 (B, F, C) views with unit channel stride; sb / sf = batch / frame strides in
 * elements.  workspace: ibm_workspace_bytes() of zeroed device memory (left zeroed on return).
 * IBM_E_ARG carries the reference's ValueError text (empty tensor, C % 3, C % vec_size, C != 6). */

/* get_squared_diff_mean_vector (…py:73-83): result[c] = mean over (b, f) of (out-lab)^2, any C. */
int ibm_sqdiff_mean_vector(const float* out_t, int64_t o_sb, int64_t o_sf, const float* lab_t, int64_t l_sb,
                           int64_t l_sf, int64_t B, int64_t F, int32_t C, float* result, void* workspace,
                           void* stream);
/* its autograd backward: grad_out[b,f,c] = 2 upstream[c] (out-lab) / (B F) (contiguous (B,F,C)); grad_lab is
 * the negative.  Either may be NULL. */
int ibm_sqdiff_mean_vector_bwd(const float* out_t, int64_t o_sb, int64_t o_sf, const float* lab_t, int64_t l_sb,
                               int64_t l_sf, int64_t B, int64_t F, int32_t C, const float* upstream,
                               float* grad_out, float* grad_lab, void* stream);
/* get_mask_by_threes (…py:85-108): mask[b,f,3g:3g+3] = (||x[b,f,3g:3g+3]||_2 > threshold) ? 1 : 0, strict >,
 * bit-exact; mask is contiguous fp32 (B,F,C); C % 3 == 0. */
int ibm_mask_by_threes(const float* x, int64_t sb, int64_t sf, int64_t B, int64_t F, int32_t C, float threshold,
                       float* mask, void* stream);
/* get_mean_norm_error (…py:119-141): result[0] = mean over (b, g) of ||(out-lab)[b, F-1, g v:(g+1) v]||_2 — last
 * frame only; C % vec_size == 0.  fold_halves = 1 is get_com_acc_error (…py:143-158): C == 6, the two 3-vectors of
 * each tensor are summed first. */
int ibm_mean_norm_error(const float* out_t, int64_t o_sb, int64_t o_sf, const float* lab_t, int64_t l_sb, int64_t l_sf,
                        int64_t B, int64_t F, int32_t C, int32_t vec_size, int32_t fold_halves, float* result,
                        void* workspace, void* stream);

/* ---- DDPM (builder-owned spec; NOT in the reference) --------------------------- */

/* x_t = sqrt_abar[t_b] x0 + sqrt_1m_abar[t_b] eps.  x0/eps/xt_f32: fp32 [B, per_win] contiguous
 * (per_win = F*30).  If eps == NULL, eps is drawn on-device from Philox4x32-10(seed, offset) and
 * (if eps_out != NULL) written there.  xt_bf16 (optional): rows30 scattered into a bf16 matrix:
 * element (m, c) at xt_bf16[m*bf16_ld + c], m = b*F + f. */
int ibm_q_sample(const float* x0, const float* eps, const int32_t* t, const float* sqrt_abar,
                 const float* sqrt_1m_abar, int64_t B, int64_t per_win, float* xt_f32,
                 void* xt_bf16, int64_t bf16_ld, uint64_t seed, uint64_t offset, float* eps_out,
                 void* stream);

/* x_{t-1} = c1[t] x0_hat + c2[t] x_t + [t>0] sigma[t] z  for one shared timestep t (read from
 * device int32 *t_dev so the sampling loop can live in a CUDA graph).  x0_hat: fp32 rows with
 * leading dim x0_ld (30 used); x_t / z / x_prev: fp32 [M,30] contiguous; z == NULL ⇒ Philox.
 * xprev_bf16 optional as in ibm_q_sample.  If t_next_dev != NULL it receives t-1. */
int ibm_ddpm_posterior_step(const float* x0_hat, int64_t x0_ld, const float* x_t, const float* z,
                            const int32_t* t_dev, const float* coef_x0, const float* coef_xt,
                            const float* sigma, int64_t M, float* x_prev, void* xprev_bf16,
                            int64_t bf16_ld, uint64_t seed, uint64_t offset, int32_t* t_next_dev,
                            void* stream);

/* sinusoidal timestep embedding, bf16 [B, dim] (ld = dim): [sin(t w_k) | cos(t w_k)],
 * w_k = 10000^(-k/(dim/2)).  t: int32[B], or if t_is_scalar a single device int32 broadcast. */
int ibm_timestep_embed(const int32_t* t, int32_t t_is_scalar, int64_t B, int32_t dim, void* out_bf16,
                       void* stream);

/* h[m,:] += temb[m / F, :] + pos[m % F, :]   (bf16 in place; temb bf16 [B,d]; pos fp32 [F,d]).
 * t_row_dev != NULL (reverse sampling: every window is at the same timestep): temb is a table with one row per
 * timestep and row t_row_dev[0] (a device int32) is added to every window, so the time MLP is evaluated once per
 * model, not once per denoise step. */
int ibm_add_time_pos(void* h_bf16, int64_t ld, const void* temb_bf16, int64_t temb_ld,
                     const float* pos, int64_t M, int32_t F, int32_t d, const int32_t* t_row_dev, void* stream);
/* backward of the above in ONE pass over dh: dtemb[b,:] = sum_f dh[b,f,:] (bf16 out); dpos[f,:] += sum_b dh[b,f,:] (fp32
 * atomic); dbias (may be NULL): fp32 [d] += sum_{b,f} dh[b,f,:] — the bias gradient of the Linear that produced h (the
 * denoiser's in-projection), which would otherwise be a second full read of dh. */
int ibm_add_time_pos_bwd(const void* dh_bf16, int64_t ld, void* dtemb_bf16, int64_t temb_ld,
                         float* dpos, int64_t M, int32_t F, int32_t d, float* dbias, void* stream);

/* ---- dense layers  (nn.Linear sites: FeedForward…py:73,113; Groundlink.py:51-62;
 *                     TransformerBaseline.py:12-18,91; nn.Conv1d: Groundlink.py:41) --------------- */

/* D[M,N] (+)= epilogue( A · B^T )  on tcgen05 tensor cores, bf16 operands, fp32 accumulation.
 *   A: bf16, logical [M,K]. a_mn_major==0: stored row-major [M,K] (ld=lda).
 *                           a_mn_major==1: stored row-major [K,M] (ld=lda)  (i.e. A^T in memory).
 *   B: bf16, logical [N,K]. b_mn_major likewise ([N,K] vs [K,N] in memory).
 *   bias: fp32[N] or NULL.  act: IBM_ACT_* applied to (acc + bias).
 *   aux: bf16 [M,N] (ld=ldaux) or NULL; aux_mode 1: out = act(acc+bias) + aux   (residual)
 *                                        aux_mode 2: out = (acc+bias) * act'(aux) (aux = saved
 *                                        activation OUTPUT; dgrad through relu/sigmoid/tanh/elu)
 *   out: out_dtype IBM_BF16 | IBM_F32, row-major [M,N] (ld=ldd).
 *   accumulate!=0 (requires IBM_F32 out, no act/aux): D += A·B^T via TMA reduce-add; the K range is
 *   split over `split_k` CTAs (weight gradients: K = tokens).  split_k<=0 ⇒ chosen by the library.
 *   taps>1: A rows are shifted by tap index: K is taps*K_tap, k-block kb reads A rows
 *   m + (kb / kb_per_tap) (implicit-GEMM temporal convolution, Groundlink.py:41).
 *   colsum_out: NULL, or fp32[N] that the column sums of the bf16 output (rows < M, as rounded) are
 *   ADDED to by the epilogue: when D is the gradient w.r.t. a layer's pre-activation this is that
 *   layer's bias gradient (nn.Linear bias, TransformerBaseline.py:15) without another pass over D.
 *   mask / ldmask / mask_mode: sign bitmask [M][ldmask bytes] (bit c&7 of byte c>>3 <-> column c; bf16 output,
 *   N %% 64 == 0, no aux).  mask_mode 1 (forward of Linear+ReLU, TransformerBaseline.py:15-16): bit = (D > 0) is
 *   written next to D.  mask_mode 2 (its dgrad): D = (acc + bias) where the bit is set, 0 elsewhere — the ReLU
 *   derivative from 1 bit per element instead of re-reading the saved activation (aux_mode 2).  0: unused. */
int ibm_gemm_bf16(const void* A, int64_t lda, int32_t a_mn_major, const void* B, int64_t ldb,
                  int32_t b_mn_major, int64_t M, int64_t N, int64_t K, const float* bias, int32_t act,
                  const void* aux, int64_t ldaux, int32_t aux_mode, void* D, int64_t ldd,
                  int32_t out_dtype, int32_t accumulate, int32_t split_k, int32_t taps, float* colsum_out, void* mask,
                  int64_t ldmask, int32_t mask_mode, void* stream);

/* out[n] (+)= sum_m X[m,n]   (bias gradients; X bf16 [M,N] ld; out fp32; accumulate via atomics) */
int ibm_colsum_bf16(const void* X, int64_t ld, int64_t M, int64_t N, float* out, void* stream);

/* y = act(x) elementwise bf16→bf16 and its backward dx = dy*act'(x) (x = pre-activation) */
int ibm_act_fwd(const void* x, void* y, int64_t n, int32_t act, void* stream);
int ibm_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int32_t act, void* stream);

/* fp32 → bf16 (RNE) flat cast; n elements */
int ibm_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
/* bf16 → fp32 flat cast */
int ibm_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream);
/* strided 2-D cast fp32 [rows, cols] (ld_src) → bf16 (ld_dst), zero-filling cols..ld_dst */
int ibm_cast_pad_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                          int64_t cols, void* stream);
/* Conv1d weight (Cout,Cin,Kt) fp32 → bf16 GEMM layout [Cout, Kt*cin_pad], (tap, ci) order */
int ibm_conv_weight_to_gemm(const float* w, int32_t cout, int32_t cin, int32_t kt, int32_t cin_pad,
                            void* dst_bf16, void* stream);
/* inverse for gradients: dW_gemm fp32 [Cout, Kt*cin_pad] → (Cout,Cin,Kt) fp32 (+= if accumulate) */
int ibm_conv_wgrad_from_gemm(const float* g, int32_t cout, int32_t cin, int32_t kt, int32_t cin_pad,
                             float* dw, int32_t accumulate, void* stream);

/* Temporal replicate padding for the implicit-GEMM convolution (Groundlink.py:41, padding_mode="replicate").
 * Padded row layout: window b owns rows [b*Tp, (b+1)*Tp), Tp = T + 2*pad, frame t at row b*Tp + pad + t.
 * replicate: pad rows <- first / last frame row.  fold (its adjoint): edge frame rows += pad rows, pad rows <- 0. */
int ibm_replicate_pad_rows(void* X_bf16, int64_t ld, int64_t n_win, int32_t T, int32_t pad, int32_t cols, void* stream);
int ibm_fold_pad_rows(void* G_bf16, int64_t ld, int64_t n_win, int32_t T, int32_t pad, int32_t cols, void* stream);
/* Inverted dropout (nn.Dropout, Groundlink.py:53,59): y = x * keep/(1-p), keep from Philox4x32-10(seed, offset);
 * the backward pass calls it on dy with the same (seed, offset).  n % 4 == 0. */
int ibm_dropout_bf16(const void* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
/* Same mask generator with the Philox offset = offset + step_mul * *step_dev, the step read from device memory at execution
 * time (graph-replayed training steps draw a fresh mask per replay; forward and backward of one step see the same value). */
int ibm_dropout_bf16_dev(const void* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset,
                         const int64_t* step_dev, int64_t step_mul, void* stream);
/* Conv1d weight (Cout,Cin,Kt) fp32 -> bf16 dgrad layout [Cin, Kt*cout_pad]: B[ci, j*cout_pad+co] = W[co,ci,Kt-1-j] */
int ibm_conv_weight_to_dgrad(const float* w, int32_t cout, int32_t cin, int32_t kt, int32_t cout_pad,
                             void* dst_bf16, void* stream);

/* ---- BatchNorm1d on the layer input  (src/models/FeedForwardRegressionBaseline.py:71-72; nn.BatchNorm1d defaults
 *      eps=1e-5, momentum=0.1, affine, track_running_stats) ---- */

/* floats of caller-owned scratch the two BatchNorm entry points need for C features */
size_t ibm_batchnorm_workspace_floats(int32_t C);
/* y = (x - mean) / sqrt(var + eps) * gamma + beta per column.  x,y bf16 [M, ld] (ld %% 8 == 0; columns C..ld of y are
 * written 0); gamma/beta fp32[C].  training != 0: mean / biased variance of THIS batch (two-pass per row chunk, chunks
 * merged with Chan's update), saved as save_mean / save_rstd fp32[C] for the backward pass; running_mean / running_var
 * (fp32[C], may be NULL) are updated in place with `momentum` and the UNBIASED batch variance; M == 1 is an argument
 * error like torch's ValueError.  training == 0: normalises with running_mean / running_var; save_* and workspace unused. */
int ibm_batchnorm_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, int32_t C, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, float* save_mean, float* save_rstd,
                      int32_t training, float momentum, float eps, float* workspace, void* stream);
/* Backward of the above.  dgamma[c] += sum_m dy*xhat, dbeta[c] += sum_m dy (either may be NULL);
 * dx = gamma*rstd*(dy - mean_m(dy) - xhat*mean_m(dy*xhat)) in training mode, gamma*rstd*dy in eval mode (dx may be NULL
 * when only the parameter gradients are wanted, may alias dy).  (mean, rstd_or_var) = (save_mean, save_rstd) when
 * training, (running_mean, running_var) otherwise.  act_out != NULL: dx is additionally multiplied by act'(.) expressed
 * through the saved activation OUTPUT act_out (the activation that produced x's layer input; relu/sigmoid/tanh/elu).
 * dx_colsum (fp32[C], may be NULL): += column sums of dx taken in fp32 before the bf16 rounding — the bias gradient of
 * the Linear that feeds this BatchNorm (a sum of cancelling terms in training mode). */
int ibm_batchnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, void* dx, int64_t lddx, int64_t M,
                      int32_t C, const float* gamma, const float* mean, const float* rstd_or_var, int32_t training,
                      float eps, const void* act_out, int64_t ldact, int32_t act, float* dgamma, float* dbeta,
                      float* dx_colsum, float* workspace, void* stream);

/* ---- residual + LayerNorm  (src/models/TransformerBaseline.py:31,36; nn.LayerNorm eps=1e-5) ---- */

/* y = LN(s) * gamma + beta over the first d columns of each row (columns d..ld are written 0).
 * s,y bf16 [M, ld]; gamma/beta fp32[d]; mean/rstd fp32[M] saved for backward (may be NULL). */
int ibm_layernorm_fwd(const void* s, void* y, int64_t ld, const float* gamma, const float* beta,
                      int64_t M, int32_t d, float eps, float* mean, float* rstd, void* stream);
/* ds = LN'(dy); dgamma += sum dy*xhat; dbeta += sum dy; if dcolsum != NULL: dcolsum += colsum(ds)
 * (bias gradient of the linear layer that produced s).  fp32 atomics into dgamma/dbeta/dcolsum. */
int ibm_layernorm_bwd(const void* dy, const void* s, int64_t ld, const float* gamma,
                      const float* mean, const float* rstd, int64_t M, int32_t d, void* ds,
                      float* dgamma, float* dbeta, float* dcolsum, void* stream);

/* ---- attention  (nn.MultiheadAttention, src/models/TransformerBaseline.py:12-13,29) ----------- */

/* Whole-sequence softmax attention per (window, head): o = softmax(scale * q k^T) v.
 * q,k,v,o: bf16 matrices with one row per (window, frame) (n_win*T rows); head h occupies columns
 * [h*hd_qk, (h+1)*hd_qk) of q and k and [h*hd_v, (h+1)*hd_v) of v and o.  For a fused QKV
 * projection pass q = qkv, k = qkv + kv_off, v = qkv + 2*kv_off with ldq = ldk = ldv.
 * Supported (hd_qk, hd_v): (64,64) (48,48) (32,32) and (112,8) — the latter is SimpleAttention
 * (TransformerBaseline.py:51-70: scale = 1, d = 108 padded to 112, value dim 3 padded to 8).
 * T <= 256 (whole K and V of a head live in shared memory). */
int ibm_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      void* o, int64_t ldo, int64_t n_win, int32_t T, int32_t H, int32_t hd_qk,
                      int32_t hd_v, float scale, void* stream);
/* dqkv (same layout as qkv: q | k | v blocks kv_off columns apart) from d_o; probabilities are
 * recomputed on chip.  T <= 64, head_dim in {32,48,64}.  dbias_qkv (may be NULL): fp32
 * [2*kv_off + H*head_dim] vector that the column sums of dqkv are ADDED to — the gradient of
 * nn.MultiheadAttention.in_proj_bias (TransformerBaseline.py:12) without a second pass over dqkv. */
int ibm_attention_bwd(const void* qkv, int64_t ld_qkv, int64_t kv_off, const void* d_o, int64_t ld_o,
                      void* dqkv, int64_t n_win, int32_t T, int32_t H, int32_t head_dim, float scale,
                      float* dbias_qkv, void* stream);

/* Backward of the same attention for whole windows of up to 256 frames (the reference's own TransformerBaseline shapes:
 * 3 heads x 36 -> 48 padded at T = window_size, TransformerBaseline.py:12-13,29; and SimpleAttention, …:51-70, with
 * (hd_qk, hd_v) = (112, 8), dv = NULL because its values are a model INPUT).  q,k,v,o,d_o as in ibm_attention_fwd
 * (o = the forward output); dq,dk,dv: bf16 outputs with their own leading dimensions (head h at columns h*hd).
 * Probabilities are recomputed on chip.  dbias_q/k/v (each may be NULL): fp32 [H*hd] vectors the column sums of
 * dq/dk/dv are ADDED to (in_proj_bias / query_linear.bias / key_linear.bias gradients).
 * Supported: (64,64) (48,48) (32,32) with dv, (112,8) without. */
int ibm_attention_bwd_long(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                           const void* o, int64_t ldo, const void* d_o, int64_t ld_do, void* dq, int64_t lddq,
                           void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t n_win, int32_t T, int32_t H,
                           int32_t hd_qk, int32_t hd_v, float scale, float* dbias_q, float* dbias_k, float* dbias_v,
                           void* stream);

/* ---- optimizers  (src/cli/train.py:183-197, 284; torch.optim defaults) ------------------------ */

/* One fused step over a flat fp32 parameter arena: reads grad (scaled by grad_scale, e.g.
 * 1/world_size after an allreduce-sum), updates param and optimizer state in place, and writes
 * the bf16 shadow copy used by the GEMMs.  kind: 0 rmsprop(alpha .99, eps 1e-8), 1 adam(.9,.999,
 * 1e-8), 2 sgd, 3 adagrad(eps 1e-10), 4 adadelta(rho .9, eps 1e-6), 5 adamax(.9,.999,1e-8).
 * state0/state1: fp32 arenas (unused ones may be NULL).  step = 1-based step count. */
int ibm_optimizer_step(int32_t kind, float* param, const float* grad, float* state0, float* state1,
                       void* param_bf16, int64_t n, float lr, float grad_scale, int64_t step,
                       void* stream);
/* Same step with the 1-based step count read from DEVICE memory (int64[1]) at execution time: a captured CUDA graph of the
 * training step can be replayed although Adam / Adamax bias corrections change every step. */
int ibm_optimizer_step_dev(int32_t kind, float* param, const float* grad, float* state0, float* state1,
                           void* param_bf16, int64_t n, float lr, float grad_scale, const int64_t* step_dev,
                           void* stream);
/* *counter_dev += inc (one thread): the device-resident step counter of graph-replayed training steps. */
int ibm_counter_add(int64_t* counter_dev, int64_t inc, void* stream);

/* ---- diagnostics ---------------------------------------------------------------------------- */

/* How many thread-block clusters of `cluster_size` CTAs of the 256 x 256 GEMM kernel can be resident at once
 * (cudaOccupancyMaxActiveClusters).  On B200: 74 pairs, 33 quads, 15 octets — why the GEMM uses CTA pairs and
 * two-tile work items rather than larger multicast clusters.  Returns -1 on error. */
int ibm_debug_gemm_max_clusters(int32_t cluster_size);

#ifdef __cplusplus
}
#endif
#endif /* IBM_B200_H_ */
