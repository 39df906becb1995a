/* genrl_b200 — C-ABI of the B200-native GenRL/Dreamer world-model training hot path.
 *
 * The reference (mazpie/genrl) has no FFI layer: its "kernels" are PyTorch library calls made from
 * agent/dreamer_utils.py, agent/dreamer.py and agent/video_utils.py.  Each entry point below replaces the group of
 * torch calls cited beside it.  Conventions:
 *   - every function returns 0 on success; on failure the message is in genrl_last_error();
 *   - `stream` is a cudaStream_t passed as void* (the host passes torch.cuda.current_stream().cuda_stream);
 *   - all pointers are DEVICE pointers owned by the caller; no function allocates or synchronises;
 *   - dtype codes: 0 = fp32, 1 = bf16 (activation storage type; statistics, logits, losses and gradients of
 *     parameters are always fp32);
 *   - matrices are row-major with an explicit leading dimension (`ld*`, in elements).
 * The ctypes binding (genrl_b200/_lib.py) is generated from this file, so it is the single source of truth.
 */
#ifndef GENRL_B200_H
#define GENRL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* genrl_last_error();
int genrl_abi_version();
int genrl_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* smem_optin);
int genrl_debug_occupancy();   /* prints kernel attributes / occupancy to stderr (debug aid) */

/* C[M,N] (+)= op(A) op(W)^T + bias.  Replaces nn.Linear / lowered nn.Conv2d / nn.ConvTranspose2d forward, dgrad and
 * wgrad (dreamer_utils.py:339-346, :733-735, :760, :586, :668).  in_dtype 1 -> tcgen05 + TMA kernel, 0 -> fp32 SIMT. */
int genrl_gemm(void* stream, int M, int N, int K, const void* A, long long lda, int transA, const void* W, long long ldw, int transW,
               int in_dtype, void* C, long long ldc, int c_dtype, const float* bias, int accumulate);

/* The same GEMM on the persistent 2-CTA kernel (tcgen05.mma.cta_group::2, 256x256 tile per CTA pair, csrc/gemm_pair.cu): bf16 operands,
 * A [M,K] K-major, W [N,K] (transW = 0) or [K,N] (transW = 1), no accumulate.  Returns 2 if the shape is not eligible. */
int genrl_gemm_pair(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, int transW, void* C,
                    long long ldc, int c_dtype, const float* bias);

/* Weight-gradient form on the same kernel: C[M,N] (fp32) += A^T W with A stored [K, M] and W stored [K, N] (K = rows of the activation slabs),
 * split over K with an fp32 atomic epilogue.  M, N multiples of 128.  Returns 2 if the shape is not eligible. */
int genrl_gemm_pair_wgrad(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, float* C, long long ldc);

/* fp32 parity mode on the tensor cores ("bf16x3"): x = hi + mid + lo with three bf16 terms (24 significant bits); the host side runs the six
 * products a_i w_j, i + j <= 2, on the bf16 tcgen05 GEMM with fp32 accumulation into C. */
int genrl_split3_bf16(void* stream, const float* src, long long lds, void* hi, void* mid, void* lo, long long ldd, long long rows, int cols);

/* Linear + LayerNorm(+SiLU) fused (bf16, N multiple of 128 up to 1024): the N/128 column-tile CTAs of a row stripe form a thread-block
 * cluster and merge their per-row (mean, M2) partials through distributed shared memory.  Replaces Linear -> NormLayer -> act
 * (dreamer_utils.py:738-744, :339-346).  pre is kept for the backward pass; stats[M][2] = (mean, rstd). */
int genrl_gemm_ln_fwd(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                      const float* gamma, const float* beta, float eps, int act, void* pre, long long ldpre, void* y, long long ldy,
                      float* stats);

/* Chain of L layers y_l = act(LayerNorm(y_{l-1} W_l^T + b_l)) in ONE persistent cluster kernel (bf16, N = 256/512/768/1024; K_0 free, K_l = N):
 * MLP trunks (dreamer_utils.py:718-747) and the single Linear+NormLayer+SiLU blocks of the RSSM (:339-346).  W / bias / gamma / beta / pre / y /
 * stats are host arrays of L device pointers; pre_l (pre-LayerNorm values) and stats_l = (mean, rstd) per row are kept for the backward. */
int genrl_mlp_chain_fwd(void* stream, int M, int N, int L, const void* x, long long ldx, int K0, const void* const* W, const long long* ldw,
                        const float* const* bias, const float* const* gamma, const float* const* beta, float eps, int act,
                        void* const* pre, long long ldpre, void* const* y, long long ldy, float* const* stats);

/* LayerNorm(+SiLU) rows: NormLayer dreamer_utils.py:844-859 / ImgChLayerNorm :1031-1040 + act (:744, :589). */
int genrl_ln_act_fwd(void* stream, const void* x, long long ldx, const float* gamma, const float* beta, float eps, int act, void* y,
                     long long ldy, float* stats, long long M, int D, int dtype);
int genrl_ln_act_bwd(void* stream, const void* dy, long long lddy, const void* x, long long ldx, const float* gamma, const float* beta,
                     const float* stats, int act, void* dx, long long lddx, float* dgamma, float* dbeta, float* dlin_bias, long long M,
                     int D, int dtype);
int genrl_ln_colgrad(void* stream, const void* dy, long long lddy, const void* x, long long ldx, const void* dx, long long lddx,
                     const float* gamma, const float* beta, const float* stats, int act, float* dgamma, float* dbeta, float* dlin_bias,
                     long long M, int D, int dtype);
int genrl_colsum(void* stream, const void* x, long long ldx, float* out, long long M, int D, int dtype);

/* GRUCell tail (LayerNorm(3D) + gates): dreamer_utils.py:777-785. */
int genrl_gru_fwd(void* stream, const void* z, long long ldz, const float* gamma, const float* beta, const float* hprev, long long ldh,
                  float* hout, long long ldho, void* hout_t, long long ldht, float* stats, long long M, int D, int dtype);
int genrl_gru_bwd(void* stream, const float* dh, long long lddh, const void* z, long long ldz, const float* gamma, const float* beta,
                  const float* hprev, long long ldh, const float* stats, void* dz, long long lddz, float* dhprev, long long lddhp,
                  float* dgamma, float* dbeta, long long M, int D, int dtype);
/* Frozen-weights GRU backward (D = 1024) whose incoming gradient is dh + e1 + e2 + e3: e1 / e2 / e3 (optional, activation dtype) are the other
 * contributions to d loss / d deter_i inside the imagination rollout's reverse loop (agent/dreamer.py:254-287): the feat-slice gradient, the
 * h half of the previous step's GRU dgrad and the dgrad of the prior's output block.  Returns 2 when the shape / alignment is not eligible. */
int genrl_gru_bwd_sum(void* stream, const float* dh, long long lddh, const void* e1, long long lde1, const void* e2, long long lde2,
                      const void* e3, long long lde3,
                      const void* z, long long ldz, const float* gamma, const float* beta, const float* hprev, long long ldh,
                      const float* stats, void* dz, long long lddz, float* dhprev, long long lddhp, long long M, int D, int dtype);

/* OneHotDist sample / mode with straight-through gradient: dreamer_utils.py:177-197, :410-416.
 * noise != NULL: q ~ Exp(1) supplied by the caller (parity mode); else Philox(seed, offset). */
int genrl_cat_sample_fwd(void* stream, const float* logit, const float* noise, unsigned long long seed, unsigned long long offset,
                         const unsigned long long* step_dev, int sample,
                         void* stoch, long long ld_stoch, float* stoch32, long long ld_stoch32, int* idx, long long rows, int S, int K,
                         int dtype);
int genrl_cat_sample_bwd(void* stream, const float* logit, const void* dstoch, long long ld_ds, int sample, float* dlogit, int accumulate,
                         long long rows, int S, int K, int dtype);
/* rollout backward (K = 32): straight-through backward of ds1 + ds2 (ds2 may be NULL), dlogit written in `dtype` with leading dimension ld_dl */
int genrl_cat_sample_bwd2(void* stream, const float* logit, const void* ds1, long long ld1, const void* ds2, long long ld2, void* dlogit,
                          long long ld_dl, long long rows, int S, int K, int dtype);

/* Philox4x32-10 standard-normal fill (Normal.rsample noise, dreamer_utils.py:814-819; LAFITE noise, video_utils.py:143-146) and
 * the device-side step counter that keys every in-kernel generator (so CUDA-graph replays draw fresh noise). */
int genrl_fill_normal(void* stream, float* out, long long n, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* step_dev);
int genrl_counter_add(void* stream, unsigned long long* counter, unsigned long long inc);

/* EnsembleRSSM.kl_loss + entropies: dreamer_utils.py:534-555, dreamer.py:249-250.  out4 = {loss, mean value, ent lhs, ent rhs}. */
int genrl_kl_fwd(void* stream, const float* lhs, const float* rhs, float free_bits, float* value, float* out4, long long rows, int S, int K);
int genrl_kl_bwd(void* stream, const float* lhs, const float* rhs, const float* value, float free_bits, float w_lhs, float w_rhs,
                 const float* gscale, float* dlhs, float* drhs, long long rows, int S, int K);

/* TwoHotDist.log_prob (mode 0) / .mean (mode 1): dreamer_utils.py:120-171. */
int genrl_twohot_fwd(void* stream, const float* logits, long long ldl, const float* x, int mode, float* out, long long rows);
int genrl_twohot_bwd(void* stream, const float* logits, long long ldl, const float* x, int mode, const float* dout, float scale,
                     float* dlogits, long long lddl, long long rows);

/* DistLayer 'normal' head + rsample + entropy: dreamer_utils.py:814-819. */
int genrl_actor_head_fwd(void* stream, const float* out, long long ldo, const float* eps, int use_mean, float* action, long long lda,
                         void* action_t, long long ldat, float* ent, long long rows, int A, int dtype);
int genrl_actor_head_bwd(void* stream, const float* out, long long ldo, const float* eps, const void* daction, long long ldda,
                         const float* dent, float dent_scale, void* dout, long long lddo, int dout_dtype, long long rows, int A, int dtype);

/* lambda_return reverse scan: dreamer_utils.py:228-253. */
int genrl_lambda_return_fwd(void* stream, const float* reward, const float* value, float gamma, float lambda, float* target, int H,
                            long long N);
int genrl_lambda_return_bwd(void* stream, const float* dtarget, const float* dt_scale, float gamma, float lambda, float* dreward,
                            float* dvalue, int H, long long N);

/* Replay window gather (tools/replay.py:223-236): batch[b] = store[start[b] : start[b] + L] for B windows of L rows of row_bytes each; store and
 * batch are device buffers, start is a device array of B first-row indices into the flat episode store of one key. */
int genrl_replay_gather(void* stream, const void* store, const long long* start, long long row_bytes, int B, int L, void* batch);

/* Video-text reward of the imagination behaviour: tools/genrl_utils.py:240-243 (max_cosine), :250-277 (compute_reward), :344-366
 * (align_sequence: best of T - n_frames window offsets, target delayed by that offset).  pa / pt: decoder-input projections of the imagined and
 * of the cached target latents, [T, N, E] time-major in `dtype`; reward fp32 [T, N]; shift int [N] (the chosen offset, kept for the backward). */
int genrl_vlm_reward_fwd(void* stream, const void* pa, const void* pt, float* reward, int* shift, int T, long long N, int E, int n_frames,
                         int align, int dtype);
int genrl_vlm_reward_bwd(void* stream, const void* pa, const void* pt, const int* shift, const float* dreward, void* dpa, int T, long long N,
                         int E, int dtype);

/* MSEDist.log_prob against uint8 observations with the /255-0.5 preprocessing fused: dreamer_utils.py:74-83, dreamer.py:294-295. */
int genrl_mse_fwd(void* stream, const float* mean, const unsigned char* obs, float* like, long long rows, int E);
int genrl_mse_bwd(void* stream, const float* mean, const unsigned char* obs, const float* dlike, const float* gscale, float scale,
                  float* dmean, long long rows, int E);

int genrl_mean_std(void* stream, const float* x, long long n, float* out2);
int genrl_mean_scale(void* stream, const float* x, long long n, float scale, float* out, int accumulate);
int genrl_copy2d(void* stream, const void* src, long long lds, int src_dtype, void* dst, long long ldd, int dst_dtype, long long rows,
                 int cols, const float* row_scale);

/* dst[r,:] (+)= scale[r] * src[idx[r],:] (idx<0 -> 0): time shift + is_first masking of obs_step (dreamer_utils.py:432-436). */
int genrl_gather_rows(void* stream, const void* src, long long lds, int src_dtype, const int* idx, const float* scale, void* dst,
                      long long ldd, int dst_dtype, long long rows, int cols, int accumulate);

/* Conv lowering: Encoder dreamer_utils.py:580-621, Decoder :656-706.  x_dtype -1 = uint8 NCHW with preprocessing fused. */
int genrl_im2col(void* stream, const void* x, int x_dtype, void* col, long long ldcol, int col_dtype, int N, int H, int W, int C, int k,
                 int Ho, int Wo);
int genrl_col2im(void* stream, const void* col, long long ldcol, int col_dtype, const float* bias, void* y, int y_dtype, int nchw_out, int N,
                 int Hi, int Wi, int C, int k, int Ho, int Wo);
int genrl_nchw_to_nhwc(void* stream, const float* x, void* y, int y_dtype, int N, int C, int H, int W);

/* Implicit-GEMM convolutions on tcgen05 (no im2col / col2im buffers), NHWC bf16 activations; replace the lowered
 * nn.Conv2d(k4,s2) / nn.ConvTranspose2d(k5|k6,s2) passes of Encoder/Decoder (dreamer_utils.py:580-589, :656-671).
 * Virtual im2col element: col[(n,yy,xx), ((a*tw+b)*C+c)] = in[n, yy*sy + a*ts, xx*sy + b*ts, c] (zero outside the image).
 *   genrl_conv_igemm_fwd  : out[row, col] = sum_k col[row,k] Wm[col,k] + bias; ncls==1: gather form (Conv2d fwd, ConvT dgrad);
 *                           ncls==4: sub-pixel form, column block cls=(py,px) goes to pixel (2yy+py, 2xx+px) (ConvT fwd, Conv2d dgrad).
 *                           out_mode 0: NHWC bf16 rows [Nimg*OH*OW, Cout]; 1: NCHW fp32.
 *   genrl_conv_igemm_wgrad: dw[s, k] += sum_rows S[row, s] col[row, k]  (fp32 atomics, split over the rows).
 *   genrl_conv_pack_parity / _pack_gather / _unpack_wgrad: filter layouts of the two forms from/to the [A,B,k,k] fp32 master. */
int genrl_conv_igemm_fwd(void* stream, const void* in, int Nimg, int IH, int IW, int C, int RH, int RW, int sy, int ts, int th, int tw,
                         const void* Wm, long long ldw, int Cpad, int Cout, int ncls, int OH, int OW, void* out, int out_mode,
                         const float* bias);
int genrl_conv_igemm_wgrad(void* stream, const void* in, int Nimg, int IH, int IW, int C, int RH, int RW, int sy, int th, int tw,
                           const void* S, long long lds, int Cs, float* dw, long long lddw);
int genrl_conv_pack_parity(void* stream, const float* w, void* dst, int A, int B, int k, int Cpad);
int genrl_conv_pack_gather(void* stream, const float* w, void* dst, int A, int B, int k, int Cp);
int genrl_conv_unpack_wgrad(void* stream, const float* dwp, long long ld, float* grad, int A, int B, int k, int Cp);
int genrl_nchw_to_nhwc_pad(void* stream, const float* x, void* y, long long N, int C, int H, int W, int Cp);
/* uint8 NCHW observation -> (x/255 - 0.5) bf16 NHWC, channels zero-padded to 8 (WorldModel.preprocess dreamer.py:294-295 + layout change). */
int genrl_u8_nchw_to_nhwc8(void* stream, const void* x, void* y, long long N, int C, int H, int W);
int genrl_permute_abp(void* stream, const void* src, int src_dtype, void* dst, int dst_dtype, long long A, int B, int P, int accumulate);

/* Optimizer.__call__: dreamer_utils.py:892-932 (clip -> weight decay -> Adam) on flat fp32 buffers. */
int genrl_grad_sumsq(void* stream, const float* g, long long n, float* sumsq_out, int zero_first);
int genrl_scale(void* stream, float* p, long long n, float s);
int genrl_adam_step(void* stream, float* p, float* g, float* m, float* v, long long n, float lr, float eps, float clip, const float* sumsq,
                    float* state, int finish, float decay);
int genrl_pack_weights(void* stream, const void* table, int n_entries, int dst_dtype);
int genrl_pack_entry_size();

/* RewardEMA + actor objective: dreamer_utils.py:1014-1029, dreamer.py:392-429. */
int genrl_quantile_ema(void* stream, const float* x, long long n, float q_lo, float q_hi, float alpha, float* ema, float* out4);
int genrl_actor_objective(void* stream, const float* target, const float* ent, const float* offscale, float ent_scale, int H, long long N,
                          float* out4, float* dcoef);

/* Persistent time-loop of EnsembleRSSM.observe / VideoSSM.update: dreamer_utils.py:362-371 + :432-473 + :771-785. */
long long genrl_rssm_scan_workspace_bytes(int B, int T, int D);
int genrl_rssm_scan_set_debug(long long* dbg);   /* profiling aid: per-phase globaltimer stamps of block 0, dbg[T][6] or NULL */
int genrl_rssm_scan_fwd(void* stream, int B, int T, int D, const float* gx, const float* Wh, long long ldw, const float* gamma,
                        const float* beta, const float* mask, const float* h0, float* deter, void* deter_t, long long ld_dt, void* hprev_t,
                        long long ld_hp, int t_dtype, float* z, float* stats, void* workspace);
int genrl_rssm_scan_bwd(void* stream, int B, int T, int D, const float* dh_ext, const float* z, const float* stats, const float* Wh,
                        long long ldw, const float* gamma, const float* beta, const float* mask, const float* deter, const float* h0,
                        void* dz_t, long long ld_dz, int t_dtype, float* dh0, float* dgamma, float* dbeta, void* workspace);

/* Weight-streaming variant of the same time loop for hidden sizes whose recurrent matrix cannot stay resident in shared memory
 * (BASELINE configs[4]: deter = 4096): W_h (bf16 [3D, D], row stride ldw) is streamed from L2 / HBM once per time step straight into
 * mma.sync A fragments, the <= 16 batch rows are the N operand.  hprev / z: bf16 outputs [B*T, D] / [B*T, 3D]; workspace as for genrl_rssm_scan_fwd. */
int genrl_rssm_scan_stream_supported(int B, int D);      /* 1 / 0, no launch */
/* backward: dz (bf16 [B*T, 3D]) = gradient of the pre-LayerNorm activations; WhT = TRANSPOSED recurrent matrix, bf16 [D, 3D]; returns 2 when
 * D % 512 != 0 (the caller then walks the steps with genrl_gru_bwd + genrl_gemm) */
int genrl_rssm_scan_stream_bwd(void* stream, int B, int T, int D, const float* dh_ext, const void* z, long long ld_z, const float* stats,
                               const void* WhT, long long ldwt, const float* gamma, const float* beta, const float* mask, const float* deter,
                               const float* h0, void* dz, long long ld_dz, float* dh0, float* dgamma, float* dbeta, void* workspace);
int genrl_rssm_scan_stream_fwd(void* stream, int B, int T, int D, const float* gx, const void* Wh, long long ldw, const float* gamma,
                               const float* beta, const float* mask, const float* h0, float* deter, void* hprev, long long ld_hp,
                               void* z, long long ld_z, float* stats, void* workspace);

#ifdef __cplusplus
}
#endif
#endif
