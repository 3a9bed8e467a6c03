/* csmae_b200 -- C-ABI of the B200-native Cross-Scale MAE pretraining hot path.
 *
 * The reference (aicip/Cross-Scale-MAE) is pure Python/PyTorch: it has no native library and no FFI.
 * Its "operator interface" for this path is the set of ATen calls made by
 * models_mae/MAE_ViT_{Shared,Baseline,MsLd,MsLdCeCd}.py, models_mae/MLP.py and util/contrast_loss.py.
 * Every entry point below replaces one group of those calls; the citation next to each one is the
 * reference file:line whose arithmetic it reproduces (paths relative to the upstream repo root).
 * The Python binding a maintainer would add (ctypes over torch tensors' data_ptr()) is shown in
 * INTEGRATION.md and implemented in cross-scale-mae_b200/csmae_b200/_native.py.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers into caller-owned memory;
 *   - activations are contiguous row-major [rows, features] with rows = image-major tokens;
 *   - "bf16" buffers hold __nv_bfloat16, residual stream / statistics / losses / gradients are f32,
 *     ids_restore is int64 (interchangeable with torch.gather users), mask is f32;
 *   - every call is asynchronous on `stream` and never synchronises with the host;
 *   - return value: 0 on success, negative on error (csm_last_error() holds the message;
 *     no exception ever crosses the boundary);
 *   - the library contains sm_100a code only: csm_device_check() refuses any other device and there
 *     is no CPU or generic-GPU fallback.
 */
#ifndef CSMAE_B200_H
#define CSMAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* csm_stream_t; /* == cudaStream_t */

/* ---- runtime ------------------------------------------------------------------------------ */
const char* csm_last_error(void);
int csm_version(void);
/* returns the SM count (> 0) when `device` is an sm_100 part, else a negative error */
int csm_device_check(int device);

/* ---- tcgen05 GEMMs: nn.Linear forward / dgrad / wgrad ---------------------------------------
 * timm Block qkv/proj/fc1/fc2 (MAE_ViT_Baseline.py:160-188), patch_embed.proj as a GEMM (:75-77,245),
 * decoder_embed (:270), decoder_pred (:295), predictor Linears (MLP.py:6,9).
 * epilogue codes: 0 out_bf16 = bf16(acc + bias)
 *                 1 h = bf16(acc + bias); out_bf16 = bf16(gelu'(h)), aux_bf16 = bf16(gelu(h)) (fc1 + nn.GELU, erf form)
 *                 2 out_f32  = aux_f32 + bf16(acc + bias)                                  (x = x + proj/fc2)
 *                 3 out_bf16 = bf16(bf16(acc) * aux_bf16), aux = the gelu'(h) kept by epilogue 1   (dgrad only)
 *                 5 out_f32  = acc + bias
 */
int csm_linear_fwd(const void* x_bf16, const void* w_bf16, const float* bias, void* out, void* aux, int M, int N,
                   int K, int epilogue, csm_stream_t stream);
/* dX[M,K] = dY[M,N] . W[N,K]  (W read in its stored [N,K] layout) ; epilogue 0, 3 or 5 */
int csm_linear_dgrad(const void* dy_bf16, const void* w_bf16, void* dx, const void* aux, int M, int N, int K,
                     int epilogue, csm_stream_t stream);
/* Optional stream-K workspace for csm_linear_fwd / csm_linear_dgrad (no reference counterpart: cuBLAS keeps its own).
 * When the 256 x BN output tiles of a problem do not fill whole rounds of the 74 SM pairs, the k-block stream is cut
 * evenly between the clusters instead and a unit cut in two is finished through an fp32 partial tile in this buffer.
 * `ws` = csm_gemm_workspace_bytes(num_sms) bytes of ZEROED device memory, 256-byte aligned, owned by the caller and kept
 * alive until csm_gemm_set_workspace(NULL, 0).  Contract: while a workspace is registered, forward / dgrad GEMMs must
 * be issued on ONE stream at a time (csm_linear_wgrad never touches it and may run concurrently).  Without a
 * workspace the GEMMs use whole-tile scheduling only. */
int csm_gemm_workspace_bytes(int num_sms);
int csm_gemm_set_workspace(void* ws, long long bytes);
/* dW[N,K] += dY[rows,N]^T . X[rows,K]   (f32, split over the token rows, red.global.add) */
int csm_linear_wgrad(const void* dy_bf16, const void* x_bf16, float* dw, int rows, int N, int K, int num_sms,
                     csm_stream_t stream);
/* db[N] += column sums of dY[rows,N]; rows with (row % skip_period == 0) are skipped when skip_period > 0 */
int csm_colsum_bf16(const void* dy_bf16, float* db, int rows, int N, int skip_period, int num_sms,
                    csm_stream_t stream);

/* ---- masking / token shuffles ---------------------------------------------------------------
 * random_masking: MAE_ViT_Shared.py:57-84 (stable-by-index argsort of the caller's noise) */
int csm_random_masking(const float* noise, int nimg, int L, int keep, long long* ids_restore, int* ids_shuffle,
                       float* mask, csm_stream_t stream);
/* in-model random-resized-crop (MAE_ViT_MsLd.py:29-35,52): the batch-wide crop box (top, left, h, w) of
 * imgs [planes = N*C, H, W] resized to [planes, S, S] with torchvision's bilinear + antialias filter */
int csm_resized_crop(const float* imgs, float* out, int planes, int H, int W, int top, int left, int h, int w, int S,
                     csm_stream_t stream);
/* fixed 2-D sin-cos position table [cls_token + G*G, embed_dim] f32 (util/pos_embed.py:16-63; w-coordinate first, cls
 * row zero), evaluated in fp64 and rounded once -- init-time (MAE_ViT_Baseline.py:203-218) */
/* zeroes nbytes of device memory asynchronously (a memset node when captured into a CUDA graph) */
int csm_zero_async(void* ptr, long long nbytes, csm_stream_t stream);
int csm_sincos_pos_embed(float* out, int embed_dim, int grid_size, int cls_token, csm_stream_t stream);
/* kept patches -> GEMM operand rows [(nimg*(keep+1)), C*p*p] in Conv2d (c,py,px) order; cls slot rows zero
 * (timm PatchEmbed conv, MAE_ViT_Baseline.py:75-77,245, fused with the masking gather :251) */
int csm_patch_gather(const float* imgs, const int* ids_shuffle, void* out_bf16, int nimg, int C, int H, int p, int L,
                     int keep, csm_stream_t stream);
/* + pos-embed of the kept patch, cls token row (MAE_ViT_Baseline.py:248,254-256) */
int csm_encoder_assemble(const void* emb_bf16, const int* ids_shuffle, const float* pos, const float* cls, float* x,
                         int nimg, int L, int keep, int D, csm_stream_t stream);
/* mask tokens + un-shuffle + decoder pos-embed (MAE_ViT_Baseline.py:273-283) and its backward */
int csm_decoder_assemble(const void* demb_bf16, const long long* ids_restore, const float* mask_token,
                         const float* dpos, float* y, int nimg, int L, int keep, int Dd, csm_stream_t stream);
int csm_decoder_assemble_bwd(const float* dy, const int* ids_shuffle, void* d_demb_bf16, float* d_mask_token,
                             int nimg, int L, int keep, int Dd, csm_stream_t stream);
/* gradient entering the (un-normed, MAE_ViT_Baseline.py:264) encoder output: decoder_embed dgrad + NT-Xent feature grad */
int csm_encoder_out_grad(const void* d_enc_bf16, const float* d_feat, float* dx, void* dx_bf16, int nimg, int Se,
                         int D, csm_stream_t stream);
int csm_cls_grad(const float* dx, float* d_cls, int nimg, int Se, int D, csm_stream_t stream);

/* ---- LayerNorm (eps 1e-6, MAE_ViT_Baseline.py:43-45) and casts -------------------------------- */
int csm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16, float* out_f32,
                      float* mean, float* rstd, int rows, int D, float eps, csm_stream_t stream);
/* dres_out = dres_in + LN'(dy_bf16 + dy2_f32); dgamma/dbeta accumulate; dcolsum (nullable) accumulates the column
 * sums of the bf16 copy dres_bf16 = the bias gradient of the Linear (attn.proj / mlp.fc2) whose dY it is */
int csm_layernorm_bwd(const void* dy_bf16, const float* dy2_f32, const float* x, const float* mean, const float* rstd,
                      const float* gamma, const float* dres_in, float* dres_out, void* dres_bf16, float* dgamma,
                      float* dbeta, float* dcolsum, int rows, int D, int num_sms, csm_stream_t stream);
int csm_cast_multi(const void* table_dev, int num_tensors, int blocks_per_tensor, csm_stream_t stream);
int csm_cast_f32_bf16(const float* src, void* dst_bf16, long long n, csm_stream_t stream);

/* ---- attention (timm 0.4.12 Attention; MAE_ViT_Baseline.py:160-188) -------------------------- */
int csm_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int H, int head_dim,
                      csm_stream_t stream);
/* dbias (nullable, f32 [3*H*head_dim]) accumulates the column sums of dqkv over the tokens = attn.qkv.bias gradient */
int csm_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                      float* delta_scratch, void* dqkv_bf16, float* dbias, int B, int S, int H, int head_dim,
                      csm_stream_t stream);

/* ---- losses ---------------------------------------------------------------------------------
 * masked per-patch MSE against the image read in patch order (MAE_ViT_Shared.py:24-39,97-120) */
int csm_recon_loss_fwd(const void* pred_full_bf16, const float* imgs, const float* mask, float* loss_sum, int nimg,
                       int C, int H, int p, int L, int norm_pix, csm_stream_t stream);
int csm_recon_loss_bwd(const void* pred_full_bf16, const float* imgs, const float* mask, void* dpred_bf16,
                       const float* grad_scalar, float coef, int nimg, int C, int H, int p, int L, int norm_pix,
                       csm_stream_t stream);
/* loss = sum_i loss_terms[i] * coefs[i], i < n <= 32: the scalar of MAE_ViT_MsLdCeCd.py:62-69 from the accumulator array
 * every loss kernel adds its raw sum into */
int csm_loss_finalize(const float* loss_terms, const float* coefs, float* loss, int n, csm_stream_t stream);
/* cross-scale decoder MSE, target not detached (MAE_ViT_MsLdCeCd.py:57-59) */
int csm_cross_mse_fwd(const void* cp_bf16, const float* tgt, float* loss_sum, int rows, int Sd, int Dd,
                      csm_stream_t stream);
int csm_cross_mse_bwd(const void* cp_bf16, const float* tgt, void* d_cp_bf16, float* d_tgt, const float* grad_scalar,
                      float coef, int rows, int Sd, int Dd, csm_stream_t stream);
/* BatchNorm1d(num_patches) over [N, L, Hp] + ReLU (MLP.py:7-8).  training = 1: batch statistics (+ running-stat update);
 * training = 0: running statistics (mean / rstd still written for the backward, which then treats them as constants:
 * dh = gamma * rstd * dy, as nn.BatchNorm1d in eval mode) */
int csm_bn_patch_fwd(const void* h_bf16, const float* gamma, const float* beta, void* out_bf16, float* mean,
                     float* rstd, float* running_mean, float* running_var, int N, int L, int Hp, float eps,
                     float momentum, int training, csm_stream_t stream);
int csm_bn_patch_bwd(const void* h_bf16, const void* out_bf16, const void* d_out_bf16, const float* gamma,
                     const float* mean, const float* rstd, void* dh_bf16, float* dgamma, float* dbeta, int N, int L,
                     int Hp, int training, csm_stream_t stream);
/* NT-Xent on the token-mean encoder features (util/contrast_loss.py:71-101; MAE_ViT_MsLdCeCd.py:62-69) */
int csm_ntxent_fwd(const float* enc_out, float* zhat, float* fnorm, float* neg, float* loss_sum, int B, int Se, int D,
                   float tau, float eps, csm_stream_t stream);
int csm_ntxent_bwd(const float* zhat, const float* fnorm, const float* neg, const float* grad_scalar, float* d_feat,
                   int B, int D, float tau, float eps, csm_stream_t stream);

/* ---- optimizer step (row f1: util/misc.py:299-355 + torch.optim.AdamW at main_pretrain.py:426-427) ------------
 * table_dev: device array of 80-byte entries {float* p; const float* g; float* m; float* v; bf16* shadow_or_null;
 * int64 n; float lr, wd, beta1, beta2, eps, bias_correction1, sqrt(bias_correction2), grad_scale};
 * chunks_dev: device array of {int tensor_index, int chunk_index} work items of 8192 elements */
int csm_adamw_multi(const void* table_dev, const void* chunks_dev, int num_chunks, const float* ctl_dev, int num_sms,
                    csm_stream_t stream);
/* ctl_dev (nullable, device f32[2]): {gradient multiplier (GradScaler's 1/scale x clip coefficient), found_inf}: the
 * multiplier is applied to every gradient on the fly and a nonzero found_inf skips the whole update -- the AMP step of
 * util/misc.py:314-326 with no host synchronisation between backward and step.
 * csm_grad_stats_f32: the matching single pass over the flat gradient buffer: stats[0] += sum((x * *mul_dev)^2)
 * (global gradient norm of the UNSCALED gradients), stats[1] = 1 when any element is not finite. */
int csm_grad_stats_f32(const float* x, long long n, const float* mul_dev, float* stats, int num_sms,
                       csm_stream_t stream);
/* scalar part of the AMP step (GradScaler.unscale_/step/update, util/misc.py:314-326) on the device: stats = the pair
 * csm_grad_stats_f32 wrote, state = {scale, growth_tracker, 1 / scale} (updated in place), ctl = the pair
 * csm_adamw_multi reads, norm_out[0] = global gradient norm; clip <= 0 disables clipping */
int csm_amp_update(const float* stats, float* state, float* ctl, float* norm_out, float clip, float growth,
                   float backoff, int interval, csm_stream_t stream);
/* out[0] += sum of squares of x[0..n)  (global gradient norm of the flat gradient buffer in one pass) */
int csm_sumsq_f32(const float* x, long long n, float* out, int num_sms, csm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CSMAE_B200_H */
