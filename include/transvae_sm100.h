/* C ABI of libtransvae_sm100.so -- the drop-in boundary of the B200-native TransVAE hot path.
 *
 * The reference (benabbouosama/DEEPL-Project) has no FFI layer: its hot path is the Python nn.Module
 * API of package `transvae` (transvae/__init__.py:5-9) and every FLOP is dispatched to torch.nn /
 * ATen.  This header is what the host-side mirror of that API (deepl-project_b200/transvae/) binds
 * with ctypes; each entry point names the reference call it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise.
 *  - the caller (PyTorch's caching allocator) owns every buffer, including workspaces.
 *  - every launch is asynchronous on the `stream` argument (a cudaStream_t passed as void*).
 *  - return 0 on success, negative on error; tvae_last_error() then describes it.  Nothing throws.
 *  - there is NO CPU fallback: without an sm_100 device the compute entry points return an error.
 *  - activations are NHWC bf16 ("pixel-major": [B, H, W, C] == token-major [B*H*W, C]).
 */
#ifndef TRANSVAE_SM100_H_
#define TRANSVAE_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVAE_ABI_VERSION 4
#define TVAE_MAX_TAPS 20
#define TVAE_MAX_PHASES 4

int tvae_abi_version(void);
const char* tvae_last_error(void);
/* 1 when the current CUDA device is compute capability 10.x, 0 otherwise (or no device). */
int tvae_device_ok(void);
int tvae_num_sms(void);
/* Leave `n` SMs free in the grids of the persistent (one CTA per SM) GEMM / weight-gradient kernels, e.g. for the
 * channels of an NCCL all-reduce running under the backward pass (DistributedDataParallel's overlap, train.py:672-674);
 * 0 restores full-width grids.  Returns the number of SMs the persistent kernels will use. */
int tvae_set_reserved_sms(int32_t n);

/* ---- pixel views ---------------------------------------------------------------------------
 * A view of a contiguous NHWC bf16 tensor [B, H, W, C].
 *   split == 0: the plain tensor.
 *   split == 1: the 2x2 "phase view" (2C, W/2, 2, H/2, B): element (b, 2h+p, 2w+q, c) is addressed as
 *               channel q*C+c of pixel (b, h, w) in phase p.  This is how stride-2 convolution,
 *               pixel_unshuffle / pixel_shuffle and nearest-2x upsampling are expressed without copies.
 * A flat [M, K] matrix is the view {B=1, H=1, W=M, C=K}. */
typedef struct {
  const void* ptr; /* NULL = absent */
  int32_t B, H, W, C;
  int32_t split;
} tvae_view;

/* One K-slab ("tap") of the implicit GEMM: the A operand for `kblocks` blocks of 64 input channels is
 * the view `map` (0 or 1) shifted by (dw, dh) pixels in phase `p`, starting at channel `c_off`; the
 * matching weights start at column `wk_off` of the packed [N, K_total] weight matrix. */
typedef struct {
  int32_t map, c_off, dw, p, dh, kblocks, wk_off;
} tvae_tap;

enum { TVAE_ACT_NONE = 0, TVAE_ACT_GELU = 1, TVAE_ACT_SILU = 2 };

/* Fused multi-tap implicit GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators):
 *
 *   acc[m, n]  = sum_taps sum_k  A_tap[m, k] * Wp[n, wk_off + k]
 *   v          = row_scale[m] * acc - row_shift[m] * col_sum[n] + bias[phase][n]    (each part optional)
 *   v          = act(v)
 *   v          = rope / q-scale (QKV projection only)
 *   out[m, n]  = v + residual[m, n]
 *
 * Backward use (act_grad != 0, no bias / affine / rope): the same launch computes an input gradient and applies the
 * derivative of the producing layer's activation, so dZ = dY * act'(Z) never needs its own pass over HBM:
 *   act_grad == 1:  out = acc * act'(res)                 (`res` holds the saved pre-activation Z)
 *   act_grad == 2:  out = (acc + res) * act'(z)           (`res` = a gradient to add first, `z` = plain [M, n_total] bf16)
 *
 * Replaces (reference file:line): nn.Linear (attention.py:43-48, conv.py:39,65), nn.Conv2d 1x1
 * (conv.py:55,59; upsample.py:42,103), nn.Conv2d 3x3 s1/s2 (blocks.py:34,37; upsample.py:34,36,95,97;
 * conv.py:57; encoder.py:52; decoder.py:49,94; transvae.py:76-77), F.pixel_unshuffle / F.pixel_shuffle /
 * nn.Upsample(nearest) (upsample.py:60,123,94), the three pre-projection LayerNorms and the RMSNorms
 * folded into the following projection (attention.py:71-73; blocks.py:146-149), RoPE2D
 * (attention.py:132-199), F.gelu / nn.SiLU and the residual adds (conv.py:86,93; blocks.py:68,146,149). */
typedef struct {
  tvae_view a0, a1;  /* A operands (a1 optional) */
  tvae_view out;     /* bf16 output view (ignored when out_f32 != NULL) */
  tvae_view res;     /* optional residual, same geometry as out */
  const void* w;     /* packed bf16 weights [n_total, k_total], K contiguous */
  int32_t n_total, k_total;
  int32_t num_phases;
  int32_t ntaps[TVAE_MAX_PHASES];
  tvae_tap taps[TVAE_MAX_PHASES][TVAE_MAX_TAPS];
  int32_t out_p[TVAE_MAX_PHASES];     /* phase coordinate of the output view per phase */
  int32_t out_c_off[TVAE_MAX_PHASES]; /* channel offset in the output view per phase */
  const float* bias;                  /* [num_phases][n_total] fp32 or NULL */
  int32_t act;
  const float* row_scale;             /* [M] or NULL */
  const float* row_shift;             /* [M] or NULL (requires col_sum) */
  const float* col_sum;               /* [n_total] */
  const float* rope_tab;              /* [max(H,W)][16][2] (cos, sin) fp32 or NULL */
  int32_t rope_C, rope_H, rope_W;     /* columns [0, 2*rope_C) are rotated; token grid rope_H x rope_W */
  float q_scale;                      /* columns [0, rope_C) are multiplied by this after RoPE */
  float* out_f32;                     /* optional: write fp32 NCHW [B, out_n, H, W] directly */
  int32_t out_n;
  int32_t act_grad;                   /* 0: forward epilogue; 1 / 2: multiply by act'(.) (see above), `act` names it */
  const void* z;                      /* act_grad == 2: pre-activation [M, n_total] bf16, rows in output-pixel order */
  /* Optional GroupNorm statistics of the OUTPUT (blocks.py:60,64 / decoder.py:128: the nn.GroupNorm that consumes this
   * convolution): gn_sums[B][gn_groups][2] receives (sum, sum of squares) of the stored bf16 values per image and
   * channel group (fp64, see tvae_groupnorm_stats), so the consumer needs tvae_groupnorm_apply only.  The library zeroes
   * the buffer.  The sums come out of the GEMM epilogue (no extra pass over HBM) when the launch runs on the CTA-pair kernel with a bias / bias+residual
   * epilogue, one image per 128-pixel tile and gn_groups <= 64; otherwise the library runs tvae_groupnorm_stats on the
   * output behind the GEMM.  Requires a bf16 output whose channel count equals n_total. */
  double* gn_sums;
  int32_t gn_groups;
  /* Optional second output of the training forward pass (act = GELU / SiLU, plain bias epilogue): `out` then receives the
   * PRE-activation acc + bias (what the backward pass differentiates through) and out_act -- bf16, same geometry as
   * `out` -- the activation of the stored bf16 values, both from one launch (conv.py:86,56,58; upsample.py:35,96). */
  void* out_act;
  /* Optional fused reduce pass of a GroupNorm backward (blocks.py:60-66): the launch is the input-gradient GEMM whose
   * output dh feeds the backward of act(GroupNorm(gnb_x)).  gnb_x: bf16, same geometry as `out`; gnb_sums: the forward
   * statistics [B][gnb_groups][2] fp64; gnb_part [B][n_total][2] fp32 receives per (image, channel) (sum dy, sum dy*xhat),
   * dy = dh * act'(gamma * xhat + beta) -- exactly what tvae_groupnorm_bwd_apply consumes.  On the CTA-pair kernel with
   * one image per 128-pixel tile the sums come out of the GEMM epilogue (x read behind an L2 tensor prefetch, per-tile partial rows
   * added in a fixed order); otherwise the library runs the stand-alone reduce pass behind the GEMM.  Plain epilogue only
   * (no bias / activation / residual). */
  const void* gnb_x;
  const double* gnb_sums;
  const float* gnb_gamma;
  const float* gnb_beta;
  float* gnb_part;
  int32_t gnb_groups;
  float gnb_eps;
  int32_t gnb_silu;
} tvae_mtgemm_desc;

int tvae_mtgemm(const tvae_mtgemm_desc* desc /* HOST pointer */, void* stream);

/* ---- attention ------------------------------------------------------------------------------
 * Flash-style attention forward, head_dim 64, non-causal (tcgen05 + TMEM + TMA).
 * qkv: bf16 [B, S, 3C] (q | k | v per token; q already rotated and scaled by 64^-0.5*log2(e), k rotated -- the
 * QKV tvae_mtgemm epilogue does both); out: bf16 [B, S, C]; lse (optional): fp32 [B, C/64, S], log2 domain.
 * Replaces F.scaled_dot_product_attention at attention.py:88-92. */
int tvae_attn_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t S, int32_t C, void* stream);

/* ---- HBM-bound kernels ------------------------------------------------------------------------ */
/* encoder.conv_in (encoder.py:52, nn.Conv2d(3, C0, 3, padding=1)) as a tensor-core GEMM: im2col of the NCHW fp32 image
 * [B,3,H,W] into bf16 cols [B*H*W, 64] = [27 taps hi | 27 taps lo | 1.0 | 0 x 9] (tap k = ci*9 + dy*3 + dx, hi + lo = the
 * fp32 pixel split in two bf16, zero outside the image).  The convolution is then tvae_mtgemm with a 1-tap K = 64 plan
 * and the packed weight [C0, 64] = [w | w | bias | 0]; its weight / bias gradient is tvae_mtgemm_wgrad on the same cols. */
int tvae_im2col_in(const float* x_nchw, void* cols, int32_t B, int32_t H, int32_t W, void* stream);
/* nn.GroupNorm(G, C) statistics (blocks.py:33,36; decoder.py:93): sums fp64 [B, G, 2] = (sum x, sum x^2).  fp32 block
 * partials (fixed order) are combined with fp64 atomics and the consumers form E[x^2] - mean^2 in fp64, so the result
 * does not depend on the order of the atomics (in fp32 the cancellation made identical runs differ by 1e-4 in rstd). */
int tvae_groupnorm_stats(const void* x_nhwc, double* sums, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream);
/* y = act(GroupNorm(x)) with act = SiLU (apply_silu=1) or identity; NHWC bf16 in/out (blocks.py:60-66). */
int tvae_groupnorm_apply(const void* x_nhwc, const double* sums, const float* gamma, const float* beta, void* y_nhwc,
                         int32_t B, int32_t HW, int32_t C, int32_t G, float eps, int32_t apply_silu, void* stream);
/* Both passes in one call (statistics, then apply).  `sums` [B, G, 2] is an OUTPUT (the backward pass reuses it).
 * Replaces nn.GroupNorm(32, C) + F.silu (blocks.py:60-66; decoder.py:128-129). */
int tvae_groupnorm_silu(const void* x_nhwc, const float* gamma, const float* beta, void* y_nhwc, double* sums, int32_t B,
                        int32_t HW, int32_t C, int32_t G, float eps, int32_t apply_silu, void* stream);
/* Per-token statistics for the folded RMSNorm (mode 0; blocks.py:149) / RMSNorm+LayerNorm (mode 1;
 * blocks.py:146 + attention.py:71-73) epilogues of tvae_mtgemm.  x: bf16 [M, C]; w1: fp32 [C] (mode 1). */
int tvae_row_stats(const void* x, const float* w1, float* out_a, float* out_b, int64_t M, int32_t C, int32_t mode,
                   void* stream);
/* NCHW fp32 [B,C,H,W] -> NHWC bf16 [B,H,W,Cpad] (channels >= C zero) and back (first C of Cs channels). */
int tvae_nchw_to_nhwc(const float* in, void* out_bf16, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                      void* stream);
int tvae_nhwc_to_nchw(const void* in_bf16, float* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cs,
                      void* stream);
/* Reparameterisation (transvae.py:186-199; patched :186-196, :244-245).  mu_out / logvar_out (optional) receive
 * the clamped tensors the patched forward returns. */
int tvae_reparam(const float* mu, const float* logvar, const float* eps, float* z, float* mu_out, float* logvar_out,
                 int64_t n, int32_t patched, void* stream);
/* L1 + KL partial sums (vae_loss.py:83,94-95; patched :80-104): acc fp32[4] = {sum|f(recon)-target|, sum KL terms,
 * count of non-finite terms, 0}; zeroed inside. */
int tvae_loss_l1_kl(const float* recon, const float* target, const float* mu, const float* logvar, float* acc,
                    int64_t n_img, int64_t n_lat, int32_t patched, float clip_lo, float clip_hi, void* stream);

/* Reconstruction metrics on the device (evaluate_transvae.py:47-77 calculate_psnr / calculate_ssim and the per-image
 * loops at :131-146; test_rope_extrapolation.py:28-51).  recon / target: fp32 NCHW [B, C, H, W]; the reconstruction
 * is first mapped through `mode` (0 identity, 1 clamp(0,1) as evaluate.py:108-110, 2 sigmoid as
 * evaluate_transvae.py:131).  acc fp32 [B, 4], zeroed inside: per image {sum squared error, sum absolute error,
 * sum of the 11x11 box-filter SSIM map (zero padded, C1 = 0.01^2, C2 = 0.03^2), 0}. */
int tvae_metrics(const float* recon, const float* target, float* acc, int32_t B, int32_t C, int32_t H, int32_t W,
                 int32_t mode, void* stream);

/* ---- backward pass ------------------------------------------------------------------------------
 * Weight gradient of tvae_mtgemm: dw[n, wk_off + k] += sum_pixels dZ[pixel, n] * A_tap[pixel, k] for every tap of
 * `desc` (same a0 / a1 / taps / n_total / k_total as the forward call; desc->out is the dZ view, i.e. the gradient
 * w.r.t. the forward pre-activation output; bias / act / rope / residual fields are ignored).  dw: fp32
 * [n_total, k_total], accumulated into (caller zeroes).  Replaces autograd's weight-gradient convolution / matmul for
 * every nn.Conv2d / nn.Linear listed at tvae_mtgemm.  Input gradients are tvae_mtgemm launches with transposed taps. */
int tvae_mtgemm_wgrad(const tvae_mtgemm_desc* desc /* HOST pointer */, float* dw, void* stream);
/* Same, and the bias gradient in the same launch: db fp32 [num_phases, n_total] += per-phase column sums of dZ (an extra
 * N = 16 tcgen05.mma against a block of ones on the CTAs of each phase's first tap).  Replaces the bias half of autograd
 * for the same layers (a separate pass over dZ otherwise). */
int tvae_mtgemm_wgrad_bias(const tvae_mtgemm_desc* desc /* HOST pointer */, float* dw, float* db, void* stream);
/* dZ = dY * act'(Z) and colsum[n] = sum_m dZ[m, n] (= bias gradient) for a row-major bf16 [M, N] matrix.
 * act == TVAE_ACT_NONE: z / dz may be NULL, only the column sums are produced. */
int tvae_bias_act_bwd(const void* dy, const void* z, void* dz, float* colsum, int64_t M, int32_t N, int32_t act,
                      void* stream);
/* Same over a 4-D contiguous bf16 tensor [R0, P, R1, Q] with colsum fp32 [P, Q] (phase views: P = 2). */
int tvae_bias_act_bwd_4d(const void* dy, const void* z, void* dz, float* colsum, int64_t R0, int32_t P, int32_t R1,
                         int32_t Q, int32_t act, void* stream);
/* y = act(z), bf16, n elements (n % 8 == 0). */
int tvae_act_fwd(const void* z, void* y, int64_t n, int32_t act, void* stream);
/* GroupNorm(+SiLU) backward (blocks.py:60-66): dx = dGN(dh) [+ add]; part fp32 [B, C, 2] = per-(image, channel)
 * (sum dy, sum dy*xhat) from which the caller reduces dgamma / dbeta.  sums = tvae_groupnorm_stats(x). */
int tvae_groupnorm_bwd(const void* x, const void* dh, const void* add, const double* sums, const float* gamma,
                       const float* beta, float* part, void* dx, int32_t B, int32_t HW, int32_t C, int32_t G, float eps,
                       int32_t apply_silu, void* stream);
/* The apply pass alone, `part` given (produced by tvae_mtgemm with the gnb_* fields, or by an earlier call). */
int tvae_groupnorm_bwd_apply(const void* x, const void* dh, const void* add, const double* sums, const float* gamma,
                             const float* beta, const float* part, void* dx, int32_t B, int32_t HW, int32_t C, int32_t G,
                             float eps, int32_t apply_silu, void* stream);
/* Materialised token norms of the training path: mode 0 RMSNorm (blocks.py:168-201), mode 1 RMSNorm followed by the
 * affine-free LayerNorm shared by norm_q/k/v (attention.py:71-73).  x, y, dy, add, dx: bf16 [M, C]; w, dw: fp32 [C]. */
int tvae_token_norm_fwd(const void* x, const float* w, void* y, int64_t M, int32_t C, int32_t mode, void* stream);
int tvae_token_norm_bwd(const void* x, const float* w, const void* dy, const void* add, void* dx, float* dw, int64_t M,
                        int32_t C, int32_t mode, void* stream);
/* Attention backward (attention.py:88-92).  delta fp32 [B, C/64, S] = rowsum(dout * out); dq_acc fp32
 * [dq_slices][B, S, C] (zeroed inside); dqkv bf16 [B, S, 3C]: k / v thirds written by tvae_attn_bwd (rotated space), q
 * third and the transposed RoPE of q and k by tvae_rope_bwd (q_scale = head_dim^-0.5).
 * dq_slices = tvae_attn_bwd_dq_slices(S): 2 when the dQ contributions of the key tiles are added in a fixed order
 * (bit-reproducible; even and odd steps go to separate accumulators so that a contribution never waits for the global
 * completion of the previous one), 1 when they are unordered reduce-adds (very long sequences). */
int tvae_attn_bwd_dq_slices(int32_t S);
int tvae_attn_delta(const void* out, const void* dout, float* delta, int32_t B, int32_t S, int32_t C, void* stream);
int tvae_attn_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq_acc, void* dqkv,
                  int32_t B, int32_t S, int32_t C, int32_t dq_slices, void* stream);
int tvae_rope_bwd(const float* dq_acc, void* dqkv, const float* rope_tab, int64_t M, int32_t C, int32_t H, int32_t W,
                  float q_scale, int32_t dq_slices, void* stream);
/* Loss backward: scal (device fp32[2]) = {dLoss * l1_weight / numel(recon), dLoss * kl_weight / kl_norm}. */
int tvae_loss_bwd(const float* recon, const float* target, const float* mu, const float* logvar, const float* scal,
                  float* drecon, float* dmu, float* dlogvar, int64_t n_img, int64_t n_lat, int32_t patched, float clip_lo,
                  float clip_hi, void* stream);
/* Reparameterisation backward: folds the gradients arriving on (z, mu', logvar') back onto (mu, logvar). */
int tvae_latent_bwd(const float* mu, const float* logvar, const float* eps, const float* dz, const float* dmu_ret,
                    const float* dlv_ret, float* dmu, float* dlogvar, int64_t n, int32_t patched, void* stream);

/* ---- depthwise ConvFFN variant (conv.py:42-50: nn.Conv2d(hidden, hidden, 3, padding=1, groups=hidden); forward
 * conv.py:89-94 `x_spatial + conv(x_spatial)`) ----------------------------------------------------------------
 * y = [u +] dwconv3x3(u; w) [+ bias] on NHWC bf16 [B, H, W, C] (C % 8 == 0); w9c fp32 [9][C] (tap k = dy*3 + dx, i.e. the
 * reference weight [C, 1, 3, 3] transposed), bias fp32 [C] or NULL.  flip != 0 uses tap 8 - k (the input gradient:
 * du = dy + dwconv3x3(dy; flipped taps)).  HBM-bound stencil: 4 bytes per element.
 * tvae_dwconv3x3_wgrad: dw fp32 [9][C] and db fp32 [C] (optional) of the same layer, zeroed inside. */
int tvae_dwconv3x3(const void* u, const float* w9c, const float* bias, void* y, int32_t B, int32_t H, int32_t W, int32_t C,
                   int32_t flip, int32_t add_input, void* stream);
int tvae_dwconv3x3_wgrad(const void* u, const void* dy, float* dw9c, float* db, int32_t B, int32_t H, int32_t W, int32_t C,
                         void* stream);

/* ---- weight-side re-layout (training step) --------------------------------------------------------------------
 * The reference stores nn.Linear weights as [out, in] and nn.Conv2d 3x3 weights as [out, in, 3, 3] fp32
 * (conv.py:39-65, blocks.py:34-37, attention.py:43-48).  tvae_weight_pack turns such a parameter w[A][B][T] (T = 1 or 9)
 * into the bf16 operands of the tensor-core kernels: fwd[A][T*B] (K index t*B + b: the forward tvae_mtgemm weight) and /
 * or dgr[B][T*A] (K index t*A + a: the weight of the input-gradient tvae_mtgemm); either output may be NULL.
 * tvae_wgrad_unpack maps a packed weight gradient g[A][T*B] fp32 (tvae_mtgemm_wgrad output) back to the parameter
 * layout [A][B][T] (T = 9); accumulate != 0 adds it to g_ref (the parameter's .grad slot: optimizer.zero_grad /
 * backward accumulation of train.py:599-603) instead of overwriting.  Both replace strided torch permute / cast copies. */
int tvae_weight_pack(const float* w, void* fwd_bf16, void* dgrad_bf16, int32_t A, int32_t B, int32_t T, void* stream);
int tvae_wgrad_unpack(const float* g_packed, float* g_ref, int32_t A, int32_t B, int32_t T, int32_t accumulate, void* stream);
/* Composite packs of the training step, one launch each way instead of hundreds of tiny torch kernels.
 * tvae_fold_qkv: LayerNorm affines of the three pre-projection norms folded into one projection (attention.py:71-79):
 *   wg fp32 [3C][C] = [Wq g_q; Wk g_k; Wv g_v] (column scale), bg fp32 [3C] = [Wq b_q; Wk b_k; Wv b_v]; w / g / b are
 *   arrays of three device pointers (q, k, v).  tvae_fold_qkv_bwd: the gradients of the nine inputs from dwg [3C][C],
 *   dbg [3C] (dw[s] [C][C], dg[s] [C], db[s] [C] are overwritten, or added to when accumulate != 0 -- the nine .grad
 *   slots directly; column sums in a fixed order).
 * tvae_upconv1_pack: Upsample's first convolution (upsample.py:94-95; nearest 2x + 3x3 == four 2x2 phase convolutions):
 *   w fp32 [O][I][3][3] -> packed fp32 [O][16*I] with coinciding taps summed (backward = 0), or the packed gradient
 *   [O][16*I] -> dw [O][I][3][3] (backward = 1). */
int tvae_fold_qkv(const float* const* w3, const float* const* g3, const float* const* b3, float* wg, float* bg, int32_t C,
                  void* stream);
int tvae_fold_qkv_bwd(const float* const* w3, const float* const* g3, const float* const* b3, const float* dwg,
                      const float* dbg, float* const* dw3, float* const* dg3, float* const* db3, int32_t C, int32_t accumulate,
                      void* stream);
int tvae_upconv1_pack(const float* src, float* dst, int32_t O, int32_t I, int32_t backward, void* stream);

/* ---- optimiser (train.py:608-620, train_2.py:266-274, 329-366: clip_grad_norm_ + fused AdamW + LambdaLR warm-up +
 *      skip of non-finite steps) -- every decision of the step is taken on the device, no host synchronisation ----
 * tvae_grad_sumsq: partials fp64 [TVAE_SUMSQ_BLOCKS] = fixed-order per-block partial sums of g^2 over a flat gradient
 *   buffer (fp32, or bf16 when g_bf16 != 0; n % 8 == 0).  No atomics: the total (re-added in a fixed order by the
 *   consumers) is bit-reproducible and identical on every rank of a data-parallel job.
 * tvae_adamw_step: one fused update over flat p / m / v (fp32) and g (fp32 or bf16).
 *   state (device fp32[8]) = {0: sum of squared gradients of this step (out), 1: clip factor (out), 2: external skip
 *   flag (in), 3: -, 4: updates APPLIED so far (in/out), 5: updates skipped (in/out), 6: learning rate used (out), 7: -}.
 *   The update index k = state[4] drives the bias correction 1 - beta^(k+1) and the warm-up lr_base * min(1, k /
 *   warmup_steps) (warmup_steps <= 0: constant); gradients are pre-scaled by grad_scale (1 / (world * accumulation)) and
 *   clipped to max_norm (<= 0: off) with clip_grad_norm_'s formula.  A non-finite norm (or state[2] != 0) skips the
 *   update and advances state[5] only -- like the reference's `continue` before optimizer.step() / scheduler.step(). */
#define TVAE_SUMSQ_BLOCKS 1024
int tvae_grad_sumsq(const void* g, int32_t g_bf16, int64_t n, double* partials, void* stream);
int tvae_adamw_step(float* p, const void* g, int32_t g_bf16, float* m, float* v, int64_t n, const double* partials,
                    float* state, float lr_base, int32_t warmup_steps, float beta1, float beta2, float eps,
                    float weight_decay, float max_norm, float grad_scale, void* stream);
/* fp32 -> bf16 cast of a gradient bucket before a bf16 all-reduce (optional deviation from DistributedDataParallel's
 * fp32 buckets, train.py:672-674; halves the NVLink payload).  n % 8 == 0. */
int tvae_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* Multi-tensor accumulate: dst[i][j] += src[i][j * src_stride[i]], j < n[i], for `count` fp32 tensors (HOST arrays of
 * device pointers; src_stride NULL = all contiguous) in ceil(count / TVAE_MTA_MAX) launches -- the gradient accumulation of the small (norm / bias / composite-weight)
 * parameters, which autograd's AccumulateGrad performs with one tiny kernel per parameter. */
#define TVAE_MTA_MAX 96
int tvae_multi_tensor_add(float* const* dst, const float* const* src, const int32_t* n, const int32_t* src_stride,
                          int32_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRANSVAE_SM100_H_ */
