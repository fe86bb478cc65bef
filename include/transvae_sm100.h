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

#define TVAE_ABI_VERSION 1
#define TVAE_MAX_TAPS 16
#define TVAE_MAX_PHASES 4

int tvae_abi_version(void);
const char* tvae_last_error(void);
/* 1 when the current CUDA device is compute capability 10.x, 0 otherwise (or no device). */
int tvae_device_ok(void);
int tvae_num_sms(void);

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
} tvae_mtgemm_desc;

int tvae_mtgemm(const tvae_mtgemm_desc* desc /* HOST pointer */, void* stream);

/* ---- attention ------------------------------------------------------------------------------
 * Flash-style attention forward, head_dim 64, non-causal (tcgen05 + TMEM + TMA).
 * qkv: bf16 [B, S, 3C] (q | k | v per token; q already rotated and scaled by 64^-0.5*log2(e), k rotated -- the
 * QKV tvae_mtgemm epilogue does both); out: bf16 [B, S, C]; lse (optional): fp32 [B, C/64, S], log2 domain.
 * Replaces F.scaled_dot_product_attention at attention.py:88-92. */
int tvae_attn_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t S, int32_t C, void* stream);

/* ---- HBM-bound kernels ------------------------------------------------------------------------ */
/* encoder.conv_in (encoder.py:52): 3x3 pad 1, NCHW fp32 [B,3,H,W] -> NHWC bf16 [B,H,W,Cout]; w fp32 OIHW. */
int tvae_conv_in(const float* x_nchw, const float* w_oihw, const float* bias, void* out_nhwc, int32_t B, int32_t Cin,
                 int32_t H, int32_t W, int32_t Cout, void* stream);
/* nn.GroupNorm(G, C) statistics (blocks.py:33,36; decoder.py:93): sums fp32 [B, G, 2] = (sum x, sum x^2). */
int tvae_groupnorm_stats(const void* x_nhwc, float* sums, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream);
/* y = act(GroupNorm(x)) with act = SiLU (apply_silu=1) or identity; NHWC bf16 in/out (blocks.py:60-66). */
int tvae_groupnorm_apply(const void* x_nhwc, const float* sums, const float* gamma, const float* beta, void* y_nhwc,
                         int32_t B, int32_t HW, int32_t C, int32_t G, float eps, int32_t apply_silu, void* stream);
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

#ifdef __cplusplus
}
#endif
#endif /* TRANSVAE_SM100_H_ */
