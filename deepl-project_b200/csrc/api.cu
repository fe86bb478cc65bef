// extern "C" surface of libtransvae_sm100.so (see include/transvae_sm100.h).
#include "../../include/transvae_sm100.h"
#include "common.cuh"

namespace tvae {
int mtgemm_run(const tvae_mtgemm_desc* d, cudaStream_t stream);
int attn_fwd_run(const void* qkv, void* out, float* lse, int B, int S, int C, cudaStream_t stream);
int im2col_in_run(const float* x, void* cols, int B, int H, int W, cudaStream_t stream);
int gn_stats_run(const void* x, double* sums, int B, int HW, int C, int G, cudaStream_t stream);
int gn_apply_run(const void* x, const double* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C,
                 int G, float eps, int apply_silu, cudaStream_t stream);
int gn_fwd_run(const void* x, double* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C, int G,
               float eps, int apply_silu, cudaStream_t stream);
int row_stats_run(const void* x, const float* w1, float* out_a, float* out_b, long long M, int C, int mode,
                  cudaStream_t stream);
int nchw_to_nhwc_run(const float* in, void* out, int B, int C, int H, int W, int Cpad, cudaStream_t stream);
int nhwc_to_nchw_run(const void* in, float* out, int B, int C, int H, int W, int Cs, cudaStream_t stream);
int reparam_run(const float* mu, const float* logvar, const float* eps, float* z, float* mu_out, float* lv_out,
                long long n, int patched, cudaStream_t stream);
int loss_run(const float* recon, const float* target, const float* mu, const float* logvar, float* acc,
             long long n_img, long long n_lat, int patched, float clip_lo, float clip_hi, cudaStream_t stream);
int mtwgrad_run(const tvae_mtgemm_desc* d, float* dw, float* db, cudaStream_t stream);
int bias_act_bwd_run(const void* dy, const void* z, void* dz, float* colsum, long long R0, int Pn, int R1, int Q, int act,
                     cudaStream_t stream);
int bias_act_bwd_matrix_run(const void* dy, const void* z, void* dz, float* colsum, long long M, int N, int act,
                            cudaStream_t stream);
int act_fwd_run(const void* z, void* y, long long n, int act, cudaStream_t stream);
int gn_bwd_run(const void* x, const void* dh, const void* add, const double* sums, const float* gamma, const float* beta,
               float* part, void* dx, int B, int HW, int C, int G, float eps, int apply_silu, cudaStream_t stream);
int gn_bwd_apply_run(const void* x, const void* dh, const void* add, const double* sums, const float* gamma, const float* beta,
                     const float* part, void* dx, int B, int HW, int C, int G, float eps, int apply_silu, cudaStream_t stream);
int token_norm_fwd_run(const void* x, const float* w, void* y, long long M, int C, int mode, cudaStream_t stream);
int token_norm_bwd_run(const void* x, const float* w, const void* dy, const void* add, void* dx, float* dw, long long M,
                       int C, int mode, cudaStream_t stream);
int attn_delta_run(const void* o, const void* dout, float* delta, int B, int S, int C, cudaStream_t stream);
int attn_bwd_dq_slices(int S);
int attn_bwd_run(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq_acc, void* dqkv, int B,
                 int S, int C, int dq_slices, cudaStream_t stream);
int rope_bwd_run(const float* dq_acc, void* dqkv, const float* tab, long long M, int C, int H, int W, float q_scale,
                 int dq_slices, cudaStream_t stream);
int loss_bwd_run(const float* recon, const float* target, const float* mu, const float* logvar, const float* scal,
                 float* drecon, float* dmu, float* dlv, long long n_img, long long n_lat, int patched, float clip_lo,
                 float clip_hi, cudaStream_t stream);
int latent_bwd_run(const float* mu, const float* logvar, const float* eps, const float* dz, const float* dmu_ret,
                   const float* dlv_ret, float* dmu, float* dlv, long long n, int patched, cudaStream_t stream);
int grad_sumsq_run(const void* g, int g_bf16, long long n, double* partials, cudaStream_t stream);
int adamw_step_run(float* p, const void* g, int g_bf16, float* m, float* v, long long n, const double* partials,
                   float* state, float lr_base, int warmup_steps, float b1, float b2, float eps, float wd, float max_norm,
                   float grad_scale, cudaStream_t stream);
int cast_f32_bf16_run(const float* in, void* out, long long n, cudaStream_t stream);
int mta_add_run(float* const* dst, const float* const* src, const int* n, const int* sstride, int count, cudaStream_t stream);
int weight_pack_run(const float* w, void* fwd, void* dgr, int A, int B, int T, cudaStream_t stream);
int wgrad_unpack_run(const float* g, float* out, int A, int B, int T, int accumulate, cudaStream_t stream);
int fold_qkv_run(const float* const* w, const float* const* g, const float* const* b, float* wg, float* bg, int C,
                 cudaStream_t stream);
int fold_qkv_bwd_run(const float* const* w, const float* const* g, const float* const* b, const float* dwg, const float* dbg,
                     float* const* dw, float* const* dg, float* const* db, int C, int accumulate, cudaStream_t stream);
int upconv1_pack_run(const float* w, float* out, int O, int I, int backward, cudaStream_t stream);
int dwconv3x3_run(const void* u, const float* w9c, const float* bias, void* y, int B, int H, int W, int C, int flip,
                  int add_input, cudaStream_t stream);
int dwconv3x3_wgrad_run(const void* u, const void* dy, float* dw, float* db, int B, int H, int W, int C, cudaStream_t stream);
int metrics_run(const float* recon, const float* target, float* acc, int B, int C, int H, int W, int mode,
                cudaStream_t stream);
}  // namespace tvae

using namespace tvae;

static int require_device() {
  if (!tvae_device_ok()) {
    set_last_error("libtransvae_sm100: no sm_100 (B200) device is current; there is no CPU fallback");
    return -10;
  }
  return 0;
}
#define S_(s) reinterpret_cast<cudaStream_t>(s)
#define GUARD()                  \
  do {                           \
    int _g = require_device();   \
    if (_g) return _g;           \
  } while (0)

extern "C" {

int tvae_abi_version(void) { return TVAE_ABI_VERSION; }
const char* tvae_last_error(void) { return get_last_error(); }
int tvae_num_sms(void) { return num_sms(); }
int tvae_set_reserved_sms(int32_t n) { set_reserved_sms(n); return persistent_sms(); }

int tvae_device_ok(void) {
  static int cached = -1;
  if (cached >= 0) return cached;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;  // not cached: a device may appear later in the process (it will not, but stay honest)
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  cached = (major == 10) ? 1 : 0;
  return cached;
}

int tvae_mtgemm(const tvae_mtgemm_desc* desc, void* stream) { GUARD(); return mtgemm_run(desc, S_(stream)); }
int tvae_attn_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t S, int32_t C, void* stream) {
  GUARD(); return attn_fwd_run(qkv, out, lse, B, S, C, S_(stream));
}
int tvae_im2col_in(const float* x_nchw, void* cols, int32_t B, int32_t H, int32_t W, void* stream) {
  GUARD(); return im2col_in_run(x_nchw, cols, B, H, W, S_(stream));
}
int tvae_groupnorm_stats(const void* x, double* sums, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream) {
  GUARD(); return gn_stats_run(x, sums, B, HW, C, G, S_(stream));
}
int tvae_groupnorm_apply(const void* x, const double* sums, const float* gamma, const float* beta, void* y, int32_t B,
                         int32_t HW, int32_t C, int32_t G, float eps, int32_t apply_silu, void* stream) {
  GUARD(); return gn_apply_run(x, sums, gamma, beta, y, B, HW, C, G, eps, apply_silu, S_(stream));
}
int tvae_groupnorm_silu(const void* x, const float* gamma, const float* beta, void* y, double* sums, int32_t B, int32_t HW,
                        int32_t C, int32_t G, float eps, int32_t apply_silu, void* stream) {
  GUARD(); return gn_fwd_run(x, sums, gamma, beta, y, B, HW, C, G, eps, apply_silu, S_(stream));
}
int tvae_row_stats(const void* x, const float* w1, float* out_a, float* out_b, int64_t M, int32_t C, int32_t mode,
                   void* stream) {
  GUARD(); return row_stats_run(x, w1, out_a, out_b, M, C, mode, S_(stream));
}
int tvae_nchw_to_nhwc(const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                      void* stream) {
  GUARD(); return nchw_to_nhwc_run(in, out, B, C, H, W, Cpad, S_(stream));
}
int tvae_nhwc_to_nchw(const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cs,
                      void* stream) {
  GUARD(); return nhwc_to_nchw_run(in, out, B, C, H, W, Cs, S_(stream));
}
int tvae_reparam(const float* mu, const float* logvar, const float* eps, float* z, float* mu_out, float* lv_out,
                 int64_t n, int32_t patched, void* stream) {
  GUARD(); return reparam_run(mu, logvar, eps, z, mu_out, lv_out, n, patched, S_(stream));
}
int tvae_loss_l1_kl(const float* recon, const float* target, const float* mu, const float* logvar, float* acc,
                    int64_t n_img, int64_t n_lat, int32_t patched, float clip_lo, float clip_hi, void* stream) {
  GUARD(); return loss_run(recon, target, mu, logvar, acc, n_img, n_lat, patched, clip_lo, clip_hi, S_(stream));
}

int tvae_mtgemm_wgrad(const tvae_mtgemm_desc* desc, float* dw, void* stream) {
  GUARD(); return mtwgrad_run(desc, dw, nullptr, S_(stream));
}
int tvae_mtgemm_wgrad_bias(const tvae_mtgemm_desc* desc, float* dw, float* db, void* stream) {
  GUARD(); return mtwgrad_run(desc, dw, db, S_(stream));
}
int tvae_bias_act_bwd(const void* dy, const void* z, void* dz, float* colsum, int64_t M, int32_t N, int32_t act, void* stream) {
  GUARD(); return bias_act_bwd_matrix_run(dy, z, dz, colsum, M, N, act, S_(stream));
}
int tvae_bias_act_bwd_4d(const void* dy, const void* z, void* dz, float* colsum, int64_t R0, int32_t P, int32_t R1, int32_t Q,
                         int32_t act, void* stream) {
  GUARD(); return bias_act_bwd_run(dy, z, dz, colsum, R0, P, R1, Q, act, S_(stream));
}
int tvae_act_fwd(const void* z, void* y, int64_t n, int32_t act, void* stream) { GUARD(); return act_fwd_run(z, y, n, act, S_(stream)); }
int tvae_groupnorm_bwd(const void* x, const void* dh, const void* add, const double* sums, const float* gamma, const float* beta,
                       float* part, void* dx, int32_t B, int32_t HW, int32_t C, int32_t G, float eps, int32_t apply_silu,
                       void* stream) {
  GUARD(); return gn_bwd_run(x, dh, add, sums, gamma, beta, part, dx, B, HW, C, G, eps, apply_silu, S_(stream));
}
int tvae_groupnorm_bwd_apply(const void* x, const void* dh, const void* add, const double* sums, const float* gamma,
                             const float* beta, const float* part, void* dx, int32_t B, int32_t HW, int32_t C, int32_t G,
                             float eps, int32_t apply_silu, void* stream) {
  GUARD(); return gn_bwd_apply_run(x, dh, add, sums, gamma, beta, part, dx, B, HW, C, G, eps, apply_silu, S_(stream));
}
int tvae_token_norm_fwd(const void* x, const float* w, void* y, int64_t M, int32_t C, int32_t mode, void* stream) {
  GUARD(); return token_norm_fwd_run(x, w, y, M, C, mode, S_(stream));
}
int tvae_token_norm_bwd(const void* x, const float* w, const void* dy, const void* add, void* dx, float* dw, int64_t M,
                        int32_t C, int32_t mode, void* stream) {
  GUARD(); return token_norm_bwd_run(x, w, dy, add, dx, dw, M, C, mode, S_(stream));
}
int tvae_attn_delta(const void* out, const void* dout, float* delta, int32_t B, int32_t S, int32_t C, void* stream) {
  GUARD(); return attn_delta_run(out, dout, delta, B, S, C, S_(stream));
}
int tvae_attn_bwd_dq_slices(int32_t S) { return attn_bwd_dq_slices(S); }
int tvae_attn_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq_acc, void* dqkv, int32_t B,
                  int32_t S, int32_t C, int32_t dq_slices, void* stream) {
  GUARD(); return attn_bwd_run(qkv, dout, lse, delta, dq_acc, dqkv, B, S, C, dq_slices, S_(stream));
}
int tvae_rope_bwd(const float* dq_acc, void* dqkv, const float* rope_tab, int64_t M, int32_t C, int32_t H, int32_t W,
                  float q_scale, int32_t dq_slices, void* stream) {
  GUARD(); return rope_bwd_run(dq_acc, dqkv, rope_tab, M, C, H, W, q_scale, dq_slices, S_(stream));
}
int tvae_loss_bwd(const float* recon, const float* target, const float* mu, const float* logvar, const float* scal,
                  float* drecon, float* dmu, float* dlogvar, int64_t n_img, int64_t n_lat, int32_t patched, float clip_lo,
                  float clip_hi, void* stream) {
  GUARD(); return loss_bwd_run(recon, target, mu, logvar, scal, drecon, dmu, dlogvar, n_img, n_lat, patched, clip_lo, clip_hi, S_(stream));
}
int tvae_latent_bwd(const float* mu, const float* logvar, const float* eps, const float* dz, const float* dmu_ret,
                    const float* dlv_ret, float* dmu, float* dlogvar, int64_t n, int32_t patched, void* stream) {
  GUARD(); return latent_bwd_run(mu, logvar, eps, dz, dmu_ret, dlv_ret, dmu, dlogvar, n, patched, S_(stream));
}
int tvae_weight_pack(const float* w, void* fwd_bf16, void* dgrad_bf16, int32_t A, int32_t B, int32_t T, void* stream) {
  GUARD(); return weight_pack_run(w, fwd_bf16, dgrad_bf16, A, B, T, S_(stream));
}
int tvae_wgrad_unpack(const float* g_packed, float* g_ref, int32_t A, int32_t B, int32_t T, int32_t accumulate, void* stream) {
  GUARD(); return wgrad_unpack_run(g_packed, g_ref, A, B, T, accumulate, S_(stream));
}
int tvae_fold_qkv(const float* const* w3, const float* const* g3, const float* const* b3, float* wg, float* bg, int32_t C,
                  void* stream) {
  GUARD(); return fold_qkv_run(w3, g3, b3, wg, bg, C, S_(stream));
}
int tvae_fold_qkv_bwd(const float* const* w3, const float* const* g3, const float* const* b3, const float* dwg,
                      const float* dbg, float* const* dw3, float* const* dg3, float* const* db3, int32_t C, int32_t accumulate,
                      void* stream) {
  GUARD(); return fold_qkv_bwd_run(w3, g3, b3, dwg, dbg, dw3, dg3, db3, C, accumulate, S_(stream));
}
int tvae_upconv1_pack(const float* src, float* dst, int32_t O, int32_t I, int32_t backward, void* stream) {
  GUARD(); return upconv1_pack_run(src, dst, O, I, backward, S_(stream));
}
int tvae_grad_sumsq(const void* g, int32_t g_bf16, int64_t n, double* partials, void* stream) {
  GUARD(); return grad_sumsq_run(g, g_bf16, n, partials, S_(stream));
}
int tvae_adamw_step(float* p, const void* g, int32_t g_bf16, float* m, float* v, int64_t n, const double* partials,
                    float* state, float lr_base, int32_t warmup_steps, float beta1, float beta2, float eps,
                    float weight_decay, float max_norm, float grad_scale, void* stream) {
  GUARD();
  return adamw_step_run(p, g, g_bf16, m, v, n, partials, state, lr_base, warmup_steps, beta1, beta2, eps, weight_decay,
                        max_norm, grad_scale, S_(stream));
}
int tvae_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream) {
  GUARD(); return cast_f32_bf16_run(in, out_bf16, n, S_(stream));
}
int tvae_multi_tensor_add(float* const* dst, const float* const* src, const int32_t* n, const int32_t* src_stride,
                          int32_t count, void* stream) {
  GUARD(); return mta_add_run(dst, src, n, src_stride, count, S_(stream));
}

int tvae_dwconv3x3(const void* u, const float* w9c, const float* bias, void* y, int32_t B, int32_t H, int32_t W, int32_t C,
                   int32_t flip, int32_t add_input, void* stream) {
  GUARD(); return dwconv3x3_run(u, w9c, bias, y, B, H, W, C, flip, add_input, S_(stream));
}
int tvae_dwconv3x3_wgrad(const void* u, const void* dy, float* dw9c, float* db, int32_t B, int32_t H, int32_t W, int32_t C,
                         void* stream) {
  GUARD(); return dwconv3x3_wgrad_run(u, dy, dw9c, db, B, H, W, C, S_(stream));
}

int tvae_metrics(const float* recon, const float* target, float* acc, int32_t B, int32_t C, int32_t H, int32_t W,
                 int32_t mode, void* stream) {
  GUARD(); return metrics_run(recon, target, acc, B, C, H, W, mode, S_(stream));
}
}  // extern "C"
