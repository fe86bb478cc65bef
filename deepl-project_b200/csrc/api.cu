// extern "C" surface of libtransvae_sm100.so (see include/transvae_sm100.h).
#include "../../include/transvae_sm100.h"
#include "common.cuh"

namespace tvae {
int mtgemm_run(const tvae_mtgemm_desc* d, cudaStream_t stream);
int attn_fwd_run(const void* qkv, void* out, float* lse, int B, int S, int C, cudaStream_t stream);
int conv_in_run(const float* x, const float* w, const float* bias, void* out, int B, int Cin, int H, int W, int Cout,
                cudaStream_t stream);
int gn_stats_run(const void* x, float* sums, int B, int HW, int C, int G, cudaStream_t stream);
int gn_apply_run(const void* x, const float* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C,
                 int G, float eps, int apply_silu, cudaStream_t stream);
int row_stats_run(const void* x, const float* w1, float* out_a, float* out_b, long long M, int C, int mode,
                  cudaStream_t stream);
int nchw_to_nhwc_run(const float* in, void* out, int B, int C, int H, int W, int Cpad, cudaStream_t stream);
int nhwc_to_nchw_run(const void* in, float* out, int B, int C, int H, int W, int Cs, cudaStream_t stream);
int reparam_run(const float* mu, const float* logvar, const float* eps, float* z, float* mu_out, float* lv_out,
                long long n, int patched, cudaStream_t stream);
int loss_run(const float* recon, const float* target, const float* mu, const float* logvar, float* acc,
             long long n_img, long long n_lat, int patched, float clip_lo, float clip_hi, cudaStream_t stream);
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
int tvae_conv_in(const float* x, const float* w, const float* bias, void* out, int32_t B, int32_t Cin, int32_t H,
                 int32_t W, int32_t Cout, void* stream) {
  GUARD(); return conv_in_run(x, w, bias, out, B, Cin, H, W, Cout, S_(stream));
}
int tvae_groupnorm_stats(const void* x, float* sums, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream) {
  GUARD(); return gn_stats_run(x, sums, B, HW, C, G, S_(stream));
}
int tvae_groupnorm_apply(const void* x, const float* sums, const float* gamma, const float* beta, void* y, int32_t B,
                         int32_t HW, int32_t C, int32_t G, float eps, int32_t apply_silu, void* stream) {
  GUARD(); return gn_apply_run(x, sums, gamma, beta, y, B, HW, C, G, eps, apply_silu, S_(stream));
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

}  // extern "C"
