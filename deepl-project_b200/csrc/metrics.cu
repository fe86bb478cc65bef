// On-GPU reconstruction metrics: per-image MSE / L1 sums and the windowed SSIM of the reference's evaluation script.
//
// Replaces the per-image Python loops of evaluate_transvae.py:47-77, :131-146 (calculate_psnr = 20 log10(1/sqrt(mse)),
// calculate_ssim = 11x11 box-filter SSIM with zero padding and count_include_pad, C1 = 0.01^2, C2 = 0.03^2, mean over
// all pixels and channels) and of test_rope_extrapolation.py:28-51 (PSNR on the raw reconstruction).
//
// HBM-bound: 8 B read per element (reconstruction logits + target, fp32 NCHW) and nothing written but 4 floats per
// image.  One CTA per 32x32 output tile of one (image, channel) plane: the tile plus its 5-pixel halo of
// f(recon) and target is staged in shared memory, the five box sums (x, y, xx, yy, xy) are formed separably.
#include "../../include/transvae_sm100.h"
#include "common.cuh"

namespace tvae {

constexpr int kMetTile = 32;
constexpr int kMetHalo = 5;
constexpr int kMetIn = kMetTile + 2 * kMetHalo;   // 42

template <int MODE>   // 0: identity, 1: clamp(0, 1), 2: sigmoid
__device__ __forceinline__ float met_xform(float v) {
  if (MODE == 1) return fminf(fmaxf(v, 0.0f), 1.0f);
  if (MODE == 2) return 1.0f / (1.0f + __expf(-v));
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256)
metrics_kernel(const float* __restrict__ recon, const float* __restrict__ target, float* __restrict__ acc, int C, int H,
               int W) {
  __shared__ float sx[kMetIn][kMetIn + 1];
  __shared__ float sy[kMetIn][kMetIn + 1];
  __shared__ float hs[5][kMetIn][kMetTile + 1];   // horizontal 11-sums of x, y, xx, yy, xy
  __shared__ float red[3][8];
  const int plane = blockIdx.z;                   // b * C + c
  const int b = plane / C;
  const int h0 = blockIdx.y * kMetTile, w0 = blockIdx.x * kMetTile;
  const float* rp = recon + (size_t)plane * H * W;
  const float* tp = target + (size_t)plane * H * W;
  for (int i = threadIdx.x; i < kMetIn * kMetIn; i += 256) {
    const int r = i / kMetIn, c = i % kMetIn;
    const int hh = h0 + r - kMetHalo, ww = w0 + c - kMetHalo;
    float x = 0.0f, y = 0.0f;                     // zero padding (avg_pool2d padding, count_include_pad=True)
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      x = met_xform<MODE>(__ldg(rp + (size_t)hh * W + ww));
      y = __ldg(tp + (size_t)hh * W + ww);
    }
    sx[r][c] = x;
    sy[r][c] = y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMetIn * kMetTile; i += 256) {
    const int r = i / kMetTile, c = i % kMetTile;
    float a = 0, bb = 0, aa = 0, b2 = 0, ab = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float x = sx[r][c + k], y = sy[r][c + k];
      a += x; bb += y; aa = fmaf(x, x, aa); b2 = fmaf(y, y, b2); ab = fmaf(x, y, ab);
    }
    hs[0][r][c] = a; hs[1][r][c] = bb; hs[2][r][c] = aa; hs[3][r][c] = b2; hs[4][r][c] = ab;
  }
  __syncthreads();
  float se = 0.0f, ae = 0.0f, ss = 0.0f;
  for (int i = threadIdx.x; i < kMetTile * kMetTile; i += 256) {
    const int r = i / kMetTile, c = i % kMetTile;
    if (h0 + r < H && w0 + c < W) {
      float s[5] = {0, 0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 11; ++k) {
#pragma unroll
        for (int q = 0; q < 5; ++q) s[q] += hs[q][r + k][c];
      }
      const float inv = 1.0f / 121.0f;
      const float mu1 = s[0] * inv, mu2 = s[1] * inv;
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float s1 = s[2] * inv - mu1_sq, s2 = s[3] * inv - mu2_sq, s12 = s[4] * inv - mu12;
      const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
      ss += ((2.0f * mu12 + C1) * (2.0f * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2));
      const float d = sx[r + kMetHalo][c + kMetHalo] - sy[r + kMetHalo][c + kMetHalo];
      se = fmaf(d, d, se);
      ae += fabsf(d);
    }
  }
  se = warp_sum(se); ae = warp_sum(ae); ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = se; red[1][warp] = ae; red[2][warp] = ss; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(acc + (size_t)b * 4 + threadIdx.x, t);
  }
}

int metrics_run(const float* recon, const float* target, float* acc, int B, int C, int H, int W, int mode,
                cudaStream_t stream) {
  TVAE_REQUIRE(recon && target && acc, "metrics: null pointer");
  TVAE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && (long long)B * C <= 65535, "metrics: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  TVAE_REQUIRE(mode >= 0 && mode <= 2, "metrics: mode %d (0 identity, 1 clamp01, 2 sigmoid)", mode);
  TVAE_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 4 * B, stream));
  dim3 grid((W + kMetTile - 1) / kMetTile, (H + kMetTile - 1) / kMetTile, B * C);
  if (mode == 0) metrics_kernel<0><<<grid, 256, 0, stream>>>(recon, target, acc, C, H, W);
  else if (mode == 1) metrics_kernel<1><<<grid, 256, 0, stream>>>(recon, target, acc, C, H, W);
  else metrics_kernel<2><<<grid, 256, 0, stream>>>(recon, target, acc, C, H, W);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
