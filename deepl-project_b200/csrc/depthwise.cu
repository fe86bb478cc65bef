// Depthwise 3x3 convolution of the ConvFFN `conv_type='depthwise'` variant (conv.py:42-50, 89-94):
//     y = u + dwconv3x3(u) + bias          (nn.Conv2d(hidden, hidden, 3, padding=1, groups=hidden) + the residual add)
// NHWC bf16 activations, fp32 weights [9][C] (tap-major, k = dy*3 + dx) and bias [C].  9 MACs per output element against
// 4 bytes of traffic: an HBM-bound stencil, not tensor-core work.  One thread owns 8 channels (a 16-byte vector) of a
// vertical strip of kRows output pixels and slides a 3-row window down the strip, so each input vector is fetched from
// L1/L2 once per three output rows of its column and the HBM traffic stays at the algorithmic read + write.
//   forward:          y = u + conv(u; w) + b          algorithmic bytes 4 / element
//   input gradient:   du = dy + conv(dy; flip(w))     the same kernel with the spatially flipped taps and no bias
//   weight gradient:  dw[k][c] = sum_pix dy[pix][c] * u[pix + off_k][c],  db[c] = sum_pix dy[pix][c]    4 B / element
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "ew_common.cuh"

namespace tvae {

constexpr int kDwRows = 8;

__device__ __forceinline__ uint4 ldg_or_zero(const uint4* __restrict__ p, bool ok) {
  return ok ? __ldg(p) : make_uint4(0u, 0u, 0u, 0u);
}

// grid: (ceil(C/8 * W / 256), ceil(H / kDwRows), B); thread -> (channel vector v, column w)
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const uint4* __restrict__ u, const float* __restrict__ wt,
                                                        const float* __restrict__ bias, uint4* __restrict__ y, int H, int W,
                                                        int C8, int flip, int add_input) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C8 * W) return;
  const int v = idx % C8, w = idx / C8;
  const int C = C8 * 8;
  const int h0 = blockIdx.y * kDwRows;
  const int b = blockIdx.z;
  // taps of this thread's 8 channels: wk[k][j], k = dy*3 + dx (flipped for the input gradient)
  float2 wk[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int ks = flip ? 8 - k : k;
    const float4 a = __ldg(reinterpret_cast<const float4*>(wt + (size_t)ks * C + v * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(wt + (size_t)ks * C + v * 8) + 1);
    wk[k][0] = make_float2(a.x, a.y); wk[k][1] = make_float2(a.z, a.w);
    wk[k][2] = make_float2(c.x, c.y); wk[k][3] = make_float2(c.z, c.w);
  }
  float2 bv[4] = {f2(0.f), f2(0.f), f2(0.f), f2(0.f)};
  if (bias != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + v * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(bias + v * 8) + 1);
    bv[0] = make_float2(a.x, a.y); bv[1] = make_float2(a.z, a.w); bv[2] = make_float2(c.x, c.y); bv[3] = make_float2(c.z, c.w);
  }
  const uint4* base = u + (size_t)b * H * W * C8;
  auto row_ptr = [&](int h, int ww) { return base + ((size_t)h * W + ww) * C8 + v; };
  // sliding window: rows r-1, r, r+1 x columns w-1, w, w+1
  uint4 win[3][3];
  auto load_row = [&](int h, uint4 (&dst)[3]) {
    const bool hok = h >= 0 && h < H;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ww = w + dx - 1;
      dst[dx] = ldg_or_zero(row_ptr(hok ? h : 0, (ww >= 0 && ww < W) ? ww : 0), hok && ww >= 0 && ww < W);
    }
  };
  load_row(h0 - 1, win[0]);
  load_row(h0, win[1]);
#pragma unroll 1
  for (int r = 0; r < kDwRows; ++r) {
    const int h = h0 + r;
    if (h >= H) break;
    load_row(h + 1, win[2]);
    float2 acc[4];
    if (add_input) {
      unpack8_2(win[1][1], acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = __fadd2_rn(acc[j], bv[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = bv[j];
    }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float2 x[4];
        unpack8_2(win[dy][dx], x);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(x[j], wk[dy * 3 + dx][j], acc[j]);
      }
    y[((size_t)b * H * W + (size_t)h * W + w) * C8 + v] = pack8_2(acc);
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      win[0][dx] = win[1][dx];
      win[1][dx] = win[2][dx];
    }
  }
}

int dwconv3x3_run(const void* u, const float* w9c, const float* bias, void* y, int B, int H, int W, int C, int flip,
                  int add_input, cudaStream_t stream) {
  TVAE_REQUIRE(u && w9c && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dwconv3x3: bad arguments (C %% 8 == 0)");
  const int C8 = C / 8;
  dim3 grid((C8 * W + 255) / 256, (H + kDwRows - 1) / kDwRows, B);
  TVAE_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "dwconv3x3: tensor too large for the launch grid");
  dwconv3x3_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(u), w9c, bias, reinterpret_cast<uint4*>(y), H, W,
                                             C8, flip, add_input);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Weight / bias gradient.  grid: (ceil(C/8 / 32), chunks of pixels); block = 32 channel vectors x 8 pixel lanes.  Each
// thread accumulates the 9 x 8 tap products + 8 column sums of its channel vector over its pixels in registers; the 8
// pixel lanes of a block meet in shared memory and one lane issues the global atomics (dw, db are zeroed by the caller).
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_kernel(const uint4* __restrict__ u, const uint4* __restrict__ dy,
                                                              float* __restrict__ dw, float* __restrict__ db, int B, int H,
                                                              int W, int C8, int pix_per_block) {
  const int lane_v = threadIdx.x & 31, lane_p = threadIdx.x >> 5;
  const int v = blockIdx.x * 32 + lane_v;
  const bool active = v < C8;
  const long long npix = (long long)B * H * W;
  const long long p0 = (long long)blockIdx.y * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  float acc[10][8];
#pragma unroll
  for (int k = 0; k < 10; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.0f;
  if (active) {
    for (long long pix = p0 + lane_p; pix < p1; pix += 8) {
      const int w = (int)(pix % W), h = (int)((pix / W) % H);
      const long long img = pix / ((long long)W * H);
      float2 g[4];
      unpack8_2(__ldg(dy + pix * C8 + v), g);
      const float gv[8] = {g[0].x, g[0].y, g[1].x, g[1].y, g[2].x, g[2].y, g[3].x, g[3].y};
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[9][j] += gv[j];
#pragma unroll
      for (int dy_ = 0; dy_ < 3; ++dy_)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int hh = h + dy_ - 1, ww = w + dx - 1;
          if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
          float2 x[4];
          unpack8_2(__ldg(u + ((img * H + hh) * W + ww) * C8 + v), x);
          const float xv[8] = {x[0].x, x[0].y, x[1].x, x[1].y, x[2].x, x[2].y, x[3].x, x[3].y};
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[dy_ * 3 + dx][j] = fmaf(gv[j], xv[j], acc[dy_ * 3 + dx][j]);
        }
    }
  }
  // reduce the 8 pixel lanes of the block: per tap, stage [pixel lane][256 channel slots] in shared memory, then thread t
  // adds column t (channel blockIdx.x * 256 + t) and issues ONE global atomic
  __shared__ float red[8][256];
  const int ch = blockIdx.x * 256 + threadIdx.x;
  const int C = C8 * 8;
#pragma unroll
  for (int k = 0; k < 10; ++k) {      // unrolled: acc[k] must stay a register array
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[lane_p][lane_v * 8 + j] = acc[k][j];
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    if (ch < C) {
      if (k < 9) atomicAdd(dw + (size_t)k * C + ch, t);
      else if (db != nullptr) atomicAdd(db + ch, t);
    }
  }
}

int dwconv3x3_wgrad_run(const void* u, const void* dy, float* dw, float* db, int B, int H, int W, int C, cudaStream_t stream) {
  TVAE_REQUIRE(u && dy && dw && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dwconv3x3_wgrad: bad arguments");
  const int C8 = C / 8;
  const long long npix = (long long)B * H * W;
  const int gx = (C8 + 31) / 32;
  long long chunks = (8LL * num_sms() + gx - 1) / gx;
  if (chunks > (npix + 63) / 64) chunks = (npix + 63) / 64;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int ppb = (int)((npix + chunks - 1) / chunks);
  TVAE_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * C * sizeof(float), stream));
  if (db != nullptr) TVAE_CHECK_CUDA(cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), stream));
  dim3 grid(gx, (unsigned)((npix + ppb - 1) / ppb));
  dwconv3x3_wgrad_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(u), reinterpret_cast<const uint4*>(dy), dw,
                                                   db, B, H, W, C8, ppb);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
