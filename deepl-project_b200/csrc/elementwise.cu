// HBM-bound kernels of the TransVAE hot path (vectorised 128-bit accesses, warp-shuffle reductions):
// first-layer direct convolution, GroupNorm(+SiLU), per-token RMS/LayerNorm statistics, layout conversion,
// reparameterisation and the L1+KL loss.  All activations are NHWC bf16; statistics are fp32.
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "ew_common.cuh"

namespace tvae {

// -------------------------------------------------------------------------------------------------
// im2col of the 3-channel input image for encoder.conv_in (encoder.py:52, 3x3 pad 1, 3 -> C0).
// NCHW fp32 [B, 3, H, W] -> bf16 [B*H*W, 64] with, per pixel,
//   columns  0..26  hi = bf16(x[ci, h+dy-1, w+dx-1])          (k = ci*9 + dy*3 + dx, zero outside the image)
//   columns 27..53  lo = bf16(x - hi)                          (the split keeps ~16 mantissa bits of the fp32 image)
//   column  54      1.0                                        (bias column)
//   columns 55..63  0
// so the first convolution and its weight gradient become K = 64 launches of the tcgen05 GEMM kernels (tvae_mtgemm /
// tvae_mtgemm_wgrad with a packed [C0, 64] weight [w | w | bias | 0]) instead of CUDA-core loops: the direct kernels
// ran at 13 % of the HBM roofline forward and took 4.4 ms per training micro-step for the weight gradient.
// Algorithmic bytes: 12 read + 128 written per pixel.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_in_kernel(const float* __restrict__ x, uint4* __restrict__ cols, int B, int H,
                                                        int W) {
  const long long npix = (long long)B * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(pix % W), hq = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
    __align__(16) __nv_bfloat16 v[64];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = hq + dy - 1, xx = wq + dx - 1;
          const float f = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (((size_t)b * 3 + ci) * H + yy) * W + xx) : 0.0f;
          const __nv_bfloat16 hi = __float2bfloat16_rn(f);
          v[(ci * 3 + dy) * 3 + dx] = hi;
          v[27 + (ci * 3 + dy) * 3 + dx] = __float2bfloat16_rn(f - __bfloat162float(hi));
        }
    v[54] = __float2bfloat16_rn(1.0f);
#pragma unroll
    for (int k = 55; k < 64; ++k) v[k] = __float2bfloat16_rn(0.0f);
    const uint4* src = reinterpret_cast<const uint4*>(v);
    uint4* dst = cols + pix * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = src[k];
  }
}

int im2col_in_run(const float* x, void* cols, int B, int H, int W, cudaStream_t stream) {
  TVAE_REQUIRE(x && cols && B > 0 && H > 0 && W > 0, "im2col_in: bad arguments");
  const long long npix = (long long)B * H * W;
  long long grid = (npix + 255) / 256;
  if (grid > num_sms() * 16LL) grid = num_sms() * 16LL;
  im2col_in_kernel<<<(int)grid, 256, 0, stream>>>(x, reinterpret_cast<uint4*>(cols), B, H, W);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// GroupNorm statistics over NHWC bf16: sums[b][g] = (sum x, sum x^2) over the group's channels and all pixels.
// Replaces the reduction half of nn.GroupNorm(32, C) (blocks.py:33,36; decoder.py:93).
// Algorithmic bytes: 2*C per pixel (one read).
// -------------------------------------------------------------------------------------------------
constexpr int kGnBatch = 8;   // independent 16-byte loads a thread keeps in flight

__global__ void __launch_bounds__(256) gn_stats_kernel(const uint4* __restrict__ x, double* __restrict__ sums, int HW,
                                                       int C, int G, int pix_per_block) {
  // block-level staging in fp64 as well: hundreds of threads add into each slot in an arbitrary order, and in fp32 that
  // order showed up as run-to-run differences of the statistics (the global sums are fp64 for the same reason)
  __shared__ double s_acc[2 * 128];
  const int nvec = C >> 3;
  const int ppb = blockDim.x / nvec;           // pixels processed per pass
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  const int v = threadIdx.x % nvec, pv = threadIdx.x / nvec;
  float2 s[4], ss[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = ss[i] = f2(0.0f);
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const uint4* base = x + (size_t)b * HW * nvec + v;
  for (int p = p0 + pv; p < p1; p += ppb * kGnBatch) {
    uint4 u[kGnBatch];
#pragma unroll
    for (int k = 0; k < kGnBatch; ++k)
      u[k] = (p + k * ppb < p1) ? __ldg(base + (size_t)(p + k * ppb) * nvec) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < kGnBatch; ++k) {
      float2 f[4];
      unpack8_2(u[k], f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i] = __fadd2_rn(s[i], f[i]);
        ss[i] = __ffma2_rn(f[i], f[i], ss[i]);
      }
    }
  }
  const int cpg = C / G;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int g0 = (v * 8 + 2 * i) / cpg, g1 = (v * 8 + 2 * i + 1) / cpg;
    atomicAdd(&s_acc[2 * g0], (double)s[i].x);
    atomicAdd(&s_acc[2 * g0 + 1], (double)ss[i].x);
    atomicAdd(&s_acc[2 * g1], (double)s[i].y);
    atomicAdd(&s_acc[2 * g1 + 1], (double)ss[i].y);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&sums[(size_t)b * 2 * G + i], s_acc[i]);
}

int gn_stats_run(const void* x, double* sums, int B, int HW, int C, int G, cudaStream_t stream) {
  TVAE_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 128 && C / 8 <= 256, "groupnorm: unsupported C=%d G=%d", C, G);
  TVAE_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * G * 2 * sizeof(double), stream));
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  // one full wave: ~6 resident blocks per SM (40 registers, 240 threads), each sweeping a contiguous pixel range; a
  // 2.3-wave grid of short blocks lost a quarter of its time to the tail
  const int slots = (num_sms() > 0 ? num_sms() : 148) * 6;
  int gx = slots / B;
  if (gx < 1) gx = 1;
  int ppb = (HW + gx - 1) / gx;
  if (ppb < 64) ppb = 64;
  dim3 grid((HW + ppb - 1) / ppb, B);
  gn_stats_kernel<<<grid, threads, 0, stream>>>(reinterpret_cast<const uint4*>(x), sums, HW, C, G, ppb);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// y = act((x - mean_g) * rstd_g * gamma_c + beta_c), act = SiLU or identity.  NHWC bf16 in / out.
// Replaces the affine half of nn.GroupNorm + F.silu (blocks.py:60-66; decoder.py:128-129).
// Algorithmic bytes: 4*C per pixel (read + write).  Packed fp32 math: 4 instructions per element.
template <bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const uint4* __restrict__ x, const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       uint4* __restrict__ y, int HW, int C, int G, float eps,
                                                       int vec_per_block) {
  extern __shared__ float s_ab[];  // scale[C], shift[C]
  const int b = blockIdx.y;
  const int cpg = C / G;
  const float inv_n = 1.0f / ((float)cpg * (float)HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float mean, rstd;
    gn_mean_rstd(sums, b, G, g, inv_n, eps, mean, rstd);
    const float a = rstd * gamma[c];
    s_ab[c] = a;
    s_ab[C + c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int nvec = C >> 3;
  const long long total = (long long)HW * nvec;
  const long long i0 = (long long)blockIdx.x * vec_per_block;
  const long long i1 = min(total, i0 + vec_per_block);
  // blockDim.x is a multiple of nvec and vec_per_block too, so each thread always sees the same 8 channels
  const int v = threadIdx.x % nvec;
  float2 a[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = make_float2(s_ab[v * 8 + 2 * i], s_ab[v * 8 + 2 * i + 1]);
    sh[i] = make_float2(s_ab[C + v * 8 + 2 * i], s_ab[C + v * 8 + 2 * i + 1]);
  }
  const uint4* xb = x + (size_t)b * total;
  uint4* yb = y + (size_t)b * total;
  const long long stride = blockDim.x;
  for (long long i = i0 + threadIdx.x; i < i1; i += stride * kGnBatch) {
    uint4 u[kGnBatch];
#pragma unroll
    for (int k = 0; k < kGnBatch; ++k)
      if (i + k * stride < i1) u[k] = __ldg(xb + i + k * stride);
#pragma unroll
    for (int k = 0; k < kGnBatch; ++k) {
      if (i + k * stride < i1) {
        float2 f[4];
        unpack8_2(u[k], f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 z = __ffma2_rn(f[q], a[q], sh[q]);
          f[q] = SILU ? silu2(z) : z;
        }
        yb[i + k * stride] = pack8_2(f);
      }
    }
  }
}

int gn_apply_run(const void* x, const double* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C,
                 int G, float eps, int apply_silu, cudaStream_t stream) {
  TVAE_REQUIRE(C % 8 == 0 && C % G == 0 && C / 8 <= 256, "groupnorm: unsupported C=%d G=%d", C, G);
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  const long long total = (long long)HW * nvec;
  long long vpb = (long long)threads * 4 * kGnBatch;
  while (vpb > threads && ((total + vpb - 1) / vpb) * B < 8LL * num_sms()) vpb >>= 1;
  vpb = (vpb / threads) * threads;
  if (vpb < threads) vpb = threads;
  dim3 grid((unsigned)((total + vpb - 1) / vpb), B);
  if (apply_silu)
    gn_apply_kernel<true><<<grid, threads, 2 * C * sizeof(float), stream>>>(
        reinterpret_cast<const uint4*>(x), sums, gamma, beta, reinterpret_cast<uint4*>(y), HW, C, G, eps, (int)vpb);
  else
    gn_apply_kernel<false><<<grid, threads, 2 * C * sizeof(float), stream>>>(
        reinterpret_cast<const uint4*>(x), sums, gamma, beta, reinterpret_cast<uint4*>(y), HW, C, G, eps, (int)vpb);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// silu(GroupNorm(x)) in one call: statistics pass + apply pass; `sums` (fp32 [B, G, 2]) is an output (the backward pass
// needs it).  (Running the two passes image by image so that the second read of x hits the 126 MB L2 was tried and is
// slower on B200: a 25 MB image is too small a launch -- 23 ms instead of 13 ms per inference step.)
int gn_stats_run(const void* x, float* sums, int B, int HW, int C, int G, cudaStream_t stream);
int gn_apply_run(const void* x, const double* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C,
                 int G, float eps, int apply_silu, cudaStream_t stream);

int gn_fwd_run(const void* x, double* sums, const float* gamma, const float* beta, void* y, int B, int HW, int C, int G,
               float eps, int apply_silu, cudaStream_t stream) {
  int rc;
  if ((rc = gn_stats_run(x, sums, B, HW, C, G, stream))) return rc;
  return gn_apply_run(x, sums, gamma, beta, y, B, HW, C, G, eps, apply_silu, stream);
}

// -------------------------------------------------------------------------------------------------
// Per-token statistics (one warp per token, row of C bf16 values):
//   mode 0 (FFN, blocks.py:149 RMSNorm):            out_a = 1/sqrt(mean(x^2) + 1e-6)
//   mode 1 (attention, blocks.py:146 RMSNorm followed by the three LayerNorms of attention.py:71-73, which
//           all see the same input h = x*w1/rms):    out_a = 1/(sigma*rms),  out_b = mu/sigma
//           with mu = mean(h), sigma = sqrt(var(h) + 1e-5).
//   mode 2 (bare attention module, no RMSNorm in front): as mode 1 with rms := 1.
// The normalised tensor is never materialised: the projection GEMM applies out_a / out_b in its epilogue
// (row_scale / row_shift of tvae_mtgemm) with the norm weights folded into the projection weights.
// Algorithmic bytes: 2*C per token.
// -------------------------------------------------------------------------------------------------
template <int VPL, int R>
__global__ void __launch_bounds__(256) row_stats_kernel(const uint4* __restrict__ x, const float* __restrict__ w1,
                                                        float* __restrict__ out_a, float* __restrict__ out_b,
                                                        long long M, int C, int mode) {
  // one warp handles R rows at a time (VPL 16-byte vectors per lane and row), all loads issued before the interleaved
  // warp reductions, persistent grid-stride loop (the one-short-lived-warp-per-row version reached 2.5 TB/s)
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const float invC = 1.0f / (float)C;
  float2 wv[VPL][4];
#pragma unroll
  for (int c = 0; c < VPL; ++c) {
    const int v = lane + c * 32;
    float4 wa = make_float4(0, 0, 0, 0), wb = wa;
    if (mode != 0 && v < nvec) {
      wa = __ldg(reinterpret_cast<const float4*>(w1) + 2 * v);
      wb = __ldg(reinterpret_cast<const float4*>(w1) + 2 * v + 1);
    }
    wv[c][0] = make_float2(wa.x, wa.y); wv[c][1] = make_float2(wa.z, wa.w);
    wv[c][2] = make_float2(wb.x, wb.y); wv[c][3] = make_float2(wb.z, wb.w);
  }
  const long long groups = (M + R - 1) / R;
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long grp = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); grp < groups; grp += warps_total) {
    const long long row0 = grp * R;
    uint4 u[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        const int v = lane + c * 32;
        u[r][c] = (v < nvec && row0 + r < M) ? __ldg(x + (row0 + r) * nvec + v) : make_uint4(0, 0, 0, 0);
      }
    float st[3 * R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 s2 = f2(0.0f), sw = f2(0.0f), sw2 = f2(0.0f);
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        float2 f[4];
        unpack8_2(u[r][c], f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          s2 = __ffma2_rn(f[q], f[q], s2);
          const float2 h = __fmul2_rn(f[q], wv[c][q]);
          sw = __fadd2_rn(sw, h);
          sw2 = __ffma2_rn(h, h, sw2);
        }
      }
      st[3 * r] = s2.x + s2.y;
      st[3 * r + 1] = sw.x + sw.y;
      st[3 * r + 2] = sw2.x + sw2.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < 3 * R; ++i) st[i] += __shfl_xor_sync(0xffffffffu, st[i], o);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row0 + r < M) {
          const float rms = (mode == 2) ? 1.0f : sqrtf(st[3 * r] * invC + 1e-6f);
          if (mode == 0) {
            out_a[row0 + r] = 1.0f / rms;
          } else {
            const float mu = st[3 * r + 1] * invC / rms;
            const float var = fmaxf(st[3 * r + 2] * invC / (rms * rms) - mu * mu, 0.0f);
            const float sigma = sqrtf(var + 1e-5f);
            out_a[row0 + r] = 1.0f / (sigma * rms);
            out_b[row0 + r] = mu / sigma;
          }
        }
      }
    }
  }
}

template <int VPL, int R>
static int launch_rs(const void* x, const float* w1, float* out_a, float* out_b, long long M, int C, int mode,
                     cudaStream_t stream) {
  const long long groups = (M + R - 1) / R;
  long long blocks = (groups + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  row_stats_kernel<VPL, R><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), w1, out_a, out_b, M, C, mode);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int row_stats_run(const void* x, const float* w1, float* out_a, float* out_b, long long M, int C, int mode,
                  cudaStream_t stream) {
  TVAE_REQUIRE(C % 8 == 0 && C / 8 <= 320, "row_stats: C=%d must be a multiple of 8, at most 2560", C);
  TVAE_REQUIRE(mode == 0 || (w1 != nullptr && out_b != nullptr), "row_stats: modes 1/2 need w1 and out_b");
  const int vpl = (C / 8 + 31) / 32;
  if (vpl <= 1) return launch_rs<1, 4>(x, w1, out_a, out_b, M, C, mode, stream);
  if (vpl == 2) return launch_rs<2, 2>(x, w1, out_a, out_b, M, C, mode, stream);
  if (vpl == 3) return launch_rs<3, 2>(x, w1, out_a, out_b, M, C, mode, stream);
  if (vpl == 4) return launch_rs<4, 2>(x, w1, out_a, out_b, M, C, mode, stream);
  if (vpl <= 6) return launch_rs<6, 1>(x, w1, out_a, out_b, M, C, mode, stream);
  return launch_rs<10, 1>(x, w1, out_a, out_b, M, C, mode, stream);
}

// -------------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 with zero channel padding (latent z -> decoder.conv_in operand).
// -------------------------------------------------------------------------------------------------
// One thread per (pixel, 8-channel group): reads are coalesced along the pixel axis of each channel plane, the 16-byte
// stores of the threads of a pixel line up into whole 128-byte rows (the first version -- one thread per bf16 element
// with a 64-bit div / mod chain -- reached 0.5 TB/s).
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, uint4* __restrict__ out, int B, int C,
                                                           int HW, int Cpad) {
  const int groups = Cpad >> 3;
  const long long total = (long long)B * HW;
  for (long long bp = (long long)blockIdx.x * blockDim.x + threadIdx.x; bp < total; bp += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(bp % HW);
    const int b = (int)(bp / HW);
    const float* src = in + (size_t)b * C * HW + p;
    for (int g = 0; g < groups; ++g) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
        v[k] = c < C ? __ldg(src + (size_t)c * HW) : 0.0f;
      }
      out[bp * groups + g] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  }
}

int nchw_to_nhwc_run(const float* in, void* out, int B, int C, int H, int W, int Cpad, cudaStream_t stream) {
  TVAE_REQUIRE(Cpad % 8 == 0 && C <= Cpad, "nchw_to_nhwc: Cpad %d must be a multiple of 8 and >= C %d", Cpad, C);
  const long long total = (long long)B * H * W;
  long long grid = (total + 255) / 256;
  if (grid > num_sms() * 16LL) grid = num_sms() * 16LL;
  nchw_to_nhwc_kernel<<<(int)grid, 256, 0, stream>>>(in, reinterpret_cast<uint4*>(out), B, C, H * W, Cpad);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// NHWC bf16 -> NCHW fp32 (first C of Cs channels); used to hand intermediate activations back to torch callers.
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B, int C, int HW,
                                    int Cs) {
  const long long total = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long bc = i / HW;
    const int c = (int)(bc % C);
    const int b = (int)(bc / C);
    out[i] = __bfloat162float(in[((size_t)b * HW + p) * Cs + c]);
  }
}

int nhwc_to_nchw_run(const void* in, float* out, int B, int C, int H, int W, int Cs, cudaStream_t stream) {
  const long long total = (long long)B * C * H * W;
  int grid = (int)((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  nhwc_to_nchw_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, B, C, H * W, Cs);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Reparameterisation (transvae.py:186-199; patched :186-196 and the clamps of :244-245).
//   patched: mu_c = clamp(mu, -50, 50), lv_c = clamp(logvar, -30, 20), z = mu_c + eps * exp(0.5 * lv_c)
//   main   : z = mu + eps * exp(0.5 * logvar)
// eps is drawn by the caller (torch generator) so that a seed reproduces the reference's sample.
// -------------------------------------------------------------------------------------------------
__global__ void reparam_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                               const float* __restrict__ eps, float* __restrict__ z, float* __restrict__ mu_out,
                               float* __restrict__ lv_out, long long n, int patched) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float m = mu[i], lv = logvar[i];
    if (patched) {
      m = fminf(fmaxf(m, -50.0f), 50.0f);
      lv = fminf(fmaxf(lv, -30.0f), 20.0f);
    }
    z[i] = m + eps[i] * expf(0.5f * lv);
    if (mu_out) mu_out[i] = m;
    if (lv_out) lv_out[i] = lv;
  }
}

int reparam_run(const float* mu, const float* logvar, const float* eps, float* z, float* mu_out, float* lv_out,
                long long n, int patched, cudaStream_t stream) {
  int grid = (int)((n + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  reparam_kernel<<<grid, 256, 0, stream>>>(mu, logvar, eps, z, mu_out, lv_out, n, patched);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// L1 + KL partial sums (vae_loss.py:83, :94-95; patched :80-104).
//   acc[0] += sum |f(recon) - target|   (f = sigmoid when patched)
//   acc[1] += sum -0.5 * (1 + lv - mu^2 - exp(lv))   (lv clamped when patched)
//   acc[2] += number of non-finite terms (cheap isfinite flag, no host sync)
// The caller divides by the element counts the reference uses.
// -------------------------------------------------------------------------------------------------
// Cross-block reduction without atomics on the values: every block stores its three partial sums, the last block to
// arrive (ticket counter) adds them in block order -- the loss terms are bit-reproducible from run to run.  (Launches of
// this kernel are ordered on one stream, so one staging array per device is enough.)
constexpr int kLossMaxBlocks = 1024;
__device__ float g_loss_part[3][kLossMaxBlocks];
__device__ unsigned int g_loss_ticket = 0;

__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ recon, const float* __restrict__ target,
                                                   const float* __restrict__ mu, const float* __restrict__ logvar,
                                                   float* __restrict__ acc, long long n_img, long long n_lat,
                                                   int patched, float clip_lo, float clip_hi) {
  float l1 = 0.0f, kl = 0.0f, bad = 0.0f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = n_img >> 2;
  for (long long i = t0; i < n4; i += stride) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(recon) + i);
    const float4 t = __ldg(reinterpret_cast<const float4*>(target) + i);
    float rv[4] = {r.x, r.y, r.z, r.w};
    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (patched) rv[k] = 1.0f / (1.0f + __expf(-rv[k]));
      const float d = fabsf(rv[k] - tv[k]);
      if (!isfinite(d)) bad += 1.0f;
      l1 += d;
    }
  }
  for (long long i = (n4 << 2) + t0; i < n_img; i += stride) {
    float rv = recon[i];
    if (patched) rv = 1.0f / (1.0f + __expf(-rv));
    l1 += fabsf(rv - target[i]);
  }
  for (long long i = t0; i < n_lat; i += stride) {
    const float m = mu[i];
    float lv = logvar[i];
    if (patched) lv = fminf(fmaxf(lv, clip_lo), clip_hi);
    const float v = -0.5f * (1.0f + lv - m * m - expf(lv));
    if (!isfinite(v)) bad += 1.0f;
    kl += v;
  }
  __shared__ float red[3][8];
  l1 = warp_sum(l1); kl = warp_sum(kl); bad = warp_sum(bad);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = l1; red[1][warp] = kl; red[2][warp] = bad; }
  __syncthreads();
  if (warp == 0) {
    l1 = lane < 8 ? red[0][lane] : 0.0f;
    kl = lane < 8 ? red[1][lane] : 0.0f;
    bad = lane < 8 ? red[2][lane] : 0.0f;
    l1 = warp_sum(l1); kl = warp_sum(kl); bad = warp_sum(bad);
    unsigned int ticket = 0;
    if (lane == 0) {
      g_loss_part[0][blockIdx.x] = l1;
      g_loss_part[1][blockIdx.x] = kl;
      g_loss_part[2][blockIdx.x] = bad;
      __threadfence();
      ticket = atomicInc(&g_loss_ticket, gridDim.x - 1);      // wraps to 0 after the last block: ready for the next launch
    }
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      float t[3] = {0.0f, 0.0f, 0.0f};
      for (int i = lane; i < (int)gridDim.x; i += 32) {
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] += __ldcg(&g_loss_part[k][i]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = warp_sum(t[k]);
      if (lane == 0) {
        acc[0] = t[0];
        acc[1] = t[1];
        acc[2] = t[2];
      }
    }
  }
}

int loss_run(const float* recon, const float* target, const float* mu, const float* logvar, float* acc,
             long long n_img, long long n_lat, int patched, float clip_lo, float clip_hi, cudaStream_t stream) {
  TVAE_CHECK_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(float), stream));
  int grid = (int)((n_img / 4 + 255) / 256);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  if (grid > kLossMaxBlocks) grid = kLossMaxBlocks;
  if (grid < 1) grid = 1;
  loss_kernel<<<grid, 256, 0, stream>>>(recon, target, mu, logvar, acc, n_img, n_lat, patched, clip_lo, clip_hi);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
