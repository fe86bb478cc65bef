// Optimiser step of the data-parallel training loop: clip_grad_norm_(max_norm) + fused AdamW over flat buffers, plus
// the small helpers of the gradient buckets (fp32 -> bf16 cast for bf16 gradient all-reduce, multi-tensor accumulate).
// Replaces (reference): torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW(fused=True) + LambdaLR warm-up + the
// skip-on-non-finite rule (train.py:608-620; train_2.py:266-274, 329-338, 349-366; train_working.py:384-397).
//
// Everything the step decides lives on the device, so a training step never synchronises with the host:
//   state fp32[8] = {0: sum of squared gradients of the last step (out), 1: clip factor of the last step (out),
//                    2: external skip flag (in, != 0 skips), 3: unused, 4: number of APPLIED updates, 5: number of
//                    skipped updates, 6: learning rate of the last applied update (out), 7: unused}
// The update index (bias correction, warm-up) is state[4]: a skipped step moves neither -- like the reference, whose
// `continue` runs before optimizer.step() and scheduler.step().
//
// The gradient norm is bit-reproducible and identical on every rank: kSumsqBlocks per-block partials, each formed in a
// fixed order (fp32 per thread, fp64 across the block), and every block of the update kernel re-adds the partial array
// in the same fixed order -- no floating-point atomics (an fp32 atomicAdd reduction made the clip factor differ in the
// last bits between ranks, which lets replicas drift apart).
#include "../../include/transvae_sm100.h"
#include "common.cuh"

namespace tvae {

constexpr int kSumsqBlocks = TVAE_SUMSQ_BLOCKS;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block reduction (256 threads) of one double per thread; result valid on thread 0
__device__ __forceinline__ double block_sum_f64(double v, double* red /*[8]*/) {
  v = warp_sum_f64(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    t = warp_sum_f64(t);
  }
  return t;
}

template <bool BF16>
__global__ void __launch_bounds__(256) sumsq_partials_kernel(const void* __restrict__ gv, long long n8,
                                                             double* __restrict__ partials) {
  // n8 = number of 8-element groups (fp32: two float4; bf16: one uint4)
  float acc = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    if constexpr (BF16) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(gv) + i);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      acc += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(gv) + 2 * i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(gv) + 2 * i + 1);
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
  }
  __shared__ double red[8];
  const double t = block_sum_f64((double)acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

int grad_sumsq_run(const void* g, int g_bf16, long long n, double* partials, cudaStream_t stream) {
  TVAE_REQUIRE(n % 8 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "grad_sumsq: buffer must be 16-byte aligned, n %% 8 == 0");
  if (g_bf16)
    sumsq_partials_kernel<true><<<kSumsqBlocks, 256, 0, stream>>>(g, n / 8, partials);
  else
    sumsq_partials_kernel<false><<<kSumsqBlocks, 256, 0, stream>>>(g, n / 8, partials);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

struct AdamArgs {
  float lr_base, b1, b2, eps, wd, max_norm, grad_scale;
  int warmup_steps;
};

// the decisions of one step, derived identically by every block from (partials, state)
struct AdamStep {
  float sumsq, clip, lr, bc1, bc2;
  bool skip;
};

__device__ __forceinline__ AdamStep adam_decide(const double* __restrict__ partials, const float* __restrict__ state,
                                                const AdamArgs& A, double* red) {
  double t = 0.0;
  for (int i = threadIdx.x; i < kSumsqBlocks; i += blockDim.x) t += partials[i];
  t = block_sum_f64(t, red);
  __shared__ AdamStep st;
  if (threadIdx.x == 0) {
    const float norm = sqrtf((float)t) * A.grad_scale;
    st.sumsq = (float)t;
    st.clip = A.max_norm > 0.0f ? fminf(1.0f, A.max_norm / (norm + 1e-6f)) : 1.0f;   // clip_grad_norm_'s formula
    st.skip = !isfinite(norm) || state[2] != 0.0f;
    const float k = state[4];                                  // updates applied so far = LambdaLR's step index
    st.lr = A.warmup_steps > 0 ? A.lr_base * fminf(1.0f, k / (float)A.warmup_steps) : A.lr_base;
    st.bc1 = 1.0f - powf(A.b1, k + 1.0f);
    st.bc2 = 1.0f - powf(A.b2, k + 1.0f);
  }
  __syncthreads();
  return st;
}

template <bool BF16>
__global__ void __launch_bounds__(256) adamw_step_kernel(float4* __restrict__ p, const void* __restrict__ gv,
                                                         float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                         const double* __restrict__ partials,
                                                         const float* __restrict__ state, const AdamArgs A) {
  __shared__ double red[8];
  const AdamStep S = adam_decide(partials, state, A, red);
  if (S.skip) return;
  const float s = A.grad_scale * S.clip;
  const float decay = 1.0f - S.lr * A.wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i], gg;
    if constexpr (BF16) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(gv) + i);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      gg = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gg = __ldg(reinterpret_cast<const float4*>(gv) + i);
    }
    float* P4 = reinterpret_cast<float*>(&pp);
    float* G4 = reinterpret_cast<float*>(&gg);
    float* M4 = reinterpret_cast<float*>(&mm);
    float* V4 = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G4[k] * s;
      M4[k] = A.b1 * M4[k] + (1.0f - A.b1) * gr;
      V4[k] = A.b2 * V4[k] + (1.0f - A.b2) * gr * gr;
      const float mh = M4[k] / S.bc1;
      const float vh = V4[k] / S.bc2;
      P4[k] = P4[k] * decay - S.lr * mh / (sqrtf(vh) + A.eps);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

// runs behind the update kernel (stream order): publishes the step's decisions and advances the counters
__global__ void __launch_bounds__(256) adamw_finish_kernel(const double* __restrict__ partials, float* __restrict__ state,
                                                           const AdamArgs A) {
  __shared__ double red[8];
  const AdamStep S = adam_decide(partials, state, A, red);
  if (threadIdx.x == 0) {
    state[0] = S.sumsq;
    state[1] = S.clip;
    if (S.skip) {
      state[5] += 1.0f;
    } else {
      state[4] += 1.0f;
      state[6] = S.lr;
    }
  }
}

int adamw_step_run(float* p, const void* g, int g_bf16, float* m, float* v, long long n, const double* partials,
                   float* state, float lr_base, int warmup_steps, float b1, float b2, float eps, float wd, float max_norm,
                   float grad_scale, cudaStream_t stream) {
  TVAE_REQUIRE(n % 8 == 0, "adamw: buffer length must be a multiple of 8");
  const AdamArgs A{lr_base, b1, b2, eps, wd, max_norm, grad_scale, warmup_steps};
  long long grid = (n / 4 + 255) / 256;
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  if (grid < 1) grid = 1;
  if (g_bf16)
    adamw_step_kernel<true><<<(int)grid, 256, 0, stream>>>(reinterpret_cast<float4*>(p), g, reinterpret_cast<float4*>(m),
                                                           reinterpret_cast<float4*>(v), n / 4, partials, state, A);
  else
    adamw_step_kernel<false><<<(int)grid, 256, 0, stream>>>(reinterpret_cast<float4*>(p), g, reinterpret_cast<float4*>(m),
                                                            reinterpret_cast<float4*>(v), n / 4, partials, state, A);
  TVAE_CHECK_CUDA(cudaGetLastError());
  adamw_finish_kernel<<<1, 256, 0, stream>>>(partials, state, A);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast of a gradient bucket (bf16 gradient all-reduce: halves the NVLink payload).  n % 8 == 0.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float4* __restrict__ in, uint4* __restrict__ out,
                                                            long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
    out[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

int cast_f32_bf16_run(const float* in, void* out, long long n, cudaStream_t stream) {
  TVAE_REQUIRE(n % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "cast_f32_bf16: 16-byte aligned buffers with n %% 8 == 0 expected");
  if (n == 0) return 0;
  long long grid = (n / 8 + 255) / 256;
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  cast_f32_bf16_kernel<<<(int)grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint4*>(out), n / 8);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Multi-tensor accumulate: dst[i][j] += src[i][j] for up to kMtaMax small fp32 tensors in ONE launch (the ~600 norm /
// bias / composite-weight gradients of a training step, which autograd otherwise accumulates with one tiny kernel each).
// One block column per tensor (blockIdx.y), grid-stride over its elements.
// -------------------------------------------------------------------------------------------------
struct MtaArgs {
  float* dst[TVAE_MTA_MAX];
  const float* src[TVAE_MTA_MAX];
  int n[TVAE_MTA_MAX];
  int sstride[TVAE_MTA_MAX];   // element stride of the source (1 = contiguous; 2 = one column of a [n, 2] matrix, ...)
};

__global__ void __launch_bounds__(256) mta_add_kernel(const __grid_constant__ MtaArgs A) {
  const int t = blockIdx.y;
  float* __restrict__ d = A.dst[t];
  const float* __restrict__ s = A.src[t];
  const int n = A.n[t];
  const int ss = A.sstride[t];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] += s[(size_t)i * ss];
}

int mta_add_run(float* const* dst, const float* const* src, const int* n, const int* sstride, int count, cudaStream_t stream) {
  TVAE_REQUIRE(count >= 0 && dst != nullptr && src != nullptr && n != nullptr, "multi_tensor_add: bad arguments");
  for (int base = 0; base < count; base += TVAE_MTA_MAX) {
    MtaArgs A;
    const int c = count - base < TVAE_MTA_MAX ? count - base : TVAE_MTA_MAX;
    int nmax = 0;
    for (int i = 0; i < c; ++i) {
      A.dst[i] = dst[base + i];
      A.src[i] = src[base + i];
      A.n[i] = n[base + i];
      A.sstride[i] = sstride != nullptr ? sstride[base + i] : 1;
      if (n[base + i] > nmax) nmax = n[base + i];
    }
    int gx = (nmax + 255) / 256;
    if (gx > 256) gx = 256;
    if (gx < 1) gx = 1;
    mta_add_kernel<<<dim3(gx, c), 256, 0, stream>>>(A);
    TVAE_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace tvae
