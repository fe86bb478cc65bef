// Packed-fp32 helpers for the HBM-bound elementwise kernels.
//
// At 6.5 TB/s an SM has to retire ~26 B per clock; with 2-byte activations that is 6-13 elements per clock, and the
// SM issues at most 128 lane-instructions per clock -- so a kernel that spends more than ~10 instructions per
// element is ISSUE-bound, not bandwidth-bound (the first GroupNorm-backward kernels ran at 35 % of the HBM peak with
// the schedulers 70 % busy).  Blackwell's packed fp32 instructions (FFMA2 / FADD2 / FMUL2: two lanes of fp32 per
// instruction) and the single-MUFU tanh form of the sigmoid roughly halve the instruction count per element.
#pragma once
#include "common.cuh"

namespace tvae {

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
// two bf16 packed in a u32 -> two fp32 (exact)
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(float2 v) { return pack_bf16(v.x, v.y); }
__device__ __forceinline__ void unpack8_2(const uint4& u, float2 (&f)[4]) {
  f[0] = bf16x2_to_f2(u.x); f[1] = bf16x2_to_f2(u.y); f[2] = bf16x2_to_f2(u.z); f[3] = bf16x2_to_f2(u.w);
}
__device__ __forceinline__ uint4 pack8_2(const float2 (&f)[4]) {
  return make_uint4(f2_to_bf16x2(f[0]), f2_to_bf16x2(f[1]), f2_to_bf16x2(f[2]), f2_to_bf16x2(f[3]));
}
// round to bf16 and back (what a consumer of the stored bf16 tensor will see)
__device__ __forceinline__ float2 round_bf16_2(float2 v) { return bf16x2_to_f2(f2_to_bf16x2(v)); }

// -------------------------------------------------------------------------------------------------
// Fixed-order cross-block reduction of one row of `n` floats per block (no atomics on the values, so two runs are
// bit-identical).  Every block has stored its row at rows[blockIdx.x * n ...]; the last block of each group of
// kOrdGroup consecutive blocks to arrive (ticket counter) adds the group's rows in block order into groups[g * n ...],
// and the last group to finish adds the group rows in group order into out.  Tickets wrap to zero (atomicInc), so the
// array is ready for the next launch.  All threads of the block call this; `s_flag` is a shared bool.
// tickets: [1 + ceil(nblocks / kOrdGroup)] zero-initialised unsigned ints.
// -------------------------------------------------------------------------------------------------
// sum of p[0], p[stride], ..., p[(count - 1) * stride] in that order, eight loads in flight at a time
__device__ __forceinline__ float ordered_column_sum(const float* p, size_t stride, unsigned int count) {
  float a = 0.0f;
  unsigned int k = 0;
  for (; k + 8 <= count; k += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + (size_t)(k + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) a += v[u];
  }
  for (; k < count; ++k) a += __ldcg(p + (size_t)k * stride);
  return a;
}

constexpr unsigned int kOrdGroup = 32;
constexpr unsigned int kOrdMaxGroups = 128;     // up to 4096 blocks
template <typename OutIndex>     // out[out_index(i)] receives column i
__device__ __forceinline__ void ordered_rows_reduce(const float* rows, float* groups, unsigned int* tickets, float* out, int n,
                                                    unsigned int nblocks, unsigned int bid, bool* s_flag, OutIndex out_index) {
  const unsigned int ngroups = (nblocks + kOrdGroup - 1) / kOrdGroup;
  const unsigned int grp = bid / kOrdGroup;
  const unsigned int gsize = min(kOrdGroup, nblocks - grp * kOrdGroup);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_flag = atomicInc(&tickets[1 + grp], gsize - 1) == gsize - 1;
  __syncthreads();
  if (!*s_flag) return;
  __threadfence();
  for (int i = threadIdx.x; i < n; i += blockDim.x) groups[(size_t)grp * n + i] = ordered_column_sum(rows + (size_t)grp * kOrdGroup * n + i, n, gsize);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_flag = atomicInc(&tickets[0], ngroups - 1) == ngroups - 1;
  __syncthreads();
  if (!*s_flag) return;
  __threadfence();
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[out_index(i)] = ordered_column_sum(groups + i, n, ngroups);
}
__device__ __forceinline__ void ordered_rows_reduce(const float* rows, float* groups, unsigned int* tickets, float* out, int n,
                                                    unsigned int nblocks, unsigned int bid, bool* s_flag) {
  ordered_rows_reduce(rows, groups, tickets, out, n, nblocks, bid, s_flag, [](int i) { return i; });
}

__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.TANH, |abs err| < 2^-10.9
  return r;
}
// silu(y) = y * sigmoid(y) = h + h * tanh(h), h = y / 2: FMUL2 + 2 MUFU + FFMA2 for two elements.
__device__ __forceinline__ float2 silu2(float2 y) {
  const float2 h = __fmul2_rn(y, f2(0.5f));
  float2 t;
  t.x = tanh_approx(h.x);
  t.y = tanh_approx(h.y);
  return __ffma2_rn(h, t, h);
}
// d silu / dy = s * (1 + y * (1 - s)), s = sigmoid(y) = 0.5 + 0.5 * tanh(y / 2)
__device__ __forceinline__ float2 silu_grad2(float2 y) {
  const float2 h = __fmul2_rn(y, f2(0.5f));
  float2 t;
  t.x = tanh_approx(h.x);
  t.y = tanh_approx(h.y);
  const float2 s = __ffma2_rn(t, f2(0.5f), f2(0.5f));
  const float2 om = __ffma2_rn(t, f2(-0.5f), f2(0.5f));
  return __fmul2_rn(s, __ffma2_rn(y, om, f2(1.0f)));
}
// erf-GELU for two elements with ONE MUFU per element: 0.5 * erfc(|x| / sqrt 2) = exp2(q(|x|)), q a degree-5 polynomial
// fitted to log2(erfc) on |x| / sqrt 2 in [0, 4] (max |error| of Phi 3.2e-7, of gelu 1.0e-6; beyond the fit range
// q keeps falling, so the tail underflows to the exact limit).  Phi(x) = 0.5 + copysign(0.5 - exp2(q), x).
// (Abramowitz-Stegun 7.1.26, used before, needs a reciprocal and an exponential: 2 MUFU per element made every GELU
// epilogue and the activation kernels MUFU-bound at 16 results / clk / SM.)
__device__ __forceinline__ float2 gelu_cdf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  float2 q = __ffma2_rn(f2(-5.188658834e-04f), ax, f2(7.387229707e-03f));
  q = __ffma2_rn(q, ax, f2(-5.253804848e-02f));
  q = __ffma2_rn(q, ax, f2(-4.592766762e-01f));
  q = __ffma2_rn(q, ax, f2(-1.151083112e+00f));
  q = __ffma2_rn(q, ax, f2(-1.000000834e+00f));
  float2 e;
  e.x = exp2f(q.x);
  e.y = exp2f(q.y);
  const float2 h = __ffma2_rn(e, f2(-1.0f), f2(0.5f));          // 0.5 - 0.5 erfc >= 0
  return make_float2(0.5f + copysignf(h.x, x.x), 0.5f + copysignf(h.y, x.y));
}
__device__ __forceinline__ float2 gelu2(float2 x) { return __fmul2_rn(x, gelu_cdf2(x)); }
// d gelu / dx = Phi(x) + x * phi(x), phi(x) = exp(-x^2 / 2) / sqrt(2 pi) = exp2(-x^2 * log2(e) / 2 - log2(sqrt(2 pi)))
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  const float2 q = __ffma2_rn(__fmul2_rn(x, x), f2(-0.7213475204444817f), f2(-1.3257480647361593f));
  float2 pdf;
  pdf.x = exp2f(q.x);
  pdf.y = exp2f(q.y);
  return __ffma2_rn(x, pdf, gelu_cdf2(x));
}

// GroupNorm coefficients from the (sum, sum of squares) pair of a group.  The sums are fp64: block partials (fp32, fixed
// order) are added into them with fp64 atomics, so the order of the atomics no longer shows in the result, and
// E[x^2] - mean^2 is formed in fp64 -- in fp32 the cancellation (|mean| >> sigma in single-channel groups) turned the
// rounding noise of the atomics into 1e-4 relative differences of rstd between identical runs, i.e. a flipped bf16
// rounding in 2-3 % of the outputs of the first ResBlock and a 1.3 % (l2) run-to-run difference of the reconstruction.
__device__ __forceinline__ void gn_mean_rstd(const double* __restrict__ sums, int b, int G, int g, float inv_n, float eps,
                                             float& mean, float& rstd) {
  const double m = sums[((size_t)b * G + g) * 2] * (double)inv_n;
  const double var = fmax(sums[((size_t)b * G + g) * 2 + 1] * (double)inv_n - m * m, 0.0);
  mean = (float)m;
  rstd = rsqrtf((float)var + eps);      // three fp64 operations per call; the square root does not need them
}

}  // namespace tvae
