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
// erf-GELU for two elements (Abramowitz-Stegun 7.1.26 as fast_erf, packed): Phi(x) = 0.5 * (1 + erf(x / sqrt 2)).
// Returns Phi in .cdf and exp(-x^2 / 2) in .e (shared by the derivative).
struct Gelu2 {
  float2 cdf, e;
};
__device__ __forceinline__ Gelu2 gelu_parts2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 d = __ffma2_rn(ax, f2(0.3275911f * 0.70710678118654752f), f2(1.0f));
  float2 t;
  t.x = rcp_approx(d.x);
  t.y = rcp_approx(d.y);
  float2 p = __ffma2_rn(f2(1.061405429f), t, f2(-1.453152027f));
  p = __ffma2_rn(p, t, f2(1.421413741f));
  p = __ffma2_rn(p, t, f2(-0.284496736f));
  p = __ffma2_rn(p, t, f2(0.254829592f));
  p = __fmul2_rn(p, t);
  const float2 q = __fmul2_rn(__fmul2_rn(x, x), f2(-0.7213475204444817f));   // -x^2/2 * log2(e)
  Gelu2 r;
  r.e.x = exp2f(q.x);
  r.e.y = exp2f(q.y);
  // erf(|x|/sqrt2) = 1 - p*e ; Phi(x) = 0.5 + 0.5 * sign(x) * erf(|x|/sqrt2)
  const float2 half_erf = __ffma2_rn(__fmul2_rn(p, r.e), f2(-0.5f), f2(0.5f));
  r.cdf = make_float2(0.5f + copysignf(half_erf.x, x.x), 0.5f + copysignf(half_erf.y, x.y));
  return r;
}
__device__ __forceinline__ float2 gelu2(float2 x) { return __fmul2_rn(x, gelu_parts2(x).cdf); }
// d gelu / dx = Phi(x) + x * phi(x), phi(x) = exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  const Gelu2 g = gelu_parts2(x);
  return __ffma2_rn(__fmul2_rn(x, f2(0.3989422804014327f)), g.e, g.cdf);
}

}  // namespace tvae
