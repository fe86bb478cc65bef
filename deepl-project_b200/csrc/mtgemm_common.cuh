// Shared pieces of the fused multi-tap implicit GEMM kernels (1-CTA: mtgemm.cu, CTA-pair: mtgemm2.cu): launch
// parameters, compile-time epilogue variants and the per-8-column epilogue math.
#pragma once
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "ew_common.cuh"
#include "tmap.cuh"

namespace tvae {

struct MtTap {
  int32_t c_off;
  int32_t wk_off;
  int16_t kblocks;
  int8_t map, dw, p, dh;
  int8_t pad_[2];
};
static_assert(sizeof(MtTap) == 16, "MtTap must be 16 bytes");

struct MtParams {
  int tiles_w, tiles_h, tiles_b, n_tiles;
  int tw, th, nb;
  int num_phases;
  int ntaps[TVAE_MAX_PHASES];
  int out_p[TVAE_MAX_PHASES];
  int out_c_off[TVAE_MAX_PHASES];
  MtTap taps[TVAE_MAX_PHASES][TVAE_MAX_TAPS];
  const float* bias;
  int n_total;
  int act;
  const float* row_scale;
  const float* row_shift;
  const float* col_sum;
  int has_residual;
  const float2* rope_tab;
  int rope_C, rope_H, rope_W;
  float q_scale;
  float* out_f32;
  int out_n;
  int vB, vH, vW;  // output view extents (pixels) for row indexing / bounds
  const __nv_bfloat16* z;   // act_grad == 2: pre-activation matrix [M, n_total]
  double* gn_sums;          // fused GroupNorm statistics of the output: [vB][gn_groups][2] (sum, sumsq), or nullptr
  int gn_groups, gn_cpg;    // groups, channels per group
  // halo mode (3x3 stride-1 convolution, 128-pixel-wide tiles): ONE (128 + 2)-pixel A tile per (kernel row, K block)
  // serves the three dx taps through row-offset UMMA descriptors.  0 = off, 1 = on
  int halo;
  // chunk-pipelined residual epilogue of the CTA-pair kernel (mtgemm2.cu); 0 = one residual barrier per tile
  int pipe_res;
  // kEpiGnBwd: the launch is the input-gradient GEMM whose output dh feeds the backward of act(GroupNorm(x)); its epilogue
  // also produces, per 128-pixel tile and channel, (sum dy, sum dy * xhat) with dy = dh * act'(gamma * xhat + beta):
  // gnb_part [pixel tiles][n_total][2] fp32, plain stores (every element written once)
  const void* gnb_x;            // the GroupNorm input, bf16, same (plain) geometry as the output
  const double* gnb_sums;       // GroupNorm statistics of x: [vB][gnb_groups][2] (sum, sumsq)
  const float* gnb_gamma;
  const float* gnb_beta;
  float* gnb_part;
  int gnb_groups, gnb_cpg, gnb_silu;
  float gnb_eps, gnb_inv_n;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kThreads = 384;   // 4 control warps + 8 epilogue warps
constexpr int kHaloPix = kBlockM + 2;                       // pixels of a halo tile
constexpr int kHaloATx = kHaloPix * kBlockK * 2;            // bytes one halo load delivers (16640)
constexpr int kHaloABytes = (kHaloATx + 1023) / 1024 * 1024;  // its shared-memory slot (17408)

template <int BLOCK_N>
struct MtCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kOutBytes = kBlockM * BLOCK_N * 2;
  static constexpr int kStages = (227 * 1024 - 2048 - kOutBytes) / (kABytes + kBBytes) > 8
                                     ? 8
                                     : (227 * 1024 - 2048 - kOutBytes) / (kABytes + kBBytes);
  static constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + kOutBytes + 1024 /*align slack*/ + 256 /*bars*/;
};

// Epilogue variants (compile-time, so each kernel's epilogue is a few hundred instructions: the first version kept
// every option as a run-time branch inside a 32x unrolled loop and was instruction-fetch bound -- ncu showed
// stall_no_inst on the epilogue warps and the MMA warp waiting on tmem_empty).
enum Epi : int {
  kEpiBias = 0,        // v = acc + bias
  kEpiBiasRes = 1,     // v = acc + bias + residual
  kEpiBiasGelu = 2,    // v = gelu(acc + bias)
  kEpiBiasSilu = 3,    // v = silu(acc + bias)
  kEpiRsBiasGelu = 4,  // v = gelu(rs[m] * acc + bias)                      (RMSNorm folded into proj_in)
  kEpiAffineRope = 5,  // v = rope(rs[m] * acc - rsh[m] * cs[n] + bias)     (RMSNorm + LayerNorm folded into QKV)
  kEpiDirect = 6,      // fp32 NCHW direct store of acc + bias
  kEpiRsBias = 7,      // v = rs[m] * acc - rsh[m] * cs[n] + bias [+ residual] (generic, tests / bare modules)
  // backward: the input-gradient GEMM applies the derivative of the producing layer's activation itself
  kEpiMulGeluGrad = 8,     // v = acc * gelu'(z),          z tile staged like a residual
  kEpiMulSiluGrad = 9,     // v = acc * silu'(z)
  kEpiResMulGeluGrad = 10, // v = (acc + residual) * gelu'(z),  residual staged; z staged (pair kernel) or read from global memory
  // training forward: ONE launch stores the pre-activation (what the backward pass needs) and the activation (what the
  // next layer reads) -- the activation used to be a separate pass over HBM (tvae_act_fwd, 86 launches per micro-step)
  kEpiBiasGeluDual = 11,   // out = acc + bias,  out_act = gelu(bf16(out))       (CTA-pair kernel only)
  kEpiBiasSiluDual = 12,   // out = acc + bias,  out_act = silu(bf16(out))
  // backward: input-gradient GEMM + the reduce pass of the GroupNorm backward that consumes its output (x tile staged)
  kEpiGnBwd = 13,          // out = acc;  gnb_part[tile][n] = (sum dy, sum dy * xhat) over the tile's pixels   (pair kernel only)
};
template <int EPI>
__host__ __device__ constexpr bool epi_is_dual() {
  return EPI == kEpiBiasGeluDual || EPI == kEpiBiasSiluDual;
}

template <int EPI>
__host__ __device__ constexpr bool epi_has_res() {
  return EPI == kEpiBiasRes || EPI == kEpiRsBias || EPI == kEpiMulGeluGrad || EPI == kEpiMulSiluGrad || EPI == kEpiResMulGeluGrad;
}

// Combine 8 accumulator values with the 8 bf16 values of the staged tile (`r`): residual add, or multiplication by the
// activation derivative; EPI 10 additionally reads the pre-activation from global memory (`zrow` -> this row, column n).
template <int EPI>
__device__ __forceinline__ void epi_combine8(float (&f)[8], const uint4& r, const __nv_bfloat16* zrow, int n) {
  float2 t[4] = {bf16x2_to_f2(r.x), bf16x2_to_f2(r.y), bf16x2_to_f2(r.z), bf16x2_to_f2(r.w)};
  if constexpr (EPI == kEpiMulGeluGrad || EPI == kEpiMulSiluGrad) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = EPI == kEpiMulGeluGrad ? gelu_grad2(t[k]) : silu_grad2(t[k]);
      f[2 * k] *= g.x;
      f[2 * k + 1] *= g.y;
    }
  } else if constexpr (EPI == kEpiResMulGeluGrad) {
    const uint4 zu = __ldg(reinterpret_cast<const uint4*>(zrow + n));
    const float2 z[4] = {bf16x2_to_f2(zu.x), bf16x2_to_f2(zu.y), bf16x2_to_f2(zu.z), bf16x2_to_f2(zu.w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = gelu_grad2(z[k]);
      f[2 * k] = (f[2 * k] + t[k].x) * g.x;
      f[2 * k + 1] = (f[2 * k + 1] + t[k].y) * g.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] += t[k].x;
      f[2 * k + 1] += t[k].y;
    }
  }
}

// f = (f + residual) * gelu'(z) with both tiles staged in shared memory (CTA-pair kernel)
__device__ __forceinline__ void epi_res_mul_gelu_grad8(float (&f)[8], const uint4& r, const uint4& zu) {
  const float2 t[4] = {bf16x2_to_f2(r.x), bf16x2_to_f2(r.y), bf16x2_to_f2(r.z), bf16x2_to_f2(r.w)};
  const float2 z[4] = {bf16x2_to_f2(zu.x), bf16x2_to_f2(zu.y), bf16x2_to_f2(zu.z), bf16x2_to_f2(zu.w)};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 g = gelu_grad2(z[k]);
    f[2 * k] = (f[2 * k] + t[k].x) * g.x;
    f[2 * k + 1] = (f[2 * k + 1] + t[k].y) * g.y;
  }
}

// act(z) of eight staged bf16 pre-activations (the rounded values: exactly what a separate pass over the stored z gives)
template <int EPI>
__device__ __forceinline__ uint4 epi_act_of_bf16(const uint4& zu) {
  float2 z[4] = {bf16x2_to_f2(zu.x), bf16x2_to_f2(zu.y), bf16x2_to_f2(zu.z), bf16x2_to_f2(zu.w)};
#pragma unroll
  for (int k = 0; k < 4; ++k) z[k] = EPI == kEpiBiasGeluDual ? gelu2(z[k]) : silu2(z[k]);
  return make_uint4(f2_to_bf16x2(z[0]), f2_to_bf16x2(z[1]), f2_to_bf16x2(z[2]), f2_to_bf16x2(z[3]));
}

enum : int { kEpiCount = 14 };

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// tcgen05.wait::ld that also "produces" the sixteen loaded registers, so the compiler cannot move a use of them above
// the wait (needed once loads for the NEXT column group are in flight while the current group is being processed)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]),
                 "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
               :
               : "memory");
}

struct EpiRow {
  float rs, rsh;
  int rope_r, rope_c;
  bool row_ok;
  int pw, phh, pb;
};

// Epilogue math for 8 consecutive output columns [n, n+8) of one row.
template <int EPI>
__device__ __forceinline__ void epi_math8(const MtParams& P, const float* __restrict__ bias, const EpiRow& R, int n,
                                          const uint32_t (&v)[8], float (&f)[8]) {
  float bv[8];
  if (bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n) + 1);
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) bv[k] = 0.0f;
  }
  if constexpr (EPI == kEpiRsBiasGelu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(__uint_as_float(v[k]), R.rs, bv[k]);
  } else if constexpr (EPI == kEpiAffineRope || EPI == kEpiRsBias) {
    if (P.row_shift != nullptr) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(P.col_sum + n));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(P.col_sum + n) + 1);
      const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) bv[k] = fmaf(-R.rsh, cv[k], bv[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(__uint_as_float(v[k]), R.rs, bv[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[k]) + bv[k];
  }
  if constexpr (EPI == kEpiBiasGelu || EPI == kEpiRsBiasGelu) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) {       // packed fp32, one MUFU per element (ew_common.cuh)
      const float2 r = gelu2(make_float2(f[k], f[k + 1]));
      f[k] = r.x;
      f[k + 1] = r.y;
    }
  } else if constexpr (EPI == kEpiBiasSilu) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float2 r = silu2(make_float2(f[k], f[k + 1]));
      f[k] = r.x;
      f[k + 1] = r.y;
    }
  }
  if constexpr (EPI == kEpiAffineRope) {
    if (P.rope_tab != nullptr && n < 2 * P.rope_C) {
      // columns n..n+7 sit in one half of a 64-wide head: the first half rotates with the row index, the second with
      // the column index (attention.py:161-170); pair (2i, 2i+1) uses the angle of slot 2i for the even output and of
      // slot 2i+1 for the odd output (attention.py:178-197).
      const int j0 = n & 63;
      const int pos = (j0 < 32) ? R.rope_r : R.rope_c;
      const float4* tab = reinterpret_cast<const float4*>(P.rope_tab + (size_t)pos * 16 + (j0 & 15));
      const float qs = (n < P.rope_C) ? P.q_scale : 1.0f;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float4 t = __ldg(tab + (k >> 1));   // (cos a, sin a, cos b, sin b)
        const float a = f[k], b = f[k + 1];
        f[k] = (a * t.x - b * t.y) * qs;
        f[k + 1] = (a * t.w + b * t.z) * qs;
      }
    }
  }
}

}  // namespace tvae
