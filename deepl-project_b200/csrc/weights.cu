// Weight-side re-layout kernels of the training step (HBM-bound, once per weight and micro-step).
//
// The reference keeps nn.Linear weights as [out, in] and nn.Conv2d weights as [out, in, kh, kw] fp32
// (transvae/modules/conv.py:39-65, blocks.py:34-37, attention.py:43-48); the tensor-core kernels want bf16 operands
// with K contiguous: the forward GEMM reads W_f[out][tap * in + i], the input-gradient GEMM reads
// W_d[in][tap * out + o].  Doing these permutes with strided torch copies cost 24 ms per training micro-step
// (1.05 G parameters, at:: elementwise kernels at a few hundred GB/s); here each is one coalesced tile transpose
// through shared memory.
//
//   weight_pack   : fp32 ref[A][B][T]  ->  bf16 fwd[A][T][B]  and / or  bf16 dgr[B][T][A]      (T = 1 or 9)
//   wgrad_unpack  : fp32 packed gradient g[A][T][B]  ->  fp32 ref-layout gradient [A][B][T]
//
// Algorithmic bytes: pack 4 + 2 (+ 2) per parameter, unpack 8 per parameter.
#include "../../include/transvae_sm100.h"
#include "common.cuh"

namespace tvae {

constexpr int kWpTile = 32;

// One block = one 32 (a) x 32 (b) tile with all T taps.  smem row = one `a`: [b][t] as in memory, padded to an odd
// number of floats so that both read-back patterns (b fastest with stride T, a fastest with the row stride) are
// conflict-free for T = 1 and T = 9.
template <int T>
__global__ void __launch_bounds__(256) weight_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                                          __nv_bfloat16* __restrict__ dgr, int A, int B) {
  constexpr int kRow = kWpTile * T + 1;   // odd
  __shared__ float tile[kWpTile][kRow];
  const int a0 = blockIdx.y * kWpTile, b0 = blockIdx.x * kWpTile;
  const int nb = min(kWpTile, B - b0), na = min(kWpTile, A - a0);
  // load: for each a the segment w[a][b0 .. b0+nb)[0 .. T) is contiguous (nb * T floats)
  for (int i = threadIdx.x; i < kWpTile * kWpTile * T; i += 256) {
    const int a = i / (kWpTile * T), r = i - a * (kWpTile * T);
    if (a < na && r < nb * T) tile[a][r] = __ldg(w + ((size_t)(a0 + a) * B + b0) * T + r);
  }
  __syncthreads();
  if (fwd != nullptr) {
    // fwd[a][t * B + b]: lanes run over pairs of b (one 4-byte store each, 64-byte runs per (a, t))
    for (int i = threadIdx.x; i < kWpTile * T * (kWpTile / 2); i += 256) {
      const int bp = i % (kWpTile / 2), at = i / (kWpTile / 2);
      const int t = at % T, a = at / T;
      const int b = 2 * bp;
      if (a < na && b < nb) {
        __nv_bfloat16* dst = fwd + ((size_t)(a0 + a) * T + t) * B + b0 + b;
        if (b + 1 < nb) {
          *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(tile[a][b * T + t], tile[a][(b + 1) * T + t]);
        } else {
          *dst = __float2bfloat16_rn(tile[a][b * T + t]);
        }
      }
    }
  }
  if (dgr != nullptr) {
    // dgr[b][t * A + a]: lanes run over pairs of a
    for (int i = threadIdx.x; i < kWpTile * T * (kWpTile / 2); i += 256) {
      const int ap = i % (kWpTile / 2), bt = i / (kWpTile / 2);
      const int t = bt % T, b = bt / T;
      const int a = 2 * ap;
      if (b < nb && a < na) {
        __nv_bfloat16* dst = dgr + ((size_t)(b0 + b) * T + t) * A + a0 + a;
        if (a + 1 < na) {
          *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(tile[a][b * T + t], tile[a + 1][b * T + t]);
        } else {
          *dst = __float2bfloat16_rn(tile[a][b * T + t]);
        }
      }
    }
  }
}

// T = 1 (nn.Linear): 64 x 64 tiles, 16-byte loads, 8-byte stores (128-byte runs on both outputs).
__global__ void __launch_bounds__(256) weight_pack_lin_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                                              __nv_bfloat16* __restrict__ dgr, int A, int B) {
  __shared__ float tile[64][65];
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int a = a0 + r0 + 16 * k;
    v[k] = (a < A && b0 + c4 < B) ? __ldg(reinterpret_cast<const float4*>(w + (size_t)a * B + b0 + c4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + 16 * k;
    tile[r][c4] = v[k].x; tile[r][c4 + 1] = v[k].y; tile[r][c4 + 2] = v[k].z; tile[r][c4 + 3] = v[k].w;
    if (fwd != nullptr && a0 + r < A && b0 + c4 < B) {
      uint2 o;
      o.x = pack_bf16(v[k].x, v[k].y);
      o.y = pack_bf16(v[k].z, v[k].w);
      *reinterpret_cast<uint2*>(fwd + (size_t)(a0 + r) * B + b0 + c4) = o;
    }
  }
  if (dgr == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int b = r0 + 16 * k;     // row of the transposed tile
    if (b0 + b < B && a0 + c4 < A) {
      uint2 o;
      o.x = pack_bf16(tile[c4][b], tile[c4 + 1][b]);
      o.y = pack_bf16(tile[c4 + 2][b], tile[c4 + 3][b]);
      *reinterpret_cast<uint2*>(dgr + (size_t)(b0 + b) * A + a0 + c4) = o;
    }
  }
}

int weight_pack_run(const float* w, void* fwd, void* dgr, int A, int B, int T, cudaStream_t stream) {
  TVAE_REQUIRE(w != nullptr && (fwd != nullptr || dgr != nullptr), "weight_pack: missing operand");
  TVAE_REQUIRE(A >= 1 && B >= 1 && (T == 1 || T == 9), "weight_pack: unsupported shape [%d][%d][%d] (T = 1 or 9)", A, B, T);
  if (T == 1) {
    // 16-byte loads / 8-byte stores need rows that are multiples of four elements
    TVAE_REQUIRE(A % 4 == 0 && B % 4 == 0, "weight_pack: A = %d and B = %d must be multiples of 4 for a matrix", A, B);
    dim3 grid((B + 63) / 64, (A + 63) / 64);
    TVAE_REQUIRE(grid.y <= 65535, "weight_pack: A = %d too large", A);
    weight_pack_lin_kernel<<<grid, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(fwd),
                                                     reinterpret_cast<__nv_bfloat16*>(dgr), A, B);
  } else {
    // the paired 4-byte stores need even row lengths
    TVAE_REQUIRE(A % 2 == 0 && B % 2 == 0, "weight_pack: A = %d and B = %d must be even", A, B);
    dim3 grid((B + kWpTile - 1) / kWpTile, (A + kWpTile - 1) / kWpTile);
    TVAE_REQUIRE(grid.y <= 65535, "weight_pack: A = %d too large", A);
    weight_pack_kernel<9><<<grid, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(fwd),
                                                    reinterpret_cast<__nv_bfloat16*>(dgr), A, B);
  }
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// g[a][t][b] -> out[a][b][t]: one block = one `a` and 128 consecutive b; reads T runs of 512 bytes, writes one run of
// 128 * T floats.
template <int T>
__global__ void __launch_bounds__(128) wgrad_unpack_kernel(const float* __restrict__ g, float* __restrict__ out, int A, int B,
                                                           int accumulate) {
  __shared__ float tile[T][128 + 1];
  const int a = blockIdx.y, b0 = blockIdx.x * 128;
  const int nb = min(128, B - b0);
  const float* src = g + (size_t)a * T * B + b0;
#pragma unroll
  for (int t = 0; t < T; ++t)
    if ((int)threadIdx.x < nb) tile[t][threadIdx.x] = __ldg(src + (size_t)t * B + threadIdx.x);
  __syncthreads();
  float* dst = out + ((size_t)a * B + b0) * T;
  for (int i = threadIdx.x; i < nb * T; i += 128) {
    const int b = i / T, t = i - b * T;
    dst[i] = accumulate ? dst[i] + tile[t][b] : tile[t][b];
  }
}

int wgrad_unpack_run(const float* g, float* out, int A, int B, int T, int accumulate, cudaStream_t stream) {
  TVAE_REQUIRE(g != nullptr && out != nullptr, "wgrad_unpack: missing operand");
  TVAE_REQUIRE(A >= 1 && A <= 65535 && B >= 1 && T == 9, "wgrad_unpack: unsupported shape [%d][%d][%d] (T = 9)", A, T, B);
  dim3 grid((B + 127) / 128, A);
  wgrad_unpack_kernel<9><<<grid, 128, 0, stream>>>(g, out, A, B, accumulate);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Composite packs of the training step.  With differentiable torch expressions these were ~830 (QKV fold, 26 blocks) and
// ~700 (Upsample conv1, 4 layers) tiny at:: launches per micro-step -- 5 ms of a 240 ms step (torch profiler,
// profiles/r2i_at_sources.txt).  One kernel each way here; all reductions run in a fixed order (bit-reproducible).
//
// fold_qkv (attention.py:71-79: q = to_q(norm_q(x)) etc. with LayerNorm affines g, b):
//     wg[s*C + n][k] = w_s[n][k] * g_s[k],      bg[s*C + n] = sum_k w_s[n][k] * b_s[k]             s = q, k, v
// backward:
//     dw_s[n][k] = dwg[s*C+n][k] * g_s[k] + dbg[s*C+n] * b_s[k]
//     dg_s[k] = sum_n dwg[s*C+n][k] * w_s[n][k],     db_s[k] = sum_n dbg[s*C+n] * w_s[n][k]
// -------------------------------------------------------------------------------------------------
struct FoldPtrs {
  const float* w[3];
  const float* g[3];
  const float* b[3];
};
struct FoldGradPtrs {
  float* dw[3];
  float* dg[3];
  float* db[3];
};

__global__ void __launch_bounds__(256) fold_qkv_fwd_kernel(FoldPtrs P, float* __restrict__ wg, float* __restrict__ bg, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);       // one warp per output row
  const int lane = threadIdx.x & 31;
  if (row >= 3 * C) return;
  const int s = row / C, n = row - s * C;
  const float* __restrict__ w = P.w[s] + (size_t)n * C;
  const float* __restrict__ g = P.g[s];
  const float* __restrict__ b = P.b[s];
  float acc = 0.0f;
  for (int k = lane * 4; k < C; k += 128) {                   // C % 4 == 0
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + k));
    const float4 bv = __ldg(reinterpret_cast<const float4*>(b + k));
    *reinterpret_cast<float4*>(wg + (size_t)row * C + k) = make_float4(wv.x * gv.x, wv.y * gv.y, wv.z * gv.z, wv.w * gv.w);
    acc = fmaf(wv.x, bv.x, fmaf(wv.y, bv.y, fmaf(wv.z, bv.z, fmaf(wv.w, bv.w, acc))));
  }
  acc = warp_sum(acc);
  if (lane == 0) bg[row] = acc;
}

// block = (32 columns, source s) x 32 row groups; the partial column sums of the row groups meet in shared memory and are
// added in a fixed order.  (With 8 row groups a thread walked C / 8 rows one dependent load batch at a time: 57 us per
// launch at C = 1536.)
constexpr int kFoldRowGroups = 32;
__global__ void __launch_bounds__(32 * kFoldRowGroups) fold_qkv_bwd_kernel(FoldPtrs P, FoldGradPtrs G, const float* __restrict__ dwg,
                                                                           const float* __restrict__ dbg, int C, int accumulate) {
  __shared__ float s_dg[kFoldRowGroups][33], s_db[kFoldRowGroups][33];
  const int s = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bool ok = c < C;
  const float* __restrict__ w = P.w[s];
  const float gc = ok ? __ldg(P.g[s] + c) : 0.0f, bc = ok ? __ldg(P.b[s] + c) : 0.0f;
  float* __restrict__ dw = G.dw[s];
  float dg = 0.0f, db = 0.0f;
  if (ok) {
#pragma unroll 8
    for (int r = ty; r < C; r += kFoldRowGroups) {
      const float dv = __ldg(dwg + ((size_t)s * C + r) * C + c);
      const float wv = __ldg(w + (size_t)r * C + c);
      const float dbn = __ldg(dbg + s * C + r);
      const float dwv = fmaf(dv, gc, dbn * bc);
      dw[(size_t)r * C + c] = accumulate ? dw[(size_t)r * C + c] + dwv : dwv;
      dg = fmaf(dv, wv, dg);
      db = fmaf(dbn, wv, db);
    }
  }
  s_dg[ty][tx] = dg;
  s_db[ty][tx] = db;
  __syncthreads();
  if (ty == 0 && ok) {
    float a = 0.0f, bsum = 0.0f;
#pragma unroll
    for (int j = 0; j < kFoldRowGroups; ++j) {
      a += s_dg[j][tx];
      bsum += s_db[j][tx];
    }
    G.dg[s][c] = accumulate ? G.dg[s][c] + a : a;
    G.db[s][c] = accumulate ? G.db[s][c] + bsum : bsum;
  }
}

int fold_qkv_run(const float* const* w, const float* const* g, const float* const* b, float* wg, float* bg, int C,
                 cudaStream_t stream) {
  TVAE_REQUIRE(C % 4 == 0, "fold_qkv: C=%d must be a multiple of 4", C);
  FoldPtrs P;
  for (int i = 0; i < 3; ++i) {
    P.w[i] = w[i];
    P.g[i] = g[i];
    P.b[i] = b[i];
  }
  fold_qkv_fwd_kernel<<<(3 * C + 7) / 8, 256, 0, stream>>>(P, wg, bg, C);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fold_qkv_bwd_run(const float* const* w, const float* const* g, const float* const* b, const float* dwg, const float* dbg,
                     float* const* dw, float* const* dg, float* const* db, int C, int accumulate, cudaStream_t stream) {
  FoldPtrs P;
  FoldGradPtrs G;
  for (int i = 0; i < 3; ++i) {
    P.w[i] = w[i];
    P.g[i] = g[i];
    P.b[i] = b[i];
    G.dw[i] = dw[i];
    G.dg[i] = dg[i];
    G.db[i] = db[i];
  }
  fold_qkv_bwd_kernel<<<dim3((C + 31) / 32, 3), 32 * kFoldRowGroups, 0, stream>>>(P, G, dwg, dbg, C, accumulate);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Upsample conv1 (upsample.py:94-95: nearest 2x then 3x3, computed as four phase-specific 2x2 convolutions on the
// low-resolution input): slab ((py*2 + px)*2 + a)*2 + b of the packed weight [O][16*I] is the sum of the 3x3 taps
// (dy, dx) with dy in rows(py, a), dx in rows(px, b), rows(0, .) = {0}, {1, 2}; rows(1, .) = {0, 1}, {2}.
// Forward: w[O][I][3][3] -> packed; backward: dw[o][i][dy][dx] = sum of the dpacked slabs that contain the tap.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int up_row_mask(int p, int a) {     // bit dy set: tap row dy belongs to rows(p, a)
  return p == 0 ? (a == 0 ? 0b001 : 0b110) : (a == 0 ? 0b011 : 0b100);
}

__global__ void __launch_bounds__(256) upconv1_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int I) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (o, i), i fastest
  if (idx >= (long long)O * I) return;
  const int o = (int)(idx / I), i = (int)(idx - (long long)o * I);
  float t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) t[k] = __ldg(w + idx * 9 + k);
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int py = s >> 3, px = (s >> 2) & 1, a = (s >> 1) & 1, b = s & 1;
    const int my = up_row_mask(py, a), mx = up_row_mask(px, b);
    float acc = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        if (((my >> dy) & 1) && ((mx >> dx) & 1)) acc += t[dy * 3 + dx];
    out[(size_t)o * 16 * I + (size_t)s * I + i] = acc;
  }
}

__global__ void __launch_bounds__(256) upconv1_unpack_kernel(const float* __restrict__ dout, float* __restrict__ dw, int O, int I) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)O * I) return;
  const int o = (int)(idx / I), i = (int)(idx - (long long)o * I);
  float t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) t[k] = 0.0f;
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int py = s >> 3, px = (s >> 2) & 1, a = (s >> 1) & 1, b = s & 1;
    const int my = up_row_mask(py, a), mx = up_row_mask(px, b);
    const float d = __ldg(dout + (size_t)o * 16 * I + (size_t)s * I + i);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        if (((my >> dy) & 1) && ((mx >> dx) & 1)) t[dy * 3 + dx] += d;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) dw[idx * 9 + k] = t[k];
}

int upconv1_pack_run(const float* w, float* out, int O, int I, int backward, cudaStream_t stream) {
  const long long n = (long long)O * I;
  const int blocks = (int)((n + 255) / 256);
  if (backward) upconv1_unpack_kernel<<<blocks, 256, 0, stream>>>(w, out, O, I);
  else upconv1_pack_kernel<<<blocks, 256, 0, stream>>>(w, out, O, I);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
