// Weight gradient of the fused multi-tap implicit GEMM (sm_100a, tcgen05 + TMA + TMEM).
//
//   dW[n, wk_off_tap + k] += sum_pixels dZ[pixel, n] * A_tap[pixel, k]
//
// Same pixel views / tap tables as the forward kernel (mtgemm.cu), but the reduction runs over PIXELS: both
// operands are TMA-loaded as [128 pixels x 64 channels] 128B-swizzled tiles and fed to the tensor core as
// MN-major operands (channels contiguous, pixels = MMA K dimension).  One CTA owns one
// (phase, tap, 128-wide n tile, KT-wide k tile) output block for a strided subset of the pixel tiles ("split-K over
// pixels"), accumulates it in TMEM over its whole loop and adds it to the fp32 dW matrix with vectorised
// red.global.add at the end.  Replaces the weight-gradient half of autograd for every nn.Conv2d / nn.Linear the
// forward kernel replaces (reference call sites: include/transvae_sm100.h, tvae_mtgemm).
#include <cstdlib>

#include "../../include/transvae_sm100.h"
#include "cluster2.cuh"
#include "common.cuh"
#include "tmap.cuh"

namespace tvae {

struct WgTap {
  int32_t c_off;
  int32_t wk_off;
  int16_t kblocks;
  int8_t map, dw, p, dh;
  int8_t ph;      // phase this tap belongs to
  int8_t pad_;
};
static_assert(sizeof(WgTap) == 16, "WgTap must be 16 bytes");

struct WgParams {
  int tiles_w, tiles_h, tiles_b;
  int tw, th, nb;
  int ntaps;                 // flattened over phases
  int n_tiles;               // ceil(n_total / 128)
  int n_total, k_total;
  int splits;
  int out_p[TVAE_MAX_PHASES];
  int out_c_off[TVAE_MAX_PHASES];
  WgTap taps[TVAE_MAX_PHASES * TVAE_MAX_TAPS];
  float* dw;
  float* db;                 // optional bias gradient [phases][n_total] (column sums of dZ), accumulated into
  int first_tap[TVAE_MAX_PHASES];   // index of the first tap of each phase (its k-tile-0 CTAs also produce db)
  int num_phases;
  int items_per_phase[TVAE_MAX_PHASES];   // transposed kernel: ceil((chunks + bias slot) / 2)
  // Deterministic split reduction: with more than one pixel split per block, split s writes its partial block with plain
  // stores into slice s of a workspace (dw / db then point INTO the workspace, slices ws_dw_stride / ws_db_stride floats
  // apart) and wgrad_reduce_kernel adds the slices to the gradient in a fixed order.  0 = one split: the single writer of
  // every element accumulates straight into the gradient (atomicAdd, but with one writer per launch the order is fixed).
  long long ws_dw_stride, ws_db_stride;
  // phases that share weight slabs (Upsample conv2: the 3x3 taps of all four output phases) would overwrite each other's
  // partials: such plans get one slice per (split, phase) -- slice = split * ws_phases + phase -- of a zero-filled workspace
  int ws_phases;
};

// accumulate (single writer) or store (workspace slice)
__device__ __forceinline__ void wg_out(float* p, float v, bool plain) {
  if (plain) *p = v;
  else atomicAdd(p, v);
}
__device__ __forceinline__ void wg_out4(float* p, float4 v, bool plain) {
  if (plain) *reinterpret_cast<float4*>(p) = v;
  else atomicAdd(reinterpret_cast<float4*>(p), v);
}

// gradient += sum over the slices, in slice order (bit-reproducible).  A slice is [weight gradient (n_dw) | bias gradient
// (n_db)], both multiples of 4 floats; one launch covers both destinations.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(float* __restrict__ dw, float* __restrict__ db,
                                                           const float* __restrict__ ws, long long n_dw, long long n_db,
                                                           int splits) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long n = n_dw + n_db, stride = n;
  if (i >= n) return;
  float* g = i < n_dw ? dw + i : db + (i - n_dw);
  float4 a = *reinterpret_cast<const float4*>(g);
  // eight independent loads in flight, added in slice order
  int s = 0;
  for (; s + 8 <= splits; s += 8) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(ws + (size_t)(s + k) * stride + i));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w;
    }
  }
  for (; s < splits; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(ws + (size_t)s * stride + i));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4*>(g) = a;
}

constexpr int kWgTile = 128 * 64 * 2;  // one [128 pixels x 64 channels] bf16 box

template <int KT>
struct WgCfg {
  static constexpr int kStageBytes = (2 + KT / 64) * kWgTile;
  static constexpr int kStages = (220 * 1024) / kStageBytes > 4 ? 4 : (220 * 1024) / kStageBytes;
  // KT accumulator columns + 16 for the bias gradient (dZ^T x ones), rounded up to a power of two
  static constexpr int kTmemCols = KT + 16 <= 128 ? 128 : (KT + 16 <= 256 ? 256 : 512);
  static constexpr int kOnesBytes = 2048;   // a [16 pixels x 64 channels] block of bf16 1.0 (layout-free: all equal)
  static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 128;
};

template <int KT>
__global__ void __launch_bounds__(256, 1)
mtwgrad_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ WgParams P) {
#ifdef TVAE_DEVICE_OK
  using Cfg = WgCfg<KT>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ones = smem + STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + Cfg::kOnesBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmDZ);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (P.db != nullptr) {
    for (int i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3f803f80u;
    fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  // ---- decode the work item of this CTA: (tap, k tile, n tile) x pixel split
  const int split = blockIdx.x % P.splits;
  int item = blockIdx.x / P.splits;
  int ti = 0, kt = 0, nt = 0;
  for (; ti < P.ntaps; ++ti) {
    const int cnt = (P.taps[ti].kblocks * 64 / KT) * P.n_tiles;
    if (item < cnt) {
      kt = item / P.n_tiles;
      nt = item % P.n_tiles;
      break;
    }
    item -= cnt;
  }
  const WgTap tap = P.taps[ti];
  // Bias gradient db[n] = sum_pixels dZ[pixel, n]: the CTAs of the first tap / first k tile of each phase run one extra
  // N = 16 MMA per 16 pixels against a block of ones and keep dZ^T x 1 in 16 more TMEM columns -- the column sums cost
  // 16 / KT more tensor work on 1 / (taps * k tiles) of the CTAs instead of a separate pass over dZ.
  const bool do_bias = P.db != nullptr && kt == 0 && ti == P.first_tap[tap.ph];
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int my_tiles = (m_tiles - split + P.splits - 1) / P.splits;   // tiles split, split+splits, ...

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mapA = tap.map ? &tmA1 : &tmA0;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m_t = split + i * P.splits;
        const int w0 = (m_t % P.tiles_w) * P.tw;
        const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
        const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
        uint8_t* s = smem + stage * Cfg::kStageBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_5d(s + j * kWgTile, &tmDZ, &full[stage], P.out_c_off[tap.ph] + nt * 128 + j * 64, w0, P.out_p[tap.ph], h0,
                      b0);
#pragma unroll
        for (int j = 0; j < KT / 64; ++j)
          tma_load_5d(s + (2 + j) * kWgTile, mapA, &full[stage], tap.c_off + kt * KT + j * 64, w0 + tap.dw, tap.p,
                      h0 + tap.dh, b0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, elected lane issues (common.cuh: warp-uniform issue)
      constexpr uint32_t idesc = umma_idesc_bf16(128, KT, 1, 1);
      constexpr uint32_t idesc_ones = umma_idesc_bf16(128, 16, 1, 1);
      const uint64_t ones_desc = umma_desc_mnmajor_sw128(smem_u32(s_ones), kWgTile, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t dz_base = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t a_base = dz_base + 2 * kWgTile;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 pixels per MMA
          umma_f16_elect(tmem_base, umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024),
                   umma_desc_mnmajor_sw128(a_base + k * 2048, kWgTile, 1024), idesc, (i | k) != 0);
        if (do_bias) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_elect(tmem_base + KT, umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024), ones_desc, idesc_ones,
                     (i | k) != 0);
        }
        umma_commit_elect(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_elect(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int n = nt * 128 + q * 32 + lane;
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const bool plain = P.ws_dw_stride != 0;
      const size_t slice = (size_t)split * P.ws_phases + (P.ws_phases > 1 ? tap.ph : 0);
      float* dst = P.dw + slice * P.ws_dw_stride + (size_t)n * P.k_total + tap.wk_off + kt * KT;
#pragma unroll 1
      for (int c = 0; c < KT / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        if (n < P.n_total) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 f = make_float4(__uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]),
                                   __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
            wg_out4(dst + c * 32 + g * 4, f, plain);
          }
        }
      }
      if (do_bias) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + KT, v);
        tmem_ld_wait();
        if (n < P.n_total) wg_out(P.db + slice * P.ws_db_stride + (size_t)tap.ph * P.n_total + n, __uint_as_float(v[0]), plain);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
#endif
}

// -------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for layers whose output-channel count is a multiple of 256 and whose taps cover
// multiples of KT2 (256 / 128) input channels: every 768- / 1536-wide linear and 3x3 convolution of the Transformer
// stages.  The 1-CTA kernel above pulls 2 dZ boxes + KT / 64 A boxes from L2 per 128-pixel step for a 128 x KT block
// (87 flop per byte at KT = 256) and is bound by the L2 -> SM path (ncu: 13.4 TB/s, 7.5x the DRAM traffic).  Here two
// CTAs of a cluster own a 256 (n) x KT2 (k) block: CTA r loads the dZ boxes of ITS 128 n channels (its half of the MMA
// M dimension) and only HALF of the A boxes (KT2 / 2 k channels: its half of the MMA N dimension); the leader's
// tcgen05.mma.cta_group::2 (M = 256, N = KT2) reads both halves across the pair -- 131 flop per byte at KT2 = 256, the
// ratio of the forward pair kernel.  Protocol as mtgemm2.cu: both producers complete on the LEADER's full[s], the
// leader's commit is multicast to empty[s] of both CTAs, each CTA drains its own accumulator rows.
// -------------------------------------------------------------------------------------------------
template <int KT2>
struct Wg2Cfg {
  static constexpr int kABoxes = KT2 / 128;                         // per CTA
  static constexpr int kStageBytes = (2 + kABoxes) * kWgTile;       // per CTA
  static constexpr int kStages = (200 * 1024) / kStageBytes > 4 ? 4 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = KT2 + 16 <= 128 ? 128 : (KT2 + 16 <= 256 ? 256 : 512);
  static constexpr int kOnesBytes = 2048;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 256;
};

template <int KT2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
mtwgrad2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ WgParams P) {
#ifdef TVAE_DEVICE_OK
  using Cfg = Wg2Cfg<KT2>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ones = smem + STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + Cfg::kOnesBytes);
  uint64_t* full = bars;                   // [STAGES] (the leader's are used)
  uint64_t* empty = bars + STAGES;         // [STAGES] (each CTA's own)
  uint64_t* acc_full = bars + 2 * STAGES;  // each CTA's own
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmDZ);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (P.db != nullptr) {
    for (int i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3f803f80u;
    fence_proxy_async_smem();
  }
  cluster_sync_all();   // both CTAs' barriers (and ones blocks) exist before any cross-CTA completion / MMA
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  // ---- work item of this CTA pair: (tap, k tile of KT2, n tile of 256) x pixel split
  const int cluster_id = blockIdx.x >> 1;
  const int split = cluster_id % P.splits;
  int item = cluster_id / P.splits;
  const int n_tiles2 = P.n_total / 256;
  int ti = 0, kt = 0, nt = 0;
  for (; ti < P.ntaps; ++ti) {
    const int cnt = (P.taps[ti].kblocks * 64 / KT2) * n_tiles2;
    if (item < cnt) {
      kt = item / n_tiles2;
      nt = item % n_tiles2;
      break;
    }
    item -= cnt;
  }
  const WgTap tap = P.taps[ti];
  const bool do_bias = P.db != nullptr && kt == 0 && ti == P.first_tap[tap.ph];
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int my_tiles = (m_tiles - split + P.splits - 1) / P.splits;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mapA = tap.map ? &tmA1 : &tmA0;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m_t = split + i * P.splits;
        const int w0 = (m_t % P.tiles_w) * P.tw;
        const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
        const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
        mbar_wait(&empty[stage], phase ^ 1);
        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
        uint8_t* s = smem + stage * Cfg::kStageBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j)         // this CTA's half of the M dimension: 128 of the 256 n channels
          tma2_load_5d(s + j * kWgTile, &tmDZ, &full[stage], P.out_c_off[tap.ph] + nt * 256 + (int)rank * 128 + j * 64, w0,
                       P.out_p[tap.ph], h0, b0);
#pragma unroll
        for (int j = 0; j < Cfg::kABoxes; ++j)   // this CTA's half of the N dimension: KT2 / 2 of the k channels
          tma2_load_5d(s + (2 + j) * kWgTile, mapA, &full[stage], tap.c_off + kt * KT2 + (int)rank * (KT2 / 2) + j * 64,
                       w0 + tap.dw, tap.p, h0 + tap.dh, b0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {   // whole warp, elected lane issues (common.cuh: warp-uniform issue)
      constexpr uint32_t idesc = umma_idesc_bf16(256, KT2, 1, 1);
      constexpr uint32_t idesc_ones = umma_idesc_bf16(256, 16, 1, 1);
      const uint64_t ones_desc = umma_desc_mnmajor_sw128(smem_u32(s_ones), kWgTile, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t dz_base = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t a_base = dz_base + 2 * kWgTile;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 pixels per MMA
          umma2_f16_elect(tmem_base, umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024),
                    umma_desc_mnmajor_sw128(a_base + k * 2048, kWgTile, 1024), idesc, (i | k) != 0);
        if (do_bias) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma2_f16_elect(tmem_base + KT2, umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024), ones_desc, idesc_ones,
                      (i | k) != 0);
        }
        umma2_commit_mc_elect(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma2_commit_mc_elect(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int n = nt * 256 + (int)rank * 128 + q * 32 + lane;     // this CTA's accumulator rows
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const bool plain = P.ws_dw_stride != 0;
      const size_t slice = (size_t)split * P.ws_phases + (P.ws_phases > 1 ? tap.ph : 0);
      float* dst = P.dw + slice * P.ws_dw_stride + (size_t)n * P.k_total + tap.wk_off + kt * KT2;
#pragma unroll 1
      for (int c = 0; c < KT2 / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 f = make_float4(__uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]),
                                 __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
          wg_out4(dst + c * 32 + g * 4, f, plain);
        }
      }
      if (do_bias) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + KT2, v);
        tmem_ld_wait();
        wg_out(P.db + slice * P.ws_db_stride + (size_t)tap.ph * P.n_total + n, __uint_as_float(v[0]), plain);
      }
    }
  }
  // the peer's smem / barriers / TMEM are touched by the leader's MMAs and multicast commits until the very end
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                 : "memory");
  }
#endif
}

template <int KT2>
static int launch_wg2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& dz, const WgParams& P, int grid,
                      cudaStream_t stream) {
  using Cfg = Wg2Cfg<KT2>;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(mtwgrad2_kernel<KT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  mtwgrad2_kernel<KT2><<<grid, 256, Cfg::kSmemBytes, stream>>>(a0, a1, dz, P);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Transposed variant for layers whose output-channel count is not a multiple of 128 (N = 64, 192: the ResBlock /
// Downsample / Upsample convolutions at 192 channels and the 64-wide heads).
//
// The kernel above puts the n (output) channels on the MMA M dimension in tiles of 128, so N = 192 runs a second
// tile that is half empty (25 % of the tensor work wasted, ResBlock wgrad ~720 TFLOP/s).  Here the roles are swapped:
//   D[M = 2 x 64 k-channel chunks, N = all n channels] += [A_chunk0 | A_chunk1]^T (pixels x 128) . dZ (pixels x N)
// and the two 64-row halves of an M tile are independent [128 pixels x 64 channels] smem tiles -- possibly of
// DIFFERENT taps -- that the MN-major UMMA descriptor stitches together through its leading-dimension byte offset.
// A phase with c chunks (taps x C_in / 64) needs ceil(c / 2) full-width MMAs instead of 2 x c half-empty ones.  The
// bias gradient (dZ^T . 1) is one more "chunk" whose tile is a constant block of ones.
// -------------------------------------------------------------------------------------------------
template <int NT>
struct WgtCfg {
  static constexpr int kStageBytes = (2 + NT / 64) * kWgTile;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 4 ? 4 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = NT <= 64 ? 64 : (NT <= 128 ? 128 : 256);
  static constexpr int kSmemBytes = kStages * kStageBytes + kWgTile /*ones*/ + 1024 + 128;
};

struct WgSlot {
  int tap;      // index into P.taps, -1: ones (bias gradient), -2: empty
  int chunk;    // 64-channel chunk within the tap
};

// slot s (0-based) of phase ph: chunks of its taps in order, then (optionally) the ones slot, then empty
__device__ __forceinline__ WgSlot wgt_slot(const WgParams& P, int ph, int s) {
  int t = P.first_tap[ph];
  const int t_end = (ph + 1 < TVAE_MAX_PHASES && P.first_tap[ph + 1] > 0) ? P.first_tap[ph + 1] : P.ntaps;
  for (; t < t_end; ++t) {
    if (s < P.taps[t].kblocks) return WgSlot{t, s};
    s -= P.taps[t].kblocks;
  }
  if (s == 0 && P.db != nullptr) return WgSlot{-1, 0};
  return WgSlot{-2, 0};
}

template <int NT>
__global__ void __launch_bounds__(256, 1)
mtwgrad_t_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ WgParams P) {
#ifdef TVAE_DEVICE_OK
  using Cfg = WgtCfg<NT>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ones = smem + STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + kWgTile);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmDZ);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (P.db != nullptr) {
    for (int i = threadIdx.x; i < kWgTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3f803f80u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  // ---- work item: (phase, pair of chunk slots) x pixel split
  const int split = blockIdx.x % P.splits;
  int item = blockIdx.x / P.splits;
  int ph = 0;
  for (; ph < P.num_phases; ++ph) {
    if (item < P.items_per_phase[ph]) break;
    item -= P.items_per_phase[ph];
  }
  const WgSlot s0 = wgt_slot(P, ph, 2 * item), s1 = wgt_slot(P, ph, 2 * item + 1);
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int my_tiles = (m_tiles - split + P.splits - 1) / P.splits;
  constexpr int kDzChunks = NT / 64;

  if (warp == 0) {
    if (lane == 0) {
      const int loaded = kDzChunks + (s0.tap >= 0 ? 1 : 0) + (s1.tap >= 0 ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m_t = split + i * P.splits;
        const int w0 = (m_t % P.tiles_w) * P.tw;
        const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
        const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], loaded * kWgTile);
        uint8_t* s = smem + stage * Cfg::kStageBytes;
#pragma unroll
        for (int j = 0; j < kDzChunks; ++j)
          tma_load_5d(s + j * kWgTile, &tmDZ, &full[stage], P.out_c_off[ph] + j * 64, w0, P.out_p[ph], h0, b0);
        if (s0.tap >= 0) {     // (an item can consist of the ones slot alone)
          const WgTap tp = P.taps[s0.tap];
          tma_load_5d(s + kDzChunks * kWgTile, tp.map ? &tmA1 : &tmA0, &full[stage], tp.c_off + s0.chunk * 64, w0 + tp.dw, tp.p,
                      h0 + tp.dh, b0);
        }
        if (s1.tap >= 0) {
          const WgTap tp = P.taps[s1.tap];
          tma_load_5d(s + (kDzChunks + 1) * kWgTile, tp.map ? &tmA1 : &tmA0, &full[stage], tp.c_off + s1.chunk * 64,
                      w0 + tp.dw, tp.p, h0 + tp.dh, b0);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, elected lane issues (common.cuh: warp-uniform issue)
      constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t dz_base = smem_u32(smem + stage * Cfg::kStageBytes);
        // first 64-row half of the A operand: a tap chunk, or the block of ones when the bias slot stands alone;
        // second half: the next tile (a tap chunk), the block of ones, or -- unpaired -- the first half again (ignored)
        const uint32_t a_base = s0.tap >= 0 ? dz_base + kDzChunks * kWgTile : smem_u32(s_ones);
        const uint32_t lbo = s1.tap >= 0 ? kWgTile : (s1.tap == -1 ? smem_u32(s_ones) - a_base : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 pixels per MMA
          umma_f16_elect(tmem_base, umma_desc_mnmajor_sw128(a_base + k * 2048, lbo, 1024),
                   umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024), idesc, (i | k) != 0);
        umma_commit_elect(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_elect(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // M row of the accumulator = slot (row / 64), channel (row % 64)
    const WgSlot sl = row < 64 ? s0 : s1;
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      float* dst = nullptr;
      size_t stride = 0;
      bool ok = false;
      if (sl.tap >= 0) {                           // dW[n][wk_off + chunk * 64 + kk], consecutive lanes -> consecutive kk
        dst = P.dw + ((size_t)split * P.ws_phases + (P.ws_phases > 1 ? ph : 0)) * P.ws_dw_stride + P.taps[sl.tap].wk_off + sl.chunk * 64 + (row & 63);
        stride = (size_t)P.k_total;
        ok = true;
      } else if (sl.tap == -1 && (row & 63) == 0) {   // all 64 rows of the ones slot hold the same column sums
        dst = P.db + ((size_t)split * P.ws_phases + (P.ws_phases > 1 ? ph : 0)) * P.ws_db_stride + (size_t)ph * P.n_total;
        stride = 1;
        ok = true;
      }
#pragma unroll 1
      for (int c = 0; c < NT / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) wg_out(dst + (size_t)(c * 32 + j) * stride, __uint_as_float(v[j]), P.ws_dw_stride != 0);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
#endif
}

// -------------------------------------------------------------------------------------------------
// Halo variant of the transposed kernel for plain 3x3 stride-1 convolutions on maps at least 128 pixels wide (the
// N = 192 ResBlock / Downsample / Upsample convolutions at 256^2 and 128^2, and the 64-wide output convolution): a pixel
// tile is 128 pixels of ONE image row, so the three dx taps of a kernel row read A tiles that overlap in 127 of 128 pixels.
// The transposed kernel above loads each (tap, 64-channel chunk) as its own [128 x 64] box -- NT / 64 dZ boxes + 2 A boxes
// per 128 x 128 x NT step (77 flop per byte at NT = 192), and every tap's CTA re-loads the same dZ tile (ncu: 12.1 GB of
// L2 -> SM traffic for 1.6 GB of DRAM reads).  Here a "slot" is (kernel row dy, 64-channel chunk, dx); a CTA owns FOUR
// consecutive slots = two full-width M = 128 MMAs per 16 pixels into two accumulators, and loads per step the dZ boxes
// ONCE plus at most two 130-pixel halo tiles [w0 - 1, w0 + 129) of the (dy, chunk) groups its slots belong to -- 155
// flop per byte.  The two 64-row halves of an MMA's A operand are (halo tile, pixel-row offset dx) pairs: in the
// MN-major SWIZZLE_128B layout a pixel row is 128 bytes, so a tap is a start address dx rows into the tile and the
// second half is reached through the descriptor's leading-dimension byte offset -- 128 B for the next tap of the same
// tile, or the distance to the other tile / the block of ones (bias gradient).  The swizzle is a function of the absolute
// shared-memory address (descriptor base offset 0): tools/experiments/umma_row_offset_mn.cu, umma_lbo_offset_mn.cu.
// -------------------------------------------------------------------------------------------------
constexpr int kWgHaloTx = 130 * 64 * 2;                         // bytes one halo load delivers
constexpr int kWgHaloBytes = (kWgHaloTx + 1023) / 1024 * 1024;  // its shared-memory slot (17408)

template <int NT>
struct WghCfg {
  static constexpr int kDzBytes = (NT / 64) * kWgTile;
  static constexpr int kStageBytes = kDzBytes + 2 * kWgHaloBytes;
  static constexpr int kStages = (200 * 1024 - kWgTile) / kStageBytes > 4 ? 4 : (200 * 1024 - kWgTile) / kStageBytes;
  static constexpr int kTmemCols = 2 * NT <= 128 ? 128 : (2 * NT <= 256 ? 256 : 512);
  static constexpr int kOnesBytes = kWgTile + 1024;   // ones for 128 pixels (+ slack: the peer half may start dx rows in)
  static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 128;
};

struct WghSlot {
  int group;   // (dy * kb + chunk), or -1: ones (bias gradient), -2: empty
  int dx;
};

template <int NT>
__global__ void __launch_bounds__(256, 1)
mtwgrad_h_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmDZ,
                 const __grid_constant__ WgParams P) {
#ifdef TVAE_DEVICE_OK
  using Cfg = WghCfg<NT>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int kDzChunks = NT / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ones = smem + STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + Cfg::kOnesBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmDZ);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (P.db != nullptr) {
    for (int i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3f803f80u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  // ---- work item: four consecutive slots x pixel split.  Slot s < 9 * kb: group s / 3 = (dy, chunk), dx = s % 3;
  // slot 9 * kb: the block of ones (when the bias gradient is wanted); beyond: empty.
  const int kb = P.taps[0].kblocks;
  const int n_tap_slots = 9 * kb;
  const int split = blockIdx.x % P.splits;
  const int item = blockIdx.x / P.splits;
  WghSlot sl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = 4 * item + j;
    if (s < n_tap_slots) sl[j] = WghSlot{s / 3, s % 3};
    else if (s == n_tap_slots && P.db != nullptr) sl[j] = WghSlot{-1, 0};
    else sl[j] = WghSlot{-2, 0};
  }
  // the (at most two) halo tiles of this item: groups g0 <= g1
  const int g0 = sl[0].group;                         // slot 4 * item is always a tap slot or the ones slot
  int g1 = g0;
#pragma unroll
  for (int j = 1; j < 4; ++j)
    if (sl[j].group >= 0 && sl[j].group != g0) g1 = sl[j].group;
  const int n_tiles_a = g0 < 0 ? 0 : (g1 != g0 ? 2 : 1);
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int my_tiles = (m_tiles - split + P.splits - 1) / P.splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m_t = split + i * P.splits;
        const int w0 = (m_t % P.tiles_w) * P.tw;
        const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
        const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], Cfg::kDzBytes + n_tiles_a * kWgHaloTx);
        uint8_t* s = smem + stage * Cfg::kStageBytes;
#pragma unroll
        for (int j = 0; j < kDzChunks; ++j)
          tma_load_5d(s + j * kWgTile, &tmDZ, &full[stage], P.out_c_off[0] + j * 64, w0, P.out_p[0], h0, b0);
        for (int a = 0; a < n_tiles_a; ++a) {
          const int g = a == 0 ? g0 : g1;
          const int dy = g / kb, chunk = g % kb;
          tma_load_5d(s + Cfg::kDzBytes + a * kWgHaloBytes, &tmAh, &full[stage], P.taps[dy * 3].c_off + chunk * 64, w0 - 1, 0,
                      h0 + P.taps[dy * 3].dh, b0);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, elected lane issues (common.cuh: warp-uniform issue)
      constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 1, 1);
      constexpr uint64_t bo_mask = ~(uint64_t(7) << 49);     // descriptor base offset 0: swizzle on absolute addresses
      const bool second = sl[2].group != -2;                 // slots 2 / 3 exist
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t dz_base = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t t_base = dz_base + Cfg::kDzBytes;
        uint32_t addr[4] = {dz_base, dz_base, dz_base, dz_base};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (sl[j].group >= 0) addr[j] = t_base + (sl[j].group == g0 ? 0 : kWgHaloBytes) + sl[j].dx * 128;
          else if (sl[j].group == -1) addr[j] = smem_u32(s_ones);
          else if (j & 1) addr[j] = addr[j - 1];              // empty second half: repeat the first (rows ignored)
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 pixels per MMA
          umma_f16_elect(tmem_base, umma_desc_mnmajor_sw128(addr[0] + k * 2048, addr[1] - addr[0], 1024) & bo_mask,
                   umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024), idesc, (i | k) != 0);
        if (second) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_elect(tmem_base + NT, umma_desc_mnmajor_sw128(addr[2] + k * 2048, addr[3] - addr[2], 1024) & bo_mask,
                     umma_desc_mnmajor_sw128(dz_base + k * 2048, kWgTile, 1024), idesc, (i | k) != 0);
        }
        umma_commit_elect(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_elect(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // accumulator row: slot (row / 64) of the MMA, channel row % 64
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        const WghSlot s = sl[2 * a + (row >> 6)];
        float* dst = nullptr;
        size_t stride = 0;
        if (s.group >= 0) {                        // dW[n][wk_off(dy, dx) + chunk * 64 + c]
          const int dy = s.group / kb, chunk = s.group % kb;
          dst = P.dw + (size_t)split * P.ws_dw_stride + P.taps[dy * 3 + s.dx].wk_off + chunk * 64 + (row & 63);
          stride = (size_t)P.k_total;
        } else if (s.group == -1 && (row & 63) == 0) {   // all 64 rows of the ones slot hold the same column sums
          dst = P.db + (size_t)split * P.ws_db_stride;
          stride = 1;
        }
        if (sl[2 * a].group == -2) continue;       // this accumulator was never written (warp-uniform)
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * NT + c * 32, v);
          tmem_ld_wait();
          if (dst != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) wg_out(dst + (size_t)(c * 32 + j) * stride, __uint_as_float(v[j]), P.ws_dw_stride != 0);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
#endif
}

template <int NT>
static int launch_wgh(const CUtensorMap& ah, const CUtensorMap& dz, const WgParams& P, int grid, cudaStream_t stream) {
  using Cfg = WghCfg<NT>;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(mtwgrad_h_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  mtwgrad_h_kernel<NT><<<grid, 256, Cfg::kSmemBytes, stream>>>(ah, dz, P);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int NT>
static int launch_wgt(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& dz, const WgParams& P, int grid,
                      cudaStream_t stream) {
  using Cfg = WgtCfg<NT>;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(mtwgrad_t_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  mtwgrad_t_kernel<NT><<<grid, 256, Cfg::kSmemBytes, stream>>>(a0, a1, dz, P);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int pow2_ceil_w(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int KT>
static int launch_wg(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& dz, const WgParams& P, int grid,
                     cudaStream_t stream) {
  using Cfg = WgCfg<KT>;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(mtwgrad_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  mtwgrad_kernel<KT><<<grid, 256, Cfg::kSmemBytes, stream>>>(a0, a1, dz, P);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// d->out is the dZ view (gradient w.r.t. the forward kernel's pre-activation output); dw: fp32 [n_total, k_total],
// accumulated into (the caller zeroes it); db: optional fp32 [num_phases, n_total] bias gradient, accumulated into.
int mtwgrad_run(const tvae_mtgemm_desc* d, float* dw, float* db, cudaStream_t stream) {
  TVAE_REQUIRE(d != nullptr && dw != nullptr, "wgrad: null argument");
  TVAE_REQUIRE(d->a0.ptr != nullptr && d->out.ptr != nullptr, "wgrad: missing operand");
  TVAE_REQUIRE(d->k_total % 64 == 0, "wgrad: k_total %d must be a multiple of 64", d->k_total);
  TVAE_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0, "wgrad: dw must be 16-byte aligned");
  WgParams P;
  memset(&P, 0, sizeof(P));
  const tvae_view& gv = d->out;
  const int vW = gv.split ? gv.W / 2 : gv.W, vH = gv.split ? gv.H / 2 : gv.H, vB = gv.B;
  P.tw = pow2_ceil_w(vW) < 128 ? pow2_ceil_w(vW) : 128;
  P.th = pow2_ceil_w(vH) < 128 / P.tw ? pow2_ceil_w(vH) : 128 / P.tw;
  P.nb = 128 / (P.tw * P.th);
  P.tiles_w = (vW + P.tw - 1) / P.tw;
  P.tiles_h = (vH + P.th - 1) / P.th;
  P.tiles_b = (vB + P.nb - 1) / P.nb;
  P.n_total = d->n_total;
  P.k_total = d->k_total;
  P.n_tiles = (d->n_total + 127) / 128;
  P.dw = dw;
  P.db = db;
  int kt = 256;
  int nt = 0;
  for (int ph = 0; ph < d->num_phases; ++ph) {
    P.out_p[ph] = d->out_p[ph];
    P.out_c_off[ph] = d->out_c_off[ph];
    P.first_tap[ph] = nt;
    for (int t = 0; t < d->ntaps[ph]; ++t) {
      const tvae_tap& s = d->taps[ph][t];
      const tvae_view& av = s.map ? d->a1 : d->a0;
      TVAE_REQUIRE(av.ptr != nullptr, "wgrad: tap uses absent view %d", s.map);
      TVAE_REQUIRE(s.kblocks >= 1 && s.wk_off >= 0 && s.wk_off + s.kblocks * 64 <= d->k_total, "wgrad: bad tap");
      WgTap& o = P.taps[nt++];
      o.c_off = s.c_off; o.wk_off = s.wk_off; o.kblocks = (int16_t)s.kblocks;
      o.map = (int8_t)s.map; o.dw = (int8_t)s.dw; o.p = (int8_t)s.p; o.dh = (int8_t)s.dh; o.ph = (int8_t)ph;
      const int c = s.kblocks * 64;
      while (c % kt) kt -= 64;
    }
  }
  P.ntaps = nt;
  P.num_phases = d->num_phases;
  for (int ph = d->num_phases; ph < TVAE_MAX_PHASES; ++ph) P.first_tap[ph] = 0;
  // N = 64 / 192: transposed kernel (no half-empty 128-row tiles); everything else: n channels on the M dimension
  static const bool allow_t = !(getenv("TVAE_WGRAD_T") && atoi(getenv("TVAE_WGRAD_T")) == 0);
  const bool transposed = allow_t && (d->n_total == 64 || d->n_total == 192);
  // plain 3x3 stride-1 convolution on a map at least 128 pixels wide: halo variant of the transposed kernel (four
  // (kernel row, chunk, dx) slots per CTA on one dZ tile and <= 2 halo tiles); TVAE_WGRAD_HALO=0: A/B switch
  static const bool allow_h = !(getenv("TVAE_WGRAD_HALO") && atoi(getenv("TVAE_WGRAD_HALO")) == 0);
  bool halo = allow_h && transposed && P.tw == 128 && d->num_phases == 1 && d->ntaps[0] == 9 && !d->a0.split && !d->out.split &&
              d->out_c_off[0] == 0;
  for (int t = 0; halo && t < 9; ++t) {
    const tvae_tap& tp = d->taps[0][t];
    halo = tp.map == 0 && tp.p == 0 && tp.c_off == 0 && tp.dw == t % 3 - 1 && tp.dh == d->taps[0][t - t % 3].dh &&
           tp.kblocks == d->taps[0][0].kblocks && tp.kblocks * 64 == d->a0.C;
  }
  // n_total a multiple of 256 and every tap a multiple of 256 input channels: CTA-pair kernel; TVAE_WGRAD_PAIR=0: A/B switch
  static const bool allow_p = !(getenv("TVAE_WGRAD_PAIR") && atoi(getenv("TVAE_WGRAD_PAIR")) == 0);
  bool pair = allow_p && !transposed && d->n_total % 256 == 0;
  for (int i = 0; pair && i < nt; ++i) pair = (P.taps[i].kblocks * 64) % 256 == 0;
  long long items = 0;
  if (halo) {
    items = (9LL * d->taps[0][0].kblocks + (db != nullptr ? 1 : 0) + 3) / 4;
  } else if (transposed) {
    for (int ph = 0; ph < d->num_phases; ++ph) {
      int chunks = db != nullptr ? 1 : 0;
      const int t_end = ph + 1 < d->num_phases ? P.first_tap[ph + 1] : nt;
      for (int t = P.first_tap[ph]; t < t_end; ++t) chunks += P.taps[t].kblocks;
      P.items_per_phase[ph] = (chunks + 1) / 2;
      items += P.items_per_phase[ph];
    }
  } else if (pair) {
    for (int i = 0; i < nt; ++i) items += (long long)(P.taps[i].kblocks * 64 / 256) * (d->n_total / 256);
  } else {
    for (int i = 0; i < nt; ++i) items += (long long)(P.taps[i].kblocks * 64 / kt) * P.n_tiles;
  }
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  int sms = persistent_sms();
  if (sms <= 0) sms = 148;
  if (pair) sms /= 2;          // work items are CTA pairs: one cluster per TPC
  // pixel splits ("split-K over pixels"): fill (at most) two full waves of CTAs -- rounding DOWN so the last wave is
  // not nearly empty -- and keep at least ~4 pixel tiles per CTA so the TMEM drain + atomics are amortised
  long long splits = (2LL * sms) / items;
  if (splits > m_tiles) splits = m_tiles;
  if (splits < 1) splits = 1;
  while (splits > 1 && m_tiles / splits < 4) --splits;
  // more work items than SMs: an unsplit grid can end in a nearly empty wave (the 768-channel 3x3 wgrad has 162 items
  // -> 148 + 14 CTAs, 55 % efficiency); if so, take the small split count with the best wave efficiency
  if (splits == 1) {
    auto eff = [&](long long sp) {
      const long long g = items * sp;
      return (double)g / (double)(((g + sms - 1) / sms) * sms);
    };
    if (eff(1) < 0.8) {
      double best = eff(1);
      for (long long sp = 2; sp <= 8 && m_tiles / sp >= 16; ++sp)
        if (eff(sp) > best + 0.02) {
          best = eff(sp);
          splits = sp;
        }
    }
  }
  // Cost model on top (TVAE_WGRAD_SPLIT_MODEL=0: the rule above alone, A/B switch): time ~ waves x (pixel tiles per CTA +
  // F), F = the CTA's pipeline fill + TMEM drain + slice store in units of one pixel tile's MMA time (~5 us = 8 tiles),
  // plus the fixed-order slice reduce, which grows with the split count.  The rule above always aims at TWO waves; when
  // one wave of twice as long CTAs fills the machine just as well it pays F once instead of twice (small-weight layers:
  // 24-36 work items) -- the model only ever LOWERS the split count.
  static const bool split_model = !(getenv("TVAE_WGRAD_SPLIT_MODEL") && atoi(getenv("TVAE_WGRAD_SPLIT_MODEL")) == 0);
  if (split_model && splits > 1) {
    static const double F = getenv("TVAE_WGRAD_SPLIT_F") ? atof(getenv("TVAE_WGRAD_SPLIT_F")) : 8.0;   // tuning switch
    const double R = 0.25;
    auto cost = [&](long long sp) {
      const long long g = items * sp;
      const long long waves = (g + sms - 1) / sms;
      const long long tiles = (m_tiles + sp - 1) / sp;
      return (double)waves * ((double)tiles + F) + (sp > 1 ? R * (double)sp : 0.0);
    };
    long long best = splits;
    double best_c = cost(splits);
    for (long long sp = splits - 1; sp >= 1; --sp) {
      const double c = cost(sp);
      if (c < best_c * 0.995) {
        best_c = c;
        best = sp;
      }
    }
    splits = best;
  }
  P.splits = (int)splits;
  const long long grid = items * splits * (pair ? 2 : 1);
  TVAE_REQUIRE(grid < (1LL << 31), "wgrad: grid too large");

  CUtensorMap mA0, mA1, mDZ;
  int rc;
  if (halo) {
    if ((rc = make_tmap_pix_halo(&mA0, d->a0.ptr, d->a0.B, d->a0.H, d->a0.W, d->a0.C))) return rc;
  } else if ((rc = make_tmap_pix(&mA0, d->a0.ptr, d->a0.B, d->a0.H, d->a0.W, d->a0.C, d->a0.split, P.tw, P.th, P.nb))) {
    return rc;
  }
  if (d->a1.ptr) {
    if ((rc = make_tmap_pix(&mA1, d->a1.ptr, d->a1.B, d->a1.H, d->a1.W, d->a1.C, d->a1.split, P.tw, P.th, P.nb))) return rc;
  } else {
    mA1 = mA0;
  }
  if ((rc = make_tmap_pix(&mDZ, d->out.ptr, d->out.B, d->out.H, d->out.W, d->out.C, d->out.split, P.tw, P.th, P.nb))) return rc;
  // more than one pixel split: partial blocks go to workspace slices, a second launch adds them in slice order
  // (TVAE_WGRAD_DET=0: atomics straight into the gradient, the A/B switch; TVAE_WGRAD_WS_POISON=1: NaN-filled workspace,
  // so a block that some split fails to write shows up in the tests)
  static const bool det = !(getenv("TVAE_WGRAD_DET") && atoi(getenv("TVAE_WGRAD_DET")) == 0);
  static const bool poison = getenv("TVAE_WGRAD_WS_POISON") && atoi(getenv("TVAE_WGRAD_WS_POISON")) != 0;
  const long long n_dw = (long long)d->n_total * P.k_total, n_db = db != nullptr ? (long long)d->num_phases * d->n_total : 0;
  float* ws = nullptr;
  bool shared_slabs = false;     // two phases writing the same weight columns
  for (int i = 0; i < nt && !shared_slabs; ++i)
    for (int j = i + 1; j < nt; ++j)
      if (P.taps[i].ph != P.taps[j].ph && P.taps[i].wk_off == P.taps[j].wk_off) {
        shared_slabs = true;
        break;
      }
  P.ws_phases = shared_slabs ? d->num_phases : 1;
  const long long slices = splits * P.ws_phases;
  if (det && slices > 1) {
    const size_t need = (size_t)slices * (size_t)(n_dw + n_db) * sizeof(float);
    if ((rc = scratch_workspace(need, reinterpret_cast<void**>(&ws)))) return rc;
    if (shared_slabs) TVAE_CHECK_CUDA(cudaMemsetAsync(ws, 0, need, stream));   // a phase writes only its own slabs
    else if (poison) TVAE_CHECK_CUDA(cudaMemsetAsync(ws, 0xFF, need, stream));
    P.dw = ws;                            // slice = [dw | db]
    P.ws_dw_stride = n_dw + n_db;
    if (db != nullptr) {
      P.db = ws + n_dw;
      P.ws_db_stride = n_dw + n_db;
    }
  } else {
    P.ws_phases = 1;
  }
  if (halo) {
    rc = d->n_total == 192 ? launch_wgh<192>(mA0, mDZ, P, (int)grid, stream) : launch_wgh<64>(mA0, mDZ, P, (int)grid, stream);
  } else if (transposed) {
    rc = d->n_total == 192 ? launch_wgt<192>(mA0, mA1, mDZ, P, (int)grid, stream)
                           : launch_wgt<64>(mA0, mA1, mDZ, P, (int)grid, stream);
  } else if (pair) {
    rc = launch_wg2<256>(mA0, mA1, mDZ, P, (int)grid, stream);
  } else {
    switch (kt) {
      case 256: rc = launch_wg<256>(mA0, mA1, mDZ, P, (int)grid, stream); break;
      case 192: rc = launch_wg<192>(mA0, mA1, mDZ, P, (int)grid, stream); break;
      case 128: rc = launch_wg<128>(mA0, mA1, mDZ, P, (int)grid, stream); break;
      default: rc = launch_wg<64>(mA0, mA1, mDZ, P, (int)grid, stream); break;
    }
  }
  if (rc) return rc;
  if (ws != nullptr) {
    wgrad_reduce_kernel<<<(unsigned)(((n_dw + n_db) / 4 + 255) / 256), 256, 0, stream>>>(dw, db, ws, n_dw, n_db, (int)slices);
    TVAE_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace tvae
