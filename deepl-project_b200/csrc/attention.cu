// Flash-style attention forward for sm_100a (head_dim 64, non-causal, no mask).
//
// Replaces F.scaled_dot_product_attention(q, k, v, scale=64^-0.5) at attention.py:88-92.  q, k, v live
// interleaved per token in the output of the fused QKV projection: qkv[b, s, 3C] bf16 with q in columns
// [0, C), k in [C, 2C), v in [2C, 3C), head h at columns h*64..h*64+63.  The QKV GEMM epilogue has already
// applied the reference's RoPE to q and k and multiplied q by scale*log2(e), so scores are in log2 units.
//
// One CTA per (256-query tile = two 128-row Q tiles, head, image), 12 warps:
//   warp 0     TMA producer : both Q tiles once, then K/V tiles through a 4-stage ring
//   warp 1     MMA issuer   : per K/V block and Q tile t:  S_t = Q_t K^T (tcgen05, M=128 N=128 K=64 -> TMEM) and
//                             O_t = P_t V (M=128 N=64 K=128; P from smem, V as an MN-major operand)
//   warps 4-7  softmax for Q tile 0, warps 8-11 for Q tile 1 : one query row per thread: tcgen05.ld S, online
//              max / sum with exp2, P -> bf16 -> swizzled smem, O accumulated in registers (rescaled by
//              exp2(m_old - m_new)), final O / l -> bf16 -> TMA store; log-sum-exp kept for the backward pass.
// The two softmax warpgroups ping-pong: while one exponentiates, the tensor pipe runs the other tile's S / PV
// and the TMEM-load latency of one warp hides behind the other warp on the same scheduler.  (The first version -- one
// Q tile, one softmax warpgroup -- reached 310-340 TFLOP/s at S=4096; profiles/r1_breakdown_*.txt.)
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "tmap.cuh"

namespace tvae {

constexpr int kAttStages = 4;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: 128 rows x 64 bf16
constexpr int kAttThreads = 384;
constexpr int kAttSmem = 2 * kTileBytes /*Q0,Q1*/ + kAttStages * 2 * kTileBytes /*K,V*/ + 2 * 2 * kTileBytes /*P0,P1*/ +
                         1024 + 256;

__device__ __forceinline__ void att_tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Row maximum of one 128-wide score block (one row per thread).  MASK: the block reaches past the sequence end.
template <bool MASK>
__device__ __forceinline__ float att_row_max(uint32_t t_s, int key0, int S) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(t_s + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 1]);
      if (MASK) {
        a = (key0 + c * 32 + i < S) ? a : -INFINITY;
        b = (key0 + c * 32 + i + 1 < S) ? b : -INFINITY;
      }
      m0 = fmaxf(m0, a);
      m1 = fmaxf(m1, b);
    }
  }
  return fmaxf(m0, m1);
}

// P = exp2(S - m_new) -> bf16 -> swizzled smem (K-major A operand of the P V product); returns the row sum.
template <bool MASK>
__device__ __forceinline__ float att_exp_store(uint32_t t_s, uint8_t* sPt, int r, int key0, int S, float m_new) {
  float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;   // four independent chains: a single running sum is a 32-deep
  const float neg_m = -m_new;                          // dependent FADD chain per chunk
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(t_s + c * 32, v);
    tmem_ld_wait();
    float p[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float e = exp2f(__uint_as_float(v[i]) + neg_m);
      if (MASK) e = (key0 + c * 32 + i < S) ? e : 0.0f;
      p[i] = e;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      l0 += p[i];
      l1 += p[i + 1];
      l2 += p[i + 2];
      l3 += p[i + 3];
    }
    uint8_t* row = sPt + (c >> 1) * kTileBytes + r * 128;
    const int cbase = (c & 1) * 4;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 o;
      o.x = pack_bf16(p[g * 8 + 0], p[g * 8 + 1]);
      o.y = pack_bf16(p[g * 8 + 2], p[g * 8 + 3]);
      o.z = pack_bf16(p[g * 8 + 4], p[g * 8 + 5]);
      o.w = pack_bf16(p[g * 8 + 6], p[g * 8 + 7]);
      *reinterpret_cast<uint4*>(row + (((cbase + g) ^ (r & 7)) << 4)) = o;
    }
  }
  return (l0 + l1) + (l2 + l3);
}

__global__ void __launch_bounds__(kAttThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                float* __restrict__ lse, int S, int C, int nh) {
#ifdef TVAE_DEVICE_OK
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // [2 tiles]
  uint8_t* sK = sQ + 2 * kTileBytes;                 // [stages]
  uint8_t* sV = sK + kAttStages * kTileBytes;        // [stages]
  uint8_t* sP = sV + kAttStages * kTileBytes;        // [2 tiles] x 2 chunks (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kTileBytes);
  uint64_t* q_full = bars;                           // 1
  uint64_t* kv_full = bars + 1;                      // [stages]
  uint64_t* kv_empty = kv_full + kAttStages;         // [stages]
  uint64_t* s_full = kv_empty + kAttStages;          // [2 tiles][2 buffers]
  uint64_t* b_free = s_full + 4;                     // [2 tiles][2 buffers]
  uint64_t* p_full = b_free + 4;                     // [2 tiles]
  uint64_t* o_full = p_full + 2;                     // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (S + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      for (int u = 0; u < 2; ++u) {
        mbar_init(&s_full[t * 2 + u], 1);
        mbar_init(&b_free[t * 2 + u], 4);
      }
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM: tile t, buffer u at columns t*256 + u*128: S_t(j) fills buffer j&1 (128 columns); O_t(j) = P_t(j) V_j is
  // written over the first 64 columns of the same buffer once the softmax warps have consumed S_t(j).  All 512 columns
  // are in use, and S_t(j+1) is computed while the softmax of block j is still running.

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      att_tma_load_3d(sQ, &tmQKV, q_full, h * 64, q0, b);
      att_tma_load_3d(sQ + kTileBytes, &tmQKV, q_full, h * 64, q0 + 128, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], 2 * kTileBytes);
        att_tma_load_3d(sK + stage * kTileBytes, &tmQKV, &kv_full[stage], C + h * 64, j * 128, b);
        att_tma_load_3d(sV + stage * kTileBytes, &tmQKV, &kv_full[stage], 2 * C + h * 64, j * 128, b);
        if (++stage == kAttStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      auto issue_qk = [&](int j, int t) {
        const int stage = j % kAttStages;
        const int u = j & 1;
        if (t == 0) mbar_wait(&kv_full[stage], (j / kAttStages) & 1);
        mbar_wait(&b_free[t * 2 + u], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_base = smem_u32(sQ + t * kTileBytes);
        const uint32_t k_base = smem_u32(sK + stage * kTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + t * 256 + u * 128, umma_desc_kmajor_sw128(q_base + k * 32),
                   umma_desc_kmajor_sw128(k_base + k * 32), idesc_qk, k != 0);
        umma_commit(&s_full[t * 2 + u]);
      };
      auto issue_pv = [&](int j, int t) {
        const int stage = j % kAttStages;
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
        const uint32_t p_base = smem_u32(sP + t * 2 * kTileBytes);
        const uint32_t v_base = smem_u32(sV + stage * kTileBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + t * 256 + (j & 1) * 128, umma_desc_kmajor_sw128(p_base + (k >> 2) * kTileBytes + (k & 3) * 32),
                   umma_desc_mnmajor_sw128(v_base + k * 2048, 1024, 1024), idesc_pv, k != 0);
        umma_commit(&o_full[t]);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < 2 && j < nblk; ++j) {
        issue_qk(j, 0);
        issue_qk(j, 1);
      }
      for (int j = 0; j < nblk; ++j) {
        issue_pv(j, 0);
        issue_pv(j, 1);
        umma_commit(&kv_empty[j % kAttStages]);
        if (j + 2 < nblk) {
          issue_qk(j + 2, 0);
          issue_qk(j + 2, 1);
        }
      }
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;       // Q tile of this warpgroup
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t t_tile = tmem_base + lane_off + t * 256;
    uint8_t* sPt = sP + t * 2 * kTileBytes;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, alpha_prev = 1.0f;

    for (int j = 0; j < nblk; ++j) {
      const uint32_t t_s = t_tile + (j & 1) * 128;
      mbar_wait(&s_full[t * 2 + (j & 1)], (j >> 1) & 1);
      tc_fence_after();
      const int key0 = j * 128;
      const bool partial = key0 + 128 > S;     // only the last block can reach past the sequence end
      // pass 1: row max
      const float m_blk = partial ? att_row_max<true>(t_s, key0, S) : att_row_max<false>(t_s, key0, S);
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = exp2f(m_run - m_new);   // m_run = -inf on the first block -> 0
      // the previous block's P V must be complete before sP is overwritten; fold it into the accumulator now
      if (j > 0) {
        mbar_wait(&o_full[t], (j - 1) & 1);
        tc_fence_after();
        const uint32_t t_o = t_tile + ((j - 1) & 1) * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_o + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
        // buffer (j-1)&1 of this tile may now receive S_t(j+1)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_free[t * 2 + ((j - 1) & 1)]);
      }
      // pass 2: P = exp2(S - m_new) -> bf16 -> swizzled smem
      const float l_blk = partial ? att_exp_store<true>(t_s, sPt, r, key0, S, m_new)
                                  : att_exp_store<false>(t_s, sPt, r, key0, S, m_new);
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      alpha_prev = alpha;
      // S buffer drained, P written
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // last block's P V
    mbar_wait(&o_full[t], (nblk - 1) & 1);
    tc_fence_after();
    const uint32_t t_o = t_tile + ((nblk - 1) & 1) * 128;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(t_o + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
    }
    const float inv_l = 1.0f / l_run;
    // stage O (bf16) in this tile's P buffer and store with TMA (rows beyond S are clipped by the tensor map)
    uint8_t* row = sPt + r * 128;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint4 o;
      o.x = pack_bf16(o_acc[g * 8 + 0] * inv_l, o_acc[g * 8 + 1] * inv_l);
      o.y = pack_bf16(o_acc[g * 8 + 2] * inv_l, o_acc[g * 8 + 3] * inv_l);
      o.z = pack_bf16(o_acc[g * 8 + 4] * inv_l, o_acc[g * 8 + 5] * inv_l);
      o.w = pack_bf16(o_acc[g * 8 + 6] * inv_l, o_acc[g * 8 + 7] * inv_l);
      *reinterpret_cast<uint4*>(row + ((g ^ (r & 7)) << 4)) = o;
    }
    const int qrow = q0 + t * 128 + r;
    if (lse != nullptr && qrow < S) lse[((size_t)b * nh + h) * S + qrow] = m_run + log2f(l_run);
    fence_proxy_async_smem();
    if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
    if (qd == 0 && lane == 0 && q0 + t * 128 < S) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(&tmO)),
                   "r"(smem_u32(sPt)), "r"(h * 64), "r"(q0 + t * 128), "r"(b)
                   : "memory");
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
#endif
}

int attn_fwd_run(const void* qkv, void* out, float* lse, int B, int S, int C, cudaStream_t stream) {
  TVAE_REQUIRE(C % 64 == 0, "attention: C=%d must be a multiple of head_dim 64", C);
  const int nh = C / 64;
  CUtensorMap mQKV, mO;
  int rc;
  if ((rc = make_tmap_3d(&mQKV, qkv, 3 * (uint64_t)C, S, B, 3 * (uint64_t)C, (uint64_t)S * 3 * C, 128))) return rc;
  if ((rc = make_tmap_3d(&mO, out, C, S, B, C, (uint64_t)S * C, 128))) return rc;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    configured = true;
  }
  dim3 grid((S + 255) / 256, nh, B);
  attn_fwd_kernel<<<grid, kAttThreads, kAttSmem, stream>>>(mQKV, mO, lse, S, C, nh);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
