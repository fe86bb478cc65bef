// Flash-style attention forward for sm_100a (head_dim 64, non-causal, no mask).
//
// Replaces F.scaled_dot_product_attention(q, k, v, scale=64^-0.5) at attention.py:88-92.  q, k, v live
// interleaved per token in the output of the fused QKV projection: qkv[b, s, 3C] bf16 with q in columns
// [0, C), k in [C, 2C), v in [2C, 3C), head h at columns h*64..h*64+63.  The QKV GEMM epilogue has already
// applied the reference's RoPE to q and k and multiplied q by scale*log2(e), so scores are in log2 units.
//
// One CTA per (256-query tile = two 128-row Q tiles, head, image), 12 warps:
//   warp 0     TMA producer : both Q tiles once, then K/V tiles through a 2-stage ring
//   warp 1     MMA issuer   : per K/V block j and Q tile t:  S_t = Q_t K_j^T (tcgen05, M=128 N=128 K=64 -> TMEM) and
//                             O_t += P_t V_j (M=128 N=64 K=128; P from smem, V as an MN-major operand), O_t
//                             ACCUMULATES in TMEM over the whole loop
//   warps 4-7  softmax for Q tile 0, warps 8-11 for Q tile 1 (one query row per thread): ONE streaming tcgen05.ld pass
//              over the 128 scores: P = exp2(S - m_ref) -> bf16 -> swizzled smem, row sum and row max on the fly.
// m_ref is the running reference maximum from earlier blocks (speculation): it is only raised -- the block redone and
// O_t rescaled in TMEM with tcgen05.ld / tcgen05.st -- when the block maximum turns out to exceed it by more than 8
// (a factor 256, harmless in bf16 P / fp32 O).  After the first few blocks this almost never happens, so per block O
// is neither read nor rescaled and S is read from TMEM exactly once.  Earlier versions (profiles/r1_ncu_attn_*)
// read S twice and folded O through registers every block: 160 KiB of TMEM reads per 128x128 block, 310-540 TFLOP/s.
// S_t(j+1) is issued as soon as the softmax warps have pulled S_t(j) into registers, so the tensor pipe runs ahead.
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "tmap.cuh"

namespace tvae {

constexpr int kAttStages = 2;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: 128 rows x 64 bf16
constexpr int kAttThreads = 384;
constexpr int kAttSmem = 2 * kTileBytes /*Q0,Q1*/ + kAttStages * 2 * kTileBytes /*K,V*/ + 2 * 2 * 2 * kTileBytes /*P double buffered*/ +
                         1024 + 256;
constexpr float kRescaleThreshold = 8.0f;   // log2 units

__device__ __forceinline__ void att_tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
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
  uint8_t* sP = sV + kAttStages * kTileBytes;        // [2 tiles][2 buffers] x 2 chunks (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 8 * kTileBytes);
  uint64_t* q_full = bars;                           // 1
  uint64_t* k_full = bars + 1;                       // [stages]  K and V have separate rings: K_j is released as soon
  uint64_t* k_empty = k_full + kAttStages;           // [stages]  as both S_t(j) are issued, V_j only after both P_t(j) V
  uint64_t* v_full = k_empty + kAttStages;           // [stages]
  uint64_t* v_empty = v_full + kAttStages;           // [stages]
  uint64_t* s_full = v_empty + kAttStages;           // [2 tiles]
  uint64_t* s_free = s_full + 2;                     // [2 tiles] S_t pulled into registers
  uint64_t* p_full = s_free + 2;                     // [2 tiles]
  uint64_t* o_full = p_full + 2;                     // [2 tiles][2]: P_t(j) V done (buffer j&1 free, O_t updated)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

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
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t * 2], 1);
      mbar_init(&o_full[t * 2 + 1], 1);
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
  // TMEM columns: S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      att_tma_load_3d(sQ, &tmQKV, q_full, h * 64, q0, b);
      att_tma_load_3d(sQ + kTileBytes, &tmQKV, q_full, h * 64, q0 + 128, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], kTileBytes);
        att_tma_load_3d(sK + stage * kTileBytes, &tmQKV, &k_full[stage], C + h * 64, j * 128, b);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], kTileBytes);
        att_tma_load_3d(sV + stage * kTileBytes, &tmQKV, &v_full[stage], 2 * C + h * 64, j * 128, b);
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
        if (t == 0) mbar_wait(&k_full[stage], (j / kAttStages) & 1);
        mbar_wait(&s_free[t], (j & 1) ^ 1);       // S_t(j-1) consumed (passes immediately for j = 0)
        tc_fence_after();
        const uint32_t q_base = smem_u32(sQ + t * kTileBytes);
        const uint32_t k_base = smem_u32(sK + stage * kTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + t * 128, umma_desc_kmajor_sw128(q_base + k * 32), umma_desc_kmajor_sw128(k_base + k * 32),
                   idesc_qk, k != 0);
        umma_commit(&s_full[t]);
        if (t == 1) umma_commit(&k_empty[stage]);
      };
      auto issue_pv = [&](int j, int t) {
        const int stage = j % kAttStages;
        if (t == 0) mbar_wait(&v_full[stage], (j / kAttStages) & 1);
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
        const uint32_t p_base = smem_u32(sP + (t * 2 + (j & 1)) * 2 * kTileBytes);
        const uint32_t v_base = smem_u32(sV + stage * kTileBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + 256 + t * 64, umma_desc_kmajor_sw128(p_base + (k >> 2) * kTileBytes + (k & 3) * 32),
                   umma_desc_mnmajor_sw128(v_base + k * 2048, 1024, 1024), idesc_pv, (j | k) != 0);
        umma_commit(&o_full[t * 2 + (j & 1)]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0, 0);
      issue_qk(0, 1);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          issue_qk(j + 1, 0);
          issue_qk(j + 1, 1);
        }
        issue_pv(j, 0);
        issue_pv(j, 1);
        umma_commit(&v_empty[j % kAttStages]);
      }
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;       // Q tile of this warpgroup
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + t * 128;
    const uint32_t t_o = tmem_base + lane_off + 256 + t * 64;
    float m_ref = -INFINITY, l_run = 0.0f;

    // one streaming pass over S_t(j): p = exp2(s - m_ref) -> smem buffer, returns (row sum, row max)
    auto pass = [&](uint8_t* sPt, int key0, float mref, float& lsum, float& mblk) {
      const float neg_m = -mref;
      const bool partial = key0 + 128 > S;
      float l0 = 0.0f, l1 = 0.0f, m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_s + c * 32, v);
        tmem_ld_wait();
        if (partial) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + c * 32 + i >= S) v[i] = 0xff800000u;   // -inf: exp2 -> 0, ignored by the max
        }
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a = __uint_as_float(v[i]), bq = __uint_as_float(v[i + 1]);
          m0 = fmaxf(m0, a);
          m1 = fmaxf(m1, bq);
          p[i] = exp2f(a + neg_m);
          p[i + 1] = exp2f(bq + neg_m);
          l0 += p[i];
          l1 += p[i + 1];
        }
        uint8_t* row = sPt + (c >> 1) * kTileBytes + r * 128;
        const int cbase = (c & 1) * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(row + (((cbase + g) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16(p[g * 8 + 0], p[g * 8 + 1]), pack_bf16(p[g * 8 + 2], p[g * 8 + 3]),
                         pack_bf16(p[g * 8 + 4], p[g * 8 + 5]), pack_bf16(p[g * 8 + 6], p[g * 8 + 7]));
      }
      lsum = l0 + l1;
      mblk = fmaxf(m0, m1);
    };

    for (int j = 0; j < nblk; ++j) {
      uint8_t* sPt = sP + (t * 2 + (j & 1)) * 2 * kTileBytes;
      const int key0 = j * 128;
      // P buffer j&1 was last read by P_t(j-2) V
      if (j >= 2) mbar_wait(&o_full[t * 2 + (j & 1)], ((j - 2) >> 1) & 1);
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float l_blk, m_blk;
      if (j == 0) {                       // no reference yet: plain max pass first
        float m0 = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + c * 32 + i < S) m0 = fmaxf(m0, __uint_as_float(v[i]));
        }
        m_ref = m0;
      }
      pass(sPt, key0, m_ref, l_blk, m_blk);
      if (__any_sync(0xffffffffu, m_blk > m_ref + kRescaleThreshold)) {
        // rare: the speculated reference was too small for some row of this warp.  Raise it for those rows, rescale
        // their O / l, and redo the block (rows that did not need it recompute identical values).
        const bool need = m_blk > m_ref + kRescaleThreshold;
        const float factor = need ? exp2f(m_ref - m_blk) : 1.0f;
        if (need) m_ref = m_blk;
        l_run *= factor;
        if (j > 0) {
          mbar_wait(&o_full[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);   // O_t final for block j-1
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(t_o + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
            tmem_st32(t_o + c * 32, v);
          }
          tmem_st_wait();
        }
        pass(sPt, key0, m_ref, l_blk, m_blk);
      }
      l_run += l_blk;
      // S_t consumed (S_t(j+1) may overwrite it), P_t(j) written
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_free[t]);
        mbar_arrive(&p_full[t]);
      }
    }
    uint8_t* sPt = sP + (t * 2) * 2 * kTileBytes;   // staging for the output tile (buffer 0 is free after the last P V)
    // ---- epilogue: O_t / l -> bf16 -> smem staging (this tile's P buffer) -> TMA store
    mbar_wait(&o_full[t * 2 + ((nblk - 1) & 1)], ((nblk - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    uint8_t* row = sPt + r * 128;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(t_o + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
        o.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
        o.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
        o.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(row + (((c * 4 + g) ^ (r & 7)) << 4)) = o;
      }
    }
    const int qrow = q0 + t * 128 + r;
    if (lse != nullptr && qrow < S) lse[((size_t)b * nh + h) * S + qrow] = m_ref + log2f(l_run);
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
