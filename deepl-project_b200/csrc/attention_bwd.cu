// Flash-style attention backward for sm_100a (head_dim 64, non-causal), tcgen05 + TMEM + TMA.
//
// Backward of F.scaled_dot_product_attention (attention.py:88-92) on the layout of attention.cu:
//   qkv [B, S, 3C] bf16 (q~ = scale*log2e * RoPE(q), k~ = RoPE(k), v), o / do [B, S, C] bf16,
//   lse [B, C/64, S] fp32 (log2 domain), delta [B, C/64, S] fp32 = rowsum(do * o).
// One CTA per (128-key tile j, head, image) keeps K_j, V_j in shared memory and walks over the query tiles i:
//   S~ = Q_i K_j^T, dP = dO_i V_j^T                       (2 MMAs -> TMEM)
//   P = exp2(S~ - lse), dZ = P * (dP - delta)             (softmax warps; bf16 P, dZ -> swizzled smem)
//   dV += P^T dO_i, dK~ += dZ^T Q_i, dQ~_i = dZ K_j       (3 MMAs; P / dZ / dO / Q / K as MN-major operands)
// dV, dK~ accumulate in TMEM over the whole loop; dQ~_i goes to an fp32 global accumulator with red.add.
// Outputs: dqkv[:, :, C:2C] = ln2 * dK~ (still in rotated space), dqkv[:, :, 2C:3C] = dV, dq_acc fp32 [B, S, C]
// (rotated space, unscaled); tvae_rope_bwd then applies the transposed RoPE and the softmax scale.
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "tmap.cuh"

namespace tvae {

constexpr int kBT = 128 * 64 * 2;  // 16 KiB tile
constexpr int kBwdThreads = 384;   // 4 control warps + 8 softmax warps (two per TMEM lane quarter, 64 key columns each)
constexpr int kBwdSmem = 2 * kBT /*K,V*/ + 2 * 2 * kBT /*Q,dO ring*/ + 2 * 2 * 2 * kBT /*P, dZ double buffered*/ + 1024 + 256;

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const float* __restrict__ lse, const float* __restrict__ delta,
                __nv_bfloat16* __restrict__ dqkv, int S, int C, int nh) {
#ifdef TVAE_DEVICE_OK
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kBT;
  uint8_t* sQ = sV + kBT;            // [2 stages]
  uint8_t* sDO = sQ + 2 * kBT;       // [2 stages]
  uint8_t* sP = sDO + 2 * kBT;       // [2 buffers] x 2 chunks (keys 0-63, 64-127)
  uint8_t* sDZ = sP + 4 * kBT;       // [2 buffers] x 2 chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDZ + 4 * kBT);
  uint64_t* kv_full = bars;          // 1
  uint64_t* qdo_full = bars + 1;     // [2]
  uint64_t* qdo_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;     // 1
  uint64_t* pds_full = bars + 6;     // 1 (8 warp arrivals)
  uint64_t* mma_done = bars + 7;     // [2] (alternating, so a waiter never lags two phases)
  uint64_t* sdp_free = bars + 9;     // 1 (8 warp arrivals): S~ / dP of the current tile sit in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nq = (S + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(sdp_free, 8);
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
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
  const uint32_t t_S = tmem_base, t_dP = tmem_base + 128, t_dV = tmem_base + 256, t_dK = tmem_base + 320,
                 t_dQ = tmem_base + 384;  // 2 x 64

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");      // control warpgroup (TMA / MMA issue)
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * kBT);
      tma_load_3d(sK, &tmQKV, kv_full, C + h * 64, k0, b);
      tma_load_3d(sV, &tmQKV, kv_full, 2 * C + h * 64, k0, b);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kBT);
        tma_load_3d(sQ + st * kBT, &tmQKV, &qdo_full[st], h * 64, i * 128, b);
        tma_load_3d(sDO + st * kBT, &tmDO, &qdo_full[st], h * 64, i * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_kk = umma_idesc_bf16(128, 128, 0, 0);   // S~, dP : both operands K-major (d contiguous)
      constexpr uint32_t id_mm = umma_idesc_bf16(128, 64, 1, 1);    // dV, dK : both MN-major (reduction over queries)
      constexpr uint32_t id_km = umma_idesc_bf16(128, 64, 0, 1);    // dQ     : A = dZ K-major, B = K_j MN-major
      const uint32_t k_base = smem_u32(sK), v_base = smem_u32(sV);
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        mbar_wait(&qdo_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t q_base = smem_u32(sQ + st * kBT), do_base = smem_u32(sDO + st * kBT);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(t_S, umma_desc_kmajor_sw128(q_base + k * 32), umma_desc_kmajor_sw128(k_base + k * 32), id_kk, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(t_dP, umma_desc_kmajor_sw128(do_base + k * 32), umma_desc_kmajor_sw128(v_base + k * 32), id_kk, k != 0);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        // S~ / dP of tile i+1 go to the tensor pipe as soon as the softmax warps hold tile i in registers, i.e.
        // they overlap the exponentiation of tile i (the first version issued them after dV / dK / dQ of tile i, so
        // tensor work and softmax strictly alternated)
        if (i + 1 < nq) {
          mbar_wait(sdp_free, i & 1);
          tc_fence_after();
          issue_sdp(i + 1);
        }
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
        const uint32_t q_base = smem_u32(sQ + st * kBT), do_base = smem_u32(sDO + st * kBT);
        const uint32_t p_base = smem_u32(sP + (i & 1) * 2 * kBT), dz_base = smem_u32(sDZ + (i & 1) * 2 * kBT);
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 16 queries per MMA
          umma_f16(t_dV, umma_desc_mnmajor_sw128(p_base + k * 2048, kBT, 1024),
                   umma_desc_mnmajor_sw128(do_base + k * 2048, kBT, 1024), id_mm, (i | k) != 0);
          umma_f16(t_dK, umma_desc_mnmajor_sw128(dz_base + k * 2048, kBT, 1024),
                   umma_desc_mnmajor_sw128(q_base + k * 2048, kBT, 1024), id_mm, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)     // 16 keys per MMA
          umma_f16(t_dQ + (i & 1) * 64, umma_desc_kmajor_sw128(dz_base + (k >> 2) * kBT + (k & 3) * 32),
                   umma_desc_mnmajor_sw128(k_base + k * 2048, kBT, 1024), id_km, k != 0);
        umma_commit(&qdo_empty[st]);
        umma_commit(&mma_done[i & 1]);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");    // softmax warpgroups: 64 S~ + 64 dP values per thread
    // Eight softmax warps: warps w and w + 4 share TMEM lane quarter (w & 3) -- i.e. the same 32 query rows -- and
    // split the 128 key columns in halves.  Nothing in the backward softmax reduces along a row (lse and delta come
    // from the forward pass), so the halves are independent; with one warp per scheduler (the first version) every
    // TMEM-load, MUFU and shared-store latency was exposed and the kernel ran at 28 % of the tensor peak.
    const int qd = warp & 3;
    const int half = (warp - 4) >> 2;          // key columns [64 * half, 64 * half + 64)
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const size_t stat_base = ((size_t)b * nh + h) * S;

    // dQ~_i (this half's 32 columns of the 128 x 64 fp32 tile) -> swizzled smem -> ONE bulk tensor reduce-add into the
    // fp32 accumulator.  The staging area is the dZ buffer of tile i, which the tensor core has finished reading.
    // (The first version issued per-thread red.global.add.v4.f32: every warp instruction touched 32 different rows,
    // 6.4 GB of scattered 16-byte atomics per launch -- the L2 atomic path, not the tensor pipe, set the pace.)
    auto drain_dq = [&](int i) {
      uint8_t* stage = sDZ + ((i & 1) * 2 + half) * kBT;
      uint32_t v[32];
      tmem_ld32(t_dQ + (i & 1) * 64 + lane_off + half * 32, v);
      tmem_ld_wait();
      uint8_t* row = stage + r * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(row + ((g ^ (r & 7)) << 4)) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      fence_proxy_async_smem();
      if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (qd == 0 && lane == 0) {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmDQ)),
                     "r"(smem_u32(stage)), "r"(h * 64 + half * 32), "r"(i * 128), "r"(b)
                     : "memory");
        tma_store_commit();
        tma_store_wait_read<0>();      // the staging buffer is rewritten by the softmax of tile i + 2
      }
      if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };

    for (int i = 0; i < nq; ++i) {
      const int qrow = i * 128 + r;
      const float l2 = (qrow < S) ? __ldg(lse + stat_base + qrow) : INFINITY;
      const float dl = (qrow < S) ? __ldg(delta + stat_base + qrow) : 0.0f;
      const float2 nl2 = make_float2(-l2, -l2), ndl = make_float2(-dl, -dl);
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      uint32_t sv[64], pv[64];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
        uint32_t(&p0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&pv[0]);
        uint32_t(&p1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&pv[32]);
        tmem_ld32(t_S + lane_off + half * 64, s0);
        tmem_ld32(t_S + lane_off + half * 64 + 32, s1);
        tmem_ld32(t_dP + lane_off + half * 64, p0);
        tmem_ld32(t_dP + lane_off + half * 64 + 32, p1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);                          // the S~ / dP accumulators may be overwritten
      if (i >= 2) mbar_wait(&mma_done[i & 1], ((i - 2) >> 1) & 1);   // P / dZ buffer i&1 no longer read by MMA(i-2)
      uint8_t* prow = sP + ((i & 1) * 2 + half) * kBT + r * 128;
      uint8_t* zrow = sDZ + ((i & 1) * 2 + half) * kBT + r * 128;
      const bool full_tile = k0 + 128 <= S;                          // no masking needed
#pragma unroll
      for (int g = 0; g < 8; ++g) {                                  // 8 columns -> one 16-byte store of P and of dZ
        uint32_t pk[4], zk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = g * 8 + 2 * k;
          float2 e = __fadd2_rn(make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), nl2);
          e.x = exp2f(e.x);
          e.y = exp2f(e.y);
          if (!full_tile) {
            if (k0 + half * 64 + j >= S) e.x = 0.0f;
            if (k0 + half * 64 + j + 1 >= S) e.y = 0.0f;
          }
          const float2 d = __fmul2_rn(e, __fadd2_rn(make_float2(__uint_as_float(pv[j]), __uint_as_float(pv[j + 1])), ndl));
          pk[k] = pack_bf16(e.x, e.y);
          zk[k] = pack_bf16(d.x, d.y);
        }
        const int off = (g ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(zrow + off) = make_uint4(zk[0], zk[1], zk[2], zk[3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (i > 0) {
        mbar_wait(&mma_done[(i - 1) & 1], ((i - 1) >> 1) & 1);   // dQ_{i-1} complete
        tc_fence_after();
        drain_dq(i - 1);
      }
    }
    mbar_wait(&mma_done[(nq - 1) & 1], ((nq - 1) >> 1) & 1);
    tc_fence_after();
    drain_dq(nq - 1);
    if (qd == 0 && lane == 0) tma_store_wait<0>();   // all reduce-adds of this CTA have landed
    // dK~ (x ln2, warps of half 0) and dV (half 1) of this key tile (row r = key k0 + r)
    const int krow = k0 + r;
    __nv_bfloat16* dst = dqkv + ((size_t)b * S + krow) * 3 * C + h * 64;
    {
      const int which = half;
      const uint32_t t = which ? t_dV : t_dK;
      const float sc = which ? 1.0f : 0.6931471805599453f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(t + lane_off + c * 32, v);
        tmem_ld_wait();
        if (krow < S) {
          uint4* o = reinterpret_cast<uint4*>(dst + (which ? 2 * C : C) + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * sc, __uint_as_float(v[g * 8 + 1]) * sc);
            u.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * sc, __uint_as_float(v[g * 8 + 3]) * sc);
            u.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * sc, __uint_as_float(v[g * 8 + 5]) * sc);
            u.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * sc, __uint_as_float(v[g * 8 + 7]) * sc);
            o[g] = u;
          }
        }
      }
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

int attn_bwd_run(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq_acc, void* dqkv, int B,
                 int S, int C, cudaStream_t stream) {
  TVAE_REQUIRE(C % 64 == 0, "attention_bwd: C=%d must be a multiple of 64", C);
  const int nh = C / 64;
  CUtensorMap mQKV, mDO;
  int rc;
  if ((rc = make_tmap_3d(&mQKV, qkv, 3 * (uint64_t)C, S, B, 3 * (uint64_t)C, (uint64_t)S * 3 * C, 128))) return rc;
  if ((rc = make_tmap_3d(&mDO, dout, C, S, B, C, (uint64_t)S * C, 128))) return rc;
  CUtensorMap mDQ;
  if ((rc = make_tmap_3d_f32(&mDQ, dq_acc, C, S, B, C, (uint64_t)S * C, 128))) return rc;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    configured = true;
  }
  TVAE_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * S * C * sizeof(float), stream));
  dim3 grid((S + 127) / 128, nh, B);
  attn_bwd_kernel<<<grid, kBwdThreads, kBwdSmem, stream>>>(mQKV, mDO, mDQ, lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv),
                                                           S, C, nh);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
