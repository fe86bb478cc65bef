// Flash-style attention backward for sm_100a (head_dim 64, non-causal), tcgen05 + TMEM + TMA.
//
// Backward of F.scaled_dot_product_attention (attention.py:88-92) on the layout of attention.cu:
//   qkv [B, S, 3C] bf16 (q~ = scale*log2e * RoPE(q), k~ = RoPE(k), v), o / do [B, S, C] bf16,
//   lse [B, C/64, S] fp32 (log2 domain), delta [B, C/64, S] fp32 = rowsum(do * o).
// One CTA per (128-key tile j, head, image) keeps K_j, V_j in shared memory and walks over the query tiles i:
//   S~ = Q_i K_j^T, dP = dO_i V_j^T                       (2 MMAs -> TMEM)
//   P = exp2(S~ - lse), dZ = P * (dP - delta)             (softmax warps; bf16 P, dZ -> swizzled smem)
//   dV += P^T dO_i, dK~ += dZ^T Q_i, dQ~_i = dZ K_j       (3 MMAs; P / dZ / dO / Q / K as MN-major operands)
// dV, dK~ accumulate in TMEM over the whole loop; dQ~_i goes to an fp32 global accumulator with a bulk tensor reduce-add
// issued by a dedicated drain warpgroup, so the softmax warps never wait for it.
// Outputs: dqkv[:, :, C:2C] = ln2 * dK~ (still in rotated space), dqkv[:, :, 2C:3C] = dV, dq_acc fp32 [B, S, C]
// (rotated space, unscaled); tvae_rope_bwd then applies the transposed RoPE and the softmax scale.
#include "../../include/transvae_sm100.h"
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "tmap.cuh"

namespace tvae {

constexpr int kBT = 128 * 64 * 2;  // 16 KiB tile
constexpr int kBwdThreads = 512;   // 4 control warps + 8 softmax warps (two per TMEM lane quarter, 64 key columns each)
                                   // + 4 drain warps (dQ~: TMEM -> smem -> bulk reduce-add)
constexpr int kQStages = 3;       // Q / dO ring
constexpr int kBwdSmem = 2 * kBT /*K,V*/ + 2 * kQStages * kBT /*Q,dO ring*/ + 2 * kBT /*P*/ + 2 * 2 * kBT /*dZ double buffered*/ + 1024 + 256;

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 2^x for two lanes on the FMA pipe (see attention.cu: round-to-nearest split with the 1.5 * 2^23 trick, degree-3 minimax
// polynomial on [-0.5, 0.5], max relative error 7.5e-5 -- far below the bf16 rounding of P; clamped at -126).
__device__ __forceinline__ float2 exp2_fma2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 xf = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
  const float2 n = __fadd2_rn(xf, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(make_float2(0.0551716685f, 0.0551716685f), r, make_float2(0.242611125f, 0.242611125f));
  p = __ffma2_rn(p, r, make_float2(0.693260968f, 0.693260968f));
  p = __ffma2_rn(p, r, make_float2(0.999928057f, 0.999928057f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(xf.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(xf.y) << 23));
  return p;
}

// POLY: every POLY-th pair of probabilities is exponentiated on the FMA pipe instead of MUFU (0 = never).
template <int POLY>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDQ1, const float* __restrict__ lse, const float* __restrict__ delta,
                __nv_bfloat16* __restrict__ dqkv, int* __restrict__ sem, int S, int C, int nh) {
#ifdef TVAE_DEVICE_OK
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kBT;
  // Q / dO ring of THREE stages: a stage is released by the last MMA that reads it (dQ of the tile) and refilled from
  // L2 -- with two stages that round trip (commit -> producer -> TMA latency -> S~ / dP -> softmax -> dV / dK / dQ) was
  // the pace of the whole kernel: 0.9 of 1.44 ms remained with every MMA, exponential, store and drain switched off.
  // The shared memory for the third stage comes from P, which is single-buffered: the dV MMAs that read it are issued
  // first and signal p_free on their own.
  uint8_t* sQ = sV + kBT;                 // [kQStages]
  uint8_t* sDO = sQ + kQStages * kBT;     // [kQStages]
  uint8_t* sP = sDO + kQStages * kBT;     // 2 chunks (keys 0-63, 64-127)
  uint8_t* sDZ = sP + 2 * kBT;            // [2 buffers] x 2 chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDZ + 4 * kBT);
  uint64_t* kv_full = bars;          // 1
  uint64_t* qdo_full = bars + 1;     // [3]
  uint64_t* qdo_empty = bars + 4;    // [3]
  uint64_t* sdp_full = bars + 7;     // 1
  uint64_t* pds_full = bars + 8;     // 1 (8 warp arrivals)
  uint64_t* mma_done = bars + 9;     // [2] (alternating, so a waiter never lags two phases)
  uint64_t* sdp_free = bars + 11;    // 1 (8 warp arrivals): S~ / dP of the current tile sit in registers
  uint64_t* dq_drained = bars + 12;  // [2]: dQ~ of tile i left its staging area (= the dZ buffer i & 1) and its TMEM buffer
  uint64_t* p_free = bars + 14;      // 1: the dV MMAs of the tile have read P
  uint64_t* dq_staged = bars + 15;   // [2] (4 warp arrivals): dQ~ of tile i sits in its staging area
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nq = (S + 127) / 128;
  // Query-tile order.  sem != nullptr (ordered mode): key-tile CTA j visits the query tiles in the rotated order
  // (j + step) mod nq, and adds its dQ~ contribution to a tile only after the earlier contributions have landed: tile i
  // receives CTA i, i-1, i-2, ... in that order, so dQ is bit-reproducible.  Even and odd steps go to two accumulators
  // (summed by tvae_rope_bwd), each with its own counter per (image, head, query tile): a contribution then waits for
  // the one TWO steps back -- with a single accumulator every step waited for the global completion of the previous
  // reduce-add (~3 us against a 1.9 us tile period: 4.1 instead of 2.7 ms).  At every step the CTAs of an (image, head)
  // touch different tiles.  The host enables this mode only when the nq CTAs of a group fit on
  // the machine several times over (a group must become co-resident for the chain to resolve).
  const int rot = sem != nullptr ? (int)blockIdx.x : 0;
  auto tile_of = [&](int step) {
    int t = step + rot;
    return t >= nq ? t - nq : t;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDQ1);
    mbar_init(kv_full, 1);
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(p_free, 1);
    mbar_init(&dq_staged[0], 4);
    mbar_init(&dq_staged[1], 4);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(sdp_free, 8);
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&dq_drained[0], 1);
    mbar_init(&dq_drained[1], 1);
    fence_mbar_init();
    // K_j / V_j and the first ring stages of Q / dO go out right here -- their barriers exist (this thread made them) and
    // nobody else touches shared memory before the block-wide sync below: the first round trip to L2 / HBM then runs under
    // the TMEM allocation, the sync and the register re-split instead of after them (a fixed cost paid by every CTA).
    mbar_arrive_expect_tx(kv_full, 2 * kBT);
    tma_load_3d(sK, &tmQKV, kv_full, C + h * 64, k0, b);
    tma_load_3d(sV, &tmQKV, kv_full, 2 * C + h * 64, k0, b);
    for (int i = 0; i < nq && i < kQStages; ++i) {
      mbar_arrive_expect_tx(&qdo_full[i], 2 * kBT);
      tma_load_3d(sQ + i * kBT, &tmQKV, &qdo_full[i], h * 64, tile_of(i) * 128, b);
      tma_load_3d(sDO + i * kBT, &tmDO, &qdo_full[i], h * 64, tile_of(i) * 128, b);
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  const uint32_t t_S = tmem_base, t_dP = tmem_base + 128, t_dV = tmem_base + 256, t_dK = tmem_base + 320,
                 t_dQ = tmem_base + 384;  // 2 x 64

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");      // control warpgroup (TMA / MMA issue)
  if (warp == 0) {
    if (lane == 0) {
      // tiles 0 .. kQStages-1 (and K_j, V_j) were requested before the block-wide sync: the ring continues at stage 0, phase 1
      int st = 0;
      uint32_t ph = 1;
      for (int i = kQStages; i < nq; ++i) {
        mbar_wait(&qdo_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kBT);
        tma_load_3d(sQ + st * kBT, &tmQKV, &qdo_full[st], h * 64, tile_of(i) * 128, b);
        tma_load_3d(sDO + st * kBT, &tmDO, &qdo_full[st], h * 64, tile_of(i) * 128, b);
        if (++st == kQStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue: the whole warp walks the loop (uniform control flow, operands in uniform registers), one elected
    // lane issues.  Descriptors are built once and advanced by constant offsets.
    constexpr uint32_t id_kk = umma_idesc_bf16(128, 128, 0, 0);   // S~, dP : both operands K-major (d contiguous)
    constexpr uint32_t id_mm = umma_idesc_bf16(128, 64, 1, 1);    // dV, dK : both MN-major (reduction over queries)
    constexpr uint32_t id_km = umma_idesc_bf16(128, 64, 0, 1);    // dQ     : A = dZ K-major, B = K_j MN-major
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dk_k = umma_desc_kmajor_sw128(smem_base);                  // K-major view of the tile at offset 0
    const uint64_t dk_m = umma_desc_mnmajor_sw128(smem_base, kBT, 1024);      // MN-major view
    constexpr uint32_t oK = 0, oV = kBT, oQ = 2 * kBT, oDO = (2 + kQStages) * kBT, oP = (2 + 2 * kQStages) * kBT,
                       oDZ = (4 + 2 * kQStages) * kBT;
    auto issue_sdp = [&](uint32_t st, uint32_t ph) {     // S~ / dP of the tile in ring stage st
      mbar_wait(&qdo_full[st], ph);
      tc_fence_after();
      const uint64_t dq = umma_desc_advance(dk_k, oQ + st * kBT), ddo = umma_desc_advance(dk_k, oDO + st * kBT);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_elect(t_S, umma_desc_advance(dq, k * 32), umma_desc_advance(dk_k, oK + k * 32), id_kk, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_elect(t_dP, umma_desc_advance(ddo, k * 32), umma_desc_advance(dk_k, oV + k * 32), id_kk, k != 0);
      umma_commit_elect(sdp_full);
    };
    mbar_wait(kv_full, 0);
    issue_sdp(0, 0);
    uint32_t st = 0, ph = 0;                              // ring position of tile i
    for (int i = 0; i < nq; ++i) {
      const uint32_t buf = i & 1;
      uint32_t st_n = st + 1, ph_n = ph;                  // ... of tile i + 1
      if (st_n == kQStages) {
        st_n = 0;
        ph_n ^= 1;
      }
      // S~ / dP of tile i+1 go to the tensor pipe as soon as the softmax warps hold tile i in registers, i.e.
      // they overlap the exponentiation of tile i (the first version issued them after dV / dK / dQ of tile i, so
      // tensor work and softmax strictly alternated)
      if (i + 1 < nq) {
        mbar_wait(sdp_free, i & 1);
        tc_fence_after();
        issue_sdp(st_n, ph_n);
      }
      mbar_wait(pds_full, i & 1);
      tc_fence_after();
      const uint64_t mq = umma_desc_advance(dk_m, oQ + st * kBT), mdo = umma_desc_advance(dk_m, oDO + st * kBT);
      const uint64_t mp = umma_desc_advance(dk_m, oP), mdz = umma_desc_advance(dk_m, oDZ + buf * 2 * kBT);
      const uint64_t kdz = umma_desc_advance(dk_k, oDZ + buf * 2 * kBT);
      const uint32_t acc = i != 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)     // dV: 16 queries per MMA; first, so that P can be rewritten early
        umma_f16_elect(t_dV, umma_desc_advance(mp, k * 2048), umma_desc_advance(mdo, k * 2048), id_mm, k ? 1u : acc);
      umma_commit_elect(p_free);
#pragma unroll
      for (int k = 0; k < 8; ++k)     // dK
        umma_f16_elect(t_dK, umma_desc_advance(mdz, k * 2048), umma_desc_advance(mq, k * 2048), id_mm, k ? 1u : acc);
#pragma unroll
      for (int k = 0; k < 8; ++k)     // dQ: 16 keys per MMA
        umma_f16_elect(t_dQ + buf * 64, umma_desc_advance(kdz, (k >> 2) * kBT + (k & 3) * 32),
                       umma_desc_advance(dk_m, oK + k * 2048), id_km, k != 0);
      umma_commit_elect(&qdo_empty[st]);
      umma_commit_elect(&mma_done[buf]);
      st = st_n;
      ph = ph_n;
    }
  } else if (lane == 0) {
    // ---- dQ issuers: warp 2 takes the even steps, warp 3 the odd ones (= one staging buffer and, in ordered mode, one
    // accumulator each).  An issuer may sit in the global-completion wait of its reduce-add for up to two tile periods
    // without holding anyone up (with the drain warps issuing, that wait was on the path of every tile: 3.6 vs 2.7 ms).
    // (Four issuers -- lanes 0 / 16 of both warps, steps mod 4, so that an ordered step has four tile periods for its
    // completion wait -- were measured slower: 3.25 vs 3.09 ms ordered, 2.85 vs 2.76 unordered at S = 4096.)
    const int par = warp - 2;
    for (int i = par; i < nq; i += 2) {
      mbar_wait(&dq_staged[par], (i >> 1) & 1);
      const int qt = tile_of(i);
      uint8_t* stage = sDZ + par * 2 * kBT;
      int* my_sem = sem != nullptr ? sem + (((size_t)par * gridDim.z + b) * nh + h) * nq + qt : nullptr;
      if (my_sem != nullptr) {                 // the earlier contributions of this parity to the tile have landed
        uint32_t spins = 0;
        uint64_t t0 = 0;
        while (ld_acquire_gpu(my_sem) != (i >> 1)) {
          if ((++spins & 0x3FF) == 0) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > TVAE_WAIT_TIMEOUT_NS) {
              printf("tvae: attention backward dQ order wait timeout block=(%d,%d,%d) step=%d\n", blockIdx.x, blockIdx.y,
                     blockIdx.z, i);
              __trap();
            }
          }
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");
      }
      // ordered mode: the contribution of step 0 / 1 is the FIRST one its accumulator receives for this tile (every tile gets
      // one of each) -- a plain tensor store, so the accumulators need no zero fill (2 x B x S x C fp32 per launch)
      if (my_sem != nullptr && i < 2) {
#pragma unroll
        for (int half = 0; half < 2; ++half)
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(par ? &tmDQ1 : &tmDQ)),
                       "r"(smem_u32(stage + half * kBT)), "r"(h * 64 + half * 32), "r"(qt * 128), "r"(b)
                       : "memory");
      } else {
#pragma unroll
        for (int half = 0; half < 2; ++half)
          asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(my_sem != nullptr && par ? &tmDQ1 : &tmDQ)),
                       "r"(smem_u32(stage + half * kBT)), "r"(h * 64 + half * 32), "r"(qt * 128), "r"(b)
                       : "memory");
      }
      tma_store_commit();
      tma_store_wait_read<0>();
      mbar_arrive(&dq_drained[par]);
      if (my_sem != nullptr) {                 // the adds are complete and visible before the next contributor starts
        tma_store_wait<0>();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        st_release_gpu(my_sem, (i >> 1) + 1);
      }
    }
    tma_store_wait<0>();                       // all reduce-adds of this issuer have landed
  }
  } else if (warp >= 12) {
    // ---- drain warpgroup: dQ~_i (128 x 64 fp32) TMEM -> swizzled smem; warps 2 / 3 then issue two bulk tensor reduce-adds
    // into the fp32 accumulator.  The staging area is the dZ buffer of tile i, which the tensor core has finished
    // reading when mma_done(i) fires; the softmax warps take it back (tile i + 2) on dq_drained.  (First version: per-thread
    // red.global.add.v4.f32 -- 6.4 GB of scattered 16-byte atomics per launch.  Second version: the softmax warps
    // staged and waited for the bulk read themselves -- 41 % of their stall samples sat in that wait and its barriers.)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    for (int i = 0; i < nq; ++i) {
      mbar_wait(&mma_done[i & 1], (i >> 1) & 1);
      tc_fence_after();
      uint8_t* stage = sDZ + (i & 1) * 2 * kBT;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(t_dQ + (i & 1) * 64 + lane_off + half * 32, v);
        tmem_ld_wait();
        const uint32_t row = smem_u32(stage + half * kBT + r * 128);
#pragma unroll
        for (int g = 0; g < 8; ++g) st_shared_v4(row + ((g ^ (r & 7)) << 4), v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_staged[i & 1]);      // 4 warp arrivals: the tile is staged, its TMEM buffer is free
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");    // softmax warpgroups: 2 x (32 S~ + 32 dP) values per thread
    // Eight softmax warps: warps w and w + 4 share TMEM lane quarter (w & 3) -- i.e. the same 32 query rows -- and
    // split the 128 key columns in halves.  Nothing in the backward softmax reduces along a row (lse and delta come
    // from the forward pass), so the halves are independent; with one warp per scheduler (the first version) every
    // TMEM-load, MUFU and shared-store latency was exposed and the kernel ran at 28 % of the tensor peak.
    const int qd = warp & 3;
    const int half = (warp - 4) >> 2;          // key columns [64 * half, 64 * half + 64)
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const size_t stat_base = ((size_t)b * nh + h) * S;
    const bool full_tile = k0 + 128 <= S;                            // no key masking needed

    float l2_next, dl_next;
    {
      const int q0 = tile_of(0) * 128 + r;
      l2_next = (q0 < S) ? __ldg(lse + stat_base + q0) : INFINITY;
      dl_next = (q0 < S) ? __ldg(delta + stat_base + q0) : 0.0f;
    }
    for (int i = 0; i < nq; ++i) {
      const float2 nl2 = make_float2(-l2_next, -l2_next), ndl = make_float2(-dl_next, -dl_next);
      {                                                              // statistics of the next tile: off the critical path
        const int qn = (i + 1 < nq ? tile_of(i + 1) : nq) * 128 + r;
        l2_next = (qn < S) ? __ldg(lse + stat_base + qn) : INFINITY;
        dl_next = (qn < S) ? __ldg(delta + stat_base + qn) : 0.0f;
      }
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      uint32_t sa[32], pa[32], sb[32], pb[32];
      tmem_ld32(t_S + lane_off + half * 64, sa);
      tmem_ld32(t_dP + lane_off + half * 64, pa);
      tmem_ld32(t_S + lane_off + half * 64 + 32, sb);
      tmem_ld32(t_dP + lane_off + half * 64 + 32, pb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      // S~ / dP of this tile sit in registers: the accumulators may be overwritten.  This arrive must stay AHEAD of the
      // arithmetic (a version that released after the first 32 columns had it sunk below all 64 exponentials by ptxas:
      // 38 % of the softmax warps' samples then waited for S~ / dP (i + 1)); the wait loop below pins it.
      if (lane == 0) mbar_arrive(sdp_free);
      if (i >= 1) mbar_wait(p_free, (i - 1) & 1);                     // P: read by the dV MMAs of tile i-1
      if (i >= 2) mbar_wait(&dq_drained[i & 1], ((i - 2) >> 1) & 1);  // dZ buffer i&1: MMA(i-2) and the dQ drain are done
      const uint32_t prow = smem_u32(sP + half * kBT + r * 128);
      const uint32_t zrow = smem_u32(sDZ + ((i & 1) * 2 + half) * kBT + r * 128);
      // masked = the key tile reaches beyond S (last tile of a ragged sequence): the test is hoisted out of the inner
      // loop (ncu: the predicated ISETP / VIADD / FSEL of the mask were 200 of the 536 instructions per tile)
      auto chunk = [&](const uint32_t(&sv)[32], const uint32_t(&pv)[32], int c, auto masked) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {                                // 8 columns -> one 16-byte store of P and of dZ
          uint32_t pk[4], zk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = g * 8 + 2 * k;
            float2 e = __fadd2_rn(make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), nl2);
            if (POLY > 0 && ((c * 16 + g * 4 + k) % (POLY > 0 ? POLY : 1)) == POLY - 1) {
              e = exp2_fma2(e);
            } else {
              e.x = exp2f(e.x);
              e.y = exp2f(e.y);
            }
            if (decltype(masked)::value) {
              if (k0 + half * 64 + c * 32 + j >= S) e.x = 0.0f;
              if (k0 + half * 64 + c * 32 + j + 1 >= S) e.y = 0.0f;
            }
            const float2 d = __fmul2_rn(e, __fadd2_rn(make_float2(__uint_as_float(pv[j]), __uint_as_float(pv[j + 1])), ndl));
            pk[k] = pack_bf16(e.x, e.y);
            zk[k] = pack_bf16(d.x, d.y);
          }
          const uint32_t off = ((c * 4 + g) ^ (r & 7)) << 4;
          st_shared_v4(prow + off, pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(zrow + off, zk[0], zk[1], zk[2], zk[3]);
        }
      };
      if (full_tile) {
        chunk(sa, pa, 0, std::false_type{});
        chunk(sb, pb, 1, std::false_type{});
      } else {
        chunk(sa, pa, 0, std::true_type{});
        chunk(sb, pb, 1, std::true_type{});
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    mbar_wait(&mma_done[(nq - 1) & 1], ((nq - 1) >> 1) & 1);
    tc_fence_after();
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

// 2: ordered (bit-reproducible) dQ accumulation into an even-step and an odd-step accumulator -- when a group of nq key-tile
// CTAs fits on the machine at least twice, so that the wait chain of a group always resolves; 1: unordered reduce-adds
// (TVAE_ATTN_BWD_ORDERED=0 forces it: the A/B switch)
int attn_bwd_dq_slices(int S) {
  static const bool ordered_env = !(getenv("TVAE_ATTN_BWD_ORDERED") && atoi(getenv("TVAE_ATTN_BWD_ORDERED")) == 0);
  const int nq = (S + 127) / 128;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  return ordered_env && nq > 1 && 2 * nq <= sms ? 2 : 1;
}

int attn_bwd_run(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq_acc, void* dqkv, int B,
                 int S, int C, int dq_slices, cudaStream_t stream) {
  TVAE_REQUIRE(dq_slices == 1 || dq_slices == 2, "attention_bwd: dq_slices must be 1 or 2");
  TVAE_REQUIRE(C % 64 == 0, "attention_bwd: C=%d must be a multiple of 64", C);
  const int nh = C / 64;
  CUtensorMap mQKV, mDO;
  int rc;
  if ((rc = make_tmap_3d(&mQKV, qkv, 3 * (uint64_t)C, S, B, 3 * (uint64_t)C, (uint64_t)S * 3 * C, 128))) return rc;
  if ((rc = make_tmap_3d(&mDO, dout, C, S, B, C, (uint64_t)S * C, 128))) return rc;
  CUtensorMap mDQ, mDQ1;
  if ((rc = make_tmap_3d_f32(&mDQ, dq_acc, C, S, B, C, (uint64_t)S * C, 128))) return rc;
  if ((rc = make_tmap_3d_f32(&mDQ1, dq_acc + (dq_slices == 2 ? (size_t)B * S * C : 0), C, S, B, C, (uint64_t)S * C, 128))) return rc;
  // TVAE_ATTN_BWD_POLY = 0 / 2 / 3 / 4: A/B switch for the share of exponentials on the FMA pipe (default 0: measured slower with any share -- the softmax warps are not MUFU-bound here)
  static const int poly = getenv("TVAE_ATTN_BWD_POLY") ? atoi(getenv("TVAE_ATTN_BWD_POLY")) : 0;
  auto kern = poly == 0 ? attn_bwd_kernel<0> : poly == 2 ? attn_bwd_kernel<2> : poly == 3 ? attn_bwd_kernel<3> : attn_bwd_kernel<4>;
  static bool configured = false;
  if (!configured) {
    for (auto k : {attn_bwd_kernel<0>, attn_bwd_kernel<2>, attn_bwd_kernel<3>, attn_bwd_kernel<4>})
      TVAE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    configured = true;
  }
  // unordered mode: the reduce-adds need a zeroed accumulator; ordered mode: the first contribution of each parity stores
  if (dq_slices == 1) TVAE_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * S * C * sizeof(float), stream));
  dim3 grid((S + 127) / 128, nh, B);
  int* sem = nullptr;
  const int nq = (int)grid.x;
  if (dq_slices == 2) {
    const size_t bytes = (size_t)2 * B * nh * nq * sizeof(int);
    if ((rc = scratch_workspace(bytes, reinterpret_cast<void**>(&sem)))) return rc;
    TVAE_CHECK_CUDA(cudaMemsetAsync(sem, 0, bytes, stream));
  }
  kern<<<grid, kBwdThreads, kBwdSmem, stream>>>(mQKV, mDO, mDQ, mDQ1, lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv), sem, S,
                                                C, nh);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
