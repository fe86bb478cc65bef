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
#include <cstdlib>

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

// -------------------------------------------------------------------------------------------------
// Forward kernel (sixth version): S held in registers, O resident in TMEM with lazy rescaling, double-buffered P
// -------------------------------------------------------------------------------------------------
// One work item per (256 queries = two 128-row Q tiles, head, image), 8 softmax warps, one query row per thread; a
// persistent CTA walks its items without draining the pipeline in between (see the kernel).  Per K/V block a softmax thread
//   * reads its 128 scores from TMEM ONCE into registers and frees the S buffer at once, so S_t(j+1) = Q_t K_{j+1}^T
//     is on the tensor pipe while block j is still being exponentiated (single S buffer per tile, 128 columns);
//   * keeps O_t in TMEM (64 columns per tile): P_t(j) V_j accumulates there directly (tcgen05.mma accumulate flag),
//     and O is only touched by the softmax warps when the running row maximum grew by more than 2^8 since the
//     reference maximum was last fixed ("lazy rescale": P = exp2(S - m_ref) stays <= 256, exact after the final
//     division by l, which is accumulated against the same m_ref);
//   * writes P into one of two smem buffers per tile, so block j never waits for the P V product of block j-1
//     (ncu on the single-buffer version: 24 % of the softmax warps' samples sat on that wait);
//   * uses packed f32x2 adds (FADD2) for the subtraction of the maximum and the row sum, FMNMX3 for the maximum.
// K and V travel through separate 2-stage rings: K_j is released after S(j) has been issued, V_j after P(j) V_j.
// Per block and thread this is ~440 issue slots instead of ~670, of which 128 are MUFU.EX2 -- the unit that bounds
// head_dim-64 attention (16 exp2 / clk / SM against 8192 MMA flop / clk / SM).
#ifndef TVAE_ATT_TOKEN_PASS
#define TVAE_ATT_TOKEN_PASS 32
#endif
constexpr int kLazyLog2 = 8;
constexpr int kKvStages6 = 2;

// Optional phase trace (build with -DTVAE_ATT_TRACE, read with tvae_debug_att_trace): SM-clock stamps of CTA (0,0,0),
// lane 0 of the first warp of each softmax warpgroup, six stamps per K/V block.
#ifdef TVAE_ATT_TRACE
__device__ long long g_att_trace[2 * 64 * 8];
#define ATT_STAMP(slot)                                                                                     \
  if (trace_on && j < 64) g_att_trace[(t * 64 + j) * 8 + (slot)] = clock64();
#else
#define ATT_STAMP(slot)
#endif

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 2^x for two lanes on the FMA pipe (no MUFU): round-to-nearest split x = n + r with the 1.5*2^23 trick, degree-3
// minimax polynomial for 2^r on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P), exponent
// added into the float bits.  x is clamped at -126 so the exponent arithmetic cannot wrap (-inf -> 1.2e-38).
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

// POLY: every POLY-th pair of scores is exponentiated on the FMA pipe instead of MUFU (0 = never).
template <int POLY>
__global__ void __launch_bounds__(kAttThreads, 1)
attn_fwd6_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                 float* __restrict__ lse, int S, int C, int nh, int nqx, int total_items) {
#ifdef TVAE_DEVICE_OK
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // [2 tiles]
  uint8_t* sK = sQ + 2 * kTileBytes;                 // [2 stages]
  uint8_t* sV = sK + kKvStages6 * kTileBytes;        // [2 stages]
  uint8_t* sP = sV + kKvStages6 * kTileBytes;        // [2 tiles][2 buffers] x 2 chunks (keys 0-63, 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 8 * kTileBytes);
  uint64_t* q_full = bars;                           // 1
  uint64_t* k_full = bars + 1;                       // [2]
  uint64_t* k_empty = k_full + 2;                    // [2]
  uint64_t* v_full = k_empty + 2;                    // [2]
  uint64_t* v_empty = v_full + 2;                    // [2]
  uint64_t* s_full = v_empty + 2;                    // [2 tiles]
  uint64_t* s_free = s_full + 2;                     // [2 tiles]
  uint64_t* p_full = s_free + 2;                     // [2 tiles][2 buffers]
  uint64_t* o_full = p_full + 4;                     // [2 tiles][2 buffers]
  uint64_t* q_empty = o_full + 4;                    // 1 (2 commits): the item's last S = Q K^T has been issued by both tiles
  uint64_t* o_drained = q_empty + 1;                 // [2 tiles] (4 warp arrivals): O_t of the finished item left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_drained + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nblk = (S + 127) / 128;
  // Persistent CTA: work item w = (256-query tile pair, head, image), query tiles fastest (the CTAs that run side by side
  // share K / V of one (head, image) in L2); this CTA takes items blockIdx.x, blockIdx.x + gridDim.x, ...  All barriers
  // and buffer indices run on the GLOBAL key-block counter g = item_index * nblk + j, so the pipeline never drains
  // between items: K / V of the next item stream in behind the current item's last blocks, its Q tiles arrive as soon as
  // the last S = Q K^T of the current item has been issued, and its first S is on the tensor pipe while the softmax warps
  // normalise and store O.  One CTA per item (gridDim.x = total_items, the non-persistent launch) paid ~6 us of launch,
  // TMEM allocation, first round trip and pipeline fill per item: 12 % of the kernel at S = 4096, 65 % at S = 256.
  const int first_item = blockIdx.x, item_stride = gridDim.x;
  const int n_items = first_item < total_items ? (total_items - first_item + item_stride - 1) / item_stride : 0;
  auto item_q0 = [&](int w) { return (w % nqx) * 256; };
  auto item_h = [&](int w) { return (w / nqx) % nh; };
  auto item_b = [&](int w) { return w / (nqx * nh); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages6; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);     // one tcgen05.commit per Q tile's MMA thread
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      for (int u = 0; u < 2; ++u) {
        mbar_init(&p_full[t * 2 + u], 4);
        mbar_init(&o_full[t * 2 + u], 1);
      }
      mbar_init(&o_drained[t], 4);
    }
    mbar_init(q_empty, 2);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  // TMEM columns: tile t: S_t at t*256 (128 columns), O_t at t*256 + 128 (64 columns)

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");     // control warpgroup: TMA / MMA issue only (72 here made ptxas spill four loop-invariant scalars of the softmax loop)
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < n_items; ++it) {
          const int w = first_item + it * item_stride;
          const int q0 = item_q0(w), h = item_h(w), b = item_b(w);
          if (it > 0) mbar_wait(q_empty, (it - 1) & 1);      // both tiles have issued the previous item's last Q K^T
          mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
          att_tma_load_3d(sQ, &tmQKV, q_full, h * 64, q0, b);
          att_tma_load_3d(sQ + kTileBytes, &tmQKV, q_full, h * 64, q0 + 128, b);
          for (int j = 0; j < nblk; ++j) {
            const int g = it * nblk + j;
            const int stage = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            mbar_wait(&k_empty[stage], ph ^ 1);
            mbar_arrive_expect_tx(&k_full[stage], kTileBytes);
            att_tma_load_3d(sK + stage * kTileBytes, &tmQKV, &k_full[stage], C + h * 64, j * 128, b);
            mbar_wait(&v_empty[stage], ph ^ 1);
            mbar_arrive_expect_tx(&v_full[stage], kTileBytes);
            att_tma_load_3d(sV + stage * kTileBytes, &tmQKV, &v_full[stage], 2 * C + h * 64, j * 128, b);
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // one MMA-issuing thread per Q tile, so a slow tile never holds back the other tile's S / P V products
      {   // whole warp, elected lane issues (common.cuh: warp-uniform issue)
        const int t = warp - 1;
        constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
        const uint32_t q_base = smem_u32(sQ + t * kTileBytes);
        auto issue_qk = [&](int j) {          // j: GLOBAL key-block index
          const uint32_t k_base = smem_u32(sK + (j & 1) * kTileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_elect(tmem_base + t * 256, umma_desc_kmajor_sw128(q_base + k * 32),
                     umma_desc_kmajor_sw128(k_base + k * 32), idesc_qk, k != 0);
          umma_commit_elect(&s_full[t]);
          umma_commit_elect(&k_empty[j & 1]);                // second arrival (other tile) releases the K stage
        };
        auto issue_pv = [&](int j, int first) {   // first: first key block of its item -> O_t is overwritten
          const uint32_t p_base = smem_u32(sP + (t * 2 + (j & 1)) * 2 * kTileBytes);
          const uint32_t v_base = smem_u32(sV + (j & 1) * kTileBytes);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_elect(tmem_base + t * 256 + 128, umma_desc_kmajor_sw128(p_base + (k >> 2) * kTileBytes + (k & 3) * 32),
                     umma_desc_mnmajor_sw128(v_base + k * 2048, 1024, 1024), idesc_pv, (first == 0) || (k != 0));
          umma_commit_elect(&o_full[t * 2 + (j & 1)]);
          umma_commit_elect(&v_empty[j & 1]);
        };
        const int total_blocks = n_items * nblk;
        if (total_blocks > 0) {
          mbar_wait(q_full, 0);
          mbar_wait(&k_full[0], 0);
          tc_fence_after();
          issue_qk(0);
          if (nblk == 1) umma_commit_elect(q_empty);
        }
        for (int g = 0, j = 0, it = 0; g < total_blocks; ++g) {       // flat loop over the key blocks of all items
          if (g + 1 < total_blocks) {
            const bool new_item = (j + 1 == nblk);       // block g + 1 opens the next item: its Q tiles must have landed
            if (new_item) mbar_wait(q_full, (it + 1) & 1);
            mbar_wait(&k_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
            mbar_wait(&s_free[t], g & 1);            // the softmax warps hold S_t(g) in registers
            tc_fence_after();
            issue_qk(g + 1);
            if ((new_item ? 0 : j + 1) == nblk - 1) umma_commit_elect(q_empty);   // last Q K^T of its item: Q may be replaced
          }
          if (j == 0 && it > 0) mbar_wait(&o_drained[t], (it - 1) & 1);   // O_t of the previous item has been read out
          mbar_wait(&v_full[g & 1], (g >> 1) & 1);
          mbar_wait(&p_full[t * 2 + (g & 1)], (g >> 1) & 1);
          tc_fence_after();
          issue_pv(g, j == 0);
          if (++j == nblk) {
            j = 0;
            ++it;
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");    // softmax warpgroups hold 128 scores per thread
    const int t = (warp - 4) >> 2;       // Q tile of this warpgroup
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + t * 256;
    const uint32_t t_o = t_s + 128;
    uint8_t* sPt = sP + t * 4 * kTileBytes;          // this tile's two P buffers
    const uint32_t sPt_row = smem_u32(sPt) + r * 128;
#ifdef TVAE_ATT_TRACE
    const bool trace_on = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && qd == 0 && lane == 0;
#endif
    if (t == 1) asm volatile("bar.arrive 3, 256;" ::: "memory");   // tile 0 exponentiates first

    int g = 0;                                       // global key-block index: barrier phases and buffer parities
    for (int it = 0; it < n_items; ++it) {
    float m_ref = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < nblk; ++j, ++g) {
      ATT_STAMP(0)
      // the O tile of the previous item was staged in P buffer (g0 - 1) & 1, which block j = 1 (j = 0 when an item is a
      // single block) is about to overwrite: its TMA store must have read it (issued a whole block ago)
      if (it > 0 && j == (nblk >= 2 ? 1 : 0)) {
        if (qd == 0 && lane == 0) tma_store_wait_read<0>();
        if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      mbar_wait(&s_full[t], g & 1);
      tc_fence_after();
      ATT_STAMP(1)
      uint32_t v[128];
      {
        uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        uint32_t(&v2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[64]);
        uint32_t(&v3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[96]);
        tmem_ld32(t_s, v0);
        tmem_ld32(t_s + 32, v1);
        tmem_ld32(t_s + 64, v2);
        tmem_ld32(t_s + 96, v3);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);      // S buffer may receive S_t(j+1)
      ATT_STAMP(2)

      const int key0 = j * 128;
      if (key0 + 128 > S) {                         // only the last block can reach past the sequence end
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (key0 + i >= S) v[i] = 0xff800000u;    // -inf
      }
      float mx[8];                                  // eight independent FMNMX3 chains (a two-chain version spent
#pragma unroll                                      // ~400 clocks per block on dependent-issue latency)
      for (int c = 0; c < 8; ++c) mx[c] = fmaxf(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
#pragma unroll
      for (int i = 16; i < 128; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          mx[c] = fmaxf(mx[c], fmaxf(__uint_as_float(v[i + 2 * c]), __uint_as_float(v[i + 2 * c + 1])));
      }
      const float m_blk = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool grow = m_blk > m_ref + static_cast<float>(kLazyLog2);
        if (__any_sync(0xffffffffu, grow)) {
          // rare: bring O_t and l to the new reference maximum (the previous P V must have landed first)
          mbar_wait(&o_full[t * 2 + ((g - 1) & 1)], ((g - 1) >> 1) & 1);
          tc_fence_after();
          const float m_new = grow ? m_blk : m_ref;
          const float sc = exp2f(m_ref - m_new);
          l_run *= sc;
          m_ref = m_new;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
            tmem_st32(t_o + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P buffer (j & 1) was last read by the P V product of block j-2
      ATT_STAMP(3)
      if (g >= 2) mbar_wait(&o_full[t * 2 + (g & 1)], ((g - 2) >> 1) & 1);
      ATT_STAMP(4)
      // MUFU token: the exp2 phases of the two warpgroups strictly alternate, so one tile exponentiates (MUFU-bound)
      // while the other loads / reduces / synchronises.  Without it the tiles ran in lockstep (trace: 1300 clocks
      // with the MUFU idle, then 2400 clocks with both tiles fighting for it).
      if (t == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
      else asm volatile("bar.sync 4, 256;" ::: "memory");
      const uint32_t prow = sPt_row + (g & 1) * 2 * kTileBytes;
      const float2 nm2 = make_float2(-m_ref, -m_ref);
      float2 la = make_float2(0.0f, 0.0f), lb = make_float2(0.0f, 0.0f);
      // Software-pipelined by hand: pair i is exponentiated kExpAhead pairs before it is summed / packed / stored, so
      // the in-order warp never sits on the MUFU result latency (the straightforward loop left the consumer 2 MUFUs
      // behind its producer and the exp phase took 1600 clocks instead of the 1024 the MUFU needs).
      constexpr int kExpAhead = 6;
      constexpr int kTokenPass = TVAE_ATT_TOKEN_PASS;   // pair index at which the token is handed on (64 = end of phase)
#pragma unroll
      for (int i = 0; i < 64 + kExpAhead; ++i) {
        if (i < 64) {
          float2 x = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), nm2);
          if (POLY > 0 && POLY < 100 && (i % (POLY > 0 ? POLY : 1)) == POLY - 1) {
            x = exp2_fma2(x);
          } else if (POLY == 102) {                  // trace experiment: no MUFU at all
            x = __ffma2_rn(x, make_float2(1e-3f, 1e-3f), make_float2(1.0f, 1.0f));
          } else {
            x.x = exp2f(x.x);
            x.y = exp2f(x.y);
          }
          v[2 * i] = __float_as_uint(x.x);
          v[2 * i + 1] = __float_as_uint(x.y);
        }
        if (i == kTokenPass) {
          // pass the MUFU token on before this exp phase is over: one warp per scheduler cannot saturate the MUFU pipe
          // (128 back-to-back EX2 from a single warp take ~1500 clocks, not 1024), so the second half of this phase
          // overlaps the first half of the other tile's -- staggered by half a phase, never in lockstep
          if (t == 0) asm volatile("bar.arrive 4, 256;" ::: "memory");
          else if (j + 1 < nblk || it + 1 < n_items) asm volatile("bar.arrive 3, 256;" ::: "memory");
        }
        const int d = i - kExpAhead;
        if (d >= 0) {
          const float2 e = make_float2(__uint_as_float(v[2 * d]), __uint_as_float(v[2 * d + 1]));
          if (d & 1) lb = __fadd2_rn(lb, e);
          else la = __fadd2_rn(la, e);
          v[d] = pack_bf16(e.x, e.y);              // packed P reuses the low registers of the score array
          if ((d & 3) == 3 && POLY != 101) {         // (101: trace experiment without the P stores)
            const int c = d >> 4, g = (d >> 2) & 3;   // 32-column chunk, 8-column group
            const uint32_t row = prow + (c >> 1) * kTileBytes;
            st_shared_v4(row + ((((c & 1) * 4 + g) ^ (r & 7)) << 4), v[d - 3], v[d - 2], v[d - 1], v[d]);
          }
        }
      }
      l_run += (la.x + la.y) + (lb.x + lb.y);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + (g & 1)]);
      ATT_STAMP(5)
    }
    // O_t is complete once the last P V has landed (MMAs complete in order)
    const int gl = g - 1;
    const int w = first_item + it * item_stride;
    const int q0 = item_q0(w), h = item_h(w), b = item_b(w);
    const bool has_next = it + 1 < n_items;
    mbar_wait(&o_full[t * 2 + (gl & 1)], (gl >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    // stage O (bf16) in the P buffer this item used LAST (all P V products are done): the next item writes the other one first
    uint8_t* sO = sPt + (gl & 1) * 2 * kTileBytes;
    const uint32_t orow = sPt_row + (gl & 1) * 2 * kTileBytes;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        st_shared_v4(orow + (((c * 4 + g) ^ (r & 7)) << 4),
                     pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l),
                     pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l),
                     pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l),
                     pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l));
      }
    }
    // O_t has left TMEM: the first P V of the next item may overwrite it
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&o_drained[t]);
    const int qrow = q0 + t * 128 + r;
    if (lse != nullptr && qrow < S) lse[((size_t)b * nh + h) * S + qrow] = m_ref + log2f(l_run);
    fence_proxy_async_smem();
    if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
    if (qd == 0 && lane == 0 && q0 + t * 128 < S) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(&tmO)),
                   "r"(smem_u32(sO)), "r"(h * 64), "r"(q0 + t * 128), "r"(b)
                   : "memory");
      tma_store_commit();
      if (!has_next) tma_store_wait<0>();          // (otherwise waited for where the staging buffer is written again)
    }
    }   // items
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
  // persistent grid: one CTA per SM walks the (query-tile pair, head, image) items; TVAE_ATTN_PERSIST=0: one CTA per item
  const int nqx = (S + 255) / 256;
  const long long total_ll = (long long)nqx * nh * B;
  TVAE_REQUIRE(total_ll < (1LL << 30), "attention: too many work items");
  const int total_items = (int)total_ll;
  static const bool persist = !(getenv("TVAE_ATTN_PERSIST") && atoi(getenv("TVAE_ATTN_PERSIST")) == 0);
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  dim3 grid(persist && total_items > sms ? sms : total_items);
  // TVAE_ATTN_POLY: every POLY-th pair of scores is exponentiated on the FMA pipe (0 = all on MUFU; tuning switch)
  static const int poly = getenv("TVAE_ATTN_POLY") ? atoi(getenv("TVAE_ATTN_POLY")) : 4;
  static bool configured6 = false;
  if (!configured6) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd6_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd6_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd6_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd6_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    configured6 = true;
  }
#ifdef TVAE_ATT_TRACE
  if (poly == 101 || poly == 102) {
    auto kern = poly == 101 ? attn_fwd6_kernel<101> : attn_fwd6_kernel<102>;
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    kern<<<grid, kAttThreads, kAttSmem, stream>>>(mQKV, mO, lse, S, C, nh, nqx, total_items);
    return 0;
  }
#endif
  if (poly == 2) attn_fwd6_kernel<2><<<grid, kAttThreads, kAttSmem, stream>>>(mQKV, mO, lse, S, C, nh, nqx, total_items);
  else if (poly == 3) attn_fwd6_kernel<3><<<grid, kAttThreads, kAttSmem, stream>>>(mQKV, mO, lse, S, C, nh, nqx, total_items);
  else if (poly == 4) attn_fwd6_kernel<4><<<grid, kAttThreads, kAttSmem, stream>>>(mQKV, mO, lse, S, C, nh, nqx, total_items);
  else attn_fwd6_kernel<0><<<grid, kAttThreads, kAttSmem, stream>>>(mQKV, mO, lse, S, C, nh, nqx, total_items);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#ifdef TVAE_ATT_TRACE
extern "C" int tvae_debug_att_trace(long long* dst) {
  return cudaMemcpyFromSymbol(dst, g_att_trace, sizeof(g_att_trace)) == cudaSuccess ? 0 : -1;
}
#endif

}  // namespace tvae
