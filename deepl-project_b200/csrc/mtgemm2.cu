// CTA-pair (cta_group::2) variant of the fused multi-tap implicit GEMM.
//
// Two CTAs of a 2-CTA cluster (one TPC) work on one 256 x BLOCK_N output tile: CTA r owns rows [128 r, 128 r + 128)
// (its own pixel tile, its own TMEM accumulator, its own epilogue / residual / store), but the weight tile B is
// SPLIT between them -- each CTA TMA-loads only BLOCK_N / 2 rows of it and the leader's
// tcgen05.mma.cta_group::2 (M = 256) reads both halves across the pair.  Per 64-wide K block a CTA therefore pulls
// 16 KiB (A) + BLOCK_N * 64 B (half of B) from L2 instead of 16 KiB + BLOCK_N * 128 B.  ncu on the 1-CTA kernel
// showed every large GEMM pinned at ~15 TB/s of L2->SM traffic (0.023-0.026 B/MAC) and, for N = 192, at the 128 B/clk
// shared-memory read port; halving B relieves both.
//
// Protocol (per pipeline stage):
//   both producers: wait own empty[s]; TMA A + B-half into OWN smem with .cta_group::2, completing on the LEADER's
//                   full[s]; the leader alone arms full[s] with the byte count of both CTAs;
//   leader MMA    : wait full[s]; 4 x tcgen05.mma.cta_group::2; tcgen05.commit ... multicast -> empty[s] of BOTH CTAs;
//                   after the last K block: commit multicast -> tmem_full[acc] of both CTAs;
//   epilogues     : each CTA drains its own accumulator half; all 16 epilogue warps arrive (remotely for the peer) on
//                   the leader's tmem_empty[acc].
// A pair whose second pixel tile lies beyond the tensor (odd tile count) runs it as a phantom: its loads are zero-filled
// by TMA, its stores clipped.
#include "mtgemm_common.cuh"
#include "cluster2.cuh"

namespace tvae {

template <int BLOCK_N, int EPI>
struct Mt2Cfg {
  static constexpr int kBHalfBytes = (BLOCK_N / 2) * kBlockK * 2;
  static constexpr int kOutBytes = kBlockM * BLOCK_N * 2;
  // kEpiResMulGeluGrad stages TWO input tiles per output tile (the gradient to add and the saved pre-activation z),
  // both through TMA.  (z used to be read from global memory by the epilogue threads, 16 bytes per thread and row: every
  // warp load touched 32 cache lines and the L1 tag stage, not HBM, set the pace -- 265 TFLOP/s / 2.1 TB/s on the
  // K = 384 input-gradient GEMMs of the ConvFFN.)
  static constexpr bool kStageZ = EPI == kEpiResMulGeluGrad;
  // the dual-output epilogues stage two OUTPUT tiles (pre-activation and activation)
  static constexpr bool kDual = epi_is_dual<EPI>();
  // kEpiGnBwd stages the GroupNorm input x next to the output tile (second INPUT tile, no residual in the first).  The
  // second tile costs the halo pipeline a stage (3 -> 2: the GEMM itself drops from 1670 to 1250 TFLOP/s at N = 192), still
  // cheaper than the stand-alone reduce pass; reading x from global memory in the epilogue instead (behind an L2 tensor
  // prefetch, three stages kept) was measured far slower: 1.88 vs 1.12 ms per launch -- LSU loads queue behind the TMA
  // operand traffic.
  static constexpr bool kStageX = EPI == kEpiGnBwd;
  static constexpr int kStagingBytes = kOutBytes * (kStageZ || kDual || kStageX ? 2 : 1);
  static constexpr int kStagesRaw = (227 * 1024 - 2048 - kStagingBytes) / (kABytes + kBHalfBytes);
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * (kABytes + kBHalfBytes) + kStagingBytes + 1024 + 512;
  // halo mode: a stage = one halo A tile + the B halves of the three dx taps, carved out of the same ring
  static constexpr int kHaloStageBytes = kHaloABytes + 3 * kBHalfBytes;
  static constexpr int kHaloStages = kStages * (kABytes + kBHalfBytes) / kHaloStageBytes;
};

constexpr int kMaxChunks = 4;   // 64-column chunks of the widest tile (BLOCK_N = 256)

template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mtgemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmAh,
               const __grid_constant__ CUtensorMap tmX2, const __grid_constant__ MtParams P) {
#ifdef TVAE_DEVICE_OK
  using Cfg = Mt2Cfg<BLOCK_N, EPI>;
  constexpr int STAGES = Cfg::kStages;
  constexpr bool kHasRes = epi_has_res<EPI>();
  constexpr bool kStageX = Cfg::kStageX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kABytes;
  uint8_t* sOut = sB + STAGES * Cfg::kBHalfBytes;
  uint8_t* sZ = sOut + Cfg::kOutBytes;            // second staged tile (kStageZ only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + Cfg::kStagingBytes);
  uint64_t* full = bars;                   // [STAGES] (leader's are used)
  uint64_t* empty = bars + STAGES;         // [STAGES] (each CTA's own)
  uint64_t* tmem_full = bars + 2 * STAGES; // [2]      (each CTA's own)
  uint64_t* tmem_empty = tmem_full + 2;    // [2]      (leader's are used, 16 arrivals)
  uint64_t* res_full = tmem_empty + 2;     // [1]
  uint64_t* out_free = res_full + 1;       // [1]
  uint64_t* res_full_c = out_free + 1;     // [kMaxChunks] per 64-column chunk (chunk-pipelined residual epilogues)
  uint64_t* out_free_c = res_full_c + kMaxChunks;   // [kMaxChunks]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_free_c + kMaxChunks);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const bool has_res = kHasRes && P.has_residual;
  constexpr int kChunks = BLOCK_N / 64;
  constexpr bool kCanPipe = EPI != kEpiDirect && kChunks >= 2;
  // fused GroupNorm statistics of the output (bias / bias+residual epilogues only): per-CTA staging of (sum, sumsq)
  constexpr bool kGn = (EPI == kEpiBias || EPI == kEpiBiasRes);
  __shared__ double s_gn[kGn ? 128 : 1];     // fp64 like the global sums: several column groups may add into one group
  const bool gn = kGn && P.gn_sums != nullptr;
  // Chunk-pipelined residual epilogue: the staged residual (and z) tile arrives, is combined in place and leaves again
  // per 64-column chunk, each chunk with its own pair of barriers, so the next tile's residual is in flight while this
  // tile's later chunks are drained -- with ONE residual barrier per tile the load (128 KB at an SM's 44 GB/s share of
  // HBM: 2.9 us), the TMEM drain and the store followed each other (K = 384 GELU-gradient GEMM: 8.5 us per tile against
  // 4.4 us of HBM time).  Not with the statistics passes, which read the whole staged tile after the drain.
  const bool pipe_res = kCanPipe && has_res && !gn && !kStageX && P.pipe_res != 0;
  if constexpr (kGn) {
    if (threadIdx.x < 128) s_gn[threadIdx.x] = 0.0;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRes);
    tma_prefetch_desc(&tmAh);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 16);
    }
    mbar_init(res_full, 1);
    mbar_init(out_free, 1);
    for (int j = 0; j < kMaxChunks; ++j) {
      mbar_init(&res_full_c[j], 1);
      mbar_init(&out_free_c[j], 1);
    }
    fence_mbar_init();
  }
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / cross-CTA TMA completion
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

  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_pairs = P.num_phases * m_pairs * P.n_tiles;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  // pair tile -> (n tile, this CTA's pixel tile, phase); the pixel tile may be a phantom beyond the tensor
#define TVAE_DECODE_PAIR(pt)                                              \
  const int n_t = (pt) % P.n_tiles;                                       \
  const int t2 = (pt) / P.n_tiles;                                        \
  const int m_t = 2 * (t2 % m_pairs) + (int)rank;                         \
  const int ph = t2 / m_pairs;                                            \
  const int w0 = (m_t % P.tiles_w) * P.tw;                                \
  const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;                  \
  const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0 && P.halo) {
      // halo mode: per (kernel row dy, K block) one 130-pixel A tile [w0 - 1, w0 + 129) and three B halves
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        TVAE_DECODE_PAIR(pt)
        (void)ph;
        const int kbs = P.taps[0][0].kblocks;
        for (int dy = 0; dy < 3; ++dy) {
          for (int kb = 0; kb < kbs; ++kb) {
            uint8_t* st = smem + stage * Cfg::kHaloStageBytes;
            mbar_wait(&empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (kHaloATx + 3 * Cfg::kBHalfBytes));
            tma2_load_5d(st, &tmAh, &full[stage], kb * kBlockK, w0 - 1, 0, h0 + P.taps[0][dy * 3].dh, b0);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma2_load_2d(st + kHaloABytes + dx * Cfg::kBHalfBytes, &tmB, &full[stage],
                           P.taps[0][dy * 3 + dx].wk_off + kb * kBlockK, n_t * BLOCK_N + (int)rank * (BLOCK_N / 2));
            if (++stage == Cfg::kHaloStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        TVAE_DECODE_PAIR(pt)
        const int nt = P.ntaps[ph];
        for (int t = 0; t < nt; ++t) {
          const MtTap tap = P.taps[ph][t];
          const CUtensorMap* mapA = tap.map ? &tmA1 : &tmA0;
          for (int kb = 0; kb < tap.kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (kABytes + Cfg::kBHalfBytes));
            tma2_load_5d(sA + stage * kABytes, mapA, &full[stage], tap.c_off + kb * kBlockK, w0 + tap.dw, tap.p,
                         h0 + tap.dh, b0);
            tma2_load_2d(sB + stage * Cfg::kBHalfBytes, &tmB, &full[stage], tap.wk_off + kb * kBlockK,
                         n_t * BLOCK_N + (int)rank * (BLOCK_N / 2));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The whole warp walks the loop in uniform control flow and one elected lane issues: operands stay in uniform
    // registers and descriptors advance by constants (common.cuh, "warp-uniform issue").
    if (leader && P.halo) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N, 0, 0);
      // descriptor base-offset field cleared: the 128-byte swizzle is a function of the shared-memory ADDRESS bits (as TMA
      // wrote it), so a start address in the middle of a 1024-byte swizzle atom needs no correction -- with base offset =
      // (addr >> 7) & 7 every row-shifted tap came out wrong, with 0 all are exact (tools/experiments/umma_row_offset.cu)
      const uint64_t desc0 = umma_desc_kmajor_sw128(smem_u32(smem));      // 1024-byte aligned: base offset 0
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int kbs = P.taps[0][0].kblocks;
      // the tap with pixel offset dw reads halo rows [dw + 1, dw + 129): the K-major SW128 layout is linear in
      // the row (128 B per row, SBO = 8 rows), so a tap is just a start address (dw + 1) rows further
      uint32_t tap_off[3][3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) tap_off[dy][dx] = uniform_u32((P.taps[0][dy * 3 + dx].dw + 1) * 128);
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll 1
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t da = umma_desc_advance(desc0, stage * Cfg::kHaloStageBytes);
            const uint64_t db = umma_desc_advance(da, kHaloABytes);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const uint64_t da_tap = umma_desc_advance(da, tap_off[dy][dx]);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                umma2_f16_elect(d_tmem, umma_desc_advance(da_tap, k * 32), umma_desc_advance(db, dx * Cfg::kBHalfBytes + k * 32),
                                idesc, (dy | kb | dx | k) != 0);
              }
            }
            umma2_commit_mc_elect(&empty[stage]);
            if (++stage == Cfg::kHaloStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        umma2_commit_mc_elect(&tmem_full[acc]);
      }
    } else if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N, 0, 0);
      const uint64_t desc_a0 = umma_desc_kmajor_sw128(smem_u32(sA)), desc_b0 = umma_desc_kmajor_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++it) {
        const int ph = (pt / P.n_tiles) / m_pairs;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        int kblocks = 0;
        for (int t = 0; t < P.ntaps[ph]; ++t) kblocks += P.taps[ph][t].kblocks;
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_advance(desc_a0, stage * kABytes), db = umma_desc_advance(desc_b0, stage * Cfg::kBHalfBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma2_f16_elect(d_tmem, umma_desc_advance(da, k * 32), umma_desc_advance(db, k * 32), idesc, (kb | k) != 0);
          }
          umma2_commit_mc_elect(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc_elect(&tmem_full[acc]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual loader (each CTA, own tile)
    if (lane == 0 && (has_res || kStageX)) {
      int it = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++it) {
        TVAE_DECODE_PAIR(pt)
        if constexpr (kCanPipe && kHasRes) {
          if (pipe_res) {
            // chunk by chunk: chunk j of this tile is fetched as soon as the store of the previous tile's chunk j has been
            // read (out_free_c[j]), i.e. while the epilogue warps are still working on that tile's later chunks
#pragma unroll
            for (int j = 0; j < kChunks; ++j) {
              mbar_wait(&out_free_c[j], (it & 1) ^ 1);
              mbar_arrive_expect_tx(&res_full_c[j], kABytes * (Cfg::kStageZ ? 2 : 1));
              tma_load_5d(sOut + j * kABytes, &tmRes, &res_full_c[j], P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0,
                          P.out_p[ph], h0, b0);
              if constexpr (Cfg::kStageZ)
                tma_load_5d(sZ + j * kABytes, &tmX2, &res_full_c[j], P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0,
                            P.out_p[ph], h0, b0);
            }
            continue;
          }
        }
        mbar_wait(out_free, (it & 1) ^ 1);
        mbar_arrive_expect_tx(res_full, kStageX ? Cfg::kOutBytes : Cfg::kStagingBytes);
        if constexpr (kStageX) {             // the GroupNorm input x of this tile (map tmX2) -> second staging tile
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_5d(sZ + j * kABytes, &tmX2, res_full, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
          continue;
        }
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_5d(sOut + j * kABytes, &tmRes, res_full, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph],
                      h0, b0);
        if constexpr (Cfg::kStageZ) {        // the saved pre-activation z: same view as the output (map tmX2)
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_5d(sZ + j * kABytes, &tmX2, res_full, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (each CTA, own 128 rows)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int half = (warp - 4) >> 2;
    const bool store_leader = (threadIdx.x == 128);
    // Chunk-pipelined stores: epilogues that only WRITE the staging tile (no residual tile arriving in it, no statistics
    // pass over it) send each 64-column chunk off as soon as the eight warps have finished it, and never wait for a
    // store at the end of a tile -- with one store + wait per tile the small-K GEMMs (K = 384: 3000 clocks of MMA per
    // tile) were bound by the epilogue's TMEM drain + store latency.  Needs at least two chunks.
    const bool pipe = kCanPipe && (!has_res || pipe_res) && !gn && !kStageX;
    int it = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++it) {
      TVAE_DECODE_PAIR(pt)
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      EpiRow R;
      const __nv_bfloat16* zrow = nullptr;
      {
        const int wi = r % P.tw, hi = (r / P.tw) % P.th, bi = r / (P.tw * P.th);
        R.pw = w0 + wi; R.phh = h0 + hi; R.pb = b0 + bi;
        R.row_ok = (R.pw < P.vW) && (R.phh < P.vH) && (R.pb < P.vB);
        const long long grow = ((long long)R.pb * P.vH + R.phh) * P.vW + R.pw;
        R.rs = 1.0f; R.rsh = 0.0f; R.rope_r = 0; R.rope_c = 0;
        if constexpr (EPI == kEpiResMulGeluGrad) zrow = P.z + (R.row_ok ? grow : 0) * P.n_total;
        if constexpr (EPI == kEpiRsBiasGelu || EPI == kEpiAffineRope || EPI == kEpiRsBias) {
          if (P.row_scale != nullptr && R.row_ok) R.rs = __ldg(P.row_scale + grow);
          if (P.row_shift != nullptr && R.row_ok) R.rsh = __ldg(P.row_shift + grow);
        }
        if constexpr (EPI == kEpiAffineRope) {
          if (P.rope_tab != nullptr) {
            const int tok = (int)(grow % ((long long)P.rope_H * P.rope_W));
            R.rope_r = tok / P.rope_W;
            R.rope_c = tok % P.rope_W;
          }
        }
      }
      const float* bias = P.bias ? P.bias + (size_t)ph * P.n_total : nullptr;
      // clipped rows must not reach the statistics / the GroupNorm-backward sums (TMA drops them from the store anyway)
      const bool gn_zero_row = (gn || kStageX) && !R.row_ok;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if ((has_res && !pipe_res) || kStageX) mbar_wait(res_full, it & 1);

      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      const int n_base = n_t * BLOCK_N;
      // software-pipelined over the 16-column groups: the tcgen05.ld of the next group is in flight while the current
      // group goes through the epilogue math (two register sets, ping-pong)
      auto process = [&](const int c16, const uint32_t (&va)[8], const uint32_t (&vb)[8]) {
        float fa[8], fb[8];
        epi_math8<EPI>(P, bias, R, n_base + c16 * 16, va, fa);
        epi_math8<EPI>(P, bias, R, n_base + c16 * 16 + 8, vb, fb);
        if constexpr (EPI == kEpiDirect) {
          if (R.row_ok) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int n = n_base + c16 * 16 + k;
              if (n < P.out_n) P.out_f32[(((long long)R.pb * P.out_n + n) * P.vH + R.phh) * P.vW + R.pw] = fa[k];
              if (n + 8 < P.out_n) P.out_f32[(((long long)R.pb * P.out_n + n + 8) * P.vH + R.phh) * P.vW + R.pw] = fb[k];
            }
          }
        } else {
          uint8_t* chunk = sOut + (c16 >> 2) * kABytes + r * 128;
          const int g0 = (c16 & 3) * 2;
          uint4* pa = reinterpret_cast<uint4*>(chunk + (((g0) ^ (r & 7)) << 4));
          uint4* pb = reinterpret_cast<uint4*>(chunk + (((g0 + 1) ^ (r & 7)) << 4));
          if constexpr (kHasRes) {
            if (has_res) {
              const uint4 ra = *pa, rb = *pb;
              if constexpr (Cfg::kStageZ) {
                const uint4 za = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(pa) + Cfg::kOutBytes);
                const uint4 zb = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(pb) + Cfg::kOutBytes);
                epi_res_mul_gelu_grad8(fa, ra, za);
                epi_res_mul_gelu_grad8(fb, rb, zb);
              } else {
                epi_combine8<EPI>(fa, ra, zrow, n_base + c16 * 16);
                epi_combine8<EPI>(fb, rb, zrow, n_base + c16 * 16 + 8);
              }
            }
          }
          if constexpr (kGn || kStageX) {
            if (gn_zero_row) {
#pragma unroll
              for (int k = 0; k < 8; ++k) fa[k] = fb[k] = 0.0f;
            }
          }
          uint4 o;
          o.x = pack_bf16(fa[0], fa[1]); o.y = pack_bf16(fa[2], fa[3]);
          o.z = pack_bf16(fa[4], fa[5]); o.w = pack_bf16(fa[6], fa[7]);
          *pa = o;
          if constexpr (Cfg::kDual) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(pa) + Cfg::kOutBytes) = epi_act_of_bf16<EPI>(o);
          o.x = pack_bf16(fb[0], fb[1]); o.y = pack_bf16(fb[2], fb[3]);
          o.z = pack_bf16(fb[4], fb[5]); o.w = pack_bf16(fb[6], fb[7]);
          *pb = o;
          if constexpr (Cfg::kDual) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(pb) + Cfg::kOutBytes) = epi_act_of_bf16<EPI>(o);
        }
      };
      {
        constexpr int kGroups = BLOCK_N / 16;
        uint32_t v0a[8], v0b[8], v1a[8], v1b[8];
        tmem_ld8(t_row + half * 16, v0a);
        tmem_ld8(t_row + half * 16 + 8, v0b);
#pragma unroll 1
        for (int c16 = half; c16 < kGroups; c16 += 4) {      // every warp owns an even number of groups
          if constexpr (kCanPipe && kHasRes) {
            if (pipe_res) mbar_wait(&res_full_c[c16 >> 2], it & 1);     // residual (and z) chunk of this tile has landed
          }
          tmem_ld_wait_dep(v0a, v0b);
          tmem_ld8(t_row + (c16 + 2) * 16, v1a);
          tmem_ld8(t_row + (c16 + 2) * 16 + 8, v1b);
          process(c16, v0a, v0b);
          tmem_ld_wait_dep(v1a, v1b);
          if (c16 + 4 < kGroups) {
            tmem_ld8(t_row + (c16 + 4) * 16, v0a);
            tmem_ld8(t_row + (c16 + 4) * 16 + 8, v0b);
          }
          process(c16 + 2, v1a, v1b);
          if constexpr (kCanPipe) {
            if (pipe) {
              // the 64-column chunk c16 >> 2 is complete in every warp after this barrier: its store goes out while
              // the next chunk is drained.  Before the barrier the leader makes sure the store that last read the NEXT
              // chunk to be written (same chunk, previous tile) is done: kChunks - 2 younger groups may still be pending.
              fence_proxy_async_smem();
              if (store_leader && !pipe_res) tma_store_wait_read<kChunks - 2>();
              asm volatile("bar.sync 1, 256;" ::: "memory");
              if (store_leader) {
                const int j = c16 >> 2;
                tma_store_5d(&tmOut, sOut + j * kABytes, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
                if constexpr (Cfg::kDual)
                  tma_store_5d(&tmX2, sZ + j * kABytes, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
                tma_store_commit();
                if constexpr (kHasRes) {
                  // the store before this one (previous chunk; for chunk 0 the previous tile's last chunk) has been read:
                  // the loader may refill that chunk with the next tile's residual
                  if (pipe_res && (it > 0 || j > 0)) {
                    tma_store_wait_read<1>();
                    mbar_arrive(&out_free_c[(j + kChunks - 1) % kChunks]);
                  }
                }
              }
            }
          }
        }
      }
      // accumulator half drained -> tell the leader's MMA warp (remote arrive from the peer CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);

      if constexpr (EPI != kEpiDirect) {
        if constexpr (kCanPipe) {
          if (pipe) continue;               // stores already issued chunk by chunk (uniform over the CTA)
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (store_leader) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            tma_store_5d(&tmOut, sOut + j * kABytes, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
            if constexpr (Cfg::kDual)
              tma_store_5d(&tmX2, sZ + j * kABytes, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
          }
          tma_store_commit();
        }
        if constexpr (kGn) {
          if (gn) {
            // column sums of the staged bf16 tile (what the consumer's GroupNorm will read).  A warp takes four 8-channel
            // column groups; its lanes are (column group, row mod 8), so the 16-byte reads of a quarter warp hit eight
            // different slots of the 128-byte swizzle (conflict-free), the eight partial sums of a column group meet in a
            // 3-step butterfly and ONE lane per column group folds its 8 channels into the per-group staging sums -- the
            // first version (thread = column group x row slot, every thread flushing through shared-memory atomics) put
            // 1200 ten-way contended CAS loops on the shared-memory port per tile and cost 8 % of the kernel
            constexpr int kVec = BLOCK_N / 8;
            const int gc = (warp - 4) * 4 + (lane >> 3);
            const int rl = lane & 7;
            if ((warp - 4) * 4 < kVec) {          // warp-uniform (kVec is a multiple of 4)
              const uint8_t* col = sOut + (gc >> 3) * kABytes + rl * 128 + ((((gc & 7) ^ rl)) << 4);
              float2 s[4], ss[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) s[i] = ss[i] = make_float2(0.0f, 0.0f);
#pragma unroll 4
              for (int rb = 0; rb < kBlockM / 8; ++rb) {
                const uint4 u = *reinterpret_cast<const uint4*>(col + rb * 1024);
                float2 f[4];
                unpack8_2(u, f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  s[i] = __fadd2_rn(s[i], f[i]);
                  ss[i] = __ffma2_rn(f[i], f[i], ss[i]);
                }
              }
#pragma unroll
              for (int m = 1; m < 8; m <<= 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  s[i].x += __shfl_xor_sync(0xffffffffu, s[i].x, m);
                  s[i].y += __shfl_xor_sync(0xffffffffu, s[i].y, m);
                  ss[i].x += __shfl_xor_sync(0xffffffffu, ss[i].x, m);
                  ss[i].y += __shfl_xor_sync(0xffffffffu, ss[i].y, m);
                }
              }
              if (rl == 0) {
                const float sv[8] = {s[0].x, s[0].y, s[1].x, s[1].y, s[2].x, s[2].y, s[3].x, s[3].y};
                const float qv[8] = {ss[0].x, ss[0].y, ss[1].x, ss[1].y, ss[2].x, ss[2].y, ss[3].x, ss[3].y};
                const int c0 = n_t * BLOCK_N + gc * 8;
                int g = c0 / P.gn_cpg, rem = c0 - g * P.gn_cpg;
                float as = 0.0f, aq = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  as += sv[k];
                  aq += qv[k];
                  if (++rem == P.gn_cpg || k == 7) {
                    if (g < P.gn_groups) {
                      atomicAdd(&s_gn[2 * g], (double)as);
                      atomicAdd(&s_gn[2 * g + 1], (double)aq);
                    }
                    as = aq = 0.0f;
                    rem = 0;
                    ++g;
                  }
                }
              }
            }
          }
        }
        if constexpr (kStageX) {
          // Reduce pass of the GroupNorm backward over the two staged tiles (dh as stored, x): lanes = (8-channel column
          // group, row mod 8) as in the forward statistics pass; dy = dh * act'(gamma * xhat + beta), per channel the sums
          // of dy and dy * xhat over the tile's 128 pixels (3-step butterfly over the row lanes, fixed order).  Rows beyond
          // the image were staged as zeros (gn_zero_row) and contribute nothing.
          constexpr int kVec = BLOCK_N / 8;
          const int gc = (warp - 4) * 4 + (lane >> 3);
          const int rl = lane & 7;
          if ((warp - 4) * 4 < kVec && b0 < P.vB) {          // warp-uniform; phantom tiles write nothing
            const int c0 = n_t * BLOCK_N + gc * 8;
            float2 ca[4], cb[4], cg[4], ce[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c0 + 2 * i;
              float m0, r0, m1, r1;
              gn_mean_rstd(P.gnb_sums, b0, P.gnb_groups, c / P.gnb_cpg, P.gnb_inv_n, P.gnb_eps, m0, r0);
              gn_mean_rstd(P.gnb_sums, b0, P.gnb_groups, (c + 1) / P.gnb_cpg, P.gnb_inv_n, P.gnb_eps, m1, r1);
              ca[i] = make_float2(r0, r1);
              cb[i] = make_float2(-m0 * r0, -m1 * r1);
              cg[i] = make_float2(__ldg(P.gnb_gamma + c), __ldg(P.gnb_gamma + c + 1));
              ce[i] = make_float2(__ldg(P.gnb_beta + c), __ldg(P.gnb_beta + c + 1));
            }
            const uint8_t* cd = sOut + (gc >> 3) * kABytes + rl * 128 + ((((gc & 7) ^ rl)) << 4);
            const uint8_t* cx = cd + Cfg::kOutBytes;
            float2 s1[4], s2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) s1[i] = s2[i] = make_float2(0.0f, 0.0f);
            const bool silu = P.gnb_silu != 0;
#pragma unroll 4
            for (int rb = 0; rb < kBlockM / 8; ++rb) {
              float2 g[4], xf[4];
              unpack8_2(*reinterpret_cast<const uint4*>(cd + rb * 1024), g);
              unpack8_2(*reinterpret_cast<const uint4*>(cx + rb * 1024), xf);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 xh = __ffma2_rn(xf[i], ca[i], cb[i]);
                const float2 dy = silu ? __fmul2_rn(g[i], silu_grad2(__ffma2_rn(xh, cg[i], ce[i]))) : g[i];
                s1[i] = __fadd2_rn(s1[i], dy);
                s2[i] = __ffma2_rn(dy, xh, s2[i]);
              }
            }
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                s1[i].x += __shfl_xor_sync(0xffffffffu, s1[i].x, m);
                s1[i].y += __shfl_xor_sync(0xffffffffu, s1[i].y, m);
                s2[i].x += __shfl_xor_sync(0xffffffffu, s2[i].x, m);
                s2[i].y += __shfl_xor_sync(0xffffffffu, s2[i].y, m);
              }
            }
            if (rl == 0) {
              float4* dst = reinterpret_cast<float4*>(P.gnb_part + ((size_t)m_t * P.n_total + c0) * 2);
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = make_float4(s1[i].x, s2[i].x, s1[i].y, s2[i].y);
            }
          }
        }
        if (store_leader) tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // the staged tile has been read by the TMA store and by the statistics pass: the residual loader may refill it
        if (store_leader && (has_res || kStageX)) mbar_arrive(out_free);
        if constexpr (kGn) {
          if (gn) {
            const int et = (int)threadIdx.x - 128;
            if (et < 2 * P.gn_groups) {
              const double v = s_gn[et];
              if (v != 0.0) {       // only the groups of this n tile; a phantom tile (b0 >= vB) contributes nothing
                atomicAdd(P.gn_sums + (size_t)b0 * 2 * P.gn_groups + et, v);
                s_gn[et] = 0.0;
              }
            }
          }
        }
      }
    }
    if (store_leader) tma_store_wait<0>();
  }
#undef TVAE_DECODE_PAIR

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

template <int BLOCK_N, int EPI>
static int launch2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                   const CUtensorMap& r, const CUtensorMap& ah, const CUtensorMap& x2, const MtParams& P, cudaStream_t stream) {
  using Cfg = Mt2Cfg<BLOCK_N, EPI>;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(mtgemm2_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    configured = true;
  }
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int total_pairs = P.num_phases * ((m_tiles + 1) / 2) * P.n_tiles;
  int clusters = persistent_sms() / 2;
  if (clusters <= 0) clusters = 74;
  if (total_pairs < clusters) clusters = total_pairs;
  mtgemm2_kernel<BLOCK_N, EPI><<<2 * clusters, kThreads, Cfg::kSmemBytes, stream>>>(a0, a1, b, o, r, ah, x2, P);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int EPI>
static int launch2_n(int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                     const CUtensorMap& r, const CUtensorMap& ah, const CUtensorMap& x2, const MtParams& P,
                     cudaStream_t stream) {
  switch (block_n) {
    case 256: return launch2<256, EPI>(a0, a1, b, o, r, ah, x2, P, stream);
    case 192: return launch2<192, EPI>(a0, a1, b, o, r, ah, x2, P, stream);
    case 64: return launch2<64, EPI>(a0, a1, b, o, r, ah, x2, P, stream);
    default: return launch2<128, EPI>(a0, a1, b, o, r, ah, x2, P, stream);
  }
}

// Called by mtgemm_run (mtgemm.cu) when the CTA-pair kernel applies (block_n >= 128).  `b` must be a weight map with
// box rows = block_n / 2.
int mtgemm2_dispatch(int epi, int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                     const CUtensorMap& o, const CUtensorMap& r, const CUtensorMap& ah, const CUtensorMap& x2,
                     const MtParams& P, cudaStream_t stream) {
  switch (epi) {
    case kEpiBias: return launch2_n<kEpiBias>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiBiasRes: return launch2_n<kEpiBiasRes>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiBiasGelu: return launch2_n<kEpiBiasGelu>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiBiasSilu: return launch2_n<kEpiBiasSilu>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiRsBiasGelu: return launch2_n<kEpiRsBiasGelu>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiAffineRope: return launch2_n<kEpiAffineRope>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiDirect: return launch2_n<kEpiDirect>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiMulGeluGrad: return launch2_n<kEpiMulGeluGrad>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiMulSiluGrad: return launch2_n<kEpiMulSiluGrad>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiResMulGeluGrad: return launch2_n<kEpiResMulGeluGrad>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiBiasGeluDual: return launch2_n<kEpiBiasGeluDual>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiBiasSiluDual: return launch2_n<kEpiBiasSiluDual>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    case kEpiGnBwd: return launch2_n<kEpiGnBwd>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
    default: return launch2_n<kEpiRsBias>(block_n, a0, a1, b, o, r, ah, x2, P, stream);
  }
}

}  // namespace tvae
