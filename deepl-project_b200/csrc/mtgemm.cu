// Fused multi-tap implicit GEMM for sm_100a.
//
//   out[tile] = epilogue( sum_taps  A_tap[128 pixels x 64k] * W[BLOCK_N x 64k]^T )
//
// One persistent CTA per SM, 12 warps, warp-specialised:
//   warp 0  TMA producer   : per K block one 5-D pixel-box load (A; zero-filled halo = conv padding) and one
//                            2-D weight box load (B) into a STAGES-deep 128B-swizzled smem ring
//   warp 1  MMA issuer     : tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16, fp32 accumulators in
//                            TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
//                            main loop of tile i+1
//   warp 2  residual loader: TMA-loads the residual tile straight into the output staging buffer
//   warp 3  idle
//   warps 4-11 epilogue    : tcgen05.ld -> row/col affine, bias, GELU/SiLU, RoPE, residual -> bf16 -> swizzled
//                            smem -> TMA store (or direct fp32 NCHW store for the 3-/64-channel heads); two warps
//                            per TMEM lane quarter, each taking every other 16-column group
//
// Every convolution, linear layer, pixel (un)shuffle and nearest-2x upsample of the reference is one launch of
// this kernel with a different tap table (see include/transvae_sm100.h and transvae/_taps.py).
#include <cstdlib>

#include "mtgemm_common.cuh"

namespace tvae {

int mtgemm2_dispatch(int epi, int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                     const CUtensorMap& o, const CUtensorMap& r, const CUtensorMap& ah, const CUtensorMap& x2,
                     const MtParams& P, cudaStream_t stream);
int act_fwd_run(const void* z, void* y, long long n, int act, cudaStream_t stream);   // backward.cu
int gn_bwd_reduce_run(const void* x, const void* dh, const double* sums, const float* gamma, const float* beta, float* part,
                      int B, int HW, int C, int G, float eps, int apply_silu, cudaStream_t stream);          // backward.cu
int gn_bwd_tiles_reduce_run(const float* part_tiles, float* groups_ws, float* part, int B, int tiles_per_image, int n,
                            cudaStream_t stream);                                                            // backward.cu
static int mtgemm1_dispatch(int epi, int block_n, const CUtensorMap& mA0, const CUtensorMap& mA1, const CUtensorMap& mB,
                            const CUtensorMap& mO, const CUtensorMap& mR, const MtParams& P, cudaStream_t stream);
int gn_stats_run(const void* x, double* sums, int B, int HW, int C, int G, cudaStream_t stream);   // elementwise.cu

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
mtgemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
              const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
              const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ MtParams P) {
#ifdef TVAE_DEVICE_OK
  using Cfg = MtCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  constexpr bool kHasRes = epi_has_res<EPI>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kABytes;
  uint8_t* sOut = sB + STAGES * Cfg::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + Cfg::kOutBytes);
  uint64_t* full = bars;                   // [STAGES]
  uint64_t* empty = bars + STAGES;         // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES; // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint64_t* res_full = tmem_empty + 2;     // [1]
  uint64_t* out_free = res_full + 1;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_free + 1);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const bool has_res = kHasRes && P.has_residual;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRes);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    mbar_init(res_full, 1);
    mbar_init(out_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int total_tiles = P.num_phases * m_tiles * P.n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_t = tile % P.n_tiles;
        const int t2 = tile / P.n_tiles;
        const int m_t = t2 % m_tiles;
        const int ph = t2 / m_tiles;
        const int w0 = (m_t % P.tiles_w) * P.tw;
        const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
        const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
        const int nt = P.ntaps[ph];
        for (int t = 0; t < nt; ++t) {
          const MtTap tap = P.taps[ph][t];
          const CUtensorMap* mapA = tap.map ? &tmA1 : &tmA0;
          for (int kb = 0; kb < tap.kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], kABytes + Cfg::kBBytes);
            tma_load_5d(sA + stage * kABytes, mapA, &full[stage], tap.c_off + kb * kBlockK, w0 + tap.dw, tap.p,
                        h0 + tap.dh, b0);
            tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full[stage], tap.wk_off + kb * kBlockK, n_t * BLOCK_N);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {   // whole warp, elected lane issues (common.cuh: warp-uniform issue)
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int ph = (tile / P.n_tiles) / m_tiles;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        int kblocks = 0;
        for (int t = 0; t < P.ntaps[ph]; ++t) kblocks += P.taps[ph][t].kblocks;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * kABytes);
          const uint32_t b_base = smem_u32(sB + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_f16_elect(d_tmem, umma_desc_kmajor_sw128(a_base + k * 32), umma_desc_kmajor_sw128(b_base + k * 32), idesc,
                     (kb | k) != 0);
          }
          umma_commit_elect(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_elect(&tmem_full[acc]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual loader
    if (lane == 0 && has_res) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int n_t = tile % P.n_tiles;
        const int t2 = tile / P.n_tiles;
        const int m_t = t2 % m_tiles;
        const int ph = t2 / m_tiles;
        const int w0 = (m_t % P.tiles_w) * P.tw;
        const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
        const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
        mbar_wait(out_free, (it & 1) ^ 1);
        mbar_arrive_expect_tx(res_full, Cfg::kOutBytes);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_5d(sOut + j * kABytes, &tmRes, res_full, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph],
                      h0, b0);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // 8 epilogue warps: warps w and w+4 share TMEM lane quarter (w & 3) and split the tile's columns between them
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row of the tile owned by this thread
    const int half = (warp - 4) >> 2;  // which interleaved half of the 16-column groups this warp handles
    const bool store_leader = (threadIdx.x == 128);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_t = tile % P.n_tiles;
      const int t2 = tile / P.n_tiles;
      const int m_t = t2 % m_tiles;
      const int ph = t2 / m_tiles;
      const int w0 = (m_t % P.tiles_w) * P.tw;
      const int h0 = ((m_t / P.tiles_w) % P.tiles_h) * P.th;
      const int b0 = (m_t / (P.tiles_w * P.tiles_h)) * P.nb;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;

      EpiRow R;
      const __nv_bfloat16* zrow = nullptr;
      {
        const int wi = r % P.tw, hi = (r / P.tw) % P.th, bi = r / (P.tw * P.th);
        R.pw = w0 + wi; R.phh = h0 + hi; R.pb = b0 + bi;
        R.row_ok = (R.pw < P.vW) && (R.phh < P.vH) && (R.pb < P.vB);
        const long long grow = ((long long)R.pb * P.vH + R.phh) * P.vW + R.pw;  // flattened output-view pixel index
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

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (has_res) mbar_wait(res_full, it & 1);

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
          // 16 columns = two 16-byte groups of this thread's 128-byte row inside 64-column chunk (c16 / 4)
          uint8_t* chunk = sOut + (c16 >> 2) * kABytes + r * 128;
          const int g0 = (c16 & 3) * 2;
          uint4* pa = reinterpret_cast<uint4*>(chunk + (((g0) ^ (r & 7)) << 4));
          uint4* pb = reinterpret_cast<uint4*>(chunk + (((g0 + 1) ^ (r & 7)) << 4));
          if constexpr (kHasRes) {
            if (has_res) {
              const uint4 ra = *pa, rb = *pb;
              epi_combine8<EPI>(fa, ra, zrow, n_base + c16 * 16);
              epi_combine8<EPI>(fb, rb, zrow, n_base + c16 * 16 + 8);
            }
          }
          uint4 o;
          o.x = pack_bf16(fa[0], fa[1]); o.y = pack_bf16(fa[2], fa[3]);
          o.z = pack_bf16(fa[4], fa[5]); o.w = pack_bf16(fa[6], fa[7]);
          *pa = o;
          o.x = pack_bf16(fb[0], fb[1]); o.y = pack_bf16(fb[2], fb[3]);
          o.z = pack_bf16(fb[4], fb[5]); o.w = pack_bf16(fb[6], fb[7]);
          *pb = o;
        }
      };
      {
        constexpr int kGroups = BLOCK_N / 16;
        uint32_t v0a[8], v0b[8], v1a[8], v1b[8];
        tmem_ld8(t_row + half * 16, v0a);
        tmem_ld8(t_row + half * 16 + 8, v0b);
#pragma unroll 1
        for (int c16 = half; c16 < kGroups; c16 += 4) {      // every warp owns an even number of groups
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
        }
      }
      // TMEM accumulator drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);

      if constexpr (EPI != kEpiDirect) {
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (store_leader) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_store_5d(&tmOut, sOut + j * kABytes, P.out_c_off[ph] + n_t * BLOCK_N + j * 64, w0, P.out_p[ph], h0, b0);
          tma_store_commit();
          tma_store_wait_read<0>();
          if (has_res) mbar_arrive(out_free);
        }
        // staging buffer may be overwritten only after the bulk store has read it
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    if (store_leader) tma_store_wait<0>();
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
// host side
// -------------------------------------------------------------------------------------------------
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BLOCK_N, int EPI>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                  const CUtensorMap& r, const MtParams& P, cudaStream_t stream) {
  using Cfg = MtCfg<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(mtgemm_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    configured = true;
  }
  const int total = P.num_phases * P.tiles_w * P.tiles_h * P.tiles_b * P.n_tiles;
  int grid = persistent_sms();
  if (grid <= 0) grid = 148;
  if (total < grid) grid = total;
  mtgemm_kernel<BLOCK_N, EPI><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(a0, a1, b, o, r, P);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int EPI>
static int launch_n(int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                    const CUtensorMap& r, const MtParams& P, cudaStream_t stream) {
  switch (block_n) {
    case 256: return launch<256, EPI>(a0, a1, b, o, r, P, stream);
    case 192: return launch<192, EPI>(a0, a1, b, o, r, P, stream);
    case 128: return launch<128, EPI>(a0, a1, b, o, r, P, stream);
    default: return launch<64, EPI>(a0, a1, b, o, r, P, stream);
  }
}

static void view_extents(const tvae_view& v, int* vW, int* vH, int* vB) {
  *vW = v.split ? v.W / 2 : v.W;
  *vH = v.split ? v.H / 2 : v.H;
  *vB = v.B;
}

int mtgemm_run(const tvae_mtgemm_desc* d, cudaStream_t stream) {
  TVAE_REQUIRE(d != nullptr, "null descriptor");
  TVAE_REQUIRE(d->a0.ptr != nullptr && d->w != nullptr, "mtgemm: missing operand");
  TVAE_REQUIRE(d->num_phases >= 1 && d->num_phases <= TVAE_MAX_PHASES, "mtgemm: bad num_phases %d", d->num_phases);
  TVAE_REQUIRE(d->n_total % 64 == 0 && d->k_total % 64 == 0, "mtgemm: n_total %d / k_total %d must be multiples of 64",
               d->n_total, d->k_total);
  const bool direct = d->out_f32 != nullptr;
  TVAE_REQUIRE(direct || d->out.ptr != nullptr, "mtgemm: no output");

  MtParams P;
  memset(&P, 0, sizeof(P));
  // the tile grid is defined on the output view (for a direct store: on a0's view, which must match)
  const tvae_view& gv = direct ? d->a0 : d->out;
  view_extents(gv, &P.vW, &P.vH, &P.vB);
  P.tw = pow2_ceil(P.vW) < 128 ? pow2_ceil(P.vW) : 128;
  P.th = pow2_ceil(P.vH) < 128 / P.tw ? pow2_ceil(P.vH) : 128 / P.tw;
  P.nb = 128 / (P.tw * P.th);
  P.tiles_w = (P.vW + P.tw - 1) / P.tw;
  P.tiles_h = (P.vH + P.th - 1) / P.th;
  P.tiles_b = (P.vB + P.nb - 1) / P.nb;

  int block_n = 64;
  if (d->n_total % 256 == 0) block_n = 256;
  else if (d->n_total % 192 == 0) block_n = 192;
  else if (d->n_total % 128 == 0) block_n = 128;
  P.n_tiles = d->n_total / block_n;
  P.n_total = d->n_total;
  P.num_phases = d->num_phases;
  for (int ph = 0; ph < d->num_phases; ++ph) {
    TVAE_REQUIRE(d->ntaps[ph] >= 1 && d->ntaps[ph] <= TVAE_MAX_TAPS, "mtgemm: bad ntaps[%d]=%d", ph, d->ntaps[ph]);
    P.ntaps[ph] = d->ntaps[ph];
    P.out_p[ph] = d->out_p[ph];
    P.out_c_off[ph] = d->out_c_off[ph];
    for (int t = 0; t < d->ntaps[ph]; ++t) {
      const tvae_tap& s = d->taps[ph][t];
      const tvae_view& av = s.map ? d->a1 : d->a0;
      TVAE_REQUIRE(av.ptr != nullptr, "mtgemm: tap uses absent view %d", s.map);
      TVAE_REQUIRE(s.kblocks >= 1 && s.c_off >= 0 && s.c_off + s.kblocks * 64 <= (av.split ? 2 : 1) * av.C,
                   "mtgemm: tap channel range [%d, +%d*64) exceeds view", s.c_off, s.kblocks);
      TVAE_REQUIRE(s.wk_off >= 0 && s.wk_off + s.kblocks * 64 <= d->k_total, "mtgemm: tap weight range out of bounds");
      MtTap& o = P.taps[ph][t];
      o.c_off = s.c_off; o.wk_off = s.wk_off; o.kblocks = (int16_t)s.kblocks;
      o.map = (int8_t)s.map; o.dw = (int8_t)s.dw; o.p = (int8_t)s.p; o.dh = (int8_t)s.dh;
    }
  }
  P.bias = d->bias;
  P.act = d->act;
  P.row_scale = d->row_scale;
  P.row_shift = d->row_shift;
  P.col_sum = d->col_sum;
  TVAE_REQUIRE(d->row_shift == nullptr || d->col_sum != nullptr, "mtgemm: row_shift needs col_sum");
  P.has_residual = d->res.ptr != nullptr;
  static const int pipe_res_env = !(getenv("TVAE_EPI_PIPE_RES") && atoi(getenv("TVAE_EPI_PIPE_RES")) == 0);   // A/B switch
  P.pipe_res = pipe_res_env;
  P.rope_tab = reinterpret_cast<const float2*>(d->rope_tab);
  P.rope_C = d->rope_C; P.rope_H = d->rope_H; P.rope_W = d->rope_W;
  P.q_scale = d->q_scale;
  P.out_f32 = d->out_f32;
  P.out_n = d->out_n;
  P.z = reinterpret_cast<const __nv_bfloat16*>(d->z);
  TVAE_REQUIRE(!(direct && P.has_residual), "mtgemm: residual not supported with direct fp32 store");

  CUtensorMap mA0, mA1, mB, mO, mR;
  int rc;
  if ((rc = make_tmap_pix(&mA0, d->a0.ptr, d->a0.B, d->a0.H, d->a0.W, d->a0.C, d->a0.split, P.tw, P.th, P.nb))) return rc;
  if (d->a1.ptr) {
    if ((rc = make_tmap_pix(&mA1, d->a1.ptr, d->a1.B, d->a1.H, d->a1.W, d->a1.C, d->a1.split, P.tw, P.th, P.nb))) return rc;
  } else {
    mA1 = mA0;
  }
  if ((rc = make_tmap_2d(&mB, d->w, d->n_total, d->k_total, d->k_total, block_n))) return rc;
  if (!direct) {
    TVAE_REQUIRE(((d->out.split ? 2 : 1) * d->out.C) % 64 == 0, "mtgemm: output channels must be a multiple of 64");
    if ((rc = make_tmap_pix(&mO, d->out.ptr, d->out.B, d->out.H, d->out.W, d->out.C, d->out.split, P.tw, P.th, P.nb))) return rc;
  } else {
    mO = mA0;
  }
  if (P.has_residual) {
    if ((rc = make_tmap_pix(&mR, d->res.ptr, d->res.B, d->res.H, d->res.W, d->res.C, d->res.split, P.tw, P.th, P.nb))) return rc;
  } else {
    mR = mO;
  }
  // pick the compile-time epilogue variant
  const bool affine = d->row_scale != nullptr || d->row_shift != nullptr;
  int epi;
  if (d->act_grad != 0) {
    TVAE_REQUIRE(!direct && !affine && d->bias == nullptr && d->rope_tab == nullptr && P.has_residual,
                 "mtgemm: act_grad excludes bias / affine / rope / direct store and needs `res`");
    if (d->act_grad == 1) {
      TVAE_REQUIRE(d->act == TVAE_ACT_GELU || d->act == TVAE_ACT_SILU, "mtgemm: act_grad needs GELU or SiLU");
      epi = d->act == TVAE_ACT_GELU ? kEpiMulGeluGrad : kEpiMulSiluGrad;
    } else {
      TVAE_REQUIRE(d->act_grad == 2 && d->act == TVAE_ACT_GELU && d->z != nullptr && !d->out.split,
                   "mtgemm: act_grad 2 = (acc + res) * gelu'(z) on a plain output view");
      epi = kEpiResMulGeluGrad;
    }
  } else if (direct) {
    TVAE_REQUIRE(!affine && d->act == TVAE_ACT_NONE && d->rope_tab == nullptr, "mtgemm: direct store supports bias only");
    epi = kEpiDirect;
  } else if (d->rope_tab != nullptr) {
    TVAE_REQUIRE(d->act == TVAE_ACT_NONE && !P.has_residual, "mtgemm: rope epilogue excludes act / residual");
    epi = kEpiAffineRope;
  } else if (affine) {
    if (d->act == TVAE_ACT_GELU && d->row_shift == nullptr && !P.has_residual) epi = kEpiRsBiasGelu;
    else {
      TVAE_REQUIRE(d->act == TVAE_ACT_NONE, "mtgemm: row-affine epilogue supports GELU (row_scale only) or no activation");
      epi = kEpiRsBias;
    }
  } else if (P.has_residual) {
    TVAE_REQUIRE(d->act == TVAE_ACT_NONE, "mtgemm: residual epilogue excludes an activation");
    epi = kEpiBiasRes;
  } else {
    epi = d->act == TVAE_ACT_GELU ? kEpiBiasGelu : (d->act == TVAE_ACT_SILU ? kEpiBiasSilu : kEpiBias);
  }
  // fused GroupNorm-backward reduce (input-gradient GEMM of a ResBlock): plain epilogue only
  const bool gnb = d->gnb_part != nullptr;
  if (gnb) {
    TVAE_REQUIRE(d->gnb_x != nullptr && d->gnb_sums != nullptr && d->gnb_gamma != nullptr && d->gnb_beta != nullptr &&
                     d->gnb_groups >= 1 && d->n_total % d->gnb_groups == 0 && d->out.C == d->n_total && !d->out.split,
                 "mtgemm: gnb_* needs x / sums / gamma / beta and a plain output view with n_total channels");
    TVAE_REQUIRE(epi == kEpiBias && d->bias == nullptr && d->gn_sums == nullptr && d->out_act == nullptr,
                 "mtgemm: the fused GroupNorm-backward reduce needs a plain epilogue");
  }
  // dual output (training forward): `out` = pre-activation, `out_act` = activation of the stored bf16 values
  const bool dual = d->out_act != nullptr;
  if (dual) {
    TVAE_REQUIRE(d->act_grad == 0 && !direct && !affine && d->rope_tab == nullptr && !P.has_residual &&
                     d->gn_sums == nullptr && (d->act == TVAE_ACT_GELU || d->act == TVAE_ACT_SILU),
                 "mtgemm: out_act needs a plain bias + GELU / SiLU epilogue");
    epi = kEpiBias;          // until the CTA-pair kernel is chosen below
  }
  // CTA-pair kernel (cta_group::2, weight tile split across the pair) whenever the tile is wide enough to matter
  static const bool use_pair = !(getenv("TVAE_2CTA") && atoi(getenv("TVAE_2CTA")) == 0);
  // halo mode: a plain 3x3 stride-1 convolution whose tiles are 128 pixels of one image row (W >= 128) loads ONE
  // 130-pixel tile per (kernel row, K block) instead of three shifted 128-pixel tiles (TVAE_HALO=0 disables: A/B switch)
  static const int halo_env = getenv("TVAE_HALO") ? atoi(getenv("TVAE_HALO")) : 1;
  bool halo_ok = halo_env && P.tw == 128 && d->num_phases == 1 && d->ntaps[0] == 9 && !d->a0.split;
  for (int t = 0; halo_ok && t < 9; ++t) {       // three kernel rows of three taps: same dh, dw a permutation of {-1, 0, 1}
    const tvae_tap& s = d->taps[0][t];
    halo_ok = s.map == 0 && s.p == 0 && s.c_off == 0 && s.dw >= -1 && s.dw <= 1 && s.dh == d->taps[0][t - t % 3].dh &&
              s.kblocks == d->taps[0][0].kblocks && s.kblocks * 64 == d->a0.C;
    if (halo_ok && t % 3 == 2)
      halo_ok = d->taps[0][t].dw + d->taps[0][t - 1].dw + d->taps[0][t - 2].dw == 0 &&
                d->taps[0][t].dw != d->taps[0][t - 1].dw && d->taps[0][t - 1].dw != d->taps[0][t - 2].dw;
  }
  // narrow outputs (N = 64: the 3-channel output convolution, padded) are bound by the L2->SM traffic of their A tiles,
  // so they take the pair kernel too when the halo tiles cut that traffic by 3x
  const bool pair = use_pair && (block_n >= 128 || halo_ok) && P.tiles_w * P.tiles_h * P.tiles_b >= 2;
  // GroupNorm statistics of the output: from the epilogue when the launch qualifies, else one tvae_groupnorm_stats pass
  bool gn_fused = false;
  if (d->gn_sums != nullptr) {
    TVAE_REQUIRE(!direct && d->gn_groups >= 1 && d->n_total % d->gn_groups == 0 && d->out.C == d->n_total,
                 "mtgemm: gn_sums needs a bf16 output with n_total = %d channels divisible into %d groups", d->n_total,
                 d->gn_groups);
    static const bool allow_fused = !(getenv("TVAE_GN_FUSED") && atoi(getenv("TVAE_GN_FUSED")) == 0);
    gn_fused = allow_fused && pair && P.nb == 1 && (epi == kEpiBias || epi == kEpiBiasRes) && d->gn_groups <= 64;
    if (gn_fused) {
      TVAE_CHECK_CUDA(cudaMemsetAsync(d->gn_sums, 0, (size_t)d->out.B * d->gn_groups * 2 * sizeof(double), stream));
      P.gn_sums = d->gn_sums;
      P.gn_groups = d->gn_groups;
      P.gn_cpg = d->n_total / d->gn_groups;
    }
  }
  if (pair) {
    CUtensorMap mBh, mAh = mA0, mX2 = mO;
    // The second staging tile of the dual epilogue costs pipeline stages (N = 256: 5 -> 3): worth it where the epilogue /
    // HBM side sets the pace (K <= 4096), not for the long-K 3x3 convolutions (K = 13824: 1240 -> 1020 TFLOP/s, against
    // a 10 us activation pass) -- those keep the deep pipeline and run tvae_act_fwd behind the GEMM.
    // fused GroupNorm-backward reduce: one image per 128-pixel tile, the n tiles cover the channels exactly
    static const bool allow_gnb = !(getenv("TVAE_GNB_FUSED") && atoi(getenv("TVAE_GNB_FUSED")) == 0);
    const bool gnb_fused = gnb && allow_gnb && P.nb == 1 && d->n_total % block_n == 0 && d->gnb_groups <= 128;
    float* gnb_ws = nullptr;
    const int m_tiles_all = P.tiles_w * P.tiles_h * P.tiles_b;
    const int tiles_per_image = P.tiles_w * P.tiles_h;
    if (gnb_fused) {
      const size_t rows = (size_t)m_tiles_all + (size_t)d->out.B * ((tiles_per_image + 31) / 32);
      if ((rc = scratch_workspace(rows * 2 * d->n_total * sizeof(float), reinterpret_cast<void**>(&gnb_ws)))) return rc;
      epi = kEpiGnBwd;
      if ((rc = make_tmap_pix(&mX2, d->gnb_x, d->out.B, d->out.H, d->out.W, d->out.C, d->out.split, P.tw, P.th, P.nb))) return rc;
      P.gnb_x = d->gnb_x;
      P.gnb_sums = d->gnb_sums;
      P.gnb_gamma = d->gnb_gamma;
      P.gnb_beta = d->gnb_beta;
      P.gnb_part = gnb_ws;
      P.gnb_groups = d->gnb_groups;
      P.gnb_cpg = d->n_total / d->gnb_groups;
      P.gnb_silu = d->gnb_silu;
      P.gnb_eps = d->gnb_eps;
      P.gnb_inv_n = 1.0f / ((float)(d->n_total / d->gnb_groups) * (float)(d->out.H * d->out.W));
    }
    const bool dual_fused = dual && d->k_total <= 4096;
    if (dual_fused) {
      epi = d->act == TVAE_ACT_GELU ? kEpiBiasGeluDual : kEpiBiasSiluDual;
      if ((rc = make_tmap_pix(&mX2, d->out_act, d->out.B, d->out.H, d->out.W, d->out.C, d->out.split, P.tw, P.th, P.nb))) return rc;
    }
    if ((rc = make_tmap_2d(&mBh, d->w, d->n_total, d->k_total, d->k_total, block_n / 2))) return rc;
    if (epi == kEpiResMulGeluGrad) {
      // the pair kernel stages z (same plain view as the output) through TMA next to the residual
      if ((rc = make_tmap_pix(&mX2, d->z, d->out.B, d->out.H, d->out.W, d->out.C, d->out.split, P.tw, P.th, P.nb))) return rc;
    }
    if (halo_ok) {
      if ((rc = make_tmap_pix_halo(&mAh, d->a0.ptr, d->a0.B, d->a0.H, d->a0.W, d->a0.C))) return rc;
      P.halo = 1;
    }
    if ((rc = mtgemm2_dispatch(epi, block_n, mA0, mA1, mBh, mO, mR, mAh, mX2, P, stream))) return rc;
    if (dual && !dual_fused)
      return act_fwd_run(d->out.ptr, d->out_act, (long long)d->out.B * d->out.H * d->out.W * d->out.C, d->act, stream);
    if (gnb_fused)
      return gn_bwd_tiles_reduce_run(gnb_ws, gnb_ws + (size_t)m_tiles_all * 2 * d->n_total, d->gnb_part, d->out.B,
                                     tiles_per_image, 2 * d->n_total, stream);
    if (gnb)
      return gn_bwd_reduce_run(d->gnb_x, d->out.ptr, d->gnb_sums, d->gnb_gamma, d->gnb_beta, d->gnb_part, d->out.B,
                               d->out.H * d->out.W, d->out.C, d->gnb_groups, d->gnb_eps, d->gnb_silu, stream);
    if (d->gn_sums != nullptr && !gn_fused)
      return gn_stats_run(d->out.ptr, d->gn_sums, d->out.B, d->out.H * d->out.W, d->out.C, d->gn_groups, stream);
    return 0;
  }
  if (d->gn_sums != nullptr) {
    if ((rc = mtgemm1_dispatch(epi, block_n, mA0, mA1, mB, mO, mR, P, stream))) return rc;
    return gn_stats_run(d->out.ptr, d->gn_sums, d->out.B, d->out.H * d->out.W, d->out.C, d->gn_groups, stream);
  }
  if ((rc = mtgemm1_dispatch(epi, block_n, mA0, mA1, mB, mO, mR, P, stream))) return rc;
  if (dual)    // one-CTA fallback: the activation as its own pass over the (complete, contiguous) output tensor
    return act_fwd_run(d->out.ptr, d->out_act, (long long)d->out.B * d->out.H * d->out.W * d->out.C, d->act, stream);
  if (gnb)     // one-CTA fallback: the stand-alone reduce pass
    return gn_bwd_reduce_run(d->gnb_x, d->out.ptr, d->gnb_sums, d->gnb_gamma, d->gnb_beta, d->gnb_part, d->out.B,
                             d->out.H * d->out.W, d->out.C, d->gnb_groups, d->gnb_eps, d->gnb_silu, stream);
  return 0;
}

static int mtgemm1_dispatch(int epi, int block_n, const CUtensorMap& mA0, const CUtensorMap& mA1, const CUtensorMap& mB,
                            const CUtensorMap& mO, const CUtensorMap& mR, const MtParams& P, cudaStream_t stream) {
  switch (epi) {
    case kEpiBias: return launch_n<kEpiBias>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiBiasRes: return launch_n<kEpiBiasRes>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiBiasGelu: return launch_n<kEpiBiasGelu>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiBiasSilu: return launch_n<kEpiBiasSilu>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiRsBiasGelu: return launch_n<kEpiRsBiasGelu>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiAffineRope: return launch_n<kEpiAffineRope>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiDirect: return launch_n<kEpiDirect>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiMulGeluGrad: return launch_n<kEpiMulGeluGrad>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiMulSiluGrad: return launch_n<kEpiMulSiluGrad>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    case kEpiResMulGeluGrad: return launch_n<kEpiResMulGeluGrad>(block_n, mA0, mA1, mB, mO, mR, P, stream);
    default: return launch_n<kEpiRsBias>(block_n, mA0, mA1, mB, mO, mR, P, stream);
  }
}

}  // namespace tvae
