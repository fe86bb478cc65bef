// HBM-bound kernels of the backward pass and the materialised training-path norms (vectorised 128-bit accesses,
// warp-shuffle / shared-memory reductions, fp32 statistics).  Activations and their gradients are NHWC bf16.
#include "../../include/transvae_sm100.h"
#include "common.cuh"
#include "ew_common.cuh"

namespace tvae {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}

// -------------------------------------------------------------------------------------------------
// dZ = dY * act'(Z)  and  colsum[p][q] += sum dZ  over a 4-D contiguous bf16 tensor [R0, P, R1, Q]
// (plain [M, N] matrix: R0 = M, P = R1 = 1, Q = N; phase view of [B, 2H, 2W, C]: R0 = B*H, P = 2, R1 = W, Q = 2C).
// Backward of the bias add + GELU / SiLU epilogues of tvae_mtgemm (conv.py:86,56,58; upsample.py:35,96).
// Algorithmic bytes per element: 2 (dY) [+ 2 (Z) + 2 (dZ) when act != none].
// -------------------------------------------------------------------------------------------------
// Processes columns [c0, c0+Qs) of the row-major [R0*Pn*R1, Qfull] matrix; row r belongs to phase (r / R1) % Pn.
constexpr int kEwBatch = 4;   // independent 16-byte loads per stream a thread keeps in flight

__device__ unsigned int g_bab_tickets[1 + kOrdMaxGroups];

template <int ACT>
__global__ void __launch_bounds__(256) bias_act_bwd_slab_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ z,
                                                                uint4* __restrict__ dz, float* __restrict__ colsum,
                                                                float* __restrict__ ws, long long nrows, int Pn, int R1,
                                                                int Qfull, int c0, int Qs, int rows_per_block) {
  extern __shared__ __align__(16) float s_col[];   // [row lanes of the block][Pn][Qs]
  __shared__ bool s_flag;
  const int nvec = Qs >> 3, nvf = Qfull >> 3, v0 = c0 >> 3;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec, rpp = blockDim.x / nvec;
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  const long long r_end = min(nrows, r_begin + rows_per_block);
  float2 acc[2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = f2(0.0f);
  if (rl < rpp) {
    for (long long r = r_begin + rl; r < r_end; r += (long long)rpp * kEwBatch) {
      uint4 ug[kEwBatch], uz[kEwBatch];
#pragma unroll
      for (int k = 0; k < kEwBatch; ++k) {
        const long long rr = r + (long long)k * rpp;
        if (rr < r_end) {
          ug[k] = __ldg(dy + rr * nvf + v0 + v);
          if (ACT != TVAE_ACT_NONE) uz[k] = __ldg(z + rr * nvf + v0 + v);
        }
      }
#pragma unroll
      for (int k = 0; k < kEwBatch; ++k) {
        const long long rr = r + (long long)k * rpp;
        if (rr < r_end) {
          const int p = (Pn == 1) ? 0 : (int)((rr / R1) % Pn);
          float2 g[4];
          unpack8_2(ug[k], g);
          if (ACT != TVAE_ACT_NONE) {
            float2 zz[4];
            unpack8_2(uz[k], zz);
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = __fmul2_rn(g[i], ACT == TVAE_ACT_GELU ? gelu_grad2(zz[i]) : silu_grad2(zz[i]));
            const uint4 o = pack8_2(g);
            dz[rr * nvf + v0 + v] = o;
            unpack8_2(o, g);   // the column sums must see the bf16-rounded dZ that the GEMMs will read
          }
          if (p & 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[1][i] = __fadd2_rn(acc[1][i], g[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][i] = __fadd2_rn(acc[0][i], g[i]);
          }
        }
      }
    }
    // column sums: fixed-order reduction (row lanes of the block in order, then the blocks: ordered_rows_reduce)
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
      if (pp < Pn) {
        float* p = s_col + ((size_t)rl * Pn + pp) * Qs + v * 8;
        *reinterpret_cast<float4*>(p) = make_float4(acc[pp][0].x, acc[pp][0].y, acc[pp][1].x, acc[pp][1].y);
        *reinterpret_cast<float4*>(p + 4) = make_float4(acc[pp][2].x, acc[pp][2].y, acc[pp][3].x, acc[pp][3].y);
      }
  }
  if (colsum == nullptr) return;          // uniform
  __syncthreads();
  const int n = Pn * Qs;
  float* rows = ws;                                            // [gridDim.x][Pn * Qs]
  float* groups = ws + (size_t)gridDim.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float a = 0.0f;
    for (int r = 0; r < rpp; ++r) a += s_col[(size_t)r * n + i];
    rows[(size_t)blockIdx.x * n + i] = a;
  }
  ordered_rows_reduce(rows, groups, g_bab_tickets, colsum, n, gridDim.x, blockIdx.x, &s_flag,
                      [=](int i) { return (size_t)(i / Qs) * Qfull + c0 + (i % Qs); });
}

int bias_act_bwd_run(const void* dy, const void* z, void* dz, float* colsum, long long R0, int Pn, int R1, int Q, int act,
                     cudaStream_t stream) {
  TVAE_REQUIRE(Q % 8 == 0, "bias_act_bwd: Q=%d must be a multiple of 8", Q);
  TVAE_REQUIRE(Pn == 1 || Pn == 2, "bias_act_bwd: P must be 1 or 2");
  TVAE_REQUIRE(act == TVAE_ACT_NONE || (z != nullptr && dz != nullptr), "bias_act_bwd: activation needs z and dz");
  TVAE_REQUIRE(act == TVAE_ACT_NONE || act == TVAE_ACT_GELU || act == TVAE_ACT_SILU, "bias_act_bwd: unknown activation %d", act);
  const long long nrows = R0 * Pn * R1;
  for (int c0 = 0; c0 < Q; c0 += 2048) {
    const int qs = (Q - c0) < 2048 ? (Q - c0) : 2048;
    const int nvec = qs / 8;
    const int threads = nvec >= 256 ? 256 : (256 / nvec) * nvec;
    const int rpp = threads / nvec > 0 ? threads / nvec : 1;
    // ~4 resident blocks per SM, each sweeping a contiguous run of rows: the column sums leave a block as one global
    // atomic per column, so fewer, longer blocks mean fewer contended atomics (thousands of short blocks put
    // 6500 atomics on each of 192 addresses)
    const long long unit = (long long)rpp * kEwBatch;
    long long rpb = ((nrows + 4LL * num_sms() - 1) / (4LL * num_sms()) + unit - 1) / unit * unit;
    if (rpb < unit) rpb = unit;
    const int grid = (int)((nrows + rpb - 1) / rpb);
    TVAE_REQUIRE(grid <= (int)(kOrdGroup * kOrdMaxGroups), "bias_act_bwd: grid %d too large", grid);
    const size_t smem = (size_t)rpp * Pn * qs * sizeof(float);          // <= 16 KiB
    float* ws = nullptr;
    if (colsum != nullptr) {
      const size_t need = (size_t)(grid + (grid + kOrdGroup - 1) / kOrdGroup) * Pn * qs * sizeof(float);
      const int rc = scratch_workspace(need, reinterpret_cast<void**>(&ws));
      if (rc) return rc;
    }
    auto args = [&](auto kern) {
      kern<<<grid, threads, smem, stream>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(z),
                                            reinterpret_cast<uint4*>(dz), colsum, ws, nrows, Pn, R1, Q, c0, qs, (int)rpb);
    };
    if (act == TVAE_ACT_GELU) args(bias_act_bwd_slab_kernel<TVAE_ACT_GELU>);
    else if (act == TVAE_ACT_SILU) args(bias_act_bwd_slab_kernel<TVAE_ACT_SILU>);
    else args(bias_act_bwd_slab_kernel<TVAE_ACT_NONE>);
    TVAE_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int bias_act_bwd_matrix_run(const void* dy, const void* z, void* dz, float* colsum, long long M, int N, int act,
                            cudaStream_t stream) {
  return bias_act_bwd_run(dy, z, dz, colsum, M, 1, 1, N, act, stream);
}

// y = act(z) elementwise (training path: the GEMM stores the pre-activation z that the backward pass needs, the
// activation output is produced by this pass).  Algorithmic bytes: 4 per element.
template <int ACT>
__global__ void __launch_bounds__(256) act_fwd_kernel(const uint4* __restrict__ z, uint4* __restrict__ y, long long n8) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride * kEwBatch) {
    uint4 u[kEwBatch];
#pragma unroll
    for (int k = 0; k < kEwBatch; ++k)
      if (i + k * stride < n8) u[k] = __ldg(z + i + k * stride);
#pragma unroll
    for (int k = 0; k < kEwBatch; ++k) {
      if (i + k * stride < n8) {
        float2 f[4];
        unpack8_2(u[k], f);
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = (ACT == TVAE_ACT_GELU) ? gelu2(f[q]) : silu2(f[q]);
        y[i + k * stride] = pack8_2(f);
      }
    }
  }
}

int act_fwd_run(const void* z, void* y, long long n, int act, cudaStream_t stream) {
  TVAE_REQUIRE(n % 8 == 0, "act_fwd: element count must be a multiple of 8");
  TVAE_REQUIRE(act == TVAE_ACT_GELU || act == TVAE_ACT_SILU, "act_fwd: unknown activation %d", act);
  long long grid = (n / 8 + 256 * kEwBatch - 1) / (256 * kEwBatch);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  if (grid < 1) grid = 1;
  if (act == TVAE_ACT_GELU)
    act_fwd_kernel<TVAE_ACT_GELU><<<(int)grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(z), reinterpret_cast<uint4*>(y), n / 8);
  else
    act_fwd_kernel<TVAE_ACT_SILU><<<(int)grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(z), reinterpret_cast<uint4*>(y), n / 8);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// GroupNorm(+SiLU) backward.  h = act(y), y = xhat*gamma + beta, xhat = (x - mean_g) * rstd_g.
//   pass A: part[b][c] = (sum_p dy, sum_p dy*xhat) with dy = dh * act'(y)            (reads x, dh)
//   pass B: dx = rstd * (dy*gamma - m1_g - xhat*m2_g) [+ add],  m1_g = mean_g(dy*gamma), m2_g = mean_g(dy*gamma*xhat)
// dgamma[c] = sum_b part[b][c][1], dbeta[c] = sum_b part[b][c][0] (tiny, reduced by the caller).
// Backward of nn.GroupNorm + F.silu (blocks.py:60-66; decoder.py:128-129).  Algorithmic bytes: 10*C per pixel.
// -------------------------------------------------------------------------------------------------
struct GnCoef {
  float mean, rstd;
};

__device__ __forceinline__ void gn_group_coef(const double* sums, int b, int G, int g, float inv_n, float eps, float& mean,
                                              float& rstd) {
  gn_mean_rstd(sums, b, G, g, inv_n, eps, mean, rstd);
}

// per-thread constants of its 8 channels, packed in pairs: xhat = x * a + b (a = rstd, b = -mean * rstd), y = xhat * ga + be
struct GnChan {
  float2 a[4], b[4], ga[4], be[4];
};
__device__ __forceinline__ void gn_load_chan(GnChan& ch, const double* sums, const float* gamma, const float* beta, int b,
                                             int G, int cpg, int v, float inv_n, float eps) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m0, r0, m1, r1;
    const int c = v * 8 + 2 * i;
    gn_group_coef(sums, b, G, c / cpg, inv_n, eps, m0, r0);
    gn_group_coef(sums, b, G, (c + 1) / cpg, inv_n, eps, m1, r1);
    ch.a[i] = make_float2(r0, r1);
    ch.b[i] = make_float2(-m0 * r0, -m1 * r1);
    ch.ga[i] = make_float2(gamma[c], gamma[c + 1]);
    ch.be[i] = make_float2(beta[c], beta[c + 1]);
  }
}

constexpr int kGnMaxImages = 4096;
__device__ unsigned int g_gn_ticket[kGnMaxImages];     // zero-initialised; every launch leaves it at zero

template <bool SILU>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dh,
                                                            const double* __restrict__ sums,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ part, float* __restrict__ ws, int HW, int C,
                                                            int G, float eps, int pix_per_block) {
  extern __shared__ __align__(16) float s_stage[];   // [pixel rows of the block][C][2]
  __shared__ bool s_last;
  const int nvec = C >> 3, cpg = C / G;
  const int b = blockIdx.y;
  const int v = threadIdx.x % nvec, pv = threadIdx.x / nvec, ppb = blockDim.x / nvec;
  const float inv_n = 1.0f / ((float)cpg * (float)HW);
  GnChan ch;
  gn_load_chan(ch, sums, gamma, beta, b, G, cpg, v, inv_n, eps);
  float2 s1[4], s2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s1[i] = s2[i] = f2(0.0f);
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const size_t base = (size_t)b * HW * nvec + v;
  for (int p = p0 + pv; p < p1; p += ppb * kEwBatch) {
    uint4 ux[kEwBatch], ug[kEwBatch];
#pragma unroll
    for (int k = 0; k < kEwBatch; ++k) {
      const int pp = p + k * ppb;
      if (pp < p1) {
        ux[k] = __ldg(x + base + (size_t)pp * nvec);
        ug[k] = __ldg(dh + base + (size_t)pp * nvec);
      } else {
        ux[k] = ug[k] = make_uint4(0, 0, 0, 0);        // dh = 0 contributes nothing
      }
    }
#pragma unroll
    for (int k = 0; k < kEwBatch; ++k) {
      float2 xf[4], g[4];
      unpack8_2(ux[k], xf);
      unpack8_2(ug[k], g);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 xh = __ffma2_rn(xf[i], ch.a[i], ch.b[i]);
        const float2 dy = SILU ? __fmul2_rn(g[i], silu_grad2(__ffma2_rn(xh, ch.ga[i], ch.be[i]))) : g[i];
        s1[i] = __fadd2_rn(s1[i], dy);
        s2[i] = __ffma2_rn(dy, xh, s2[i]);
      }
    }
  }
  // Fixed-order reduction, no atomics on the values (two backward passes are bit-identical): the threads of a block stage
  // their 16 partials, one thread per (channel, statistic) adds the block's pixel rows in order and stores the block
  // partial; the last block of the image to arrive (ticket counter) adds the block partials in block order.  (History:
  // per-thread global atomics held the kernel at 2.1 TB/s; shared + one global atomic per block ran at 4.4 TB/s but made
  // GroupNorm-backward the first kernel of the step whose output changed from run to run.)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* pp = s_stage + (size_t)pv * 2 * C + (v * 8 + 2 * i) * 2;
    *reinterpret_cast<float4*>(pp) = make_float4(s1[i].x, s2[i].x, s1[i].y, s2[i].y);
  }
  __syncthreads();
  float* mine = ws + ((size_t)b * gridDim.x + blockIdx.x) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float a = 0.0f;
    for (int r = 0; r < ppb; ++r) a += s_stage[(size_t)r * 2 * C + i];
    mine[i] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(&g_gn_ticket[b], gridDim.x - 1);   // wraps to 0: ready for the next launch
    s_last = t == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const float* all = ws + (size_t)b * gridDim.x * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
      part[(size_t)b * C * 2 + i] = ordered_column_sum(all + i, 2 * (size_t)C, gridDim.x);
  }
}

template <bool SILU, bool ADD>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dh,
                                                           const uint4* __restrict__ add, const double* __restrict__ sums,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ part, uint4* __restrict__ dx, int HW,
                                                           int C, int G, float eps, int vec_per_block) {
  extern __shared__ float s_g[];  // per group: m1, m2
  const int nvec = C >> 3, cpg = C / G;
  const int b = blockIdx.y;
  const float inv_n = 1.0f / ((float)cpg * (float)HW);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float m1 = 0.0f, m2 = 0.0f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      m1 = fmaf(gamma[c], part[((size_t)b * C + c) * 2], m1);
      m2 = fmaf(gamma[c], part[((size_t)b * C + c) * 2 + 1], m2);
    }
    s_g[2 * g] = m1 * inv_n;
    s_g[2 * g + 1] = m2 * inv_n;
  }
  __syncthreads();
  const int v = threadIdx.x % nvec;
  GnChan ch;
  gn_load_chan(ch, sums, gamma, beta, b, G, cpg, v, inv_n, eps);
  // dx = rstd * (dy * gamma - m1 - xhat * m2) = dy * A - M1 - xhat * M2 with A = rstd * gamma, M1 = rstd * m1, M2 = rstd * m2
  float2 A[4], nM1[4], nM2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = v * 8 + 2 * i, g0 = c / cpg, g1 = (c + 1) / cpg;
    A[i] = __fmul2_rn(ch.a[i], ch.ga[i]);
    nM1[i] = make_float2(-ch.a[i].x * s_g[2 * g0], -ch.a[i].y * s_g[2 * g1]);
    nM2[i] = make_float2(-ch.a[i].x * s_g[2 * g0 + 1], -ch.a[i].y * s_g[2 * g1 + 1]);
  }
  const long long total = (long long)HW * nvec;
  const long long i0 = (long long)blockIdx.x * vec_per_block, i1 = min(total, i0 + vec_per_block);
  const size_t off = (size_t)b * total;
  const long long stride = blockDim.x;
  for (long long i = i0 + threadIdx.x; i < i1; i += stride * kEwBatch) {
    uint4 ux[kEwBatch], ug[kEwBatch], ur[kEwBatch];
#pragma unroll
    for (int k = 0; k < kEwBatch; ++k) {
      const long long ii = i + k * stride;
      if (ii < i1) {
        ux[k] = __ldg(x + off + ii);
        ug[k] = __ldg(dh + off + ii);
        if (ADD) ur[k] = __ldg(add + off + ii);
      }
    }
#pragma unroll
    for (int k = 0; k < kEwBatch; ++k) {
      const long long ii = i + k * stride;
      if (ii < i1) {
        float2 xf[4], g[4], r[4];
        unpack8_2(ux[k], xf);
        unpack8_2(ug[k], g);
        if (ADD) unpack8_2(ur[k], r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 xh = __ffma2_rn(xf[q], ch.a[q], ch.b[q]);
          const float2 dy = SILU ? __fmul2_rn(g[q], silu_grad2(__ffma2_rn(xh, ch.ga[q], ch.be[q]))) : g[q];
          float2 o = __ffma2_rn(dy, A[q], nM1[q]);
          o = __ffma2_rn(xh, nM2[q], o);
          g[q] = ADD ? __fadd2_rn(o, r[q]) : o;
        }
        dx[off + ii] = pack8_2(g);
      }
    }
  }
}

static int gn_bwd_check(int B, int C, int G) {
  TVAE_REQUIRE(C % 8 == 0 && C % G == 0 && C / 8 <= 256 && G <= 128, "groupnorm_bwd: unsupported C=%d G=%d", C, G);
  TVAE_REQUIRE(B <= kGnMaxImages, "groupnorm_bwd: batch %d exceeds %d", B, kGnMaxImages);
  return 0;
}

// reduce pass: part [B][C][2] = per-(image, channel) (sum dy, sum dy * xhat)
int gn_bwd_reduce_run(const void* x, const void* dh, const double* sums, const float* gamma, const float* beta, float* part,
                      int B, int HW, int C, int G, float eps, int apply_silu, cudaStream_t stream) {
  if (int rc = gn_bwd_check(B, C, G)) return rc;
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  int ppb = 1024;
  while (ppb > 64 && (long long)((HW + ppb - 1) / ppb) * B < 8LL * num_sms()) ppb >>= 1;
  dim3 g1((HW + ppb - 1) / ppb, B);
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  const uint4* dp = reinterpret_cast<const uint4*>(dh);
  const size_t smem_r = (size_t)(threads / nvec) * 2 * C * sizeof(float);       // <= 16 KiB
  float* ws = nullptr;
  {
    const int rc = scratch_workspace((size_t)B * g1.x * 2 * C * sizeof(float), reinterpret_cast<void**>(&ws));
    if (rc) return rc;
  }
  if (apply_silu) gn_bwd_reduce_kernel<true><<<g1, threads, smem_r, stream>>>(xp, dp, sums, gamma, beta, part, ws, HW, C, G, eps, ppb);
  else gn_bwd_reduce_kernel<false><<<g1, threads, smem_r, stream>>>(xp, dp, sums, gamma, beta, part, ws, HW, C, G, eps, ppb);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Fixed-order sum of per-tile partial rows (the fused reduce of the input-gradient GEMM epilogue, mtgemm2.cu kEpiGnBwd):
// tiles [B][tiles_per_image][n] -> out [B][n] in two levels (groups of kGnTileGroup tiles, then the groups).
constexpr int kGnTileGroup = 32;
__global__ void __launch_bounds__(256) gn_bwd_tiles_reduce_kernel(const float* __restrict__ in, float* __restrict__ out, int n,
                                                                  int rows_per_image, int group) {
  const int b = blockIdx.z, g = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r0 = g * group;
  const int cnt = min(group, rows_per_image - r0);
  const float* p = in + ((size_t)b * rows_per_image + r0) * n + i;
  out[((size_t)b * gridDim.y + g) * n + i] = ordered_column_sum(p, (size_t)n, (unsigned int)cnt);
}

int gn_bwd_tiles_reduce_run(const float* part_tiles, float* groups_ws, float* part, int B, int tiles_per_image, int n,
                            cudaStream_t stream) {
  const int ngroups = (tiles_per_image + kGnTileGroup - 1) / kGnTileGroup;
  dim3 g1((n + 255) / 256, ngroups, B), g2((n + 255) / 256, 1, B);
  if (ngroups == 1) {
    gn_bwd_tiles_reduce_kernel<<<g2, 256, 0, stream>>>(part_tiles, part, n, tiles_per_image, tiles_per_image);
  } else {
    gn_bwd_tiles_reduce_kernel<<<g1, 256, 0, stream>>>(part_tiles, groups_ws, n, tiles_per_image, kGnTileGroup);
    gn_bwd_tiles_reduce_kernel<<<g2, 256, 0, stream>>>(groups_ws, part, n, ngroups, ngroups);
  }
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// apply pass: dx = dGN(dh) [+ add] from the per-(image, channel) sums in `part`
int gn_bwd_apply_run(const void* x, const void* dh, const void* add, const double* sums, const float* gamma, const float* beta,
                     const float* part, void* dx, int B, int HW, int C, int G, float eps, int apply_silu, cudaStream_t stream) {
  if (int rc = gn_bwd_check(B, C, G)) return rc;
  const int nvec = C / 8;
  const int threads = (256 / nvec) * nvec;
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  const uint4* dp = reinterpret_cast<const uint4*>(dh);
  const uint4* ap = reinterpret_cast<const uint4*>(add);
  const long long total = (long long)HW * nvec;
  long long vpb = (long long)threads * kEwBatch * 4;
  while (vpb > threads && ((total + vpb - 1) / vpb) * B < 8LL * num_sms()) vpb >>= 1;
  vpb = (vpb / threads) * threads;
  if (vpb < threads) vpb = threads;
  dim3 g2((unsigned)((total + vpb - 1) / vpb), B);
  const size_t smem = 2 * G * sizeof(float);
  uint4* op = reinterpret_cast<uint4*>(dx);
#define TVAE_GN_APPLY(S, A) \
  gn_bwd_apply_kernel<S, A><<<g2, threads, smem, stream>>>(xp, dp, ap, sums, gamma, beta, part, op, HW, C, G, eps, (int)vpb)
  if (apply_silu) {
    if (add) TVAE_GN_APPLY(true, true); else TVAE_GN_APPLY(true, false);
  } else {
    if (add) TVAE_GN_APPLY(false, true); else TVAE_GN_APPLY(false, false);
  }
#undef TVAE_GN_APPLY
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int gn_bwd_run(const void* x, const void* dh, const void* add, const double* sums, const float* gamma, const float* beta,
               float* part, void* dx, int B, int HW, int C, int G, float eps, int apply_silu, cudaStream_t stream) {
  if (int rc = gn_bwd_reduce_run(x, dh, sums, gamma, beta, part, B, HW, C, G, eps, apply_silu, stream)) return rc;
  return gn_bwd_apply_run(x, dh, add, sums, gamma, beta, part, dx, B, HW, C, G, eps, apply_silu, stream);
}

// -------------------------------------------------------------------------------------------------
// Token norms of the training path (materialised; the inference path folds them into the GEMM epilogue).
//   mode 0: y = x * rstd * w                                  (RMSNorm, blocks.py:168-201)
//   mode 1: h = x * rstd * w;  y = (h - mean(h)) / sqrt(var(h) + 1e-5)   (RMSNorm followed by the shared, affine-free
//           part of the three LayerNorms of attention.py:71-73; their gamma / beta are folded into the projection)
// One warp per token.  Forward bytes: 4*C per token; backward: 8*C (+2*C when `add`).
// -------------------------------------------------------------------------------------------------
constexpr int kMaxVecPerLane = 10;  // C <= 2560

// Interleaved warp reductions of N independent values (the shuffles of different rows / statistics overlap).
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}

// One warp normalises R rows at a time (VPL 16-byte vectors per lane and row, C <= 256 * VPL): all R * VPL loads are
// issued before the first reduction and the warp reductions of the R rows are interleaved, in a persistent
// grid-stride loop.  (The first version -- one short-lived warp per row, 16 K blocks per launch -- reached 1.6 TB/s.)
template <int VPL, int R>
__global__ void __launch_bounds__(256) token_norm_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                             uint4* __restrict__ y, long long M, int C, int mode) {
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const float invC = 1.0f / (float)C;
  float2 wv[VPL][4];
#pragma unroll
  for (int c = 0; c < VPL; ++c) {
    const int v = lane + c * 32;
    if (v < nvec) {
      const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
      const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
      wv[c][0] = make_float2(wa.x, wa.y); wv[c][1] = make_float2(wa.z, wa.w);
      wv[c][2] = make_float2(wb.x, wb.y); wv[c][3] = make_float2(wb.z, wb.w);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) wv[c][q] = f2(0.0f);
    }
  }
  const long long groups = (M + R - 1) / R;
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long grp = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); grp < groups; grp += warps_total) {
    const long long row0 = grp * R;
    uint4 u[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        const int v = lane + c * 32;
        u[r][c] = (v < nvec && row0 + r < M) ? __ldg(x + (row0 + r) * nvec + v) : make_uint4(0, 0, 0, 0);
      }
    float2 h[R][VPL][4];
    float s2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 acc = f2(0.0f);
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        unpack8_2(u[r][c], h[r][c]);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc = __ffma2_rn(h[r][c][q], h[r][c][q], acc);
      }
      s2[r] = acc.x + acc.y;
    }
    warp_sum_n<R>(s2);
    float st[2 * R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 rstd = f2(rsqrtf(s2[r] * invC + 1e-6f));
      float2 sm = f2(0.0f), sq = f2(0.0f);
#pragma unroll
      for (int c = 0; c < VPL; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          h[r][c][q] = __fmul2_rn(h[r][c][q], __fmul2_rn(rstd, wv[c][q]));
          sm = __fadd2_rn(sm, h[r][c][q]);
          sq = __ffma2_rn(h[r][c][q], h[r][c][q], sq);
        }
      st[2 * r] = sm.x + sm.y;
      st[2 * r + 1] = sq.x + sq.y;
    }
    if (mode == 1) {
      warp_sum_n<2 * R>(st);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float mu = st[2 * r] * invC;
        const float rs = rsqrtf(fmaxf(st[2 * r + 1] * invC - mu * mu, 0.0f) + 1e-5f);
        const float2 a = f2(rs), bsh = f2(-mu * rs);
#pragma unroll
        for (int c = 0; c < VPL; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) h[r][c][q] = __ffma2_rn(h[r][c][q], a, bsh);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        const int v = lane + c * 32;
        if (v < nvec && row0 + r < M) y[(row0 + r) * nvec + v] = pack8_2(h[r][c]);
      }
  }
}

template <int VPL, int R>
static int launch_tnf(const void* x, const float* w, void* y, long long M, int C, int mode, cudaStream_t stream) {
  const long long groups = (M + R - 1) / R;
  long long blocks = (groups + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  token_norm_fwd_kernel<VPL, R><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), w,
                                                                 reinterpret_cast<uint4*>(y), M, C, mode);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int token_norm_fwd_run(const void* x, const float* w, void* y, long long M, int C, int mode, cudaStream_t stream) {
  TVAE_REQUIRE(C % 8 == 0 && C / 8 <= 32 * kMaxVecPerLane, "token_norm: C=%d unsupported", C);
  const int vpl = (C / 8 + 31) / 32;
  if (vpl <= 1) return launch_tnf<1, 4>(x, w, y, M, C, mode, stream);
  if (vpl == 2) return launch_tnf<2, 2>(x, w, y, M, C, mode, stream);
  if (vpl == 3) return launch_tnf<3, 1>(x, w, y, M, C, mode, stream);
  if (vpl == 4) return launch_tnf<4, 1>(x, w, y, M, C, mode, stream);
  if (vpl <= 6) return launch_tnf<6, 1>(x, w, y, M, C, mode, stream);
  return launch_tnf<kMaxVecPerLane, 1>(x, w, y, M, C, mode, stream);
}

// dx = d(norm)/dx^T dy [+ add];  dw[c] += sum_rows (...).
// Register-resident version: one warp per token row, VPL 16-byte vectors per lane (C <= 256*VPL), x / dy / the dw
// partials stay in registers, dw is reduced across the block's warps in shared memory and flushed with one atomic per
// column per block.  (The first version indexed per-lane arrays dynamically -> local memory; ncu census of a training
// step: 32.5 ms of 225 ms.)
__device__ unsigned int g_tnb_tickets[1 + kOrdMaxGroups];

template <int VPL>
__global__ void __launch_bounds__(256, VPL <= 2 ? 2 : 1) token_norm_bwd_reg_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                                 const uint4* __restrict__ dy, const uint4* __restrict__ add,
                                                                 uint4* __restrict__ dx, float* __restrict__ dw,
                                                                 float* __restrict__ ws, long long M, int C, int mode) {
  extern __shared__ __align__(16) float s_dw[];   // [8 warps][C]
  __shared__ bool s_flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = C >> 3;
  float2 wv[VPL][4], dwacc[VPL][4];
#pragma unroll
  for (int c = 0; c < VPL; ++c) {
    const int v = lane + c * 32;
    float4 wa = make_float4(0, 0, 0, 0), wb = wa;
    if (v < nvec) {
      wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
      wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
    }
    wv[c][0] = make_float2(wa.x, wa.y); wv[c][1] = make_float2(wa.z, wa.w);
    wv[c][2] = make_float2(wb.x, wb.y); wv[c][3] = make_float2(wb.z, wb.w);
#pragma unroll
    for (int q = 0; q < 4; ++q) dwacc[c][q] = f2(0.0f);
  }
  const float invC = 1.0f / (float)C;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  // software prefetch (narrow rows only -- wide rows have enough loads per lane and no registers to spare): the next
  // row's x / dy are in flight while this row is reduced
  constexpr bool kPrefetch = VPL <= 3;
  uint4 nx[VPL], ng[VPL];
  if (kPrefetch) {
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      const int v = lane + c * 32;
      const bool ok = v < nvec && row < M;
      nx[c] = ok ? __ldg(x + row * nvec + v) : make_uint4(0, 0, 0, 0);
      ng[c] = ok ? __ldg(dy + row * nvec + v) : make_uint4(0, 0, 0, 0);
    }
  }
  for (; row < M; row += warps_total) {
    float2 xr[VPL][4], g[VPL][4];
    uint4 ua[VPL];
    float s2[1];
    if (!kPrefetch) {
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        const int v = lane + c * 32;
        nx[c] = (v < nvec) ? __ldg(x + row * nvec + v) : make_uint4(0, 0, 0, 0);
        ng[c] = (v < nvec) ? __ldg(dy + row * nvec + v) : make_uint4(0, 0, 0, 0);
      }
    }
    {
      float2 acc = f2(0.0f);
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        unpack8_2(nx[c], xr[c]);
        unpack8_2(ng[c], g[c]);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc = __ffma2_rn(xr[c][q], xr[c][q], acc);
      }
      s2[0] = acc.x + acc.y;
    }
    if (kPrefetch) {
      const long long nrow = row + warps_total;
#pragma unroll
      for (int c = 0; c < VPL; ++c) {
        const int v = lane + c * 32;
        const bool ok = v < nvec && nrow < M;
        nx[c] = ok ? __ldg(x + nrow * nvec + v) : make_uint4(0, 0, 0, 0);
        ng[c] = ok ? __ldg(dy + nrow * nvec + v) : make_uint4(0, 0, 0, 0);
        if (add != nullptr) ua[c] = (v < nvec) ? __ldg(add + row * nvec + v) : make_uint4(0, 0, 0, 0);
      }
    }
    warp_sum_n<1>(s2);
    const float rstd = rsqrtf(s2[0] * invC + 1e-6f);
#pragma unroll
    for (int c = 0; c < VPL; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) xr[c][q] = __fmul2_rn(xr[c][q], f2(rstd));          // xhat_r
    if (mode == 1) {
      float st[2];
      {
        float2 sm = f2(0.0f), sq = f2(0.0f);
#pragma unroll
        for (int c = 0; c < VPL; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 h = __fmul2_rn(xr[c][q], wv[c][q]);
            sm = __fadd2_rn(sm, h);
            sq = __ffma2_rn(h, h, sq);
          }
        st[0] = sm.x + sm.y;
        st[1] = sq.x + sq.y;
      }
      warp_sum_n<2>(st);
      const float mu = st[0] * invC;
      const float rs = rsqrtf(fmaxf(st[1] * invC - mu * mu, 0.0f) + 1e-5f);
      const float2 rs2 = f2(rs), sh2 = f2(-mu * rs);
      float a[2];
      {
        float2 a1 = f2(0.0f), a2 = f2(0.0f);
#pragma unroll
        for (int c = 0; c < VPL; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 yh = __ffma2_rn(__fmul2_rn(xr[c][q], wv[c][q]), rs2, sh2);
            a1 = __fadd2_rn(a1, g[c][q]);
            a2 = __ffma2_rn(g[c][q], yh, a2);
          }
        a[0] = a1.x + a1.y;
        a[1] = a2.x + a2.y;
      }
      warp_sum_n<2>(a);
      // LayerNorm (no affine) backward: dh = rs * (g - mean(g) - yhat * mean(g * yhat))
      const float2 na1 = f2(-a[0] * invC), na2rs = f2(-a[1] * invC * rs);
#pragma unroll
      for (int c = 0; c < VPL; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 yh = __ffma2_rn(__fmul2_rn(xr[c][q], wv[c][q]), rs2, sh2);
          g[c][q] = __ffma2_rn(yh, na2rs, __fmul2_rn(__fadd2_rn(g[c][q], na1), rs2));
        }
    }
    // RMSNorm backward: dw += dh * xhat_r ; dx = rstd * (dh * w - xhat_r * mean(dh * w * xhat_r)) [+ add]
    float a3[1];
    {
      float2 acc = f2(0.0f);
#pragma unroll
      for (int c = 0; c < VPL; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dwacc[c][q] = __ffma2_rn(g[c][q], xr[c][q], dwacc[c][q]);
          g[c][q] = __fmul2_rn(g[c][q], wv[c][q]);
          acc = __ffma2_rn(g[c][q], xr[c][q], acc);
        }
      a3[0] = acc.x + acc.y;
    }
    warp_sum_n<1>(a3);
    const float2 na3 = f2(-a3[0] * invC), rstd2 = f2(rstd);
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      const int v = lane + c * 32;
      if (v < nvec) {
        float2 r[4];
        if (add != nullptr) {
          if (!kPrefetch) ua[c] = __ldg(add + row * nvec + v);
          unpack8_2(ua[c], r);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 o = __fmul2_rn(__ffma2_rn(xr[c][q], na3, g[c][q]), rstd2);
          if (add != nullptr) o = __fadd2_rn(o, r[q]);
          g[c][q] = o;
        }
        dx[row * nvec + v] = pack8_2(g[c]);
      }
    }
  }
  // weight gradient: fixed-order reduction (warps of the block in warp order, then the blocks: ordered_rows_reduce)
#pragma unroll
  for (int c = 0; c < VPL; ++c) {
    const int v = lane + c * 32;
    if (v < nvec) {
      float* p = s_dw + (size_t)warp * C + v * 8;
      *reinterpret_cast<float4*>(p) = make_float4(dwacc[c][0].x, dwacc[c][0].y, dwacc[c][1].x, dwacc[c][1].y);
      *reinterpret_cast<float4*>(p + 4) = make_float4(dwacc[c][2].x, dwacc[c][2].y, dwacc[c][3].x, dwacc[c][3].y);
    }
  }
  __syncthreads();
  float* rows = ws;                                            // [gridDim.x][C]
  float* groups = ws + (size_t)gridDim.x * C;                  // [groups][C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float a = 0.0f;
#pragma unroll
    for (int wdx = 0; wdx < 8; ++wdx) a += s_dw[(size_t)wdx * C + i];
    rows[(size_t)blockIdx.x * C + i] = a;
  }
  ordered_rows_reduce(rows, groups, g_tnb_tickets, dw, C, gridDim.x, blockIdx.x, &s_flag);
}

template <int VPL>
static int launch_tnb(const void* x, const float* w, const void* dy, const void* add, void* dx, float* dw, long long M, int C,
                      int mode, cudaStream_t stream) {
  // One resident wave of blocks, at least eight rows per warp: every block ends with its dw partial row going through the
  // fixed-order cross-block reduce, whose cost grows with the block count.  The first rule (a block per 8 rows, capped at
  // 4 blocks per SM) gave the stage-3 launches (M = 8192 rows of 1536 channels, 100 MB) 592 blocks of 1.7 rows per warp:
  // 67 us against 15 us of HBM time (ncu launch list, profiles/r2fin_launches_train_mb32.csv).  TVAE_TNB_ONE_WAVE=0: A/B.
  static const bool one_wave = !(getenv("TVAE_TNB_ONE_WAVE") && atoi(getenv("TVAE_TNB_ONE_WAVE")) == 0);
  long long blocks = (M + 7) / 8;
  long long cap = (long long)num_sms() * 4;
  if (one_wave) {
    blocks = (M + 63) / 64;
    cap = (long long)num_sms() * (VPL <= 2 ? 2 : 1);
  }
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const long long groups = (blocks + kOrdGroup - 1) / kOrdGroup;
  float* ws = nullptr;
  {
    const int rc = scratch_workspace((size_t)(blocks + groups) * C * sizeof(float), reinterpret_cast<void**>(&ws));
    if (rc) return rc;
  }
  const size_t smem = 8 * (size_t)C * sizeof(float);
  static bool configured = false;
  if (!configured) {
    TVAE_CHECK_CUDA(cudaFuncSetAttribute(token_norm_bwd_reg_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 8 * VPL * 4));
    configured = true;
  }
  token_norm_bwd_reg_kernel<VPL><<<(int)blocks, 256, smem, stream>>>(
      reinterpret_cast<const uint4*>(x), w, reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(add),
      reinterpret_cast<uint4*>(dx), dw, ws, M, C, mode);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Fallback for rows wider than 1536 channels (giant variant): per-lane arrays in local memory.
__global__ void __launch_bounds__(256) token_norm_bwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                             const uint4* __restrict__ dy, const uint4* __restrict__ add,
                                                             uint4* __restrict__ dx, float* __restrict__ dw, long long M,
                                                             int C, int mode, int rows_per_warp) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nvec = C >> 3;
  float dwacc[kMaxVecPerLane][8];
#pragma unroll
  for (int c = 0; c < kMaxVecPerLane; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) dwacc[c][k] = 0.0f;
  const float invC = 1.0f / (float)C;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const long long row = warp_id * rows_per_warp + rr;
    if (row >= M) break;
    float xr[kMaxVecPerLane][8], g[kMaxVecPerLane][8];
    float s2 = 0.0f;
    int cnt = 0;
    for (int v = lane; v < nvec; v += 32, ++cnt) {
      unpack8(__ldg(x + row * nvec + v), xr[cnt]);
      unpack8(__ldg(dy + row * nvec + v), g[cnt]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s2 = fmaf(xr[cnt][k], xr[cnt][k], s2);
    }
    s2 = warp_sum(s2);
    const float rstd = rsqrtf(s2 * invC + 1e-6f);
    // xr <- xhat_r = x*rstd ; for mode 1 also need h = xhat_r*w, mu, rs
    float sm = 0.0f, sq = 0.0f;
    cnt = 0;
    for (int v = lane; v < nvec; v += 32, ++cnt) {
      const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
      const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xr[cnt][k] *= rstd;
        const float h = xr[cnt][k] * wv[k];
        sm += h;
        sq = fmaf(h, h, sq);
      }
    }
    float mu = 0.0f, rs = 1.0f;
    if (mode == 1) {
      sm = warp_sum(sm);
      sq = warp_sum(sq);
      mu = sm * invC;
      rs = rsqrtf(fmaxf(sq * invC - mu * mu, 0.0f) + 1e-5f);
      // LayerNorm (no affine) backward: dh = rs * (g - mean(g) - yhat * mean(g*yhat))
      float a1 = 0.0f, a2 = 0.0f;
      cnt = 0;
      for (int v = lane; v < nvec; v += 32, ++cnt) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float yh = (xr[cnt][k] * wv[k] - mu) * rs;
          a1 += g[cnt][k];
          a2 = fmaf(g[cnt][k], yh, a2);
        }
      }
      a1 = warp_sum(a1) * invC;
      a2 = warp_sum(a2) * invC;
      cnt = 0;
      for (int v = lane; v < nvec; v += 32, ++cnt) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float yh = (xr[cnt][k] * wv[k] - mu) * rs;
          g[cnt][k] = rs * (g[cnt][k] - a1 - yh * a2);   // g is now dh
        }
      }
    }
    // RMSNorm backward: dx = rstd * (dh*w - xhat_r * mean(dh*w*xhat_r));  dw += dh * xhat_r
    float a3 = 0.0f;
    cnt = 0;
    for (int v = lane; v < nvec; v += 32, ++cnt) {
      const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
      const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dwacc[cnt][k] = fmaf(g[cnt][k], xr[cnt][k], dwacc[cnt][k]);
        g[cnt][k] *= wv[k];
        a3 = fmaf(g[cnt][k], xr[cnt][k], a3);
      }
    }
    a3 = warp_sum(a3) * invC;
    cnt = 0;
    for (int v = lane; v < nvec; v += 32, ++cnt) {
      float r[8];
      if (add != nullptr) unpack8(__ldg(add + row * nvec + v), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float o = rstd * (g[cnt][k] - xr[cnt][k] * a3);
        if (add != nullptr) o += r[k];
        g[cnt][k] = o;
      }
      dx[row * nvec + v] = pack8(g[cnt]);
    }
  }
  int cnt = 0;
  for (int v = lane; v < nvec; v += 32, ++cnt)
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(dw + v * 8 + k, dwacc[cnt][k]);
}

int token_norm_bwd_run(const void* x, const float* w, const void* dy, const void* add, void* dx, float* dw, long long M,
                       int C, int mode, cudaStream_t stream) {
  TVAE_REQUIRE(C % 8 == 0 && C / 8 <= 32 * kMaxVecPerLane, "token_norm_bwd: C=%d unsupported", C);
  const int vpl = (C / 8 + 31) / 32;
  if (vpl <= 1) return launch_tnb<1>(x, w, dy, add, dx, dw, M, C, mode, stream);
  if (vpl == 2) return launch_tnb<2>(x, w, dy, add, dx, dw, M, C, mode, stream);
  if (vpl == 3) return launch_tnb<3>(x, w, dy, add, dx, dw, M, C, mode, stream);
  if (vpl == 4) return launch_tnb<4>(x, w, dy, add, dx, dw, M, C, mode, stream);
  if (vpl <= 6) return launch_tnb<6>(x, w, dy, add, dx, dw, M, C, mode, stream);
  TVAE_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * sizeof(float), stream));     // wide-row fallback: atomics
  int rpw = 16;
  while (rpw > 1 && (M + rpw - 1) / rpw < 8LL * 4 * num_sms()) rpw >>= 1;
  const long long warps = (M + rpw - 1) / rpw;
  const int grid = (int)((warps * 32 + 255) / 256);
  token_norm_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), w, reinterpret_cast<const uint4*>(dy),
                                                  reinterpret_cast<const uint4*>(add), reinterpret_cast<uint4*>(dx), dw, M, C,
                                                  mode, rpw);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Attention backward helpers.
//   delta[b][h][s] = sum_d dO[b,s,h,d] * O[b,s,h,d]                                  (one warp per 4 (token, head) pairs)
//   rope_bwd: dqkv[:, 0:C]   = q_scale * R^T dQ_rot  (dQ accumulated in fp32 by the attention backward kernel)
//             dqkv[:, C:2C]  = R^T dK_rot   (in place, bf16)
//   with R the reference's per-pair matrix [[cos a, -sin a], [sin b, cos b]] (attention.py:178-197).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const uint4* __restrict__ o, const uint4* __restrict__ dout,
                                                         float* __restrict__ delta, int B, int S, int C) {
  // 8 lanes cover one head's 64 channels (8 x uint4)
  const int nh = C / 64;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = gid >> 3;            // (token, head) index
  const int sub = (int)(gid & 7);
  const long long total = (long long)B * S * nh;
  float acc = 0.0f;
  if (pair < total) {
    const long long tok = pair / nh;
    const int h = (int)(pair % nh);
    float a[8], g[8];
    unpack8(__ldg(o + tok * (C / 8) + h * 8 + sub), a);
    unpack8(__ldg(dout + tok * (C / 8) + h * 8 + sub), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(a[k], g[k], acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0 && pair < total) {
    const long long tok = pair / nh;
    const int h = (int)(pair % nh);
    const int b = (int)(tok / S), s = (int)(tok % S);
    delta[((size_t)b * nh + h) * S + s] = acc;
  }
}

int attn_delta_run(const void* o, const void* dout, float* delta, int B, int S, int C, cudaStream_t stream) {
  const long long threads = (long long)B * S * (C / 64) * 8;
  attn_delta_kernel<<<(int)((threads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(o),
                                                                     reinterpret_cast<const uint4*>(dout), delta, B, S, C);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) rope_bwd_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv,
                                                       const float2* __restrict__ tab, long long M, int C, int H, int W,
                                                       float q_scale, int dq_slices) {
  // one thread per (token, part in {q,k}, 8-channel vector)
  const int nvec = C >> 3;
  const long long total = M * 2 * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int part = (int)((i / nvec) % 2);
    const long long tok = i / (2 * nvec);
    const int t = (int)(tok % ((long long)H * W));
    const int j0 = (v * 8) & 63;                       // channel inside the head
    const int pos = (j0 < 32) ? t / W : t % W;
    const float2* tb = tab + (size_t)pos * 16;
    float g[8];
    __nv_bfloat16* dst = dqkv + tok * 3 * C + part * C + v * 8;
    if (part == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(dq_acc + tok * C + v * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(dq_acc + tok * C + v * 8) + 1);
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
      if (dq_slices == 2) {             // even-step + odd-step accumulators of the ordered attention backward
        const float* p1 = dq_acc + (size_t)M * C + tok * C + v * 8;
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(p1));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p1) + 1);
        g[0] += a1.x; g[1] += a1.y; g[2] += a1.z; g[3] += a1.w; g[4] += b1.x; g[5] += b1.y; g[6] += b1.z; g[7] += b1.w;
      }
    } else {
      unpack8(*reinterpret_cast<const uint4*>(dst), g);
    }
    const float sc = part == 0 ? q_scale : 1.0f;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float2 ca = __ldg(tb + ((j0 + k) & 15));
      const float2 cb = __ldg(tb + ((j0 + k + 1) & 15));
      const float d1 = g[k], d2 = g[k + 1];
      g[k] = (d1 * ca.x + d2 * cb.y) * sc;
      g[k + 1] = (d2 * cb.x - d1 * ca.y) * sc;
    }
    *reinterpret_cast<uint4*>(dst) = pack8(g);
  }
}

int rope_bwd_run(const float* dq_acc, void* dqkv, const float* tab, long long M, int C, int H, int W, float q_scale,
                 int dq_slices, cudaStream_t stream) {
  const long long total = M * 2 * (C / 8);
  int grid = (int)((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  rope_bwd_kernel<<<grid, 256, 0, stream>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(dqkv),
                                            reinterpret_cast<const float2*>(tab), M, C, H, W, q_scale, dq_slices);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// -------------------------------------------------------------------------------------------------
// Loss / reparameterisation backward (transvae.py:186-199, :244-245 patched; vae_loss.py:80-104 patched / :83,:94 main).
//   drecon = g_l1 * sign(f(r) - t) * f'(r)                      f = sigmoid (patched) or identity
//   dmu    = g_kl * mu            * [clamp passes]  + dz * [..]
//   dlv    = g_kl * -0.5*(1 - e^lv) * [..]          + dz * eps * 0.5 * exp(0.5*lv) * [..]
// g_l1 = l1_weight / numel(recon) * dLoss, g_kl = kl_weight / norm * dLoss (device scalars, no host sync).
// -------------------------------------------------------------------------------------------------
__global__ void loss_bwd_kernel(const float* __restrict__ recon, const float* __restrict__ target,
                                const float* __restrict__ scal /*[g_l1, g_kl]*/, float* __restrict__ drecon, long long n,
                                int patched) {
  const float g = scal[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r = recon[i];
    float d = 1.0f;
    if (patched) {
      const float s = 1.0f / (1.0f + __expf(-r));
      d = s * (1.0f - s);
      r = s;
    }
    const float diff = r - target[i];
    drecon[i] = g * d * (diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f));
  }
}

__global__ void latent_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                  const float* __restrict__ eps, const float* __restrict__ dz,
                                  const float* __restrict__ dmu_ret, const float* __restrict__ dlv_ret,
                                  float* __restrict__ dmu, float* __restrict__ dlv, long long n, int patched) {
  // (z, mu_c, lv_c) = reparam(mu, logvar): gradients arriving on all three outputs are folded back onto (mu, logvar)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], lv = logvar[i];
    float pm = 1.0f, pl = 1.0f, lvc = lv;
    if (patched) {
      pm = (m >= -50.0f && m <= 50.0f) ? 1.0f : 0.0f;
      pl = (lv >= -30.0f && lv <= 20.0f) ? 1.0f : 0.0f;
      lvc = fminf(fmaxf(lv, -30.0f), 20.0f);
    }
    const float gz = dz ? dz[i] : 0.0f;
    float gm = gz, gl = gz * eps[i] * 0.5f * expf(0.5f * lvc);
    if (dmu_ret) gm += dmu_ret[i];
    if (dlv_ret) gl += dlv_ret[i];
    dmu[i] = gm * pm;
    dlv[i] = gl * pl;
  }
}

__global__ void kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                              const float* __restrict__ scal, float* __restrict__ dmu, float* __restrict__ dlv,
                              long long n, int patched, float clip_lo, float clip_hi) {
  const float g = scal[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float lv = logvar[i];
    float pl = 1.0f, lvc = lv;
    if (patched) {
      pl = (lv >= clip_lo && lv <= clip_hi) ? 1.0f : 0.0f;
      lvc = fminf(fmaxf(lv, clip_lo), clip_hi);
    }
    dmu[i] = g * mu[i];
    dlv[i] = g * -0.5f * (1.0f - expf(lvc)) * pl;
  }
}

static int grid_for(long long n) {
  int g = (int)((n + 255) / 256);
  const int cap = num_sms() * 8;
  return g > cap ? cap : (g < 1 ? 1 : g);
}

int loss_bwd_run(const float* recon, const float* target, const float* mu, const float* logvar, const float* scal,
                 float* drecon, float* dmu, float* dlv, long long n_img, long long n_lat, int patched, float clip_lo,
                 float clip_hi, cudaStream_t stream) {
  loss_bwd_kernel<<<grid_for(n_img), 256, 0, stream>>>(recon, target, scal, drecon, n_img, patched);
  TVAE_CHECK_CUDA(cudaGetLastError());
  kl_bwd_kernel<<<grid_for(n_lat), 256, 0, stream>>>(mu, logvar, scal, dmu, dlv, n_lat, patched, clip_lo, clip_hi);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int latent_bwd_run(const float* mu, const float* logvar, const float* eps, const float* dz, const float* dmu_ret,
                   const float* dlv_ret, float* dmu, float* dlv, long long n, int patched, cudaStream_t stream) {
  latent_bwd_kernel<<<grid_for(n), 256, 0, stream>>>(mu, logvar, eps, dz, dmu_ret, dlv_ret, dmu, dlv, n, patched);
  TVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tvae
