// LBO probe: the two 64-channel halves of an M = 128 MN-major SWIZZLE_128B A operand taken from arbitrary places -- the
// SAME halo tile at two different pixel-row offsets (LBO = 128 B: taps dx and dx + 1 of a 3x3 convolution), or another
// tile at a row offset (LBO = tile stride + 128 B, not a multiple of the 1024-byte swizzle atom).  Needed by the
// three-taps-per-CTA weight gradient (wgrad.cu, mtwgrad_h_kernel).
// Derived from umma_lbo_offset_mn.cu: companion of umma_row_offset.cu for MN-major SWIZZLE_128B operands (the weight-gradient kernel: pixels are the MMA K
// index, 64 channels = 128 B per pixel row, 8 pixel rows per 1024-byte swizzle atom).  Can the A operand start at an
// arbitrary PIXEL ROW j of a larger (halo) tile?
//   D_j[m][n] = sum_{p < 128} A[p + j][m] * B[p][n],   M = 128 (two 64-channel chunks, LBO apart), N = 64, K = 16 per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/experiments/umma_lbo_offset_mn tools/experiments/umma_lbo_offset_mn.cu
#define TVAE_DEVICE_OK 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepl-project_b200/csrc/common.cuh"
using namespace tvae;

constexpr int kRows = 136, kJ = 4, kChunk = kRows * 128;   // 17408 B per 64-channel chunk (a multiple of 1024)
// half 0 = (chunk ca, pixel-row offset oa), half 1 = (chunk cb, pixel-row offset ob)
__constant__ int c_ca[kJ] = {0, 0, 0, 1}, c_oa[kJ] = {0, 1, 2, 0}, c_cb[kJ] = {0, 0, 1, 1}, c_ob[kJ] = {1, 2, 0, 2};

__device__ __host__ inline float aval(int p, int m) { return (float)(((p * 7 + m * 3) % 11) - 5); }
__device__ __host__ inline float bval(int p, int n) { return (float)(((p * 5 + n * 2) % 7) - 3); }

__global__ void __launch_bounds__(128) probe(float* out, int zero_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                  // two chunks: channels [0, 64) and [64, 128), each kRows x 128 B
  uint8_t* sB = base + 2 * kChunk;     // one chunk
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kChunk);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < kRows * 128; i += 128) {
    const int p = i / 128, m = i % 128, c = m & 63;
    *reinterpret_cast<__nv_bfloat16*>(sA + (m >> 6) * kChunk + p * 128 + (((c >> 3) ^ (p & 7)) << 4) + (c & 7) * 2) =
        __float2bfloat16(aval(p, m));
  }
  for (int i = threadIdx.x; i < kRows * 64; i += 128) {
    const int p = i / 64, n = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sB + p * 128 + (((n >> 3) ^ (p & 7)) << 4) + (n & 7) * 2) = __float2bfloat16(bval(p, n));
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    for (int j = 0; j < kJ; ++j) {
      for (int s = 0; s < 8; ++s) {      // 8 x 16 pixels
        const uint32_t a0 = smem_u32(sA) + c_ca[j] * kChunk + c_oa[j] * 128, a1 = smem_u32(sA) + c_cb[j] * kChunk + c_ob[j] * 128;
        uint64_t da = umma_desc_mnmajor_sw128(a0 + s * 2048, a1 - a0, 1024);
        if (zero_base_offset) da &= ~(uint64_t(7) << 49);
        umma_f16(tmem + j * 64, da, umma_desc_mnmajor_sw128(smem_u32(sB) + s * 2048, kChunk, 1024), idesc, s != 0);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int j = 0; j < kJ; ++j) {
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + j * 64 + h * 32, v);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) out[(j * 128 + threadIdx.x) * 64 + h * 32 + c] = __uint_as_float(v[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

int main() {
  float* d;
  cudaMalloc(&d, kJ * 128 * 64 * sizeof(float));
  const int smem = 3 * kChunk + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int ca[kJ] = {0, 0, 0, 1}, oa[kJ] = {0, 1, 2, 0}, cb[kJ] = {0, 0, 1, 1}, ob[kJ] = {1, 2, 0, 2};
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d, 0, kJ * 128 * 64 * sizeof(float));
    probe<<<1, 128, smem>>>(d, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(kJ * 128 * 64);
    cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    for (int j = 0; j < kJ; ++j) {
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          const int ch = (m < 64 ? ca[j] : cb[j]) * 64 + (m & 63), off = m < 64 ? oa[j] : ob[j];
          for (int p = 0; p < 128; ++p) ref += aval(p + off, ch) * bval(p, n);
          if (ref != h[(j * 128 + m) * 64 + n]) ++bad;
        }
      printf("MN-major, base_offset %s, halves (chunk %d row %d | chunk %d row %d), LBO %d B: %d / %d wrong%s\n",
             variant ? "forced 0" : "= (addr>>7)&7", ca[j], oa[j], cb[j], ob[j], (cb[j] - ca[j]) * kChunk + (ob[j] - oa[j]) * 128, bad,
             128 * 64, bad ? "" : "  -> exact");
    }
  }
  return 0;
}
