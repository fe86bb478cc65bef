// Experiment for the halo-tile plan (known headroom): can a K-major SWIZZLE_128B UMMA operand start at an
// arbitrary ROW of a larger tile that was laid out with the TMA swizzle (16-byte chunk index XOR (row & 7), rows 128 B
// apart)?  If yes, one (128 + 2)-pixel halo tile per kernel row serves the three dx taps of a 3x3 convolution.
//   D_j[r][n] = sum_k A[r + j][k] * B[n][k],  j = row offset of the descriptor start, M = 128, N = 64, K = 64.
// Two descriptor variants: base-offset field = (addr >> 7) & 7 (as common.cuh builds it) and base-offset = 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/experiments/umma_row_offset tools/experiments/umma_row_offset.cu
#define TVAE_DEVICE_OK 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepl-project_b200/csrc/common.cuh"
using namespace tvae;

constexpr int kRows = 144, kJ = 4;
__constant__ int c_offs[kJ] = {0, 1, 2, 5};

__device__ __host__ inline float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 11) - 5); }
__device__ __host__ inline float bval(int n, int k) { return (float)(((n * 5 + k * 2) % 7) - 3); }

__global__ void __launch_bounds__(128) probe(float* out, int zero_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                       // kRows x 128 B
  uint8_t* sB = base + kRows * 128;         // 64 x 128 B (kRows * 128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < kRows * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sA + r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(aval(r, k));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(bval(n, k));
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
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    for (int j = 0; j < kJ; ++j) {
      for (int k = 0; k < 4; ++k) {
        uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA) + c_offs[j] * 128 + k * 32);
        if (zero_base_offset) da &= ~(uint64_t(7) << 49);
        umma_f16(tmem + j * 64, da, umma_desc_kmajor_sw128(smem_u32(sB) + k * 32), idesc, k != 0);
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
  const int smem = kRows * 128 + 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int offs[kJ] = {0, 1, 2, 5};
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d, 0, kJ * 128 * 64 * sizeof(float));
    probe<<<1, 128, smem>>>(d, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(kJ * 128 * 64);
    cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    for (int j = 0; j < kJ; ++j) {
      int bad = 0, first = -1;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += aval(r + offs[j], k) * bval(n, k);
          if (ref != h[(j * 128 + r) * 64 + n]) { if (first < 0) first = r * 64 + n; ++bad; }
        }
      printf("base_offset %s, row offset %d: %d / %d wrong%s\n", variant ? "forced 0" : "= (addr>>7)&7", offs[j], bad, 128 * 64,
             bad ? "" : "  -> exact");
      if (bad) printf("   first mismatch at row %d col %d\n", first / 64, first % 64);
    }
  }
  return 0;
}
