// Library runtime: last-error string, device queries, TMA tensor-map encoding.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "tmap.cuh"

namespace tvae {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    n = p.multiProcessorCount;
  }
  return n;
}

// SMs the persistent (one CTA per SM) kernels may occupy: all of them, minus the ones left to a concurrent NCCL
// collective while gradient buckets are all-reduced under the backward pass (tvae_set_reserved_sms).
static int g_reserved_sms = 0;
// Scratch workspace of the fixed-order reductions (weight-gradient split slices, GroupNorm-backward block partials): grown
// on demand, kept per device.  Every user fills and consumes it with kernels enqueued by ONE C-ABI call on the one compute
// stream, so consecutive users are ordered; growing synchronises (cudaFree).
int scratch_workspace(size_t bytes, void** out) {
  constexpr int kMaxDev = 16;
  static void* ws[kMaxDev] = {};
  static size_t cap[kMaxDev] = {};
  int dev = 0;
  TVAE_CHECK_CUDA(cudaGetDevice(&dev));
  TVAE_REQUIRE(dev >= 0 && dev < kMaxDev, "scratch_workspace: device index %d out of range", dev);
  if (cap[dev] < bytes) {
    if (ws[dev] != nullptr) TVAE_CHECK_CUDA(cudaFree(ws[dev]));
    ws[dev] = nullptr;
    cap[dev] = 0;
    const size_t want = bytes + bytes / 4;
    TVAE_CHECK_CUDA(cudaMalloc(&ws[dev], want));
    cap[dev] = want;
  }
  *out = ws[dev];
  return 0;
}

void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }
int persistent_sms() {
  const int n = num_sms();
  const int r = n - g_reserved_sms;
  return r < 2 ? (n < 2 ? n : 2) : r;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode(CUtensorMap* out, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                  const cuuint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  EncodeTiledFn fn = get_encode();
  TVAE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  TVAE_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base address %p not 16-byte aligned", ptr);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, dtype, rank, const_cast<void*>(ptr), dims, strides_b, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=(%llu,%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u,%u)",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                   (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                   rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return -3;
  }
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                 uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t str[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  TVAE_REQUIRE(box_rows <= 256 && box_cols * 2 <= 128, "bad 2-D TMA box %u x %u", box_rows, box_cols);
  TVAE_REQUIRE((row_stride_elems * 2) % 16 == 0, "TMA row stride %llu not a multiple of 16 bytes",
               (unsigned long long)(row_stride_elems * 2));
  return encode(out, ptr, 2, dims, str, box);
}

int make_tmap_pix(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int split, int tw, int th, int nb) {
  TVAE_REQUIRE(C % 8 == 0, "NHWC channel count %d must be a multiple of 8", C);
  cuuint64_t dims[5];
  cuuint64_t str[4];
  const uint64_t e = 2;
  if (split) {
    TVAE_REQUIRE(H % 2 == 0 && W % 2 == 0, "phase view needs even H, W (got %d x %d)", H, W);
    dims[0] = 2 * (uint64_t)C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    str[0] = 2 * (uint64_t)C * e;            // w2
    str[1] = (uint64_t)W * C * e;            // p
    str[2] = 2 * (uint64_t)W * C * e;        // h2
    str[3] = (uint64_t)H * W * C * e;        // b
  } else {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
    str[0] = (uint64_t)C * e;
    str[1] = (uint64_t)W * C * e;
    str[2] = (uint64_t)W * C * e;
    str[3] = (uint64_t)H * W * C * e;
  }
  cuuint32_t box[5] = {64, (cuuint32_t)tw, 1, (cuuint32_t)th, (cuuint32_t)nb};
  TVAE_REQUIRE(tw * th * nb == 128 && tw <= 256 && th <= 256 && nb <= 256, "bad pixel box %d x %d x %d", tw, th, nb);
  return encode(out, ptr, 5, dims, str, box);
}

// Plain pixel view with a (64 channels, 128 + 2 pixels of one image row) box: the halo tile of the 3x3 convolutions
// (mtgemm2 halo mode).  Out-of-bounds pixels (w = -1, w = W, h = -1, h = H) are zero-filled: the convolution padding.
int make_tmap_pix_halo(CUtensorMap* out, const void* ptr, int B, int H, int W, int C) {
  TVAE_REQUIRE(C % 8 == 0, "NHWC channel count %d must be a multiple of 8", C);
  const uint64_t e = 2;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t str[4] = {(uint64_t)C * e, (uint64_t)W * C * e, (uint64_t)W * C * e, (uint64_t)H * W * C * e};
  cuuint32_t box[5] = {64, 130, 1, 1, 1};
  return encode(out, ptr, 5, dims, str, box);
}

int make_tmap_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                 uint64_t stride2_elems, uint32_t box1) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t str[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {64, box1, 1};
  TVAE_REQUIRE((stride1_elems * 2) % 16 == 0 && (stride2_elems * 2) % 16 == 0, "3-D TMA strides must be 16-byte multiples");
  return encode(out, ptr, 3, dims, str, box);
}

int make_tmap_3d_f32(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                     uint64_t stride2_elems, uint32_t box1) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t str[2] = {stride1_elems * 4, stride2_elems * 4};
  cuuint32_t box[3] = {32, box1, 1};
  TVAE_REQUIRE((stride1_elems * 4) % 16 == 0 && (stride2_elems * 4) % 16 == 0, "3-D TMA strides must be 16-byte multiples");
  return encode(out, ptr, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

}  // namespace tvae
