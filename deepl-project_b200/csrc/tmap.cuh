// Host-side construction of TMA tensor maps.  cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda (it must load -- and
// export its symbols -- on a machine without a driver; compute entry points then fail loudly).
#pragma once
#include "common.cuh"

namespace tvae {

// 2-D row-major bf16 matrix [rows, cols]; box = [box_rows, 64 cols], 128-byte swizzle.
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                 uint32_t box_rows, uint32_t box_cols = 64);

// 5-D "pixel view" of an NHWC bf16 tensor [B, H, W, C].
//   split == 0: dims (C, W, 1, H, B)           -- plain view
//   split == 1: dims (2C, W/2, 2, H/2, B)      -- 2x2 phase view: element (b, 2h+p, 2w+q, c) sits at
//                                                 coordinate (q*C + c, w, p, h, b)
// box = (64, tw, 1, th, nb), 128-byte swizzle, zero fill out of bounds (this is the conv halo).
int make_tmap_pix(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int split, int tw, int th, int nb);

// plain view, box = (64, 130, 1, 1, 1): 128 + 2 pixels of one image row (halo tile of the 3x3 convolutions)
int make_tmap_pix_halo(CUtensorMap* out, const void* ptr, int B, int H, int W, int C);

// 3-D bf16 tensor (d0 contiguous); box = (64, box1, 1), 128-byte swizzle.  Used for per-image token matrices
// [B, S, C] so that rows beyond S are zero-filled / clipped instead of running into the next image.
int make_tmap_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                 uint64_t stride2_elems, uint32_t box1);

// 3-D fp32 tensor (d0 contiguous); box = (32, box1, 1) = 128 bytes wide, 128-byte swizzle.  Used for bulk reduce-add
// (cp.reduce.async.bulk.tensor) of fp32 accumulator tiles.
int make_tmap_3d_f32(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                     uint64_t stride2_elems, uint32_t box1);

}  // namespace tvae
