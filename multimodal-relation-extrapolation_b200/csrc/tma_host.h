// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled fetched through the runtime, so the library
// does not link against libcuda).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace mre {
// Row-major float32 matrix [rows, cols] with row stride `ld` floats (ld * 4 a multiple of 16 bytes, base 16-byte
// aligned); box = box_rows x box_cols elements; 128-byte swizzle when box_cols * 4 == 128, none otherwise.
int make_tmap_f32_2d(CUtensorMap *out, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
// the same for a row-major bfloat16 matrix (ld in elements, ld * 2 a multiple of 16 bytes); 128-byte swizzle when box_cols == 64
int make_tmap_bf16_2d(CUtensorMap *out, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
}  // namespace mre
