#include "tma_host.h"

#include <cuda_runtime.h>

#include "common.h"

namespace mre {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_tmap_2d(CUtensorMap *out, CUtensorMapDataType dt, int esz, const void *base, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows, int box_cols) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return MRE_ERR_CUDA;
    }
    MRE_CHECK_ARG(((uintptr_t)base & 15) == 0 && (ld * esz) % 16 == 0, "TMA needs a 16-byte aligned base and row pitch");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols * esz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : box_cols * esz == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                          : box_cols * esz == 32  ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, dt, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%dx%d)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld, box_rows, box_cols);
        return MRE_ERR_CUDA;
    }
    return MRE_OK;
}

int make_tmap_f32_2d(CUtensorMap *out, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rows, cols, ld, box_rows, box_cols);
}

int make_tmap_bf16_2d(CUtensorMap *out, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld, box_rows, box_cols);
}

}  // namespace mre
