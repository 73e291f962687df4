// Host side of TMA: tensor-map encoding.  cuTensorMapEncodeTiled lives in the driver; it is resolved through the
// runtime (cudaGetDriverEntryPoint) on first use, so the library neither links libcuda nor needs it to load.
#include "tc_common.cuh"

namespace samble {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

int make_tile_map(CUtensorMap* map, const void* base, int inner, long long ld, int rows, int batch, int box_rows, int elem_bytes) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return SAMBLE_E_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)ld * rows * elem_bytes};
  const cuuint32_t box[3] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows, 1u};      // 128-byte rows
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a (%d, %d, %d) array of %d-byte elements, box (1, %d, 128 B)", (int)r,
              batch, rows, inner, elem_bytes, box_rows);
    return SAMBLE_E_CUDA;
  }
  return SAMBLE_OK;
}

}  // namespace samble
