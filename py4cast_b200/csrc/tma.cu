// Host side of the TMA path: tensor maps through cuTensorMapEncodeTiled, fetched from the driver at run time
// (cudaGetDriverEntryPoint: the library does not link libcuda).
#include "tma.cuh"

#include <cuda_runtime.h>
#include <stdlib.h>

namespace p4c {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool tma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("P4C_TMA");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1 && encode_fn() != nullptr;
}

bool tma_rows_map(CUtensorMap* out, const void* base, long long inner, long long rows, long long ld, long long batch,
                  long long batch_stride) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || inner <= 0 || rows <= 0 || batch <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0) return false;
  const bool shared = batch == 1 || batch_stride == 0;
  if (!shared && ((batch_stride * 2) % 16 != 0 || batch_stride < rows * ld)) return false;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows),
                              static_cast<cuuint64_t>(shared ? 1 : batch)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld * 2),
                                 static_cast<cuuint64_t>(shared ? rows * ld * 2 : batch_stride * 2)};
  const cuuint32_t box[3] = {64, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace p4c
