// TMA tensor-map construction. cuTensorMapEncodeTiled is resolved through the runtime
// (cudaGetDriverEntryPoint) so libldn.so has no link-time dependency on libcuda and can be
// dlopen'ed on a machine without a driver (symbol-export test).
#include "common.h"

#include <mutex>

namespace ldn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  LDN_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  return fn;
}

CUtensorMap make_tmap_2d(const bf16* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  LDN_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tmap2d: base must be 16B aligned");
  LDN_CHECK((ld * 2) % 16 == 0, "tmap2d: row pitch must be a multiple of 16 bytes");
  LDN_CHECK(box_rows >= 1 && box_rows <= 256, "tmap2d: box rows out of range");
  CUtensorMap tm;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r));
  return tm;
}

CUtensorMap make_tmap_nhwc(const bf16* base, int B, int H, int W, int C, int bb, int bh, int bw) {
  LDN_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tmap4d: base must be 16B aligned");
  LDN_CHECK(C % 8 == 0, "tmap4d: C must be a multiple of 8");
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
  return tm;
}

}  // namespace ldn
