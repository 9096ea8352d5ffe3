// Host-side TMA tensor-map construction.  The driver entry point is resolved at run time
// through the CUDA runtime so the library carries no link-time dependency on libcuda
// (it must still dlopen on a CPU-only box, where only symbol checks run).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace b200sr {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled& tmap_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  return fn;
}

inline bool tmap_init() {
  if (tmap_fn()) return true;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return false;
  tmap_fn() = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  return true;
}

// NHWC bf16 activation tensor [N][H][W][Cpitch]; box = {cbox channels, boxw pixels, 1 row, 1 image}.
// Out-of-bounds elements (negative or >= extent coordinates) are filled with zeros, which is
// exactly the convolution's zero padding at the (tile) border.
inline bool tmap_encode_act(CUtensorMap* out, const void* base, int N, int H, int W, int Cpitch, int cbox, int boxw,
                            int swizzle_bytes) {
  if (!tmap_init()) return false;
  cuuint64_t dims[4] = {(cuuint64_t)Cpitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Cpitch * 2, (cuuint64_t)W * Cpitch * 2, (cuuint64_t)H * W * Cpitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)cbox, (cuuint32_t)boxw, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = tmap_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Nearest-2x upsampled VIEW of a low-resolution NHWC tensor [N][H][W][Cpitch]: dims {C, 2, W, H, N} with a ZERO
// stride on the second dimension, so a box {cbox, 2, boxw, 1, 1} lands in shared memory as 2*boxw pixel rows in
// which every low-resolution pixel appears twice (F.interpolate(scale_factor=2, mode='nearest') along x; along y
// the caller addresses row Y >> 1).  Out-of-bounds pixels/rows are zero-filled as in tmap_encode_act.  Verified on
// B200 by csrc/tools/probe_dup.cu.
inline bool tmap_encode_act_up2(CUtensorMap* out, const void* base, int N, int H, int W, int Cpitch, int cbox, int boxw,
                                int swizzle_bytes) {
  if (!tmap_init()) return false;
  cuuint64_t dims[5] = {(cuuint64_t)Cpitch, 2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[4] = {0, (cuuint64_t)Cpitch * 2, (cuuint64_t)W * Cpitch * 2, (cuuint64_t)H * W * Cpitch * 2};
  cuuint32_t box[5] = {(cuuint32_t)cbox, 2, (cuuint32_t)boxw, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = tmap_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace b200sr
