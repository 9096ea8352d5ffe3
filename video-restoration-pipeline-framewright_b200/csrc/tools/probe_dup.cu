// Development probe: can a TMA tiled tensor map carry a zero-stride dimension (each pixel delivered twice), so that
// a nearest-2x upsampled activation row can be loaded straight from the low-resolution tensor?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_dup probe_dup.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ptx.cuh"
#include "../tmap.h"

using namespace b200sr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

constexpr int BOXW = 66;   // LR pixels per box -> 132 smem rows

__global__ void k(const __grid_constant__ CUtensorMap map, int w0, int y, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, BOXW * 2 * 128);
    tma_load_5d(&map, &bar, smem, 0, 0, w0, y, 0);
  }
  mbar_wait(&bar, 0);
  // un-swizzle: row r, 16-byte chunk j lives at chunk j ^ (r & 7)
  for (int i = threadIdx.x; i < BOXW * 2 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int chunk = (c / 8) ^ (r & 7);
    out[i] = *reinterpret_cast<__nv_bfloat16*>(smem + r * 128 + chunk * 16 + (c % 8) * 2);
  }
}

int main() {
  const int H = 4, W = 70, C = 64;
  std::vector<__nv_bfloat16> h(H * W * C);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = __float2bfloat16(static_cast<float>(y * 100 + x + (c == 5 ? 0.5f : 0.f)));
  __nv_bfloat16 *d, *dout;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout, BOXW * 2 * 64 * 2));
  if (!tmap_init()) { printf("no driver entry point\n"); return 2; }
  CUtensorMap map;
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)W, (cuuint64_t)H, 1};
  cuuint64_t strides[4] = {0, (cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)C, 2, BOXW, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = tmap_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode with zero stride: CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024));
  for (int w0 : {0, -1, 10}) {
    k<<<1, 128, 20 * 1024>>>(map, w0, 2, dout);
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> o(BOXW * 2 * 64);
    CK(cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int rr = 0; rr < BOXW * 2; ++rr) {
      const int x = w0 + rr / 2;
      for (int c = 0; c < 64; ++c) {
        const float want = (x < 0 || x >= W) ? 0.f : static_cast<float>(200 + x + (c == 5 ? 0.5f : 0.f));
        const float got = __bfloat162float(o[rr * 64 + c]);
        if (got != __bfloat162float(__float2bfloat16(want))) {
          if (bad < 5) printf("  w0=%d row %d ch %d: got %g want %g\n", w0, rr, c, got, want);
          ++bad;
        }
      }
    }
    printf("w0=%d: %d mismatches of %d (rows 0..5 ch0: %g %g %g %g %g %g)\n", w0, bad, BOXW * 2 * 64, __bfloat162float(o[0]),
           __bfloat162float(o[64]), __bfloat162float(o[128]), __bfloat162float(o[192]), __bfloat162float(o[256]), __bfloat162float(o[320]));
  }
  return 0;
}
