// Development probe (not part of the product library): per-SM store bandwidth to L2 for the epilogue's store
// patterns, and how much concurrent stores slow the TMA loads of the same SM.
//   warp 0      : streams 16 KB bulk loads (L2-resident source) through a 4-deep shared-memory ring
//   warps 1..8  : store 32-pixel x 64-byte "rows" (NHWC, 384 B pixel pitch) until the loads are done
//     mode 0  no stores
//     mode 1  LSU, lane = pixel, 2 x 32 B per lane                 (the conv1-4 epilogue)
//     mode 2  LSU, contiguous 2 KB per warp-row                    (upper bound for LSU stores)
//     mode 3  TMA tensor store of a [32 px][64 B] box staged in shared memory (one instruction per warp-row)
//     mode 4  TMA 1-D bulk store, contiguous 2 KB per warp-row
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_store probe_store.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ptx.cuh"
#include "../tmap.h"
#include "../conv3x3_tc.cuh"

using namespace b200sr;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int BOX = 16384;
constexpr int NST = 4;
constexpr int ROWS_PER_WARP = 4;    // store footprint per warp (rows of 32 px x 384 B), cycled

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __launch_bounds__(288, 1)
t10_kernel(const __grid_constant__ CUtensorMap smap, const uint8_t* lsrc, uint8_t* sdst, int mode, int nboxes,
           int nstore_warps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;                          // NST x 16 KB
  uint8_t* stage = smem + NST * BOX;             // 8 warps x 2 buffers x 2 KB
  __shared__ uint64_t bar[NST];
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(&bar[i], 1);
    done = 0;
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* src = lsrc + static_cast<size_t>(blockIdx.x) * (16 * BOX);
      const long long t0 = clock64();
      for (int i = 0; i < nboxes + NST; ++i) {
        const int st = i % NST;
        if (i >= NST) mbar_wait(&bar[st], ((i / NST) - 1) & 1);
        if (i < nboxes) {
          mbar_arrive_expect_tx(&bar[st], BOX);
          bulk_load_1d(&bar[st], ring + st * BOX, src + (i & 15) * BOX, BOX);
        }
      }
      out[blockIdx.x * 2] = clock64() - t0;
      done = 1;
    }
  } else if (warp - 1 < nstore_warps && mode != 0) {
    const int w = warp - 1;
    uint8_t* base = sdst + (static_cast<size_t>(blockIdx.x) * 8 + w) * (ROWS_PER_WARP * 32 * 384);
    const int img_row0 = (blockIdx.x * 8 + w) * ROWS_PER_WARP;   // tensor-map view: [rows][32 px][192 ch]
    uint8_t* mystage = stage + w * 4096;
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 8 + i;
    long long rows = 0;
    while (!done) {
      const int r = static_cast<int>(rows % ROWS_PER_WARP);
      uint8_t* rowp = base + static_cast<size_t>(r) * 32 * 384;
      if (mode == 1) {
        st_global_256(rowp + lane * 384 + 128, v);
        st_global_256(rowp + lane * 384 + 160, v);
      } else if (mode == 2) {
        st_global_256(rowp + lane * 32, v);
        st_global_256(rowp + 1024 + lane * 32, v);
      } else {
        uint8_t* sb = mystage + (rows & 1) * 2048;
        if (rows >= 2) {
          if (lane == 0) bulk_wait_read1();   // the store that last used this buffer has read it
          __syncwarp();
        }
        // lane = pixel: 64 B per lane, 16-byte chunks rotated by the lane to spread the banks
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int jj = (j + (lane >> 1)) & 3;
          *reinterpret_cast<uint4*>(sb + lane * 64 + jj * 16) = make_uint4(v[0], v[1], v[2], v[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (mode == 3)
            tma_store_4d(&smap, sb, 64, 0, img_row0 + r, 0);
          else
            bulk_store_1d(rowp, sb, 2048);
          bulk_commit();
        }
      }
      ++rows;
    }
    if (mode >= 3 && lane == 0) bulk_wait_read0();
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&out[blockIdx.x * 2 + 1]), static_cast<unsigned long long>(rows));
  }
}

int main() {
  const int nsm = 148;
  uint8_t *lsrc, *sdst;
  const size_t lbytes = static_cast<size_t>(nsm) * 16 * BOX;                       // 38 MB, L2-resident
  const size_t sbytes = static_cast<size_t>(nsm) * 8 * ROWS_PER_WARP * 32 * 384;   // 58 MB
  CK(cudaMalloc(&lsrc, lbytes));
  CK(cudaMalloc(&sdst, sbytes));
  CK(cudaMemset(lsrc, 1, lbytes));
  CK(cudaMemset(sdst, 0, sbytes));
  long long* dout;
  CK(cudaMalloc(&dout, nsm * 16));
  CUtensorMap smap;
  // [N=1][H = rows][W = 32 px][192 ch] bf16, box {32 ch (64 B), 32 px}
  if (!tmap_encode_act(&smap, sdst, 1, nsm * 8 * ROWS_PER_WARP, 32, 192, 32, 32, 0)) {
    printf("tensor map failed\n");
    return 2;
  }
  const int smem_bytes = NST * BOX + 8 * 4096 + 1024;
  CK(cudaFuncSetAttribute(t10_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const char* names[] = {"no stores", "LSU lane=pixel 2x32B strided", "LSU contiguous", "TMA tensor store [32px][64B]",
                         "TMA 1-D bulk store 2 KB"};
  const int nboxes = 4000;
  for (int nw : {8, 4}) {
    printf("T10 %d storing warps per SM, loads: 16 KB bulk boxes from L2, ring depth %d\n", nw, NST);
    for (int mode = 0; mode < 5; ++mode) {
      std::vector<long long> h(nsm * 2);
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(dout, 0, nsm * 16));
        t10_kernel<<<nsm, 288, smem_bytes>>>(smap, lsrc, sdst, mode, nboxes, nw, dout);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), dout, nsm * 16, cudaMemcpyDeviceToHost));
      double cyc = 0, rows = 0;
      for (int i = 0; i < nsm; ++i) {
        cyc += h[2 * i];
        rows += h[2 * i + 1];
      }
      cyc /= nsm;
      rows /= nsm;
      printf("T10   %-32s : loads %.1f B/cyc/SM (%.0f cyc/box), stores %.1f B/cyc/SM\n", names[mode],
             static_cast<double>(nboxes) * BOX / cyc, cyc / nboxes, rows * 2048 / cyc);
    }
  }
  printf("probe done\n");
  return 0;
}
