// Development probe (not part of the product library): tcgen05.mma.cta_group::2 (CTA pair, M = 256).
//  T12a  correctness of the operand split this repo would use: A = 128 pixel rows per CTA (own shared memory),
//        B = N/2 weight rows per CTA, D = each CTA's own 128 TMEM lanes x N columns;
//  T12b  issue cost per M256 x N x K16 instruction for N = 96 and N = 192 (one CTA pair per two SMs, all SMs busy),
//        against the cta_group::1 cost measured by probe_umma.cu T2: max(N/2, (4096 + 32 N)/128).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_2cta probe_2cta.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ptx.cuh"

using namespace b200sr;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major SWIZZLE_128B tile writer: element (row, k) of a [rows][64] bf16 tile with 128-byte rows
__device__ __forceinline__ void put_sw128(uint8_t* tile, int row, int k, float v) {
  const int chunk = (k >> 3) ^ (row & 7);
  *reinterpret_cast<__nv_bfloat16*>(tile + row * 128 + chunk * 16 + (k & 7) * 2) = __float2bfloat16(v);
}
__host__ __device__ inline float a_val(int m, int k) { return static_cast<float>((m * 3 + k) % 7 - 3); }
__host__ __device__ inline float b_val(int n, int k) { return static_cast<float>((n * 5 + k * 2) % 5 - 2); }

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
t12_kernel(float* out, int nloop, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // [128][64] bf16
  uint8_t* sB = smem + 16384;    // [N/2][64] bf16
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x >> 1;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) put_sw128(sA, i >> 6, i & 63, a_val(rank * 128 + (i >> 6), i & 63));
  for (int i = threadIdx.x; i < (N / 2) * 64; i += blockDim.x)
    put_sw128(sB, i >> 6, i & 63, b_val(rank * (N / 2) + (i >> 6), i & 63));
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  constexpr uint32_t NCOLS = N <= 128 ? 128 : 256;
  if (warp == 0) {
    tmem_alloc2(&s_tmem, NCOLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint64_t adesc = make_smem_desc(smem_u32(sA), 1024, SWZ_128B, 0);
  const uint64_t bdesc = make_smem_desc(smem_u32(sB), 1024, SWZ_128B, 0);
  const uint32_t idesc = make_idesc_16(256, N, false);
  if (rank == 0 && threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < nloop; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma2_bf16(tmem, adesc + static_cast<uint64_t>((k * 32) >> 4), bdesc + static_cast<uint64_t>((k * 32) >> 4), idesc,
                   (it > 0 || k > 0) ? 1u : 0u);
    }
    umma2_commit_mc(&bar, 0b11);
    if (cyc) cyc[pair] = clock64() - t0;
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (out && pair == 0) {
    // each CTA: 128 TMEM lanes (its pixels) x N columns
    const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tl + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(rank * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, NCOLS);
}

template <int N>
void run() {
  float* dout;
  long long* dc;
  CK(cudaMalloc(&dout, 256 * N * 4));
  CK(cudaMalloc(&dc, 74 * 8));
  const int smem_bytes = 16384 + 16384 + 1024;
  CK(cudaFuncSetAttribute(t12_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  // correctness: one pair, one pass over K = 64
  t12_kernel<N><<<2, 128, smem_bytes>>>(dout, 1, nullptr);
  CK(cudaDeviceSynchronize());
  std::vector<float> h(256 * N);
  CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < 64; ++k) ref += a_val(m, k) * b_val(n, k);
      if (h[m * N + n] != ref) {
        if (bad < 4) printf("  N=%d m=%d n=%d got %g want %g\n", N, m, n, h[m * N + n], ref);
        ++bad;
      }
    }
  printf("T12a cta_group::2 M=256 N=%d K=64: %d mismatches of %d\n", N, bad, 256 * N);
  // issue cost: 74 pairs = all 148 SMs
  const int nloop = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    t12_kernel<N><<<148, 128, smem_bytes>>>(nullptr, nloop, dc);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> hc(74);
  CK(cudaMemcpy(hc.data(), dc, 74 * 8, cudaMemcpyDeviceToHost));
  double s = 0;
  for (auto v : hc) s += v;
  const double per = s / 74 / (nloop * 4.0);
  printf("T12b cta_group::2 M=256 N=%d: %.1f cycles per MMA (math bound N/2 = %d; cta_group::1 M=128 measured max(N/2,(4096+32N)/128) = %.0f)\n",
         N, per, N / 2, (N / 2.0 > (4096 + 32.0 * N) / 128 ? N / 2.0 : (4096 + 32.0 * N) / 128));
  cudaFree(dout);
  cudaFree(dc);
}

int main() {
  run<96>();
  run<192>();
  run<64>();
  printf("probe done\n");
  return 0;
}
