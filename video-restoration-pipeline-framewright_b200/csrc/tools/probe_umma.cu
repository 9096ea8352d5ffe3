// Development probe (not part of the product library): pins down on real sm_100a hardware
//  T1  whether a K-major swizzled UMMA operand may start at an arbitrary ROW of a TMA-written
//      tile (the "shifted halo view" the conv kernel relies on), and what base_offset must be;
//  T2  tcgen05.mma issue cost per instruction vs N (is SS-mode A-read from smem the bound?);
//  T3  tcgen05.ld throughput;
//  T4  TMA throughput for the conv kernel's activation box shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma probe_umma.cu -I..
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

static inline float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ----------------------------------------------------------------------------- T1
struct T1Params {
  int row_bytes;   // 64 (SW64) or 128 (SW128)
  int shift;       // row shift of the A view
  int bo_mode;     // 0: base_offset = 0 ; 1: (addr >> 7) & 7
  int c0, x0;      // TMA coords
  int a_rows;      // rows in the A box
  uint32_t a_bytes, b_bytes;
};

__global__ void __launch_bounds__(128, 1)
t1_kernel(const __grid_constant__ CUtensorMap amap, const uint8_t* __restrict__ bimg, float* __restrict__ dout,
          T1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t swz = p.row_bytes == 128 ? SWZ_128B : SWZ_64B;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, p.a_bytes + p.b_bytes);
    tma_load_4d(&amap, &bar_full, sA, p.c0, p.x0, 1, 0);
    bulk_load_1d(&bar_full, sB, bimg, p.b_bytes);
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 32);
    const int ksteps = p.row_bytes / 32;
    for (int k = 0; k < ksteps; ++k) {
      uint32_t a_addr = smem_u32(sA) + p.shift * p.row_bytes + k * 32;
      uint32_t b_addr = smem_u32(sB) + k * 32;
      uint32_t bo = p.bo_mode ? ((a_addr >> 7) & 7) : 0;
      uint64_t ad = make_smem_desc(a_addr, 8 * p.row_bytes, swz, bo);
      uint64_t bd = make_smem_desc(b_addr, 8 * p.row_bytes, swz, 0);
      umma_bf16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(&bar_done);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  const int m = warp * 32 + lane;
  for (int n = 0; n < 32; ++n) dout[m * 32 + n] = __uint_as_float(v[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// ----------------------------------------------------------------------------- T2
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) t2_kernel(int niter, long long* cyc_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t abase = make_smem_desc(smem_u32(smem), 1024, SWZ_128B, 0);
    const uint64_t bbase = make_smem_desc(smem_u32(smem) + 4 * 17408, 1024, SWZ_128B, 0);
    long long t0 = clock64();
    for (int it = 0; it < niter; it += 24) {
#pragma unroll
      for (int j = 0; j < 24; ++j) {
        const uint32_t aoff = ((j & 3) * 17408 + (j % 3) * 128 + ((j >> 2) & 3) * 32) >> 4;
        const uint32_t boff = (((j >> 2) & 3) * 32) >> 4;
        umma_bf16(tmem + (N <= 128 ? (j % NACC) * 128 : (j % (NACC > 2 ? 2 : NACC)) * 256), abase + aoff, bbase + boff, idesc, 1);
      }
    }
    umma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    long long t1 = clock64();
    cyc_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int NACC>
static void run_t2(int grid, long long* dc) {
  const int niter = 24 * 256;
  CK(cudaFuncSetAttribute(t2_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int rep = 0; rep < 2; ++rep) {
    t2_kernel<N, NACC><<<grid, 128, 200 * 1024>>>(niter, dc);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> hc(grid);
  CK(cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0, mn = 1LL << 60;
  for (auto c : hc) {
    mx = c > mx ? c : mx;
    mn = c < mn ? c : mn;
  }
  printf("T2 grid=%3d SW128 N=%3d accumulators=%d : %.1f cyc/mma (min %.1f)  math-floor=%d  smem(A+B)@128B/cyc=%d\n", grid, N, NACC,
         (double)mx / niter, (double)mn / niter, N / 2, (4096 + N * 32) / 128);
}

// ----------------------------------------------------------------------------- T3
__global__ void __launch_bounds__(128, 1) t3_kernel(int niter, int mode, long long* cyc_out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < niter; ++it) {
      uint32_t v[32];
      tmem_ld32(tmem + (it & 15) * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i];
    }
  } else {
    for (int it = 0; it < niter; it += 2) {
      uint32_t v[32], w[32];
      tmem_ld32(tmem + (it & 15) * 32, v);
      tmem_ld32(tmem + ((it + 1) & 15) * 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i] ^ w[i];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc_out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// ----------------------------------------------------------------------------- T4
__global__ void __launch_bounds__(32, 1)
t4_kernel(const __grid_constant__ CUtensorMap amap, int box_bytes, int niter, int depth, int W, int H, int cin,
          int cbox, long long* cyc_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const int xt = W / 128;
    const int nchunk = cin / cbox;
    long long t0 = clock64();
    for (int it = 0; it < niter + depth; ++it) {
      int s = it % depth;
      if (it >= depth) mbar_wait(&bar[s], ((it / depth) - 1) & 1);
      if (it < niter) {
        int lin = blockIdx.x * 17 + it;
        int c = lin % nchunk;
        int y = (lin / nchunk) % H;
        int x = ((lin / nchunk / H) % xt) * 128 - 1;
        mbar_arrive_expect_tx(&bar[s], box_bytes);
        tma_load_4d(&amap, &bar[s], smem + s * (cbox == 64 ? 17408 : 8704), c * cbox, x, y, 0);
      }
    }
    long long t1 = clock64();
    cyc_out[blockIdx.x] = t1 - t0;
  }
}


// ----------------------------------------------------------------------------- T5
// MMA rate (N=96, one accumulator range) with concurrent TMA fills of other smem stages and/or
// concurrent TMEM loads by 4..8 "epilogue" warps.  mode bit0: TMA traffic, bit1: LDTM traffic, bit2: 8 polling warps.
__global__ void __launch_bounds__(320, 1)
t5_kernel(const __grid_constant__ CUtensorMap amap, int niter, int mode, int tma_per_12, long long* cyc_out,
          float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done, bar_tma[4], bar_never;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (100 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_never, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_tma[i], 1);
    stop_flag = 0;
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint8_t* sB = smem + 4 * 17408;              // 24 KB of B
  uint8_t* sT = smem + 4 * 17408 + 32 * 1024;  // 4 TMA landing stages (not read by the MMAs)
  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 96);
      const uint64_t abase = make_smem_desc(smem_u32(smem), 1024, SWZ_128B, 0);
      const uint64_t bbase = make_smem_desc(smem_u32(sB), 1024, SWZ_128B, 0);
      long long t0 = clock64();
      for (int it = 0; it < niter; it += 24) {
#pragma unroll
        for (int j = 0; j < 24; ++j) {
          const uint32_t aoff = ((j & 3) * 17408 + (j % 3) * 128 + ((j >> 2) & 3) * 32) >> 4;
          const uint32_t boff = (((j >> 2) & 3) * 32 + (j % 3) * 6144) >> 4;
          umma_bf16(tmem + 128, abase + aoff, bbase + boff, idesc, 1);
        }
      }
      umma_commit(&bar_done);
      mbar_wait(&bar_done, 0);
      long long t1 = clock64();
      cyc_out[blockIdx.x] = t1 - t0;
      stop_flag = 1;
    }
  } else if (warp == 1) {
    if (lane == 0 && (mode & 1)) {
      // TMA traffic: tma_per_12 boxes per 12 MMAs (~ per 672 cycles at full rate); unpaced if 0
      int it = 0;
      long long tstart = clock64();
      while (!stop_flag) {
        int s = it & 3;
        if (it >= 4) mbar_wait(&bar_tma[s], ((it >> 2) - 1) & 1);
        if (tma_per_12 > 0) {
          long long target = tstart + (long long)it * 672 / tma_per_12;
          while (clock64() < target && !stop_flag) {
          }
        }
        mbar_arrive_expect_tx(&bar_tma[s], 130 * 128);
        tma_load_4d(&amap, &bar_tma[s], sT + s * 17408, ((it + blockIdx.x) % 3) * 64, ((it * 7 + blockIdx.x) % 9) * 128 - 1,
                    (it + blockIdx.x * 5) % 64, 0);
        ++it;
      }
      // drain
      for (int k = (it > 4 ? it - 4 : 0); k < it; ++k) mbar_wait(&bar_tma[k & 3], (k >> 2) & 1);
    }
  } else {
    if ((mode & 2) && warp < 6) {
      uint32_t acc = 0;
      const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      while (!stop_flag) {
        uint32_t v[32];
        tmem_ld32(tl + 256, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
      }
      sink[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
    } else if (mode & 4) {
      // polling warps: spin on a barrier that never completes (what idle epilogue warps do)
      while (!stop_flag) {
        if (mbar_try_wait(&bar_never, 0)) break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// ----------------------------------------------------------------------------- T6
// MMA bursts as the conv kernel issues them: 12 x (M128 x N96 x K16) then `ncommit` tcgen05.commit to distinct
// mbarriers; accumulator column offset `col0 + 32 * (burst % nrows)`.
template <int BURST>
__global__ void __launch_bounds__(128, 1) t6_kernel(int nburst, int ncommit, int col0, int nrows, long long* cyc_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done, bar_c[4];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (100 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_c[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 96);
    const uint64_t abase = make_smem_desc(smem_u32(smem), 1024, SWZ_128B, 0);
    const uint64_t bbase = make_smem_desc(smem_u32(smem) + 4 * 17408, 1024, SWZ_128B, 0);
    long long t0 = clock64();
    for (int b = 0; b < nburst; ++b) {
      const uint32_t dcol = tmem + col0 + 32 * (b % nrows);
      const uint64_t ad0 = abase + (uint64_t)(((b & 3) * 17408) >> 4);
      if (elect_one_sync()) {
#pragma unroll
        for (int i = 0; i < BURST; ++i) {
          const int st = i / 12, dx = (i % 12) >> 2, k = i & 3;
          umma_bf16(dcol, ad0 + (uint64_t)((((st + b) & 3) * 17408 - (b & 3) * 17408 + dx * 128 + k * 32) >> 4),
                    bbase + (uint64_t)((dx * 12288 + k * 32) >> 4), idesc, 1);
        }
        for (int c = 0; c < ncommit; ++c) umma_commit(&bar_c[c]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) {
      umma_commit(&bar_done);
      mbar_wait(&bar_done, 0);
      long long t1 = clock64();
      cyc_out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// ----------------------------------------------------------------------------- T7
// TMA activation-box pipeline: `depth` boxes [130 px][64 ch] in flight per SM, lean issue loop (coordinates
// advance by adds only).  Reports cycles per box and the latency of a single box (depth 1).
template <int DEPTH>
__global__ void __launch_bounds__(32, 1)
t7_kernel(const __grid_constant__ CUtensorMap amap, int niter, int H, int ystride, long long* cyc_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[DEPTH];
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    int y = (blockIdx.x * 37) % H;
    const int x = (blockIdx.x % 10) * 128 - 1;
    int c = 0;
    uint32_t par = 0;
    long long t0 = clock64();
    for (int it = 0; it < niter; it += DEPTH) {
#pragma unroll
      for (int s = 0; s < DEPTH; ++s) {
        if (it > 0) mbar_wait(&bar[s], par ^ 1);
        mbar_arrive_expect_tx(&bar[s], 130 * 128);
        tma_load_4d(&amap, &bar[s], smem + s * 17408, c, x, y, 0);
        y += ystride;
        if (y >= H) y -= H;
      }
      par ^= 1;
    }
#pragma unroll
    for (int s = 0; s < DEPTH; ++s) mbar_wait(&bar[s], par ^ 1);
    long long t1 = clock64();
    cyc_out[blockIdx.x] = t1 - t0;
  }
}
template <int DEPTH>
static void run_t7(const CUtensorMap& amap, int H, int ystride, const char* what, long long* dc) {
  CK(cudaFuncSetAttribute(t7_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int niter = 1536;
  for (int rep = 0; rep < 2; ++rep) {
    t7_kernel<DEPTH><<<148, 32, 220 * 1024>>>(amap, niter, H, ystride, dc);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> hc(148);
  CK(cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
  long long mx = 0;
  double mean = 0;
  for (auto v : hc) { mx = v > mx ? v : mx; mean += v / 148.0; }
  printf("T7 %-22s depth=%2d : %.0f cyc/box (mean %.0f) -> %.1f B/cyc/SM, chip %.2f TB/s @1.7GHz\n", what, DEPTH,
         (double)mx / niter, mean / niter, 16640.0 * niter / mx, 16640.0 * niter / mean * 148 * 1.7e9 / 1e12);
}


// ----------------------------------------------------------------------------- T8
// How far can the issuing thread run ahead of the tensor pipe?  Issue `n` back-to-back MMAs (N = 96), record the
// time when issue returns and when the commit fires.
__global__ void __launch_bounds__(128, 1) t8_kernel(int n, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (100 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 96);
    const uint64_t abase = make_smem_desc(smem_u32(smem), 1024, SWZ_128B, 0);
    const uint64_t bbase = make_smem_desc(smem_u32(smem) + 4 * 17408, 1024, SWZ_128B, 0);
    long long t0 = clock64();
    for (int it = 0; it < n; it += 12) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const uint32_t aoff = ((j & 3) * 17408 + (j % 3) * 128 + ((j >> 2) & 3) * 32) >> 4;
        umma_bf16(tmem + 128, abase + aoff, bbase + (((j >> 2) & 3) * 32 >> 4), idesc, 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// ----------------------------------------------------------------------------- T9
// Epilogue store patterns: 8 warps per CTA each store `rows` rows of 32 pixels (NHWC, 384 B pixel pitch).
//  mode 0: lane = pixel, 2 x 32 B per lane (64 B slice)           -> 2 instr, 32 lines each
//  mode 1: lane pair = pixel (64 B contiguous per pair)           -> 2 instr, 16 lines each
//  mode 2: fully contiguous 1 KB per instruction (reference)      -> 2 instr
//  mode 3: lane = pixel, 4 x 32 B per lane (128 B slice, conv5)   -> 4 instr, 32 lines each
//  mode 4: 4 lanes = pixel (one full 128 B line per 4 lanes)      -> 4 instr, 8 lines each
__global__ void __launch_bounds__(256, 1) t9_kernel(uint8_t* out, int rows, int mode, long long* cyc_out, int wrap, int nwarps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t warp_base = ((size_t)blockIdx.x * 8 + warp) * (size_t)wrap * 32 * 384;
  if (warp >= nwarps) rows = 0;
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 8 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < rows; ++r) {
    uint8_t* rowp = out + warp_base + (size_t)(r % wrap) * 32 * 384;
    if (mode == 0) {
      st_global_256(rowp + lane * 384, v);
      st_global_256(rowp + lane * 384 + 32, v);
    } else if (mode == 1) {
      st_global_256(rowp + (lane >> 1) * 384 + (lane & 1) * 32, v);
      st_global_256(rowp + (16 + (lane >> 1)) * 384 + (lane & 1) * 32, v);
    } else if (mode == 2) {
      st_global_256(rowp + lane * 32, v);
      st_global_256(rowp + 1024 + lane * 32, v);
    } else if (mode == 3) {
#pragma unroll
      for (int j = 0; j < 4; ++j) st_global_256(rowp + lane * 384 + j * 32, v);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) st_global_256(rowp + (j * 8 + (lane >> 2)) * 384 + (lane & 3) * 32, v);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc_out[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (!tmap_init()) {
    printf("cuTensorMapEncodeTiled entry point not found\n");
    return 3;
  }
  // ------------------------------ T1
  if (getenv("PROBE_ALL")) {
    const int W = 200, H = 2, C = 64;
    std::vector<__nv_bfloat16> hx((size_t)H * W * C);
    srand(1);
    for (auto& v : hx) v = __float2bfloat16((float)((rand() % 9) - 4));
    __nv_bfloat16* dx;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    float* dd;
    CK(cudaMalloc(&dd, 128 * 32 * 4));
    uint8_t* db;
    CK(cudaMalloc(&db, 32 * 128));
    CK(cudaFuncSetAttribute(t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int rb : {64, 128}) {
      const int cbox = rb / 2;
      const int a_rows = 144;
      CUtensorMap amap;
      if (!tmap_encode_act(&amap, dx, 1, H, W, C, cbox, a_rows, rb == 128 ? 128 : 64)) {
        printf("T1 tensor map encode failed rb=%d\n", rb);
        continue;
      }
      // weights B[n][k], k < cbox, pre-swizzled image: row n at n*rb, 16B chunk j at j ^ (line & mask)
      std::vector<float> hb(32 * cbox);
      for (auto& v : hb) v = (float)((rand() % 7) - 3);
      std::vector<uint8_t> bimg(32 * rb, 0);
      for (int n = 0; n < 32; ++n)
        for (int k = 0; k < cbox; ++k) {
          int j = k / 8, e = k % 8;
          uint32_t off = n * rb + j * 16;
          uint32_t mask = rb == 128 ? 7 : 3;
          uint32_t sw = off ^ (((off >> 7) & mask) << 4);
          __nv_bfloat16 b = __float2bfloat16(hb[n * cbox + k]);
          memcpy(&bimg[sw + e * 2], &b, 2);
        }
      CK(cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
      for (int c0 : {0, cbox == 32 ? 32 : 0}) {
        for (int bo_mode = 0; bo_mode < 1; ++bo_mode) {
          for (int shift : {0, 1, 2, 3, 5, 8, 9}) {
            T1Params p;
            p.row_bytes = rb;
            p.shift = shift;
            p.bo_mode = bo_mode;
            p.c0 = c0;
            p.x0 = -1;
            p.a_rows = a_rows;
            p.a_bytes = a_rows * rb;
            p.b_bytes = 32 * rb;
            CK(cudaMemset(dd, 0xff, 128 * 32 * 4));
            t1_kernel<<<1, 128, 64 * 1024>>>(amap, db, dd, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("T1 rb=%d c0=%d bo=%d shift=%d : CUDA error %s\n", rb, c0, bo_mode, shift, cudaGetErrorString(e));
              return 4;
            }
            std::vector<float> hd(128 * 32);
            CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0, firstbad = -1;
            for (int m = 0; m < 128; ++m)
              for (int n = 0; n < 32; ++n) {
                float ref = 0;
                int x = m + shift - 1;  // box starts at x0 = -1, row y = 1
                for (int k = 0; k < cbox; ++k) {
                  float a = (x >= 0 && x < W) ? bf2f(hx[((size_t)1 * W + x) * C + c0 + k]) : 0.f;
                  ref += a * hb[n * cbox + k];
                }
                if (hd[m * 32 + n] != ref) {
                  if (firstbad < 0) firstbad = m;
                  ++bad;
                }
              }
            printf("T1 rowbytes=%3d c0=%2d base_offset_mode=%d shift=%d : %s (mismatches=%d firstbadrow=%d)\n", rb, c0,
                   bo_mode, shift, bad ? "FAIL" : "ok", bad, firstbad);
          }
        }
      }
    }
    cudaFree(dx);
    cudaFree(dd);
    cudaFree(db);
  }
  // ------------------------------ T2
  if (getenv("PROBE_ALL")) {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    for (int grid : {148}) {
      run_t2<32, 1>(grid, dc);
      run_t2<32, 2>(grid, dc);
      run_t2<32, 4>(grid, dc);
      run_t2<64, 1>(grid, dc);
      run_t2<64, 2>(grid, dc);
      run_t2<96, 1>(grid, dc);
      run_t2<96, 2>(grid, dc);
      run_t2<96, 3>(grid, dc);
      run_t2<96, 4>(grid, dc);
      run_t2<128, 1>(grid, dc);
      run_t2<128, 2>(grid, dc);
      run_t2<192, 1>(grid, dc);
      run_t2<192, 2>(grid, dc);
      run_t2<256, 1>(grid, dc);
      run_t2<256, 2>(grid, dc);
    }
    cudaFree(dc);
  }
  // ------------------------------ T3
  if (getenv("PROBE_ALL")) {
    long long* dc;
    float* sink;
    CK(cudaMalloc(&dc, 148 * 8));
    CK(cudaMalloc(&sink, 148 * 128 * 4));
    for (int mode : {0, 1}) {
      const int niter = 2048;
      t3_kernel<<<148, 128>>>(niter, mode, dc, sink);
      CK(cudaDeviceSynchronize());
      std::vector<long long> hc(148);
      CK(cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto c : hc) mx = c > mx ? c : mx;
      double per = (double)mx / niter;
      printf("T3 tcgen05.ld 32x32b.x32 x4 warps (+32 FADD), lds per wait=%d : %.1f cyc per (4 warps x 4 KB) -> %.1f B/cyc/SM\n",
             mode + 1, per, 16384.0 / per);
    }
    cudaFree(dc);
    cudaFree(sink);
  }
  // ------------------------------ T4
  if (getenv("PROBE_ALL")) {
    const int W = 1280, H = 64, C = 192;
    __nv_bfloat16* dx;
    CK(cudaMalloc(&dx, (size_t)H * W * C * 2));
    CK(cudaMemset(dx, 0, (size_t)H * W * C * 2));
    long long* dc;
    CK(cudaMalloc(&dc, 296 * 8));
    CK(cudaFuncSetAttribute(t4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int cbox : {32, 64}) {
      for (int boxw : {130, 66, 34}) {
        CUtensorMap amap;
        if (!tmap_encode_act(&amap, dx, 1, H, W, C, cbox, boxw, cbox == 64 ? 128 : 64)) {
          printf("T4 encode failed\n");
          continue;
        }
        const int box_bytes = boxw * cbox * 2;
        for (int depth : {2, 4, 8, 12}) {
          for (int grid : {1, 148}) {
            const int niter = 2048;
            for (int rep = 0; rep < 2; ++rep) {
              t4_kernel<<<grid, 32, 220 * 1024>>>(amap, box_bytes, niter, depth, W, H, C, cbox, dc);
              CK(cudaDeviceSynchronize());
            }
            std::vector<long long> hc(grid);
            CK(cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (auto c : hc) mx = c > mx ? c : mx;
            printf("T4 TMA box {%dch x %dpx} grid=%3d depth=%2d : %.1f cyc/box (%.2f cyc/row) -> %.1f B/cyc/SM\n", cbox,
                   boxw, grid, depth, (double)mx / niter, (double)mx / niter / boxw, (double)box_bytes * niter / mx);
          }
        }
      }
    }
    cudaFree(dx);
    cudaFree(dc);
  }
  // ------------------------------ T5
  if (getenv("PROBE_ALL")) {
    const int W = 1280, H = 64, C = 192;
    __nv_bfloat16* dx;
    CK(cudaMalloc(&dx, (size_t)H * W * C * 2));
    CK(cudaMemset(dx, 0, (size_t)H * W * C * 2));
    long long* dc;
    float* sink;
    CK(cudaMalloc(&dc, 148 * 8));
    CK(cudaMalloc(&sink, 148 * 320 * 4));
    CK(cudaFuncSetAttribute(t5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CUtensorMap amap;
    if (tmap_encode_act(&amap, dx, 1, H, W, C, 64, 130, 128)) {
      const int niter = 24 * 256;
      struct Cfg { int mode, rate; const char* name; } cfgs[] = {
          {0, 0, "MMA only"}, {4, 0, "MMA + 8 polling warps"}, {1, 1, "MMA + TMA 1 box / 12 MMA"},
          {1, 2, "MMA + TMA 2 boxes / 12 MMA"}, {1, 0, "MMA + TMA unpaced"}, {2, 0, "MMA + 4 warps LDTM"},
          {3, 1, "MMA + TMA 1/12 + LDTM"}, {7, 1, "MMA + TMA 1/12 + LDTM + polling"}};
      for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
          t5_kernel<<<148, 320, 220 * 1024>>>(amap, niter, c.mode, c.rate, dc, sink);
          CK(cudaDeviceSynchronize());
        }
        std::vector<long long> hc(148);
        CK(cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : hc) mx = v > mx ? v : mx;
        printf("T5 N=96 %-34s : %.1f cyc/mma\n", c.name, (double)mx / niter);
      }
    }
    cudaFree(dx);
    cudaFree(dc);
    cudaFree(sink);
  }
  // ------------------------------ T6
  if (getenv("PROBE_ALL")) {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    auto run = [&](auto kern, int burst, int ncommit) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      const int nburst = 6144 / burst;
      for (int rep = 0; rep < 2; ++rep) {
        kern<<<148, 128, 200 * 1024>>>(nburst, ncommit, 32, 12, dc);
        CK(cudaDeviceSynchronize());
      }
      std::vector<long long> hc(148);
      CK(cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto v : hc) mx = v > mx ? v : mx;
      printf("T6 bursts of %2d x N96, commits/burst=%d : %.1f cyc/mma (%.0f cyc overhead per burst)\n", burst, ncommit,
             (double)mx / (nburst * burst), (double)mx / nburst - 56.1 * burst);
    };
    for (int nc : {0, 1, 2, 4}) {
      run(t6_kernel<12>, 12, nc);
      run(t6_kernel<24>, 24, nc);
      run(t6_kernel<48>, 48, nc);
    }
    cudaFree(dc);
  }
  // ------------------------------ T7
  if (getenv("PROBE_ALL")) {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    for (int big = 0; big < 2; ++big) {
      const int W = 1280, H = big ? 2880 : 64, C = 192;   // 64 rows = 31 MB (L2 resident), 2880 rows = 1.4 GB (DRAM)
      __nv_bfloat16* dx;
      CK(cudaMalloc(&dx, (size_t)H * W * C * 2));
      CK(cudaMemset(dx, 0, (size_t)H * W * C * 2));
      CUtensorMap amap;
      if (tmap_encode_act(&amap, dx, 1, H, W, C, 64, 130, 128)) {
        const char* what = big ? "DRAM (1.4 GB tensor)" : "L2 (31 MB tensor)";
        const int ys = big ? 19 : 1;
        run_t7<1>(amap, H, ys, what, dc);
        run_t7<2>(amap, H, ys, what, dc);
        run_t7<4>(amap, H, ys, what, dc);
        run_t7<6>(amap, H, ys, what, dc);
        run_t7<8>(amap, H, ys, what, dc);
        run_t7<12>(amap, H, ys, what, dc);
      }
      cudaFree(dx);
    }
    cudaFree(dc);
  }
  // ------------------------------ T8
  if (getenv("PROBE_ALL")) {
    long long* dc;
    CK(cudaMalloc(&dc, 16));
    CK(cudaFuncSetAttribute(t8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int n : {12, 24, 48, 96, 192, 384, 768}) {
      for (int rep = 0; rep < 2; ++rep) {
        t8_kernel<<<1, 128, 200 * 1024>>>(n, dc);
        CK(cudaDeviceSynchronize());
      }
      long long h[2];
      CK(cudaMemcpy(h, dc, 16, cudaMemcpyDeviceToHost));
      printf("T8 %4d MMAs N96: issue returned after %6lld cyc (%.1f/mma), all complete after %6lld cyc (%.1f/mma)\n", n, h[0],
             (double)h[0] / n, h[1], (double)h[1] / n);
    }
    cudaFree(dc);
  }
  // ------------------------------ T9
  {
    const int rows = 256;
    uint8_t* buf;
    const size_t bytes = (size_t)148 * 8 * rows * 32 * 384;
    CK(cudaMalloc(&buf, bytes));
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    const char* names[] = {"lane=pixel 2x32B (now, conv1-4)", "lane pair=pixel 64B", "contiguous 1KB (reference)",
                           "lane=pixel 4x32B (now, conv5)", "4 lanes=pixel full 128B line"};
    for (int cfg = 0; cfg < 4; ++cfg) {
      const int wrap = cfg == 0 ? rows : 2;
      const int nwarps = cfg <= 1 ? 8 : (cfg == 2 ? 2 : 1);
      printf("T9 footprint %s, %d storing warps per SM\n", cfg == 0 ? "3.7 GB (DRAM)" : "29 MB (L2)", nwarps);
      for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
          t9_kernel<<<148, 256>>>(buf, rows, mode, dc, wrap, nwarps);
          CK(cudaDeviceSynchronize());
        }
        std::vector<long long> hc(148);
        CK(cudaMemcpy(hc.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : hc) mx = v > mx ? v : mx;
        const double bytes_row = (mode >= 3 ? 4096.0 : 2048.0);
        printf("T9   %-34s : %.0f cyc per warp-row -> %.1f B/cyc/SM\n", names[mode], (double)mx / rows,
               bytes_row * nwarps * rows / mx);
      }
    }
    cudaFree(buf);
    cudaFree(dc);
  }
  printf("probe done\n");
  return 0;
}
