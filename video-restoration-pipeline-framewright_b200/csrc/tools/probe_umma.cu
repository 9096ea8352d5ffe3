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
template <int N>
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
        umma_bf16(tmem + ((j & 1) ? 256 : 0), abase + aoff, bbase + boff, idesc, 1);
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

template <int N>
static void run_t2(int grid, long long* dc) {
  const int niter = 24 * 256;
  CK(cudaFuncSetAttribute(t2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int rep = 0; rep < 2; ++rep) {
    t2_kernel<N><<<grid, 128, 200 * 1024>>>(niter, dc);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> hc(grid);
  CK(cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0, mn = 1LL << 60;
  for (auto c : hc) {
    mx = c > mx ? c : mx;
    mn = c < mn ? c : mn;
  }
  printf("T2 grid=%3d SW128 N=%3d : %.1f cyc/mma (min %.1f)  math-floor=%d  smem(A+B)@128B/cyc=%d\n", grid, N,
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
  {
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
  {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    for (int grid : {1, 148}) {
      run_t2<16>(grid, dc);
      run_t2<32>(grid, dc);
      run_t2<48>(grid, dc);
      run_t2<64>(grid, dc);
      run_t2<96>(grid, dc);
      run_t2<128>(grid, dc);
      run_t2<144>(grid, dc);
      run_t2<192>(grid, dc);
      run_t2<256>(grid, dc);
    }
    cudaFree(dc);
  }
  // ------------------------------ T3
  {
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
  {
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
  printf("probe done\n");
  return 0;
}
