// Bandwidth-bound kernels around the tensor-core convs: the fused input stage
// (u8 HWC BGR -> normalise -> [reflect pre_pad / mod-pad] -> [pixel-unshuffle] -> first 3x3 conv)
// and the nearest x2 upsample.  They replace RealESRGANer.pre_process + conv_first /
// F.interpolate in the upstream forward (call site: pytorch_realesrgan.py:223).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>
#include "conv3x3_tc.cuh"

namespace b200sr {

struct FirstArgs {
  const uint8_t* src;   // [N][Hs][Ws][3] u8 BGR
  int src16;            // samples are uint16 (normalised by 65535) instead of uint8 (by 255)
  int N, Hs, Ws;
  int H1, W1;           // Hs + pre_pad, Ws + pre_pad
  int oy, ox;           // origin of this region in the padded image
  int s;                // pixel-unshuffle factor (1 or 2)
  int H, W;             // conv-domain extent of the region (= region / s)
  int cin;              // 3 * s * s
  const float* w;       // [9 taps][cin][64]
  const float* bias;    // [64]
  const float* prelu;   // [64] or nullptr
  __nv_bfloat16* out;   // 16-bit NHWC (bf16, or fp16 when out_fp16)
  int out_pitch;
  int out_fp16;
  uint8_t* lo;          // optional: e5m2 residual of the 16-bit result (TrunkLo layout, conv3x3_tc.cuh)
  float* f0;            // optional: fp32 copy of the result (tile-interleaved layout)
  float* inrgb;         // optional normalised network input [N][H][W][4] (RGB0), s == 1 only
};

// F.pad(..., 'reflect') applied twice (pre_pad, then mod-pad), right/bottom only.
__device__ __forceinline__ int reflect_src(int p, int S, int S1) {
  if (p >= S1) p = 2 * (S1 - 1) - p;
  if (p >= S) p = 2 * (S - 1) - p;
  return p;
}

template <int CIN>
__global__ void __launch_bounds__(128) first_conv_kernel(const FirstArgs a) {
  extern __shared__ float s_w[];  // [9][CIN][64]
  __shared__ float s_b[64], s_p[64];
  for (int i = threadIdx.x; i < 9 * CIN * 64; i += blockDim.x) s_w[i] = a.w[i];
  if (threadIdx.x < 64) {
    s_b[threadIdx.x] = a.bias[threadIdx.x];
    s_p[threadIdx.x] = a.prelu ? a.prelu[threadIdx.x] : 1.f;
  }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (x >= a.W) return;
  constexpr int S = (CIN == 3) ? 1 : 2;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = s_b[c];
  const size_t img = static_cast<size_t>(n) * a.Hs * a.Ws * 3;   // first sample of frame n
  float centre[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) continue;  // conv zero padding
    float in[CIN];
#pragma unroll
    for (int dy = 0; dy < S; ++dy)
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        const int sy = reflect_src(a.oy + yy * S + dy, a.Hs, a.H1);
        const int sx = reflect_src(a.ox + xx * S + dx, a.Ws, a.W1);
        const size_t p = img + (static_cast<size_t>(sy) * a.Ws + sx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)   // BGR -> RGB, img / max_range as upstream's pre_process
          in[c * S * S + dy * S + dx] = a.src16 ? static_cast<float>(reinterpret_cast<const uint16_t*>(a.src)[p + 2 - c]) / 65535.0f
                                                : static_cast<float>(a.src[p + 2 - c]) / 255.0f;
      }
    if (S == 1 && tap == 4) {
      centre[0] = in[0];
      centre[1] = in[CIN > 1 ? 1 : 0];
      centre[2] = in[CIN > 2 ? 2 : 0];
    }
    const float* wt = s_w + tap * CIN * 64;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float v = in[ci];
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wt + ci * 64 + c4 * 4);
        acc[c4 * 4 + 0] = fmaf(v, w4.x, acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(v, w4.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(v, w4.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(v, w4.w, acc[c4 * 4 + 3]);
      }
    }
  }
  if (a.prelu) {
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = acc[c] > 0.f ? acc[c] : acc[c] * s_p[c];
  }
  const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
  if (a.lo)
    store_trunk_pair(a.out + pix * a.out_pitch, a.lo + lo_off(n, y, x, a.H, a.W), acc);
  else
    store_bf16_row<64>(a.out + pix * a.out_pitch, acc, a.out_fp16);
  if (a.f0) {
    float* d = a.f0 + trunk_off(n, y, x, a.H, a.W);   // tile-interleaved fp32 layout
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(acc[g * 8 + i]);
      st_global_256(d + g * TRUNK_GSTRIDE, o);
    }
  }
  if (a.inrgb) *reinterpret_cast<float4*>(a.inrgb + pix * 4) = make_float4(centre[0], centre[1], centre[2], 0.f);
}

// out[n][Y][X][:] = in[n][Y/2][X/2][:], 64 bf16 channels (F.interpolate(scale_factor=2, mode='nearest')).
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N,
                                                          int H, int W) {
  const size_t total = static_cast<size_t>(N) * (2 * H) * (2 * W) * 8;  // 8 x 16 B per pixel
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i & 7);
    size_t p = i >> 3;
    const int X = static_cast<int>(p % (2 * W));
    p /= (2 * W);
    const int Y = static_cast<int>(p % (2 * H));
    const int n = static_cast<int>(p / (2 * H));
    out[i] = in[((static_cast<size_t>(n) * H + (Y >> 1)) * W + (X >> 1)) * 8 + v];
  }
}

}  // namespace b200sr
