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

// ---------------------------------------------------------------------------------------------------------------
// first_conv_tiled_kernel: the same arithmetic (same fp32 FMA order per output value: bias, then taps row-major,
// input channels innermost -> bit-identical results), laid out for the memory system.  The stage is HBM-bound
// (3 B in, 448 B out per pixel), so what matters is that every store instruction of a warp writes whole, adjacent
// lines and that enough loads are in flight:
//   * a block owns 128 pixels x ROWS rows; the source pixels it needs ((ROWS + 2) x 130, after reflect padding /
//     pixel-unshuffle / normalisation) are staged ONCE in shared memory as floats, the weights once per block;
//   * 4 threads per pixel PAIR: thread (pair, g) accumulates channels [16 g, 16 g + 16) of pixels 2 pair, 2 pair + 1,
//     so a weight vector read from shared memory feeds two pixels, and the four threads of a pixel together write its
//     whole 128-byte hi line (a warp writes 8 pixel pairs = 2 KB contiguous), 32-byte lo run and fp32 groups.
template <int CIN>
struct FirstTiled {
  static constexpr int ROWS = (CIN == 3) ? 8 : 4;
  static constexpr int SW_FLOATS = 9 * CIN * 64;
  static constexpr int SIN_PITCH = 130 * CIN + 2;                       // floats per staged row (+2: bank skew)
  static constexpr int SMEM_BYTES = (SW_FLOATS + (ROWS + 2) * SIN_PITCH + 128) * 4;
};

template <int CIN>
__global__ void __launch_bounds__(256) first_conv_tiled_kernel(const FirstArgs a) {
  using T = FirstTiled<CIN>;
  constexpr int S = (CIN == 3) ? 1 : 2;
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                      // [9][CIN][64]
  float* s_b = s_dyn + T::SW_FLOATS;       // [64]
  float* s_p = s_b + 64;                   // [64]
  float* s_in = s_p + 64;                  // [ROWS + 2][130][CIN] (row pitch SIN_PITCH)
  const int tx0 = blockIdx.x * 128, y0 = blockIdx.y * T::ROWS, n = blockIdx.z;
  for (int i = threadIdx.x * 4; i < T::SW_FLOATS; i += 256 * 4)
    *reinterpret_cast<float4*>(s_w + i) = *reinterpret_cast<const float4*>(a.w + i);
  if (threadIdx.x < 64) {
    s_b[threadIdx.x] = a.bias[threadIdx.x];
    s_p[threadIdx.x] = a.prelu ? a.prelu[threadIdx.x] : 1.f;
  }
  // stage the inputs: conv-domain pixel (y0 - 1 + r, tx0 - 1 + px), zero outside the region (conv zero padding)
  const size_t img = static_cast<size_t>(n) * a.Hs * a.Ws * 3;
  for (int i = threadIdx.x; i < (T::ROWS + 2) * 130 * S * S; i += 256) {
    const int sub = i % (S * S);
    const int px = (i / (S * S)) % 130;
    const int r = i / (S * S * 130);
    const int yy = y0 - 1 + r, xx = tx0 - 1 + px;
    float v[3] = {0.f, 0.f, 0.f};
    if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
      const int sy = reflect_src(a.oy + yy * S + sub / S, a.Hs, a.H1);
      const int sx = reflect_src(a.ox + xx * S + sub % S, a.Ws, a.W1);
      const size_t p = img + (static_cast<size_t>(sy) * a.Ws + sx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)   // BGR -> RGB, img / max_range as upstream's pre_process
        v[c] = a.src16 ? static_cast<float>(reinterpret_cast<const uint16_t*>(a.src)[p + 2 - c]) / 65535.0f
                       : static_cast<float>(a.src[p + 2 - c]) / 255.0f;
    }
    float* d = s_in + r * T::SIN_PITCH + px * CIN;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c * S * S + sub] = v[c];
  }
  __syncthreads();
  const int g = threadIdx.x & 3;            // channel group: channels [16 g, 16 g + 16)
  const int pair = threadIdx.x >> 2;        // pixels tx0 + 2 pair, + 1
  const int xl = 2 * pair;                  // tile-local x of the first pixel
  const float* wg = s_w + g * 16;
#pragma unroll 1
  for (int r = 0; r < T::ROWS; ++r) {
    const int y = y0 + r;
    if (y >= a.H) break;
    float acc[2][16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[0][c] = acc[1][c] = s_b[g * 16 + c];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float* row = s_in + (r + ky) * T::SIN_PITCH + xl * CIN;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float v0 = row[kx * CIN + ci], v1 = row[(kx + 1) * CIN + ci];
          const float* wt = wg + ((ky * 3 + kx) * CIN + ci) * 64;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wt + c4 * 4);
            acc[0][c4 * 4 + 0] = fmaf(v0, w4.x, acc[0][c4 * 4 + 0]);
            acc[0][c4 * 4 + 1] = fmaf(v0, w4.y, acc[0][c4 * 4 + 1]);
            acc[0][c4 * 4 + 2] = fmaf(v0, w4.z, acc[0][c4 * 4 + 2]);
            acc[0][c4 * 4 + 3] = fmaf(v0, w4.w, acc[0][c4 * 4 + 3]);
            acc[1][c4 * 4 + 0] = fmaf(v1, w4.x, acc[1][c4 * 4 + 0]);
            acc[1][c4 * 4 + 1] = fmaf(v1, w4.y, acc[1][c4 * 4 + 1]);
            acc[1][c4 * 4 + 2] = fmaf(v1, w4.z, acc[1][c4 * 4 + 2]);
            acc[1][c4 * 4 + 3] = fmaf(v1, w4.w, acc[1][c4 * 4 + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int x = tx0 + xl + j;
      if (x >= a.W) continue;
      float (&v)[16] = acc[j];
      if (a.prelu) {
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = v[c] > 0.f ? v[c] : v[c] * s_p[g * 16 + c];
      }
      const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
      __nv_bfloat16* o = a.out + pix * a.out_pitch + g * 16;
      uint32_t p[8];
      if (a.lo) {   // residual-stream pair: bf16 hi + e5m2 lo of the rounding residual (store_trunk_pair's split)
        float res[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          p[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          res[2 * i] = v[2 * i] - bf16lo_f32(p[i]);
          res[2 * i + 1] = v[2 * i + 1] - bf16hi_f32(p[i]);
        }
        st_global_256(o, p);
        uint4 l;
        l.x = f32x4_e5m2(res[0], res[1], res[2], res[3]);
        l.y = f32x4_e5m2(res[4], res[5], res[6], res[7]);
        l.z = f32x4_e5m2(res[8], res[9], res[10], res[11]);
        l.w = f32x4_e5m2(res[12], res[13], res[14], res[15]);
        *reinterpret_cast<uint4*>(a.lo + lo_off(n, y, x, a.H, a.W) + (g >> 1) * LO_GSTRIDE + (g & 1) * 16) = l;
      } else {
        if (a.out_fp16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) p[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) p[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        }
        st_global_256(o, p);
      }
      if (a.f0) {
        float* d = a.f0 + trunk_off(n, y, x, a.H, a.W) + static_cast<size_t>(g * 2) * TRUNK_GSTRIDE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) q[i] = __float_as_uint(v[h * 8 + i]);
          st_global_256(d + h * TRUNK_GSTRIDE, q);
        }
      }
      if (a.inrgb && g == 0) {
        const float* c = s_in + (r + 1) * T::SIN_PITCH + (xl + j + 1) * CIN;   // centre tap (S == 1 only)
        *reinterpret_cast<float4*>(a.inrgb + pix * 4) = make_float4(c[0], c[CIN > 1 ? 1 : 0], c[CIN > 2 ? 2 : 0], 0.f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// first_conv3_const_kernel (3 input channels: RRDBNet x4 and SRVGGNetCompact).  ncu on the tiled kernel above showed
// it bound by SHARED-MEMORY instruction issue (mio_throttle + short_scoreboard: 0.82 wavefronts per cycle per SM for
// the weight reads), not by HBM.  Here the 27 x 64 weights, the bias and the PReLU slopes travel as a KERNEL PARAMETER
// (7.4 KB, constant bank 0): with the loops fully unrolled every FFMA takes its weight as a constant-bank operand,
// so the inner loop is pure FFMA -- no load instructions, no shared memory -- and one thread per pixel keeps its 27
// input values in registers.  Same fp32 FMA order per output value as the other two kernels -> identical bits.
struct FirstWeights3 {
  float w[27][64];     // [(ky * 3 + kx) * 3 + ci][cout]
  float b[64];
  float p[64];         // PReLU slopes (SRVGG) when has_prelu
  int has_prelu;
};

// x / 255 and x / 65535 for integer x, correctly rounded, without the IEEE-division slow path: q0 = x * RN(1/d),
// r = fma(-q0, d, x), q = fma(r, RN(1/d), q0) equals RN(x / d) for EVERY x in 0..255 / 0..65535 (checked
// exhaustively with exact rational arithmetic, tests/test_precision_model.py), so the bytes do not change.
__device__ __forceinline__ float norm_sample(float x, float d, float rcp) {
  const float q0 = __fmul_rn(x, rcp);
  const float r = __fmaf_rn(-q0, d, x);
  return __fmaf_rn(r, rcp, q0);
}

// One pixel per thread, both 32-channel halves fully unrolled: every FFMA2 (two IEEE fma.rn per instruction -- same
// bits as scalar fmaf) takes its weight pair from a uniform register (LDCU.128), the input value is broadcast; 864
// FFMA2 + 661 LDCU per pixel.  (A two-pixels-per-thread form with the channel loop rolled -- indexed LDC.64 weight
// loads, register spills -- measured 0.62 ms for 4 x 720p but 4.8-7.2 ms for an SRVGG batch of 64 x 640x480; with the
// division-free normalisation this form takes 0.49 ms and 2.5 ms)
__global__ void __launch_bounds__(128, 5) first_conv3_const_kernel(const FirstArgs a, const __grid_constant__ FirstWeights3 cw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (x >= a.W) return;
  const size_t img = static_cast<size_t>(n) * a.Hs * a.Ws * 3;
  const float d = a.src16 ? 65535.0f : 255.0f;
  const float rcp = a.src16 ? (1.0f / 65535.0f) : (1.0f / 255.0f);
  float in[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;   // conv zero padding outside the region
      const int sy = reflect_src(a.oy + (ok ? yy : 0), a.Hs, a.H1);
      const int sx = reflect_src(a.ox + (ok ? xx : 0), a.Ws, a.W1);
      const size_t p = img + (static_cast<size_t>(sy) * a.Ws + sx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {   // BGR -> RGB, img / max_range as upstream's pre_process
        const float v = a.src16 ? static_cast<float>(__ldg(reinterpret_cast<const uint16_t*>(a.src) + p + 2 - c))
                                : static_cast<float>(__ldg(a.src + p + 2 - c));
        in[(ky * 3 + kx) * 3 + c] = ok ? norm_sample(v, d, rcp) : 0.f;
      }
    }
  const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
  const size_t loff = a.lo ? lo_off(n, y, x, a.H, a.W) : 0;
  float* f0 = a.f0 ? a.f0 + trunk_off(n, y, x, a.H, a.W) : nullptr;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float2 acc2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc2[c] = make_float2(cw.b[half * 32 + 2 * c], cw.b[half * 32 + 2 * c + 1]);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float2 vv = make_float2(in[t], in[t]);
#pragma unroll
      for (int c = 0; c < 16; ++c)
        acc2[c] = __ffma2_rn(vv, make_float2(cw.w[t][half * 32 + 2 * c], cw.w[t][half * 32 + 2 * c + 1]), acc2[c]);
    }
    float acc[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      acc[2 * c] = acc2[c].x;
      acc[2 * c + 1] = acc2[c].y;
    }
    if (cw.has_prelu) {
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = acc[c] > 0.f ? acc[c] : acc[c] * cw.p[half * 32 + c];
    }
    __nv_bfloat16* o = a.out + pix * a.out_pitch + half * 32;
    if (a.lo) {   // residual-stream pair: bf16 hi + e5m2 lo of the rounding residual (store_trunk_pair's split)
      uint32_t l[8];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t p[8];
        float res[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v0 = acc[g * 16 + 2 * i], v1 = acc[g * 16 + 2 * i + 1];
          p[i] = pack_bf16x2(v0, v1);
          res[2 * i] = v0 - bf16lo_f32(p[i]);
          res[2 * i + 1] = v1 - bf16hi_f32(p[i]);
        }
        st_global_256(o + g * 16, p);
#pragma unroll
        for (int i = 0; i < 4; ++i) l[g * 4 + i] = f32x4_e5m2(res[4 * i], res[4 * i + 1], res[4 * i + 2], res[4 * i + 3]);
      }
      st_global_256(a.lo + loff + half * LO_GSTRIDE, l);
    } else {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          p[i] = a.out_fp16 ? pack_f16x2(acc[g * 16 + 2 * i], acc[g * 16 + 2 * i + 1])
                            : pack_bf16x2(acc[g * 16 + 2 * i], acc[g * 16 + 2 * i + 1]);
        st_global_256(o + g * 16, p);
      }
    }
    if (f0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = __float_as_uint(acc[g * 8 + i]);
        st_global_256(f0 + static_cast<size_t>(half * 4 + g) * TRUNK_GSTRIDE, q);
      }
    }
  }
  if (a.inrgb) *reinterpret_cast<float4*>(a.inrgb + pix * 4) = make_float4(in[12], in[13], in[14], 0.f);   // centre tap
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
