// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on the sm_100a tensor cores.
//
// Replaces (per layer) what the reference executes through PyTorch -> cuDNN inside
// `RRDBNet.forward` / `SRVGGNetCompact.forward` (constructed at
// /root/reference/src/framewright/processors/pytorch_realesrgan.py:107-127, run by
// `upsampler.enhance` at :223), together with the elementwise ops that follow each conv
// (bias, LeakyReLU/PReLU, torch.cat, the 0.2 residual scalings, clamp/round/quantise).
//
// Formulation ("row streaming, accumulator stationary"):
//   * NHWC bf16 activations; a CTA owns an output tile of 128 pixels (one image-row segment)
//     x TH rows.  GEMM M = the 128 pixels, N = output channels, K = input channels x taps.
//   * One TMA box = one input row segment [130 px][64 ch] (128-byte rows, SWIZZLE_128B).
//     The three horizontal taps are three row-shifted UMMA descriptor views of that single
//     box (start address + dx*128 B; hardware swizzle is a function of the absolute smem
//     address, verified by csrc/tools/probe_umma.cu), so every activation byte is staged
//     into shared memory exactly once.  Zero padding at the image (or tile) border is the
//     TMA out-of-bounds fill.
//   * The three vertical taps are stacked along N: input row y feeds output rows y-1, y, y+1,
//     whose fp32 accumulators sit in adjacent TMEM column blocks, so one
//     tcgen05.mma M128 x N(3*Cout) x K16 updates all three (A is read from smem once per
//     3*Cout columns instead of once per Cout: the SS-mode smem read of A is the bound for
//     N <= 128).
//   * All TH accumulator rows (TH*Cout <= 512 TMEM columns) stay resident while the input
//     channels are swept in 64-wide chunks (weights of one chunk resident in smem, double
//     buffered).  Rows finish one by one during the last chunk; per-row full/empty
//     mbarriers let the epilogue warps drain row Y while the MMA warp is already working on
//     later rows / the next tile.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 =
//     epilogue (TMEM -> registers -> fused pointwise -> global).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "ptx.cuh"

namespace b200sr {

enum EpiMode : int {
  EPI_ACT_BF16 = 0,   // y = act(acc + b), act = leaky(slope) (slope = 1 -> identity); bf16 slice store
  EPI_PRELU_BF16,     // y = prelu(acc + b, a[c]); bf16 store
  EPI_RDB5,           // x <- (acc + b) * 0.2 + x                  (x kept as a bf16 hi + e5m2 lo pair)
  EPI_RDB5_RRDB,      // x <- ((acc + b) * 0.2 + x) * 0.2 + x0 ; x0 <- x
  EPI_ADD_F32,        // y = acc + b + f[c]; bf16 store   (conv_body: feat + body_feat)
  EPI_LAST_U8,        // 3 real channels: clamp(acc + b, 0, 1) -> round(255 v) -> u8 BGR (cropped)
  EPI_SRVGG_LAST      // 48 ch: pixel-shuffle(4) + nearest(x) residual -> clamp/round -> u8 BGR
};

struct ConvArgs {
  int N, H, W;        // spatial extent of this conv (input == output)
  int nchunks;        // ceil(Cin / 64)
  int last_ksteps;    // K16 steps in the last chunk: 4, or 2 when Cin % 64 == 32
  int TH;             // output rows per tile, <= 512 / COUT
  int xtiles, ytiles, ntiles;
  const uint8_t* wpack;   // packed weights: [chunk][dx][3*COUT rows][128 B] (SWIZZLE_128B image)
  const float* bias;      // COUT floats
  const float* prelu;     // COUT floats (EPI_PRELU_BF16)
  float slope;            // leaky slope (EPI_ACT_BF16)
  int in_planes;          // 0: input chunk c = channels [64c, 64c+64) of one NHWC tensor.  P > 0: chunk-planar
                          // input, chunk c is its own [N][H][W][64] tensor and image n of it is TMA image c*N + n
  int in_up2;             // the input is the nearest-2x upsampling of a tensor of half this conv's H and W, read
                          // through the duplicated-pixel TMA view (tmap.h::tmap_encode_act_up2): box = 66 source
                          // pixels = 132 rows starting at x0 - 2, so every tap view starts one row later
  int sub;                // sub-pixel phase of "nearest-2x upsample, then this conv": the conv runs on the LOW-resolution
                          // grid (N, H, W are its dims) with the 2x2 taps of phase (sub_a, sub_b) -- the weights that
                          // read the same source pixel are pre-summed and embedded in a 3x3 image whose unused ky / kx
                          // blocks are never issued: vertical blocks [sub_vlo, sub_vhi], dx tiles [sub_dlo, sub_dhi] --
                          // and output pixel (y, x) lands at (2 y + sub_a, 2 x + sub_b) of the 2H x 2W tensor `out`
  int sub_a, sub_b, sub_vlo, sub_vhi, sub_dlo, sub_dhi;
  int w_resident;         // single-chunk conv (Cin <= 64): the weight chunk is loaded ONCE per CTA and stays in shared
                          // memory; the second weight buffer's space becomes extra activation stages (more bytes in
                          // flight per SM: the HR-tail convs are bound by DRAM latency x bytes in flight)
  int abl;                // timing ablations (dev option "abl", results are WRONG when set): 1 = no epilogue stores,
                          // 2 = no MMAs issued (commits only), 4 = no TMA activation loads
  int in_fp16;            // A (activations) and B (weights) are fp16 instead of bf16
  int out_fp16;           // 16-bit output tensor is fp16 instead of bf16
  __nv_bfloat16* out;     // 16-bit NHWC destination (bf16 or fp16 per out_fp16)
  int out_pitch;          // channels per pixel in `out`
  int out_choff;          // first channel written
  // residual stream x = hi + lo (TrunkLo comment below): hi is the bf16 NHWC tensor the next convs read
  const __nv_bfloat16* hi_in;   // x.hi of the RDB input  (NHWC, pitch out_pitch, channels 0..63)
  const uint8_t* lo_in;         // x.lo of the RDB input  (e5m2, tile-interleaved)
  uint8_t* lo_out;              // x.lo of the result     (may alias lo_in or xb_lo: same pixel, same thread)
  const __nv_bfloat16* xb_hi;   // RRDB input x0.hi (EPI_RDB5_RRDB; may alias `out`)
  const uint8_t* xb_lo;         // RRDB input x0.lo
  const float* fadd;      // fp32 addend (EPI_ADD_F32) / normalised network input RGBx (EPI_SRVGG_LAST)
  uint8_t* dst;           // BGR destination frame(s)  [N][dst_h][dst_w][3], uint8 or (dst16) uint16 samples
  int dst16;              // 16-bit output samples: round(clamp(v, 0, 1) * 65535)  (upstream's max_range = 65535 branch)
  int dst_h, dst_w;       // destination frame size
  int crop_y0, crop_x0;   // first conv-output pixel kept
  int crop_h, crop_w;     // kept extent
  int dst_y0, dst_x0;     // where the kept window lands in the destination frame
};

// CTAs per SM.  1: one persistent CTA per SM with all 512 TMEM columns, 8 epilogue warps (default).
// 2: two co-resident CTAs with half of TMEM / shared memory each (4 epilogue warps, single weight buffer for
// Cout >= 48).  Measured on B200 at 720p x4 the two are within 1 % (20.8 vs 21.0 frames/s): the RDB convs are
// limited by DRAM traffic, not by tensor-pipe issue bubbles.
#ifndef B200SR_CTAS_PER_SM
#define B200SR_CTAS_PER_SM 1
#endif

template <int COUT>
struct ConvCfg {
  static constexpr int CTAS_PER_SM = B200SR_CTAS_PER_SM;
  static constexpr int TMEM_COLS = 512 / CTAS_PER_SM;
  static constexpr int MAXTH = TMEM_COLS / COUT;
  static constexpr int WTILE_BYTES = 3 * COUT * 128;    // (chunk, dx): 3 dy-blocks x COUT rows x 128 B
  static constexpr int WCHUNK_BYTES = 3 * WTILE_BYTES;
  static constexpr int A_ROWS = 130;
  static constexpr int A_BOX_BYTES = A_ROWS * 128;
  static constexpr int A_STAGE_BYTES = 17 * 1024;
  static constexpr int NWBUF = (CTAS_PER_SM == 1 || COUT <= 32) ? 2 : 1;   // weight-chunk buffers
  static constexpr int SMEM_BUDGET = (CTAS_PER_SM == 1 ? 222 : 110) * 1024;
  static constexpr int NSTAGES_FIT = (SMEM_BUDGET - NWBUF * WCHUNK_BYTES) / A_STAGE_BYTES;
  static constexpr int NSTAGES = NSTAGES_FIT > 6 ? 6 : NSTAGES_FIT;
  static_assert(NSTAGES >= 2, "not enough shared memory for two activation stages");
  // resident-weights mode (ConvArgs::w_resident): one weight buffer, everything else is activation stages
  static constexpr int MAXSTAGES = 12;
  static constexpr int NSTAGES_RES_FIT = (SMEM_BUDGET - WCHUNK_BYTES) / A_STAGE_BYTES;
  static constexpr int NSTAGES_RES = NSTAGES_RES_FIT > MAXSTAGES ? MAXSTAGES : NSTAGES_RES_FIT;
  static_assert(NSTAGES_RES >= NSTAGES && NSTAGES <= MAXSTAGES, "stage counts");
  static constexpr int SMEM_BYTES_DB = NWBUF * WCHUNK_BYTES + NSTAGES * A_STAGE_BYTES + 1024 /*align slack*/;
  static constexpr int SMEM_BYTES_RES = WCHUNK_BYTES + NSTAGES_RES * A_STAGE_BYTES + 1024;
  static constexpr int SMEM_BYTES = SMEM_BYTES_DB > SMEM_BYTES_RES ? SMEM_BYTES_DB : SMEM_BYTES_RES;
  static constexpr int NEPI_WARPS = CTAS_PER_SM == 1 ? 8 : 4;   // epilogue warps (multiple of 4)
  static constexpr int NTHREADS = 32 * (2 + NEPI_WARPS);        // warp 0 TMA, warp 1 MMA, rest epilogue
};

#ifdef B200SR_ABL_STORE_SCRATCH
__device__ uint8_t g_abl_scratch[8 << 20];
#endif
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
#ifdef B200SR_ABL_STORE_SCRATCH   // timing ablation (tools only): same store instructions, 8 MB L2-resident target
  if (B200SR_ABL_STORE_SCRATCH != 2) p = g_abl_scratch + (reinterpret_cast<uintptr_t>(p) & ((8u << 20) - 32));
#endif
#ifdef B200SR_ABL_NOSTORE   // timing ablation (tools only): keep the math alive, skip the store
  if (reinterpret_cast<uintptr_t>(p) != 1) return;
#endif
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p)
               : "memory");
}
// Evict-first variants for data that is not touched again during the launch (streams through L2 without pushing
// out the rows other CTAs are about to read).
__device__ __forceinline__ void st_global_256_ef(void* p, const uint32_t (&v)[8]) {
#if defined(B200SR_ABL_STORE_SCRATCH)
  if (B200SR_ABL_STORE_SCRATCH != 3) p = g_abl_scratch + (reinterpret_cast<uintptr_t>(p) & ((8u << 20) - 32));
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
#elif defined(B200SR_ABL_NOSTORE) || defined(B200SR_ABL_NOHINT)
  st_global_256(p, v);
#else
  asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
#endif
}
__device__ __forceinline__ void ld_global_256_ef(const void* p, uint32_t (&v)[8]) {
#if defined(B200SR_ABL_NOHINT)
  ld_global_256(p, v);
#else
  asm volatile("ld.global.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p)
               : "memory");
#endif
}
// Evict-last variant (EXPERIMENT -DB200SR_RDB_STORE_LAST: the dense-block intermediates, read again 1..7 steps later)
#ifndef B200SR_RDB_STORE_LAST
#define B200SR_RDB_STORE_LAST 0
#endif
__device__ __forceinline__ void st_global_256_el(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// Non-volatile variant: the compiler may hoist and batch these (used for read-only / read-before-write data).
__device__ __forceinline__ void ld_global_256_nv(const void* p, uint32_t (&v)[8]) {
  asm("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// B200SR_EPI_XPOSE (fused RDB kernel epilogues, full 128-pixel column tiles; same bytes in memory).  The per-pixel threads
// of a warp touch 32 DIFFERENT cache lines per 32-byte vector store (pixel pitch 128 B).  With the switch the lanes
// (2i, 2i+1) = pixels (x, x+1) exchange one of each two 32-byte blocks through the register file, so that every store
// instruction covers 64 CONTIGUOUS bytes per lane pair: 16 lines and half the L1 -> L2 requests per instruction, for
// 8 SHFL + 24 SEL per 64 bytes.  bit 0: conv1-4's 64 bytes per pixel (-0.4 % step time, four A/B pairs on two boxes);
// bit 1: conv5's 128 bytes of hi where the output is hi alone (another -0.1..-0.5 %, at the noise level).  Output bytes
// identical.  The same trick on conv5's hi LOADS lost 9 % (shuffles on the epilogue's critical path), and on the stores of
// the un-specialised epilogue 3.4 % (register spills at the 168-register cap): profiles/r02c_epi_xpose_ab.txt.
#ifndef B200SR_EPI_XPOSE
#define B200SR_EPI_XPOSE 3
#endif
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// The fp32 tensor F0 (conv_first output, added back after conv_body) uses a tile-interleaved layout,
// [n][y][x / 128][channel / 8][x % 128][channel % 8], so that the per-pixel threads of a warp read and write
// CONSECUTIVE 32-byte chunks (1 KB per warp instruction) instead of 32 different cache lines.  trunk_off() is the
// float offset of (pixel, channel group 0); group g is TRUNK_GSTRIDE * g floats further.
constexpr int TRUNK_GSTRIDE = 1024;
__device__ __forceinline__ size_t trunk_off(int n, int y, int x, int H, int W) {
  const int xt = (W + 127) >> 7;
  return ((static_cast<size_t>(n) * H + y) * xt + (x >> 7)) * (8 * TRUNK_GSTRIDE) + static_cast<size_t>(x & 127) * 8;
}

// TrunkLo.  The RRDB residual stream x is carried between RDBs as a PAIR: hi = bf16(x), which is at the same time
// the NHWC conv input of the next RDB, and lo = e5m2(x - hi), one byte per value.  hi + lo keeps ~12 significant
// bits of x where bf16 alone keeps 8; measured against the fp32 oracle (tests/test_precision_model.py) the pair is
// indistinguishable from an fp32 stream (57.1 vs 57.2 dB) while bf16 alone fails the 1-LSB gate.  The point is
// bytes: an SM stores only ~21 B/cycle to L2 (csrc/tools/probe_umma.cu T9), and conv5's epilogue was bound by
// exactly that when the stream was fp32 (384-640 B per pixel stored; now 192).
// Where the pair is kept: at the RRDB boundaries (the RRDB input x0 / output, added back with gain 1 at every RRDB end)
// -- INSIDE an RRDB the outputs of its first two RDBs are carried as hi alone (lo_in / lo_out == nullptr): their bf16
// rounding enters the RRDB output with gain 0.2, which the CPU emulation (tests/emulate.py, trunk_mode "hybrid")
// and the GPU parity tests put at 56.0-56.4 dB / 100 % within 1 LSB against 57.1 dB for the pair everywhere -- and
// saves 256 of the 1344 B per pixel the three conv5 epilogues of an RRDB move (option trunk_lo = 1: pair everywhere).
// lo layout: [n][y][x / 128][channel / 32][x % 128][channel % 32] bytes -> a warp's 32 pixels x 32 B are contiguous.
constexpr int LO_GSTRIDE = 128 * 32;   // bytes between the two 32-channel groups of a pixel
__device__ __forceinline__ size_t lo_off(int n, int y, int x, int H, int W) {
  const int xt = (W + 127) >> 7;
  return ((static_cast<size_t>(n) * H + y) * xt + (x >> 7)) * (2 * LO_GSTRIDE) + static_cast<size_t>(x & 127) * 32;
}
__device__ __forceinline__ float bf16lo_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// four e5m2 bytes -> four floats (e5m2 is the upper byte of an fp16)
__device__ __forceinline__ void e5m2x4_f32(uint32_t w, float (&f)[4]) {
  const uint32_t h01 = __byte_perm(w, 0, 0x1404), h23 = __byte_perm(w, 0, 0x3424);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h01));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h23));
  f[0] = a.x;
  f[1] = a.y;
  f[2] = b.x;
  f[3] = b.y;
}
__device__ __forceinline__ uint32_t f32x4_e5m2(float a, float b, float c, float d) {
  const uint32_t p01 = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E5M2);
  const uint32_t p23 = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E5M2);
  return p01 | (p23 << 16);
}
// Splits 64 fp32 values into the pair and stores it: hi -> NHWC bf16 `hi_dst` (64 channels), lo -> `lo_dst`.
template <bool LO_OUT = true>
__device__ __forceinline__ void store_trunk_pair(__nv_bfloat16* hi_dst, uint8_t* lo_dst, const float (&v)[64],
                                                 bool xp = false, int pitch = 0);
__device__ __forceinline__ uint8_t quant_u8(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return static_cast<uint8_t>(__float2int_rn(v * 255.0f));
}
__device__ __forceinline__ uint16_t quant_u16(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return static_cast<uint16_t>(__float2int_rn(v * 65535.0f));
}
// one output sample (channel ch of the BGR pixel at element index `px3`), 8- or 16-bit
__device__ __forceinline__ void store_sample(const ConvArgs& a, size_t px3, int ch, float v);

// Load this thread's COUT accumulator columns of one tile row.
template <int COUT>
__device__ __forceinline__ void load_acc_row(uint32_t taddr, float (&acc)[COUT]) {
  if constexpr (COUT == 16) {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[i]);
  } else if constexpr (COUT == 32) {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(v[i]);
  } else if constexpr (COUT == 48) {
    uint32_t v[32], w[16];
    tmem_ld32(taddr, v);
    tmem_ld16(taddr + 32, w);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(v[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[32 + i] = __uint_as_float(w[i]);
  } else {
    static_assert(COUT == 64, "unsupported COUT");
    uint32_t v[32], w[32];
    tmem_ld32(taddr, v);
    tmem_ld32(taddr + 32, w);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(v[i]);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[32 + i] = __uint_as_float(w[i]);
  }
}

// B200SR_EPI_XPOSE: stores the two 32-byte blocks `b` of this lane's pixel (`dst` = its first block, `pitch` = bytes to
// pixel x + 1) in the pair-coalesced order.  All 32 lanes must call it (shuffles); lane parity = pixel parity.
template <int HINT = 0>   // 1: L2::evict_first
__device__ __forceinline__ void store_blocks_paired(uint8_t* dst, ptrdiff_t pitch, const uint32_t (&b)[2][8]) {
  const bool odd = (threadIdx.x & 1) != 0;
  uint32_t r[8], dE[8], dO[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __shfl_xor_sync(0xffffffffu, odd ? b[0][i] : b[1][i], 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dE[i] = odd ? r[i] : b[0][i];    // the even lane's pixel: block 0 (own) | block 1 (received)
    dO[i] = odd ? b[1][i] : r[i];    // the odd lane's pixel:  block 0 (received) | block 1 (own)
  }
  if constexpr (HINT == 1) {
    st_global_256_ef(odd ? dst - pitch + 32 : dst, dE);
    st_global_256_ef(odd ? dst + 32 : dst + pitch, dO);
  } else {
    st_global_256(odd ? dst - pitch + 32 : dst, dE);
    st_global_256(odd ? dst + 32 : dst + pitch, dO);
  }
}

template <int COUT>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* dst, const float (&v)[COUT], int fp16 = 0, bool xp = false,
                                               int pitch = 0) {
  if constexpr (COUT == 32 && (B200SR_EPI_XPOSE & 1) != 0) {
    if (xp) {   // warp-uniform: a full column tile of the fused RDB kernel (bf16 intermediates)
      uint32_t b[2][8];
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) b[g][i] = pack_bf16x2(v[g * 16 + 2 * i], v[g * 16 + 2 * i + 1]);
      store_blocks_paired(reinterpret_cast<uint8_t*>(dst), static_cast<ptrdiff_t>(pitch) * 2, b);
      return;
    }
  }
#pragma unroll
  for (int g = 0; g < COUT / 16; ++g) {
    uint32_t p[8];
    if (fp16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = pack_f16x2(v[g * 16 + 2 * i], v[g * 16 + 2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = pack_bf16x2(v[g * 16 + 2 * i], v[g * 16 + 2 * i + 1]);
    }
    if constexpr (COUT == 32 && B200SR_RDB_STORE_LAST != 0)
      st_global_256_el(dst + g * 16, p);
    else
      st_global_256(dst + g * 16, p);
  }
}

// LO_OUT = false (fused RDB kernel, compile-time): the stream leaves as hi alone -- no rounding residual is computed.
// xp (warp-uniform, B200SR_EPI_XPOSE bit 1): the 128 bytes of hi leave as pair-coalesced stores (store_blocks_paired).
template <bool LO_OUT>
__device__ __forceinline__ void store_trunk_pair(__nv_bfloat16* hi_dst, uint8_t* lo_dst, const float (&v)[64], bool xp,
                                                 int pitch) {
  if constexpr (!LO_OUT) {
    if ((B200SR_EPI_XPOSE & 2) != 0 && xp) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {   // 64 bytes of the pixel at a time
        uint32_t b[2][8];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) b[g][i] = pack_bf16x2(v[(2 * k + g) * 16 + 2 * i], v[(2 * k + g) * 16 + 2 * i + 1]);
        store_blocks_paired<1>(reinterpret_cast<uint8_t*>(hi_dst) + k * 64, static_cast<ptrdiff_t>(pitch) * 2, b);
      }
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pack_bf16x2(v[g * 16 + 2 * i], v[g * 16 + 2 * i + 1]);
        st_global_256_ef(hi_dst + g * 16, p);
      }
    }
  } else {
  float r[64];   // residual after the bf16 rounding
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 16 + 2 * i;
      p[i] = pack_bf16x2(v[c], v[c + 1]);
      r[c] = v[c] - bf16lo_f32(p[i]);
      r[c + 1] = v[c + 1] - bf16hi_f32(p[i]);
    }
    st_global_256_ef(hi_dst + g * 16, p);
  }
  if (lo_dst == nullptr) return;   // inside an RRDB the stream is carried as hi alone (TrunkLo comment above)
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 32 + 4 * i;
      p[i] = f32x4_e5m2(r[c], r[c + 1], r[c + 2], r[c + 3]);
    }
    st_global_256_ef(lo_dst + g * LO_GSTRIDE, p);
  }
  }   // LO_OUT
}

// The RDB tail of one pixel.  `hi`/`lo` hold the pixel's x pair (RDB input) as loaded from hi_in / lo_in:
//   x = hi + lo ; v = (acc + b) * 0.2 + x ; RRDB end: v = v * 0.2 + (x0.hi + x0.lo) ; result stored as a pair.
// Shared by the per-conv kernel and the fused RDB kernel so that both produce the same bits.
// LO_IN / LO_OUT = false (the fused RDB kernel's compile-time modes): the RDB input / output is hi alone -- `lo` is not
// read (x = hi + 0, the same bits as decoding all-zero lo bytes) / no residual is split off.
template <bool RRDB, bool LO = true, bool LO_OUT = true>
__device__ __forceinline__ void trunk_pixel(const ConvArgs& a, const float* s_bias, float (&acc)[64],
                                            const uint32_t (&hi)[4][8], const uint32_t (&lo)[2][8],
                                            const uint32_t (&h0)[4][8], const uint32_t (&l0)[2][8], int n, int y,
                                            int x, bool xp = false) {
  const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
  const size_t loff = lo_off(n, y, x, a.H, a.W);
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float lf[4], l0f[4];
      if constexpr (LO) e5m2x4_f32(lo[g][i], lf);
      if constexpr (RRDB) e5m2x4_f32(l0[g][i], l0f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 32 + i * 4 + j;
        const uint32_t hw = hi[c >> 4][(c & 15) >> 1];
        // (!LO: "+ 0.f" stays -- the same bits as decoding all-zero lo bytes, also for x = -0; dropping it measured flat)
        const float xv = ((c & 1) ? bf16hi_f32(hw) : bf16lo_f32(hw)) + (LO ? lf[j] : 0.f);
        float v = (acc[c] + s_bias[c]) * 0.2f + xv;
        if constexpr (RRDB) {
          const uint32_t h0w = h0[c >> 4][(c & 15) >> 1];
          v = v * 0.2f + (((c & 1) ? bf16hi_f32(h0w) : bf16lo_f32(h0w)) + l0f[j]);
        }
        acc[c] = v;
      }
    }
  store_trunk_pair<LO_OUT>(a.out + pix * a.out_pitch + a.out_choff, (LO_OUT && a.lo_out) ? a.lo_out + loff : nullptr, acc,
                           xp, a.out_pitch);
}
// one pixel's pair (64 channels) from an NHWC hi tensor + the tile-interleaved lo bytes
// (LO = false, the fused RDB kernel's compile-time modes: the stream is hi alone and `lo` is left untouched)
template <bool LO = true>
__device__ __forceinline__ void load_trunk_pair(const __nv_bfloat16* hi_px, const uint8_t* lo_px, uint32_t (&hi)[4][8],
                                                uint32_t (&lo)[2][8]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) ld_global_256_ef(hi_px + g * 16, hi[g]);
  if constexpr (!LO) return;
  if (lo_px != nullptr) {
#pragma unroll
    for (int g = 0; g < 2; ++g) ld_global_256_ef(lo_px + g * LO_GSTRIDE, lo[g]);
  } else {   // hi-only stream: lo = +0 (all-zero e5m2 bytes)
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int i = 0; i < 8; ++i) lo[g][i] = 0u;
  }
}

__device__ __forceinline__ void store_sample(const ConvArgs& a, size_t px3, int ch, float v) {
  if (a.dst16)
    reinterpret_cast<uint16_t*>(a.dst)[px3 + ch] = quant_u16(v);
  else
    a.dst[px3 + ch] = quant_u8(v);
}

// Fused pointwise tail of one output pixel (one thread).
template <int COUT, int EPI>
__device__ __forceinline__ void epilogue_pixel(const ConvArgs& a, const float* s_bias, const float* s_prelu,
                                               float (&acc)[COUT], int n, int y, int x, bool xp = false) {
  const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
  if constexpr (EPI == EPI_ACT_BF16) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      float v = acc[c] + s_bias[c];
      acc[c] = v > 0.f ? v : v * a.slope;
    }
    const size_t opix = a.sub ? (static_cast<size_t>(n) * (2 * a.H) + (2 * y + a.sub_a)) * (2 * a.W) + (2 * x + a.sub_b) : pix;
    store_bf16_row<COUT>(a.out + opix * a.out_pitch + a.out_choff, acc, a.out_fp16, xp, a.out_pitch);
  } else if constexpr (EPI == EPI_PRELU_BF16) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      float v = acc[c] + s_bias[c];
      acc[c] = v > 0.f ? v : v * s_prelu[c];
    }
    store_bf16_row<COUT>(a.out + pix * a.out_pitch + a.out_choff, acc, a.out_fp16);
  } else if constexpr (EPI == EPI_RDB5 || EPI == EPI_RDB5_RRDB) {
    static_assert(COUT == 64, "trunk epilogues are 64-channel");
    uint32_t hi[4][8], lo[2][8], h0[4][8], l0[2][8];
    const size_t loff = lo_off(n, y, x, a.H, a.W);
    load_trunk_pair(a.hi_in + pix * a.out_pitch, a.lo_in ? a.lo_in + loff : nullptr, hi, lo);
    if constexpr (EPI == EPI_RDB5_RRDB) load_trunk_pair(a.xb_hi + pix * a.out_pitch, a.xb_lo + loff, h0, l0);
    trunk_pixel<EPI == EPI_RDB5_RRDB>(a, s_bias, acc, hi, lo, h0, l0, n, y, x);
  } else if constexpr (EPI == EPI_ADD_F32) {
    const float* f = a.fadd + trunk_off(n, y, x, a.H, a.W);
#pragma unroll
    for (int hh = 0; hh < COUT / 32; ++hh) {
      uint32_t r[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) ld_global_256_nv(f + (hh * 4 + g) * TRUNK_GSTRIDE, r[g]);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = hh * 32 + g * 8 + i;
          acc[c] = acc[c] + s_bias[c] + __uint_as_float(r[g][i]);
        }
    }
    store_bf16_row<COUT>(a.out + pix * a.out_pitch + a.out_choff, acc, a.out_fp16);
  } else if constexpr (EPI == EPI_LAST_U8) {
    const int cy = y - a.crop_y0, cx = x - a.crop_x0;
    if (cy >= 0 && cy < a.crop_h && cx >= 0 && cx < a.crop_w) {
      const size_t d = ((static_cast<size_t>(n) * a.dst_h + (a.dst_y0 + cy)) * a.dst_w + (a.dst_x0 + cx)) * 3;
      store_sample(a, d, 0, acc[2] + s_bias[2]);  // B
      store_sample(a, d, 1, acc[1] + s_bias[1]);  // G
      store_sample(a, d, 2, acc[0] + s_bias[0]);  // R
    }
  } else {
    static_assert(EPI == EPI_SRVGG_LAST && COUT == 48, "SRVGG tail needs 48 channels");
    const float4 in = *reinterpret_cast<const float4*>(a.fadd + pix * 4);  // normalised RGB of this LR pixel
    const float base[3] = {in.x, in.y, in.z};
    // The pixel's 4 x 4 HR block: 12 contiguous bytes per HR row.  When the whole block lies inside the crop window
    // (always, except at tile-mode crop borders) and samples are 8-bit, each HR row is three aligned 32-bit stores
    // instead of twelve byte stores (the byte form made this epilogue half of the kernel's time); the byte offset
    // 3 * (dst_x0 - crop_x0 + 4 x) is a multiple of 4 whenever dst_x0 - crop_x0 is a multiple of 4 (it is 4 * pad).
    const int cy0 = y * 4 - a.crop_y0, cx0 = x * 4 - a.crop_x0;
    const bool whole = !a.dst16 && cy0 >= 0 && cy0 + 3 < a.crop_h && cx0 >= 0 && cx0 + 3 < a.crop_w &&
                       (((a.dst_x0 + cx0) * 3) & 3) == 0 && ((a.dst_w * 3) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.dst) & 3) == 0;
    if (whole) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t by[12];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int o = c * 16 + i * 4 + j;
            by[3 * j + (2 - c)] = quant_u8(acc[o] + s_bias[o] + base[c]);
          }
        const size_t d = ((static_cast<size_t>(n) * a.dst_h + (a.dst_y0 + cy0 + i)) * a.dst_w + (a.dst_x0 + cx0)) * 3;
        uint32_t* dw = reinterpret_cast<uint32_t*>(a.dst + d);
#pragma unroll
        for (int w4 = 0; w4 < 3; ++w4)
          dw[w4] = by[4 * w4] | (by[4 * w4 + 1] << 8) | (by[4 * w4 + 2] << 16) | (by[4 * w4 + 3] << 24);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cy = y * 4 + i - a.crop_y0;
      if (cy < 0 || cy >= a.crop_h) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cx = x * 4 + j - a.crop_x0;
        if (cx < 0 || cx >= a.crop_w) continue;
        const size_t d = ((static_cast<size_t>(n) * a.dst_h + (a.dst_y0 + cy)) * a.dst_w + (a.dst_x0 + cx)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int o = c * 16 + i * 4 + j;
          store_sample(a, d, 2 - c, acc[o] + s_bias[o] + base[c]);
        }
      }
    }
  }
}

template <int COUT, int EPI>
__global__ void __launch_bounds__(ConvCfg<COUT>::NTHREADS, ConvCfg<COUT>::CTAS_PER_SM)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap amap, const ConvArgs args) {
  using Cfg = ConvCfg<COUT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const bool wres = args.w_resident != 0;                 // weights loaded once, one buffer (single-chunk convs)
  const int NST = wres ? Cfg::NSTAGES_RES : Cfg::NSTAGES;  // activation stages in the ring
  uint8_t* sW = smem;                                     // NWBUF (resident: 1) x WCHUNK_BYTES
  uint8_t* sA = smem + (wres ? 1 : Cfg::NWBUF) * Cfg::WCHUNK_BYTES;   // NST x A_STAGE_BYTES

  __shared__ uint64_t bar_full[Cfg::MAXSTAGES], bar_empty[Cfg::MAXSTAGES];
  __shared__ uint64_t bar_wfull[Cfg::NWBUF], bar_wempty[Cfg::NWBUF];
  __shared__ uint64_t bar_rfull[Cfg::MAXTH], bar_rempty[Cfg::MAXTH];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_bias[COUT];
  __shared__ float s_prelu[COUT];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int TH = args.TH;

  if (threadIdx.x < COUT) {
    s_bias[threadIdx.x] = args.bias[threadIdx.x];
    s_prelu[threadIdx.x] = (EPI == EPI_PRELU_BF16) ? args.prelu[threadIdx.x] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::MAXSTAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < Cfg::NWBUF; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], 1);
    }
    for (int i = 0; i < Cfg::MAXTH; ++i) {
      mbar_init(&bar_rfull[i], 1);
      mbar_init(&bar_rempty[i], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&amap);
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int tiles_per_img = args.xtiles * args.ytiles;

  // Roles 0 and 1 run with the WHOLE warp converged (values stay in uniform registers, no
  // per-instruction lane-election loops); a single elected lane issues the TMA / MMA instructions.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0, phase = 0, wb = 0, wphase = 0;
    bool wload = true;   // resident weights: only the first (tile, chunk) loads them
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
      const int n = t / tiles_per_img;
      const int r = t - n * tiles_per_img;
      const int ty = r / args.xtiles;
      const int tx = r - ty * args.xtiles;
      const int x0 = tx * 128 - 1;
      const int y0 = ty * TH;
      for (int c = 0; c < args.nchunks; ++c) {
        if (!wres) mbar_wait(&bar_wempty[wb], wphase ^ 1);
        if (wload && elect_one_sync()) {
          mbar_arrive_expect_tx(&bar_wfull[wb], Cfg::WCHUNK_BYTES);
          const uint8_t* wsrc = args.wpack + static_cast<size_t>(c) * Cfg::WCHUNK_BYTES;
#pragma unroll
          for (int d = 0; d < 3; ++d)
            bulk_load_1d(&bar_wfull[wb], sW + wb * Cfg::WCHUNK_BYTES + d * Cfg::WTILE_BYTES,
                         wsrc + d * Cfg::WTILE_BYTES, Cfg::WTILE_BYTES);
        }
        __syncwarp();
        if (wres) wload = false;
        for (int y = -1; y <= TH; ++y) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          if (args.abl & 4) {
            if (elect_one_sync()) mbar_arrive(&bar_full[stage]);
          } else if (elect_one_sync()) {
            mbar_arrive_expect_tx(&bar_full[stage], args.in_up2 ? 132 * 128 : Cfg::A_BOX_BYTES);
            if (args.in_up2)   // rows y0+y = -1 and >= H map to source rows -1 and >= H/2: zero-filled
              tma_load_5d(&amap, &bar_full[stage], sA + stage * Cfg::A_STAGE_BYTES, c * 64, 0, (x0 + 1) / 2 - 1,
                          (y0 + y) >> 1, n);
            else if (args.in_planes)
              tma_load_4d(&amap, &bar_full[stage], sA + stage * Cfg::A_STAGE_BYTES, 0, x0, y0 + y, c * args.N + n);
            else
              tma_load_4d(&amap, &bar_full[stage], sA + stage * Cfg::A_STAGE_BYTES, c * 64, x0, y0 + y, n);
          }
          __syncwarp();
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (!wres && ++wb == Cfg::NWBUF) {
          wb = 0;
          wphase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const bool f16 = args.in_fp16 != 0;
    const uint32_t idesc1 = make_idesc_16(128, COUT, f16);
    const uint32_t idesc2 = make_idesc_16(128, 2 * COUT, f16);
    const uint32_t idesc3 = make_idesc_16(128, 3 * COUT, f16);
    const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 1024, SWZ_128B, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(sW), 1024, SWZ_128B, 0);
    int stage = 0, phase = 0, wb = 0, wphase = 0;
    uint32_t tile_iter = 0;
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x, ++tile_iter) {
      const uint32_t rparity = (tile_iter & 1) ^ 1;  // accumulator row drained by the previous tile's epilogue
      for (int c = 0; c < args.nchunks; ++c) {
        const int ks = (c == args.nchunks - 1) ? args.last_ksteps : 4;
        const bool first_chunk = (c == 0);
        const bool last_chunk = (c == args.nchunks - 1);
        if (!wres || tile_iter == 0) mbar_wait(&bar_wfull[wb], wphase);   // resident: loaded once, never released
        const uint64_t bdesc_w = bdesc0 + static_cast<uint64_t>((wb * Cfg::WCHUNK_BYTES) >> 4);
        if (args.sub) {
          // sub-pixel phase: vertical blocks [vlo, vhi] and dx tiles [dlo, dhi] only (the other weights are zero and
          // are never issued).  Accumulator row R is first touched by input row R + 1 - vhi and complete after input
          // row R + 1 - vlo.
          const int vlo = args.sub_vlo, vhi = args.sub_vhi, dlo = args.sub_dlo, dhi = args.sub_dhi;
          for (int y = -1; y <= TH; ++y) {
            int blk_lo = (y < 1) ? (1 - y) : 0;
            int blk_hi = (TH - y < 2) ? (TH - y) : 2;
            blk_lo = blk_lo > vlo ? blk_lo : vlo;
            blk_hi = blk_hi < vhi ? blk_hi : vhi;
            const int nblk = blk_hi - blk_lo + 1;
            const bool new_row = first_chunk && nblk >= 1 && blk_hi == vhi;   // row y-1+vhi touched for the first time
            if (new_row) mbar_wait(&bar_rempty[y - 1 + vhi], rparity);
            mbar_wait(&bar_full[stage], phase);
            tc_fence_after();
            const uint32_t dcol = tmem_base + static_cast<uint32_t>((y - 1 + blk_lo) * COUT);
            const uint64_t ad0 = adesc0 + static_cast<uint64_t>((stage * Cfg::A_STAGE_BYTES) >> 4);
            const uint64_t bd0 = bdesc_w + static_cast<uint64_t>((blk_lo * COUT * 128) >> 4);
            const uint32_t idesc_n = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
            if (elect_one_sync()) {
              if (nblk >= 1) {
                bool first = true;
                for (int dx = dlo; dx <= dhi; ++dx)
                  for (int k = 0; k < ks; ++k) {
                    const uint64_t ad = ad0 + static_cast<uint64_t>((dx * 128 + k * 32) >> 4);
                    const uint64_t bd = bd0 + static_cast<uint64_t>((dx * Cfg::WTILE_BYTES + k * 32) >> 4);
                    if (first && new_row) {
                      if (nblk > 1) umma_bf16(dcol, ad, bd, nblk == 3 ? idesc2 : idesc1, 1);
                      umma_bf16(dcol + (nblk - 1) * COUT, ad, bd + static_cast<uint64_t>(((nblk - 1) * COUT * 128) >> 4),
                                idesc1, 0);
                    } else {
                      umma_bf16(dcol, ad, bd, idesc_n, 1);
                    }
                    first = false;
                  }
              }
              umma_commit(&bar_empty[stage]);
              const int R = y - 1 + vlo;                                   // the row this input row completes
              if (last_chunk && R >= 0 && R < TH) umma_commit(&bar_rfull[R]);
              if (y == TH && !wres) umma_commit(&bar_wempty[wb]);
            }
            __syncwarp();
            if (++stage == NST) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else {
        for (int y = -1; y <= TH; ++y) {
          const int blk_lo = (y < 1) ? (1 - y) : 0;        // output row y-1+blk must be >= 0
          const int blk_hi = (TH - y < 2) ? (TH - y) : 2;  // and < TH
          const int nblk = blk_hi - blk_lo + 1;
          const bool new_row = first_chunk && blk_hi == 2;  // accumulator row y+1 is touched for the first time
          if (new_row) mbar_wait(&bar_rempty[y + 1], rparity);
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          const uint32_t dcol = tmem_base + static_cast<uint32_t>((y - 1 + blk_lo) * COUT);
          const uint64_t ad0 = adesc0 + static_cast<uint64_t>((stage * Cfg::A_STAGE_BYTES + (args.in_up2 ? 128 : 0)) >> 4);
          const uint64_t bd0 = bdesc_w + static_cast<uint64_t>((blk_lo * COUT * 128) >> 4);
          const uint32_t idesc_n = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
          if (elect_one_sync()) {
            // first K16 step of the stage (may start a new accumulator row), then 11 (or 5) plain steps;
            // two straight-line variants so no per-MMA predicates are needed
            if (args.abl & 2) {
            } else if (new_row) {
              if (nblk > 1) umma_bf16(dcol, ad0, bd0, nblk == 3 ? idesc2 : idesc1, 1);
              umma_bf16(dcol + (nblk - 1) * COUT, ad0, bd0 + static_cast<uint64_t>(((nblk - 1) * COUT * 128) >> 4),
                        idesc1, 0);
            } else {
              umma_bf16(dcol, ad0, bd0, idesc_n, 1);
            }
            if (args.abl & 2) {
            } else if (ks == 4) {
#pragma unroll
              for (int i = 1; i < 12; ++i) {
                const int dx = i >> 2, k = i & 3;
                umma_bf16(dcol, ad0 + static_cast<uint64_t>((dx * 128 + k * 32) >> 4),
                          bd0 + static_cast<uint64_t>((dx * Cfg::WTILE_BYTES + k * 32) >> 4), idesc_n, 1);
              }
            } else {
#pragma unroll
              for (int i = 1; i < 6; ++i) {
                const int dx = i >> 1, k = i & 1;
                umma_bf16(dcol, ad0 + static_cast<uint64_t>((dx * 128 + k * 32) >> 4),
                          bd0 + static_cast<uint64_t>((dx * Cfg::WTILE_BYTES + k * 32) >> 4), idesc_n, 1);
              }
            }
            umma_commit(&bar_empty[stage]);                             // stage reusable once these MMAs retire
            if (last_chunk && y >= 1) umma_commit(&bar_rfull[y - 1]);   // output row y-1 is complete
            if (y == TH && !wres) umma_commit(&bar_wempty[wb]);         // weight buffer reusable
          }
          __syncwarp();
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        }
        if (!wres && ++wb == Cfg::NWBUF) {
          wb = 0;
          wphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2 ..)
    // Groups of four warps (one per TMEM lane quarter); group g drains tile rows Y = g, g + NGRP, ...
    constexpr int NGRP = Cfg::NEPI_WARPS / 4;
    const int eg = (warp - 2) >> 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;             // pixel within the 128-wide tile == TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t tile_iter = 0;
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x, ++tile_iter) {
      const int n = t / tiles_per_img;
      const int r = t - n * tiles_per_img;
      const int ty = r / args.xtiles;
      const int tx = r - ty * args.xtiles;
      const int x = tx * 128 + m;
      const int y0 = ty * TH;
      if constexpr (EPI == EPI_ADD_F32) {
        // pull this tile's fp32 addend rows towards L2 while the MMAs run (each warp: 8 groups x 1 KB per row)
        for (int Y = eg; Y < TH; Y += NGRP) {
          const int y = y0 + Y;
          if (y >= args.H) break;
          const size_t wbase = trunk_off(n, y, tx * 128 + q * 32, args.H, args.W) + (lane & 7) * 32;
#pragma unroll
          for (int t2 = 0; t2 < 2; ++t2)
            prefetch_l2(args.fadd + wbase + static_cast<size_t>(t2 * 4 + (lane >> 3)) * TRUNK_GSTRIDE);
        }
      }
      for (int Y = eg; Y < TH; Y += NGRP) {
        mbar_wait(&bar_rfull[Y], tile_iter & 1);
        tc_fence_after();
        float acc[COUT];
        load_acc_row<COUT>(tlane + static_cast<uint32_t>(Y * COUT), acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_rempty[Y]);
        const int y = y0 + Y;
        if (y < args.H && x < args.W && !(args.abl & 1)) epilogue_pixel<COUT, EPI>(args, s_bias, s_prelu, acc, n, y, x);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace b200sr
