// Single-chunk 3x3 convs (Cin = 64: conv_body, conv_up1/up2, conv_hr, conv_last, every SRVGG body conv and its tail)
// -- the product path of every conv outside the fused residual dense blocks.
//
// Same arithmetic and the same MMA order per accumulator as conv3x3_tc_kernel (row streaming, accumulator
// stationary, ky taps stacked on N, kx taps as shifted descriptor views: bit-identical results), restructured around
// what a timing ablation of that kernel showed on B200: with the MMAs, the TMA loads AND the
// epilogue stores all removed it still took 550 cycles per input row -- two barrier waits + two tcgen05.commit per row
// in the MMA warp -- next to 960 cycles of tensor work per row (Cout = 64).  Here
//   * the weights (one chunk) are loaded once per CTA and stay resident; the freed buffer holds more activation rows;
//   * a pipeline stage holds a PAIR of input rows (two TMA boxes, one mbarrier), an accumulator hand-over covers a
//     PAIR of output rows (the two epilogue groups take one row each): per pair the MMA warp waits on two barriers
//     and issues two commits instead of four and four.  Input pair p = rows (2p - 1, 2p) first touches output pair p
//     = rows (2p, 2p + 1) and completes output pair p - 1, so pairs align on both sides.
//   * EPI_LAST9_U8: conv_last (64 -> 3) with the kx taps stacked on N as well: the "output channels" of a ky block
//     are [kx0: 3 + 5 pad | kx1 | kx2 | 8 pad] = 32 columns, ONE unshifted A view per K16 step (4 MMAs of N = 96 per
//     input row instead of 12 of N = 48 at the 44.6-cycle SS-mode floor), and the epilogue adds the three partial sums
//     across lanes: out[x] = P0[x - 1] + P1[x] + P2[x + 1] with P_kx[q] = in[q] . w[., kx] living in TMEM lane q - x0,
//     i.e. lane j + kx for output j (shuffles inside a warp, a 9-float shared-memory exchange at warp boundaries).
//     A tile therefore yields 126 output pixels per 128 lanes (x pitch 126).
#pragma once
#include "conv3x3_tc.cuh"

namespace b200sr {

constexpr int EPI_LAST9_U8 = 7;   // (continues enum EpiMode) stacked-kx conv_last, COUT = 32 virtual channels

template <int COUT, int EPI>
struct ScCfg {
  using Base = ConvCfg<COUT>;
  static constexpr int NDX = (EPI == EPI_LAST9_U8) ? 1 : 3;         // kx tap views issued per K16 step
  static constexpr int XPITCH = (EPI == EPI_LAST9_U8) ? 126 : 128;  // output pixels per column tile
  static constexpr int WBYTES = NDX * Base::WTILE_BYTES;            // resident weight image
  static constexpr int WREGION = (WBYTES + 1023) / 1024 * 1024;
  static constexpr int NPS_FIT = (Base::SMEM_BUDGET - WREGION) / (2 * Base::A_STAGE_BYTES);
  static constexpr int NPS = NPS_FIT > 6 ? 6 : NPS_FIT;             // pair stages
  static_assert(NPS >= 2, "not enough shared memory for two pair stages");
  static constexpr int SMEM_BYTES = WREGION + NPS * 2 * Base::A_STAGE_BYTES + 1024;
  static constexpr int MAXP = Base::MAXTH / 2 + 1;                  // output-row pairs per tile
  static constexpr int NTHREADS = Base::NTHREADS;
  static constexpr int NGRP = Base::NEPI_WARPS / 4;
};

template <int COUT, int EPI>
__global__ void __launch_bounds__(ScCfg<COUT, EPI>::NTHREADS, ConvCfg<COUT>::CTAS_PER_SM)
conv3x3_sc_kernel(const __grid_constant__ CUtensorMap amap, const ConvArgs args) {
  using Cfg = ScCfg<COUT, EPI>;
  using Base = ConvCfg<COUT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + Cfg::WREGION;   // NPS x 2 x A_STAGE_BYTES

  __shared__ uint64_t bar_full[Cfg::NPS], bar_empty[Cfg::NPS];
  __shared__ uint64_t bar_w;
  __shared__ uint64_t bar_rfull[Cfg::MAXP], bar_rempty[Cfg::MAXP];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_bias[COUT];
  __shared__ float s_prelu[COUT];
  __shared__ float s_xchg[2][2][4][12];   // EPI_LAST9_U8: [group][row parity][lane quarter][9 floats]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int TH = args.TH;
  const int npairs_in = (TH + 3) >> 1;    // input rows -1 .. TH in pairs (the last pair may hold one row)
  const int npairs_out = (TH + 1) >> 1;

  if (threadIdx.x < COUT) {
    if constexpr (EPI == EPI_LAST9_U8)
      s_bias[threadIdx.x] = threadIdx.x < 3 ? args.bias[threadIdx.x] : 0.f;
    else
      s_bias[threadIdx.x] = args.bias[threadIdx.x];
    s_prelu[threadIdx.x] = (EPI == EPI_PRELU_BF16) ? args.prelu[threadIdx.x] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::NPS; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(&bar_w, 1);
    for (int i = 0; i < Cfg::MAXP; ++i) {
      mbar_init(&bar_rfull[i], 1);
      mbar_init(&bar_rempty[i], Base::NEPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&amap);
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, Base::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int tiles_per_img = args.xtiles * args.ytiles;
  const uint32_t box_bytes = args.in_up2 ? 132 * 128 : Base::A_BOX_BYTES;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&bar_w, Cfg::WBYTES);
#pragma unroll
      for (int d = 0; d < Cfg::NDX; ++d)
        bulk_load_1d(&bar_w, sW + d * Base::WTILE_BYTES, args.wpack + d * Base::WTILE_BYTES, Base::WTILE_BYTES);
    }
    __syncwarp();
    int ps = 0, phase = 0;
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
      const int n = t / tiles_per_img;
      const int r = t - n * tiles_per_img;
      const int ty = r / args.xtiles;
      const int tx = r - ty * args.xtiles;
      const int x0 = tx * Cfg::XPITCH - 1;
      const int y0 = ty * TH;
      for (int p = 0; p < npairs_in; ++p) {
        mbar_wait(&bar_empty[ps], phase ^ 1);   // rows 2p - 1 and 2p; the second exists while 2p <= TH
        if (args.abl & 4) {   // (timing ablation: no activation loads)
          if (elect_one_sync()) mbar_arrive(&bar_full[ps]);
        } else if (elect_one_sync()) {
          mbar_arrive_expect_tx(&bar_full[ps], (2 * p <= TH ? 2u : 1u) * box_bytes);
          for (int j = 0; j < 2; ++j) {
            const int y = 2 * p - 1 + j;
            if (y > TH) break;
            uint8_t* dst = sA + (ps * 2 + j) * Base::A_STAGE_BYTES;
            if (args.in_up2)   // rows y0+y = -1 and >= H map to source rows -1 and >= H/2: zero-filled
              tma_load_5d(&amap, &bar_full[ps], dst, 0, 0, (x0 + 1) / 2 - 1, (y0 + y) >> 1, n);
            else
              tma_load_4d(&amap, &bar_full[ps], dst, 0, x0, y0 + y, n);
          }
        }
        __syncwarp();
        if (++ps == Cfg::NPS) {
          ps = 0;
          phase ^= 1;
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
    const uint32_t a_skew = args.in_up2 ? 128u : 0u;   // duplicated-pixel boxes start one pixel row earlier
    mbar_wait(&bar_w, 0);
    int ps = 0, phase = 0;
    uint32_t tile_iter = 0;
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x, ++tile_iter) {
      const uint32_t rparity = (tile_iter & 1) ^ 1;   // accumulator pair drained by the previous tile's epilogue
      for (int p = 0; p < npairs_in; ++p) {
        if (p < npairs_out) mbar_wait(&bar_rempty[p], rparity);   // output rows 2p, 2p + 1 are touched for the first time
        mbar_wait(&bar_full[ps], phase);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int y = 2 * p - 1 + j;
            if (y <= TH && !(args.abl & 2)) {   // (abl & 2: timing ablation, no MMAs)
              const int blk_lo = (y < 1) ? (1 - y) : 0;        // output row y-1+blk must be >= 0
              const int blk_hi = (TH - y < 2) ? (TH - y) : 2;  // and < TH
              const int nblk = blk_hi - blk_lo + 1;
              const bool new_row = blk_hi == 2;                 // accumulator row y+1 is touched for the first time
              const uint32_t dcol = tmem_base + static_cast<uint32_t>((y - 1 + blk_lo) * COUT);
              const uint64_t ad0 = adesc0 + static_cast<uint64_t>(((ps * 2 + j) * Base::A_STAGE_BYTES + a_skew) >> 4);
              const uint64_t bd0 = bdesc0 + static_cast<uint64_t>((blk_lo * COUT * 128) >> 4);
              const uint32_t idesc_n = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
              if (new_row) {
                if (nblk > 1) umma_bf16(dcol, ad0, bd0, nblk == 3 ? idesc2 : idesc1, 1);
                umma_bf16(dcol + (nblk - 1) * COUT, ad0, bd0 + static_cast<uint64_t>(((nblk - 1) * COUT * 128) >> 4),
                          idesc1, 0);
              } else {
                umma_bf16(dcol, ad0, bd0, idesc_n, 1);
              }
#pragma unroll
              for (int i = 1; i < 4 * Cfg::NDX; ++i) {
                const int dx = i >> 2, k = i & 3;
                umma_bf16(dcol, ad0 + static_cast<uint64_t>((dx * 128 + k * 32) >> 4),
                          bd0 + static_cast<uint64_t>((dx * Base::WTILE_BYTES + k * 32) >> 4), idesc_n, 1);
              }
            }
          }
          umma_commit(&bar_empty[ps]);                    // both rows of the stage consumed
          if (p >= 1) umma_commit(&bar_rfull[p - 1]);     // output rows 2p-2, 2p-1 are complete
        }
        __syncwarp();
        if (++ps == Cfg::NPS) {
          ps = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2 ..)
    constexpr int NGRP = Cfg::NGRP;
    const int eg = (warp - 2) >> 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;             // pixel within the tile == TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t tile_iter = 0;
    int xpar = 0;   // EPI_LAST9_U8: exchange-buffer parity
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x, ++tile_iter) {
      const int n = t / tiles_per_img;
      const int r = t - n * tiles_per_img;
      const int ty = r / args.xtiles;
      const int tx = r - ty * args.xtiles;
      const int x = tx * Cfg::XPITCH + m;
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
      for (int P = 0; P < npairs_out; ++P) {
        mbar_wait(&bar_rfull[P], tile_iter & 1);
        tc_fence_after();
        if constexpr (NGRP == 2) {
          const int Y = 2 * P + eg;
          float acc[COUT];
          if (Y < TH) load_acc_row<COUT>(tlane + static_cast<uint32_t>(Y * COUT), acc);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_rempty[P]);
          if (Y < TH) {
            const int y = y0 + Y;
            if constexpr (EPI == EPI_LAST9_U8) {
              // out[j] = P0[lane j] + P1[lane j + 1] + P2[lane j + 2], three real channels per kx block
              float* sx = s_xchg[eg][xpar][q];
              if (lane == 0) {
                sx[0] = acc[8]; sx[1] = acc[9]; sx[2] = acc[10];
                sx[3] = acc[16]; sx[4] = acc[17]; sx[5] = acc[18];
              } else if (lane == 1) {
                sx[6] = acc[16]; sx[7] = acc[17]; sx[8] = acc[18];
              }
              asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
              float v[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float p1 = __shfl_down_sync(0xffffffffu, acc[8 + c], 1);
                float p2 = __shfl_down_sync(0xffffffffu, acc[16 + c], 2);
                if (q < 3) {
                  const float* nx = s_xchg[eg][xpar][q + 1];
                  if (lane == 31) {
                    p1 = nx[c];
                    p2 = nx[6 + c];
                  } else if (lane == 30) {
                    p2 = nx[3 + c];
                  }
                }
                v[c] = (acc[c] + p1) + p2 + s_bias[c];
              }
              xpar ^= 1;
              const int cy = y - args.crop_y0, cx = x - args.crop_x0;
              if (m < Cfg::XPITCH && y < args.H && x < args.W && cy >= 0 && cy < args.crop_h && cx >= 0 &&
                  cx < args.crop_w && !(args.abl & 1)) {
                const size_t d = ((static_cast<size_t>(n) * args.dst_h + (args.dst_y0 + cy)) * args.dst_w + (args.dst_x0 + cx)) * 3;
                store_sample(args, d, 0, v[2]);  // B
                store_sample(args, d, 1, v[1]);  // G
                store_sample(args, d, 2, v[0]);  // R
              }
            } else {
              if (y < args.H && x < args.W && !(args.abl & 1)) epilogue_pixel<COUT, EPI>(args, s_bias, s_prelu, acc, n, y, x);
            }
          }
        } else {
          // one epilogue group (2 CTAs / SM build): both rows of the pair, one after the other
          float acc0[COUT], acc1[COUT];
          const int Ya = 2 * P, Yb = 2 * P + 1;
          load_acc_row<COUT>(tlane + static_cast<uint32_t>(Ya * COUT), acc0);
          if (Yb < TH) load_acc_row<COUT>(tlane + static_cast<uint32_t>(Yb * COUT), acc1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_rempty[P]);
          static_assert(NGRP == 2 || EPI != EPI_LAST9_U8, "stacked conv_last needs two epilogue groups");
          if constexpr (EPI != EPI_LAST9_U8) {
            if (y0 + Ya < args.H && x < args.W) epilogue_pixel<COUT, EPI>(args, s_bias, s_prelu, acc0, n, y0 + Ya, x);
            if (Yb < TH && y0 + Yb < args.H && x < args.W) epilogue_pixel<COUT, EPI>(args, s_bias, s_prelu, acc1, n, y0 + Yb, x);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Base::TMEM_COLS);
}

}  // namespace b200sr
