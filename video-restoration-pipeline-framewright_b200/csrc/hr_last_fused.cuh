// conv_hr (64 -> 64, LeakyReLU) and conv_last (64 -> 3, clamp / quantise) of the RRDBNet HR tail as ONE kernel: the
// 4x-resolution tensor between them (128 B per HR pixel: 1.9 GB per 720p frame written and read back) never leaves the
// SM.  Epilogue stores and TMA loads of that size are what the power-capped step pays for;
// these two launches moved 7.6 of the tail's 8.9 GB per frame.
//
// Same arithmetic, same MMA order per accumulator and same fp16 rounding point as conv3x3_sc_kernel<64, EPI_ACT_BF16>
// followed by conv3x3_sc_kernel<32, EPI_LAST9_U8> (identical bytes: tests/test_gpu_parity.py), organised as a ROLLING
// pipeline down a column strip so that no conv_hr row is computed twice:
//   * a unit = (frame, 126-pixel column tile, strip of R output rows).  conv_hr is evaluated on the 128 pixels
//     [126 tx - 1, 126 tx + 126] (one MMA M tile, input box = 130 pixels from 126 tx - 2) and on rows ys-1 .. ye: exactly
//     what the stacked-kx conv_last (conv3x3_sc.cuh) needs for 126 x R outputs; hr pixels / rows outside the image are
//     written as zeros (conv_last's zero padding).
//   * rows travel in PAIRS: input pair P = rows (2P, 2P+1) [unit-relative, first input row = ys - 2] first-touches hr
//     pair P and completes hr pair P-1; hr pair Q first-touches output pair Q and completes output pair Q-1.
//   * TMEM: a ring of 3 hr pair slots (2 rows x 64 columns each, columns 0..383) + a ring of 2 output pair slots
//     (2 rows x 32 columns, columns 384..511).  The ky-stacked MMA writes 3 consecutive rows; where those straddle the
//     ring's wrap it is issued as two MMAs.
//   * warps: 0 TMA producer, 1 MMA issuer, 2-9 epilogue in two groups of four, group g owning row g of every pair in
//     both roles: "A" (hr accumulators -> bias, LeakyReLU, fp16 -> the swizzled shared-memory tile conv_last's MMAs read
//     as their A operand) and "B" (output accumulators -> cross-lane sum of the three kx partials -> clamp / round ->
//     u8 BGR).
//   * per iteration the MMA warp issues conv_hr for input pair P, then conv_last for hr pair P-2, whose tile epilogue A
//     wrote while the conv_hr MMAs of pair P were running.
#pragma once
#include "conv3x3_sc.cuh"

namespace b200sr {

struct HrLastArgs {
  int N, H, W;              // HR extent (conv_hr input == output == conv_last output)
  int xtiles;               // ceil(W / 126)
  int strips;               // row strips per column tile
  int strip_rows;           // rows per strip (even)
  int nunits;               // N * xtiles * strips
  const uint8_t* w_hr;      // conv_hr packed weights (pack_weights: 3 dx tiles x [3 blk x 64 rows] x 128 B), fp16
  const uint8_t* w_last;    // conv_last stacked-kx image (pack_weights_last9: [3 blk x 32 rows] x 128 B), fp16
  const float* bias_hr;     // 64
  const float* bias_last;   // 3
  float slope;              // LeakyReLU slope of conv_hr
  uint8_t* dst;             // BGR destination frame(s)
  int dst16;
  int dst_h, dst_w, crop_y0, crop_x0, crop_h, crop_w, dst_y0, dst_x0;
};

constexpr int HL_W_HR_BYTES = 3 * 3 * 64 * 128;       // 73,728
constexpr int HL_W_LAST_BYTES = 3 * 32 * 128;         // 12,288
constexpr int HL_ROW_BYTES = 17 * 1024;               // one input row stage (130 x 128 B box)
constexpr int HL_NPS = 3;                             // input pair stages
constexpr int HL_HTILE_BYTES = 128 * 128;             // one hr row as conv_last's A tile
constexpr int HL_OFF_WLAST = HL_W_HR_BYTES;
constexpr int HL_OFF_IN = HL_W_HR_BYTES + HL_W_LAST_BYTES;                 // 86,016 = 84 KB
constexpr int HL_OFF_H = HL_OFF_IN + HL_NPS * 2 * HL_ROW_BYTES;            // 190,464
constexpr int HL_SMEM_BYTES = HL_OFF_H + 2 * HL_HTILE_BYTES + 1024;        // 224,256
constexpr int HL_NTHREADS = 32 * 10;
constexpr uint32_t HL_LAST_COL0 = 384;                // first TMEM column of the output-pair ring

// Issues the MMAs of one A row whose ky-blocks 0..2 accumulate into the TMEM columns col[0..2] (blocks outside
// [blo, bhi] are skipped).  Block 2, if present, is touched for the first time: its first MMA overwrites (accumulate
// = 0).  Neighbouring blocks whose column ranges are consecutive share one MMA (N = nblk * COUT); at the wrap of a
// TMEM ring (at most one break among the three blocks) the row is issued as two MMAs per K step.  Straight-line code
// per case: the issuing thread must stay ahead of a tensor pipe that retires an N = 192 MMA every 96 cycles (a rolled
// loop with per-step predicates measured 2.3x slower than the MMAs themselves).
template <int COUT, int NDX>
__device__ __forceinline__ void hl_issue_row(uint64_t adesc, uint64_t bdesc, uint32_t wtile_bytes, const uint32_t (&col)[3],
                                             int blo, int bhi, bool f16) {
  constexpr uint64_t BLK = static_cast<uint64_t>((COUT * 128) >> 4);   // descriptor offset of one ky block of B
  const uint32_t id1 = make_idesc_16(128, COUT, f16), id2 = make_idesc_16(128, 2 * COUT, f16),
                 id3 = make_idesc_16(128, 3 * COUT, f16);
  // runs for the plain K steps
  int brk = -1;   // (selects instead of indexed reads: col[] stays in registers)
  if (blo <= 0 && bhi >= 1 && col[1] != col[0] + COUT) brk = 0;
  if (blo <= 1 && bhi >= 2 && col[2] != col[1] + COUT) brk = 1;
  const bool two = brk >= 0;
  const int n0 = two ? brk - blo + 1 : bhi - blo + 1;
  const uint32_t d0 = blo == 0 ? col[0] : (blo == 1 ? col[1] : col[2]);
  const uint64_t o0 = static_cast<uint64_t>(blo) * BLK;
  const uint32_t i0 = n0 == 3 ? id3 : (n0 == 2 ? id2 : id1);
  const int s1 = two ? brk + 1 : 2;
  const uint32_t d1 = s1 == 1 ? col[1] : col[2];
  const uint64_t o1 = static_cast<uint64_t>(s1) * BLK;
  const uint32_t i1 = (bhi - s1 + 1) == 2 ? id2 : id1;
  // first K step
  if (bhi == 2) {   // block 2 is new
    if (blo == 0) {
      if (col[1] == col[0] + COUT) {
        umma_bf16(col[0], adesc, bdesc, id2, 1);
      } else {
        umma_bf16(col[0], adesc, bdesc, id1, 1);
        umma_bf16(col[1], adesc, bdesc + BLK, id1, 1);
      }
    } else if (blo == 1) {
      umma_bf16(col[1], adesc, bdesc + BLK, id1, 1);
    }
    umma_bf16(col[2], adesc, bdesc + 2 * BLK, id1, 0);
  } else {
    umma_bf16(d0, adesc, bdesc + o0, i0, 1);
    if (two) umma_bf16(d1, adesc, bdesc + o1, i1, 1);
  }
  if (!two) {
#pragma unroll
    for (int i = 1; i < 4 * NDX; ++i) {
      const int dx = i >> 2, k = i & 3;
      umma_bf16(d0, adesc + static_cast<uint64_t>((dx * 128 + k * 32) >> 4),
                bdesc + o0 + static_cast<uint64_t>((dx * wtile_bytes + k * 32) >> 4), i0, 1);
    }
  } else {
#pragma unroll
    for (int i = 1; i < 4 * NDX; ++i) {
      const int dx = i >> 2, k = i & 3;
      const uint64_t a = adesc + static_cast<uint64_t>((dx * 128 + k * 32) >> 4);
      const uint64_t bq = bdesc + static_cast<uint64_t>((dx * wtile_bytes + k * 32) >> 4);
      umma_bf16(d0, a, bq + o0, i0, 1);
      umma_bf16(d1, a, bq + o1, i1, 1);
    }
  }
}

__global__ void __launch_bounds__(HL_NTHREADS, 1)
hr_last_fused_kernel(const __grid_constant__ CUtensorMap amap, const HrLastArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sWh = smem;
  uint8_t* sWl = smem + HL_OFF_WLAST;
  uint8_t* sIn = smem + HL_OFF_IN;
  uint8_t* sH = smem + HL_OFF_H;

  __shared__ uint64_t bar_w;
  __shared__ uint64_t in_full[HL_NPS], in_empty[HL_NPS];
  __shared__ uint64_t hr_full[3], hr_empty[3];     // hr pair slots (TMEM)
  __shared__ uint64_t h_ready, h_empty;            // the shared-memory hr tile (one pair)
  __shared__ uint64_t l_full[2], l_empty[2];       // output pair slots (TMEM)
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_bias_hr[64];
  __shared__ float s_bias_last[4];
  __shared__ float s_xchg[2][2][4][12];   // [group][parity][lane quarter][9 floats]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x < 64) s_bias_hr[threadIdx.x] = args.bias_hr[threadIdx.x];
  if (threadIdx.x < 3) s_bias_last[threadIdx.x] = args.bias_last[threadIdx.x];
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < HL_NPS; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&hr_full[i], 1);
      mbar_init(&hr_empty[i], 8);
    }
    mbar_init(&h_ready, 8);
    mbar_init(&h_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&l_full[i], 1);
      mbar_init(&l_empty[i], 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&amap);
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int units_per_img = args.xtiles * args.strips;
  // unit u -> (n, tx, strip): strips of one column tile are consecutive units' neighbours in x, so that the CTAs of a
  // wave read neighbouring boxes of the same rows
  auto unit_geom = [&](int u, int& n, int& tx, int& ys, int& R) {
    n = u / units_per_img;
    const int r = u - n * units_per_img;
    const int st = r / args.xtiles;
    tx = r - st * args.xtiles;
    ys = st * args.strip_rows;
    const int ye = (ys + args.strip_rows < args.H) ? ys + args.strip_rows : args.H;
    R = ye - ys;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&bar_w, HL_W_HR_BYTES + HL_W_LAST_BYTES);
#pragma unroll
      for (int d = 0; d < 3; ++d) bulk_load_1d(&bar_w, sWh + d * 24576, args.w_hr + d * 24576, 24576);
      bulk_load_1d(&bar_w, sWl, args.w_last, HL_W_LAST_BYTES);
    }
    __syncwarp();
    uint32_t gp = 0;   // input pairs so far
    for (int u = blockIdx.x; u < args.nunits; u += gridDim.x) {
      int n, tx, ys, R;
      unit_geom(u, n, tx, ys, R);
      const int x0 = tx * 126 - 2;
      const int npin = (R >> 1) + 2;
      for (int P = 0; P < npin; ++P, ++gp) {
        const int st = gp % HL_NPS;
        mbar_wait(&in_empty[st], ((gp / HL_NPS) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&in_full[st], 2 * 130 * 128);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(&amap, &in_full[st], sIn + (st * 2 + j) * HL_ROW_BYTES, 0, x0, ys - 2 + 2 * P + j, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint64_t adesc_in = make_smem_desc(smem_u32(sIn), 1024, SWZ_128B, 0);
    const uint64_t adesc_h = make_smem_desc(smem_u32(sH), 1024, SWZ_128B, 0);
    const uint64_t bdesc_hr = make_smem_desc(smem_u32(sWh), 1024, SWZ_128B, 0);
    const uint64_t bdesc_last = make_smem_desc(smem_u32(sWl), 1024, SWZ_128B, 0);
    mbar_wait(&bar_w, 0);
    uint32_t gp = 0, gq = 0, gj = 0, gql = 0;   // input pairs, hr pairs (at unit start), output pairs (at unit start), hr tiles consumed
    for (int u = blockIdx.x; u < args.nunits; u += gridDim.x) {
      int n, tx, ys, R;
      unit_geom(u, n, tx, ys, R);
      const int nq = (R >> 1) + 1;     // hr pairs of this unit
      const int nj = R >> 1;           // output pairs
      const int npin = nq + 1;         // input pairs
      // conv_last of hr pair Q (its tile is in shared memory once epilogue A has arrived on h_ready)
      auto do_last = [&](int Q) {
        mbar_wait(&h_ready, gql & 1);
        if (Q < nj) {   // output pair Q is touched for the first time
          const uint32_t g = gj + Q;
          mbar_wait(&l_empty[g & 1], ((g >> 1) & 1) ^ 1);
        }
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int uu = 2 * Q + j;                 // hr row (unit-relative); feeds output rows uu-2, uu-1, uu
            uint32_t col[3];
            int blo = 3, bhi = -1;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const int v = uu - 2 + b;
              col[b] = 0;
              if (v >= 0 && v < R) {
                col[b] = tmem_base + HL_LAST_COL0 + ((gj + (v >> 1)) & 1) * 64 + (v & 1) * 32;
                if (b < blo) blo = b;
                bhi = b;
              }
            }
            if (bhi >= blo)
              hl_issue_row<32, 1>(adesc_h + static_cast<uint64_t>((j * HL_HTILE_BYTES) >> 4), bdesc_last, 0, col, blo, bhi,
                                  /*f16=*/true);
          }
          umma_commit(&h_empty);
          if (Q >= 1) umma_commit(&l_full[(gj + Q - 1) & 1]);
        }
        __syncwarp();
        ++gql;
      };
      for (int P = 0; P < npin; ++P, ++gp) {
        if (P < nq) {   // hr pair P is touched for the first time
          const uint32_t g = gq + P;
          mbar_wait(&hr_empty[g % 3], ((g / 3) & 1) ^ 1);
        }
        const int st = gp % HL_NPS;
        mbar_wait(&in_full[st], (gp / HL_NPS) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int t = 2 * P + j;                  // input row (unit-relative); feeds hr rows t-2, t-1, t
            uint32_t col[3];
            int blo = 3, bhi = -1;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const int uu = t - 2 + b;
              col[b] = 0;
              if (uu >= 0 && uu < 2 * nq) {
                col[b] = tmem_base + ((gq + (uu >> 1)) % 3) * 128 + (uu & 1) * 64;
                if (b < blo) blo = b;
                bhi = b;
              }
            }
            if (bhi >= blo)
              hl_issue_row<64, 3>(adesc_in + static_cast<uint64_t>(((st * 2 + j) * HL_ROW_BYTES) >> 4), bdesc_hr, 24576, col,
                                  blo, bhi, /*f16=*/true);
          }
          umma_commit(&in_empty[st]);
          if (P >= 1) umma_commit(&hr_full[(gq + P - 1) % 3]);
        }
        __syncwarp();
        if (P >= 2) do_last(P - 2);
      }
      do_last(nq - 1);
      gq += nq;
      gj += nj;
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9
    // Two groups of four warps (one per TMEM lane quarter); group g owns row g of every pair, in both roles:
    //   A(Q): hr accumulator row 2Q+g -> bias, LeakyReLU, fp16 -> row g of the shared-memory tile (conv_last's A operand)
    //   B(J): output accumulator row 2J+g -> cross-lane sum of the three kx partials -> clamp / round -> u8 BGR
    // in the fixed order A(0) A(1) [A(Q) B(Q-2)]... B(nj-1): B(J) needs conv_last of hr pair J+1, i.e. A(J+1) of all
    // eight warps, which every warp has already done when it reaches B(J).
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;             // pixel within the tile == TMEM lane == A-tile row
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t gq = 0, gj = 0;                 // hr pairs / output pairs so far (all units)
    int xpar = 0;
    for (int u = blockIdx.x; u < args.nunits; u += gridDim.x) {
      int n, tx, ys, R;
      unit_geom(u, n, tx, ys, R);
      const int nq = (R >> 1) + 1, nj = R >> 1;
      const int xh = tx * 126 - 1 + m;       // hr pixel of this lane
      const bool px_ok = xh >= 0 && xh < args.W;
      const int x = tx * 126 + m;            // output pixel of this lane (m < 126)
      auto role_b = [&](int J) {
        const uint32_t gg = gj + J;
        const int slot = gg & 1;
        mbar_wait(&l_full[slot], (gg >> 1) & 1);
        tc_fence_after();
        float acc[32];
        load_acc_row<32>(tlane + HL_LAST_COL0 + static_cast<uint32_t>(slot * 64 + g * 32), acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&l_empty[slot]);
        // out[x] = P0[lane m] + P1[lane m + 1] + P2[lane m + 2] (conv3x3_sc.cuh, EPI_LAST9_U8: same order of additions)
        float* sx = s_xchg[g][xpar][q];
        if (lane == 0) {
          sx[0] = acc[8]; sx[1] = acc[9]; sx[2] = acc[10];
          sx[3] = acc[16]; sx[4] = acc[17]; sx[5] = acc[18];
        } else if (lane == 1) {
          sx[6] = acc[16]; sx[7] = acc[17]; sx[8] = acc[18];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float p1 = __shfl_down_sync(0xffffffffu, acc[8 + c], 1);
          float p2 = __shfl_down_sync(0xffffffffu, acc[16 + c], 2);
          if (q < 3) {
            const float* nx = s_xchg[g][xpar][q + 1];
            if (lane == 31) {
              p1 = nx[c];
              p2 = nx[6 + c];
            } else if (lane == 30) {
              p2 = nx[3 + c];
            }
          }
          v[c] = (acc[c] + p1) + p2 + s_bias_last[c];
        }
        xpar ^= 1;
        const int y = ys + 2 * J + g;
        const int cy = y - args.crop_y0, cx = x - args.crop_x0;
        if (m < 126 && y < args.H && x < args.W && cy >= 0 && cy < args.crop_h && cx >= 0 && cx < args.crop_w) {
          const size_t d = ((static_cast<size_t>(n) * args.dst_h + (args.dst_y0 + cy)) * args.dst_w + (args.dst_x0 + cx)) * 3;
          if (args.dst16) {
            uint16_t* o = reinterpret_cast<uint16_t*>(args.dst);
            o[d + 0] = quant_u16(v[2]);
            o[d + 1] = quant_u16(v[1]);
            o[d + 2] = quant_u16(v[0]);
          } else {
            args.dst[d + 0] = quant_u8(v[2]);   // B
            args.dst[d + 1] = quant_u8(v[1]);   // G
            args.dst[d + 2] = quant_u8(v[0]);   // R
          }
        }
      };
      for (int Q = 0; Q < nq; ++Q) {
        const uint32_t gg = gq + Q;
        const int slot = gg % 3;
        mbar_wait(&hr_full[slot], (gg / 3) & 1);
        tc_fence_after();
        uint32_t pk[32];
        {
          float acc[64];
          load_acc_row<64>(tlane + static_cast<uint32_t>(slot * 128 + g * 64), acc);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&hr_empty[slot]);
          const int h = ys - 1 + 2 * Q + g;
          const bool ok = px_ok && h >= 0 && h < args.H;   // outside the image: conv_last's zero padding
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            float v0 = acc[c] + s_bias_hr[c], v1 = acc[c + 1] + s_bias_hr[c + 1];
            v0 = v0 > 0.f ? v0 : v0 * args.slope;
            v1 = v1 > 0.f ? v1 : v1 * args.slope;
            pk[c >> 1] = ok ? pack_f16x2(v0, v1) : 0u;
          }
        }
        mbar_wait(&h_empty, (gg & 1) ^ 1);   // conv_last of the previous pair has consumed the tile
        uint8_t* row = sH + g * HL_HTILE_BYTES + m * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)   // 16-byte chunk c of this pixel, SWIZZLE_128B: chunk ^ (row & 7)
          *reinterpret_cast<uint4*>(row + ((c ^ (m & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready);
        if (Q >= 2) role_b(Q - 2);
      }
      role_b(nj - 1);
      gq += nq;
      gj += nj;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace b200sr
