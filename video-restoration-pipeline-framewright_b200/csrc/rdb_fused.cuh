// One persistent kernel per Residual Dense Block: the five 3x3 convs of an RDB
// (basicsr ResidualDenseBlock.forward; constructed through RRDBNet at
// /root/reference/src/framewright/processors/pytorch_realesrgan.py:107-127) run as ONE launch so that the
// dense-block intermediates x1..x4 are produced and consumed while still resident in the 126 MB L2.
//
// Why: un-fused, every conv of an RDB re-reads the whole dense buffer from HBM; at 720p that is 1.8 KB of
// DRAM traffic per pixel per RDB and puts all five convs below the roofline ridge (DESIGN.md sections 4.4, 6).
//
// How: the work is cut into items (conv k, frame, 128-pixel column tx, 16 rows; 8 rows for conv5).  Strip s of
// conv k covers rows [16 s - 8 k, 16 s + 16 - 8 k): every conv is shifted UP by 8 rows against its predecessor,
// so everything an item reads from a lower conv (its own rows +- 1) is produced by items of earlier strips.  The
// host orders the list by step t, conv k working on strip t - RDB_STEP_OFF[k] (b200sr.cu::build_rdb_items).
// Items are CLAIMED dynamically in list order (one atomicAdd by the producer warp when it is ready to start the
// item), so an item never starts before a lower-numbered one; a CTA runs its items with the same TMA ->
// tcgen05 -> epilogue pipeline as conv3x3_tc_kernel (per-item layer parameters).  Cross-CTA dependencies are
// completion counters per (frame, conv, block of RDB_FLAG_ROWS rows) in global memory: an epilogue warp does
// __syncwarp + red.release.gpu.add once its rows of a block are stored; the TMA producer of a consuming item
// polls the counters of the blocks a dependent channel chunk reads with relaxed loads -- lazily, when the row
// loop reaches a block -- and then executes fence.acq_rel.gpu + fence.proxy.async.global before the TMA loads
// that depend on it.  Dependencies always point to lower-numbered, i.e. already claimed and running, items, so
// the schedule cannot deadlock, whatever the number of resident CTAs.
#pragma once
#include "conv3x3_tc.cuh"

// timing ablations (tools only; results are wrong when set): skip the epilogue math + stores / the dependency waits
#ifndef B200SR_ABL_NOEPI
#define B200SR_ABL_NOEPI 0
#endif
#ifndef B200SR_ABL_NODEP
#define B200SR_ABL_NODEP 0
#endif

// -DB200SR_DEBUG: bounds traps on everything the cross-CTA protocol indexes with (compute-sanitizer is closed on this
// pool, so the checks live in the kernel): claimed item fields, completion-counter indices, TMA coordinates.  A
// violation prints the site and traps (the launch fails with an error instead of corrupting memory silently).
#ifdef B200SR_DEBUG
#define RDB_ASSERT(cond, what, a, b)                                                                        \
  do {                                                                                                      \
    if (!(cond)) {                                                                                          \
      printf("B200SR_DEBUG trap: %s (%d, %d) block %d thread %d\n", what, static_cast<int>(a),              \
             static_cast<int>(b), static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));             \
      __trap();                                                                                             \
    }                                                                                                       \
  } while (0)
#else
#define RDB_ASSERT(cond, what, a, b) \
  do {                               \
  } while (0)
#endif

namespace b200sr {

struct RdbItem {      // 32 bytes, built on the host (b200sr.cu::build_rdb_items)
  int k;              // conv index 0..4 (conv1..conv5)
  int n;              // frame
  int y0;             // first output row
  int rows;           // output rows in this item (<= 16 for k < 4, <= 8 for k == 4)
  int tx;             // 128-pixel column
  int flag_base;      // index of the completion counter of block 0 of (n, k); -1 for conv5
  int dep_base[2];    // per dependent chunk (c = 1, 2): counter base of the conv whose output that chunk needs
                      // (-1 = none); block b's counter is dep_base + b
};

struct RdbArgs {
  ConvArgs L[5];          // per-conv parameters (wpack, bias, nchunks, last_ksteps, out/out_choff, residual pair, ...)
  const RdbItem* items;
  int nitems;
  int* flags;             // [frame][conv1..4][block], zeroed before the launch
  int* counter;           // next unclaimed item (zeroed before the launch): items are claimed in list order
  int nflags;             // number of completion counters (bounds of every flag index; checked in debug builds)
  int flag_target;        // counter value of a complete block: column tiles x epilogue warps
  int half64;             // the 32-channel last chunk of conv2 / conv4 is loaded as a 32-channel TMA box (64-byte rows,
                          // SWIZZLE_64B, tensor map `amap_h`) with a matching weight image, instead of over-reading a
                          // 64-channel box of which only half feeds the MMAs
  int rrdb_end;           // conv5 epilogue also applies the RRDB-level skip
  long long* stats;       // optional [grid][16] cycle counters (dev tool; nullptr = off)
  long long* trace;       // optional [nitems][10] globaltimer stamps per item (dev tool; nullptr = off)
  long long* trace2;      // optional [nitems][16][3] per-row stamps: MMA commit, epilogue wait passed, row stored
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic -> async proxy ordering for GLOBAL memory only (the unrestricted form also covers shared memory)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Dev instrumentation (per-role wait / issue cycle counters, tools/rdb_stats.py): compiled in only with
// -DB200SR_RDB_STATS.  It costs registers and local-memory traffic, so product builds leave it out.
#ifdef B200SR_RDB_STATS
#define RDB_STATS_ON 1
#define RDB_TIMED(slot, ...)                           \
  do {                                                 \
    if (st_on) {                                       \
      const long long t__ = clock64();                 \
      __VA_ARGS__;                                     \
      st_acc[slot] += clock64() - t__;                 \
    } else {                                           \
      __VA_ARGS__;                                     \
    }                                                  \
  } while (0)
#define RDB_COUNT(slot, v) st_acc[slot] += (v)
#define RDB_STAMP(it, slot)                                                   \
  do {                                                                        \
    if (args.trace != nullptr && (threadIdx.x & 31) == 0) {                   \
      unsigned long long t__;                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                 \
      args.trace[static_cast<size_t>(it) * 10 + (slot)] = (long long)t__;    \
    }                                                                         \
  } while (0)
#define RDB_STAMP2(it, row, slot)                                                        \
  do {                                                                                   \
    if (args.trace2 != nullptr) {                                                        \
      unsigned long long t__;                                                            \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                            \
      args.trace2[(static_cast<size_t>(it) * 16 + (row)) * 3 + (slot)] = (long long)t__; \
    }                                                                                    \
  } while (0)
#else
#define RDB_STAMP2(it, row, slot) \
  do {                            \
  } while (0)
#define RDB_STAMP(it, slot) \
  do {                      \
  } while (0)
#define RDB_STATS_ON 0
#define RDB_TIMED(slot, ...) \
  do {                       \
    __VA_ARGS__;             \
  } while (0)
#define RDB_COUNT(slot, v) \
  do {                     \
  } while (0)
#endif

// EXPERIMENT: dense-block intermediates (planes 1, 2) addressed modulo a ring of this many rows, so that dead rows are
// overwritten while still in L2 instead of being written back to DRAM (0 = off: full-frame planes)
#ifndef B200SR_RDB_RING
#define B200SR_RDB_RING 0
#endif
constexpr int RDB_RING = B200SR_RDB_RING;
// EXPERIMENT: conv5's epilogue discards (discard.global.L2: drop without write-back) the dense-block intermediates that
// only its own item read -- rows y0+1 .. y0+rows-2, tile-interior pixels -- once all MMAs of the item have completed
#ifndef B200SR_RDB_DISCARD
#define B200SR_RDB_DISCARD 0
#endif
#ifndef B200SR_RDB_QD
#define B200SR_RDB_QD 2
#endif
#ifndef B200SR_RDB_POLL_NS
#define B200SR_RDB_POLL_NS 64
#endif
constexpr int RDB_QD = B200SR_RDB_QD;
// CTAs per SM.  1 (default): one persistent CTA per SM with all 512 TMEM columns -- items of 16 rows (conv5: 8), two
// weight-chunk buffers, four activation stages, eight epilogue warps.  2: two co-resident CTAs with 256 TMEM columns
// each -- items of 8 rows (conv5: 4), ONE weight buffer, two activation stages, four epilogue warps -- so that the
// issue bubbles of one CTA (barrier waits between bursts, weight reloads, dependency waits) are filled by the other
// CTA's MMAs on the shared tensor pipe, and the L2 working set of the skewed schedule halves.
#ifndef B200SR_RDB_CTAS
#define B200SR_RDB_CTAS 1
#endif
constexpr int RDB_CTAS = B200SR_RDB_CTAS;
static_assert(RDB_CTAS == 1 || RDB_CTAS == 2, "B200SR_RDB_CTAS must be 1 or 2");
// EXPERIMENT: -DB200SR_RDB_TMEM_COLS=256 with one CTA per SM = items of 8 rows (conv5: 4): half the bytes per step of the
// skewed schedule, i.e. half the L2 working window, for more halo rows and weight reloads per output row
#ifndef B200SR_RDB_TMEM_COLS
#define B200SR_RDB_TMEM_COLS (512 / B200SR_RDB_CTAS)
#endif
constexpr int RDB_TMEM_COLS = B200SR_RDB_TMEM_COLS;
constexpr int RDB_TH4 = RDB_TMEM_COLS / 32;        // output rows of a conv1..4 item (one 32-column slot per row)
constexpr int RDB_TH5 = RDB_TMEM_COLS / 64;        // output rows of a conv5 item
constexpr int RDB_STRIP_ROWS = RDB_TH4;            // rows per strip of the skewed schedule
constexpr int RDB_SHIFT_ROWS = RDB_TH4 / 2;        // conv k is shifted up by k * RDB_SHIFT_ROWS rows
#ifndef B200SR_RDB_FLAG_SHIFT
#define B200SR_RDB_FLAG_SHIFT (B200SR_RDB_TMEM_COLS == 512 ? 3 : 2)
#endif
constexpr int RDB_FLAG_SHIFT = B200SR_RDB_FLAG_SHIFT;   // completion counters per 2^shift output rows of a conv
constexpr int RDB_FLAG_ROWS = 1 << RDB_FLAG_SHIFT;
static_assert(RDB_FLAG_ROWS <= RDB_SHIFT_ROWS, "item rows must start on a completion-counter block boundary");
constexpr int RDB_MAX_DEP_BLOCKS = (RDB_TH4 + 2) / RDB_FLAG_ROWS + 2;
constexpr int RDB_NSTAGES = RDB_CTAS == 1 ? 4 : 2;
constexpr int RDB_NWBUF = RDB_CTAS == 1 ? 2 : 1;
constexpr int RDB_WBUF_BYTES = 9 * 64 * 128;                 // weight chunk buffer sized for Cout = 64
constexpr int RDB_A_STAGE_BYTES = 17 * 1024;
constexpr int RDB_SMEM_BYTES = RDB_NWBUF * RDB_WBUF_BYTES + RDB_NSTAGES * RDB_A_STAGE_BYTES + 1024;
constexpr int RDB_NEPI_WARPS = RDB_CTAS == 1 ? 8 : 4;
constexpr int RDB_NGRP = RDB_NEPI_WARPS / 4;                 // epilogue groups (one warp per TMEM lane quarter each)
constexpr int RDB_NTHREADS = 32 * (2 + RDB_NEPI_WARPS);

// MODE (compile time; the launch picks the instantiation, b200sr.cu::launch_rdb_fused) describes conv5's epilogue:
// bit 0 = RRDB end (args.rrdb_end), bit 1 = the RDB input carries a lo part (L[4].lo_in != nullptr), bit 2 = the output
// does (L[4].lo_out != nullptr).  With the default residual format two of the three launches of an RRDB are MODE 0 --
// hi in, hi out: no e5m2 decode, no rounding residual, 48 fewer live registers -- and the third is MODE 5.
constexpr int RDB_MODE_RRDB_END = 1, RDB_MODE_LO_IN = 2, RDB_MODE_LO_OUT = 4;
template <int MODE>
__global__ void __launch_bounds__(RDB_NTHREADS, RDB_CTAS)
rdb_fused_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap amap_h,
                 const __grid_constant__ RdbArgs args) {
  constexpr bool RRDB_END = (MODE & RDB_MODE_RRDB_END) != 0, LO_IN = (MODE & RDB_MODE_LO_IN) != 0,
                 LO_OUT = (MODE & RDB_MODE_LO_OUT) != 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + RDB_NWBUF * RDB_WBUF_BYTES;

  __shared__ uint64_t bar_full[RDB_NSTAGES], bar_empty[RDB_NSTAGES];
  __shared__ uint64_t bar_wfull[RDB_NWBUF], bar_wempty[RDB_NWBUF];
  __shared__ uint64_t bar_rfull[16], bar_rempty[16];   // one per 32-column TMEM slot
  __shared__ uint64_t bar_qfull[RDB_QD], bar_qempty[RDB_QD];   // claimed-item queue: producer -> MMA + epilogue warps
  __shared__ int s_q[RDB_QD];
  __shared__ RdbItem s_qitem[RDB_QD];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_bias[5][64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 5 * 64; i += blockDim.x) {
    const int k = i >> 6, c = i & 63;
    s_bias[k][c] = (k < 4 && c >= 32) ? 0.f : args.L[k].bias[c];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < RDB_NSTAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < RDB_NWBUF; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], 1);
    }
    for (int i = 0; i < 16; ++i) {
      mbar_init(&bar_rfull[i], 1);
      mbar_init(&bar_rempty[i], 4);
    }
    for (int i = 0; i < RDB_QD; ++i) {
      mbar_init(&bar_qfull[i], 1);
      mbar_init(&bar_qempty[i], 1 + RDB_NEPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&amap_h);
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem_base, RDB_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const bool st_on = RDB_STATS_ON && args.stats != nullptr;
  long long st_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long st_t0 = st_on ? clock64() : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Items are claimed dynamically, in list order (atomic counter): an item never starts before a lower-numbered
    // one, so a producer item that precedes its consumer by >= one "round" of CTAs has already finished.
    int stage = 0, phase = 0, wb = 0, wphase = 0, qi = 0, qphase = 0;
    const uint64_t pol_first = make_l2_policy_evict_first(), pol_normal = make_l2_policy_evict_normal();
    const uint64_t pol_last = make_l2_policy_evict_last();
    // (Claiming an item ahead of time was measured to be much worse: a claimed-but-not-started item delays all its
    // consumers.  Items are claimed exactly when the producer warp is ready to start them.)
    while (true) {
      mbar_wait(&bar_qempty[qi], qphase ^ 1);
      if (lane == 0) {
        const int claimed = atomicAdd(args.counter, 1);
        if (claimed < args.nitems) {
          s_qitem[qi] = args.items[claimed];
          s_q[qi] = claimed;
        } else {
          s_q[qi] = -1;
        }
        mbar_arrive(&bar_qfull[qi]);
      }
      __syncwarp();
      const int it = s_q[qi];
      const RdbItem item = s_qitem[qi];
      if (++qi == RDB_QD) {
        qi = 0;
        qphase ^= 1;
      }
      if (it < 0) break;
      RDB_ASSERT(item.k >= 0 && item.k < 5, "item.k", item.k, it);
      RDB_ASSERT(item.rows >= 1 && item.rows <= (item.k < 4 ? RDB_TH4 : RDB_TH5), "item.rows", item.rows, item.k);
      RDB_ASSERT(item.y0 >= 0 && item.y0 + item.rows <= args.L[item.k].H, "item.y0 + rows", item.y0, item.rows);
      RDB_ASSERT(item.n >= 0 && item.n < args.L[item.k].N, "item.n", item.n, args.L[item.k].N);
      RDB_ASSERT(item.tx >= 0 && item.tx * 128 < args.L[item.k].W, "item.tx", item.tx, args.L[item.k].W);
      RDB_ASSERT(item.k == 4 ? item.flag_base == -1 : (item.flag_base >= 0 && item.flag_base < args.nflags),
                 "item.flag_base", item.flag_base, args.nflags);
      RDB_STAMP(it, 0);
#ifdef B200SR_RDB_STATS
      if (args.trace != nullptr && lane == 0) args.trace[static_cast<size_t>(it) * 10 + 8] = blockIdx.x;
#endif
      const ConvArgs& L = args.L[item.k];
      const int cout = item.k < 4 ? 32 : 64;
      const uint32_t wtile = 3u * cout * 128u;
      const int x0 = item.tx * 128 - 1;
      // conv5 is the last reader of the dense block's rows in this launch: let them leave L2 first
#ifdef B200SR_ABL_NOHINT
      const uint64_t pol_item = pol_normal;
#else
#if defined(B200SR_RDB_INTER_LAST)   // EXPERIMENT: keep the intermediates' rows in L2 (evict_last) until conv5 has read them
      const uint64_t pol_item = item.k == 4 ? pol_first : pol_last;
#else
      const uint64_t pol_item = item.k == 4 ? pol_first : pol_normal;
#endif
#endif
      for (int c = 0; c < L.nchunks; ++c) {
        RDB_TIMED(2, mbar_wait(&bar_wempty[wb], wphase ^ 1));
        // half chunk (conv2 / conv4, last chunk = 32 channels): 64-byte rows everywhere (weights, activation box)
        const bool half = args.half64 && c == L.nchunks - 1 && L.last_ksteps == 2;
        const uint32_t wt = half ? wtile / 2 : wtile;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&bar_wfull[wb], 3 * wt);
          const uint8_t* wsrc = L.wpack + static_cast<size_t>(c) * 3 * wtile;
#pragma unroll
          for (int d = 0; d < 3; ++d)
            bulk_load_1d(&bar_wfull[wb], sW + wb * RDB_WBUF_BYTES + d * wt, wsrc + d * wt, wt);
        }
        __syncwarp();
        // The blocks of the predecessor conv this chunk reads (rows y0-1 .. y0+rows, at most 4 blocks) must be
        // complete -- checked LAZILY, block by block, as the row loop reaches them: the lower blocks were produced
        // two steps ago and are ready, the upper ones belong to an item that may still be running, and its last
        // block is needed only for this chunk's final (halo) row.  First one batch of relaxed loads for the blocks
        // that are already complete, then ONE acquire fence + one generic->async proxy fence per wait before the
        // TMA loads that depend on it (the weights, which do not depend on other CTAs, are already in flight).
        const int dep = (c > 0 && !B200SR_ABL_NODEP) ? item.dep_base[c - 1] : -1;
        int ready_blk = 1 << 30;   // highest block of the predecessor known complete (and fenced)
        if (dep >= 0) {
          RDB_STAMP(it, 2 * c - 1);
          const int bl = (item.y0 > 0 ? item.y0 - 1 : 0) >> RDB_FLAG_SHIFT;
          const int r_hi = item.y0 + item.rows < L.H ? item.y0 + item.rows : L.H - 1;
          const int nb = (r_hi >> RDB_FLAG_SHIFT) - bl + 1;
          RDB_ASSERT(nb >= 1 && nb <= RDB_MAX_DEP_BLOCKS, "dependency block count", nb, item.rows);
          RDB_ASSERT(dep + bl >= 0 && dep + bl + nb <= args.nflags, "dependency flag range", dep + bl, nb);
          int m = 0;   // leading complete blocks
          RDB_TIMED(4, {
            if (lane == 0) {
              const int* f = args.flags + dep + bl;
              int v[RDB_MAX_DEP_BLOCKS];
#pragma unroll
              for (int j2 = 0; j2 < RDB_MAX_DEP_BLOCKS; ++j2) v[j2] = j2 < nb ? ld_relaxed_gpu(f + j2) : args.flag_target;
#pragma unroll
              for (int j2 = 0; j2 < RDB_MAX_DEP_BLOCKS; ++j2)
                if (m == j2 && j2 < nb && v[j2] >= args.flag_target) m = j2 + 1;
              if (m > 0) {
                fence_acq_rel_gpu();
                fence_proxy_async_global();
              }
            }
            m = __shfl_sync(0xffffffffu, m, 0);
          });
          ready_blk = bl + m - 1;
        }
        // x.hi (chunk 0) is read by all five convs, ~5 steps apart, with ~100 MB of other traffic in between: keep
        // it in L2 (evict_last) until conv5, its last reader, has passed (evict_first)
#ifdef B200SR_ABL_NOHINT
        const uint64_t pol = pol_item;
#else
        const uint64_t pol = c == 0 ? pol_last : pol_item;   // (conv5's epilogue reads x.hi once more)
#endif
        for (int y = -1; y <= item.rows; ++y) {
          const int r = item.y0 + y;
          if (r >= 0 && r < L.H && (r >> RDB_FLAG_SHIFT) > ready_blk) {   // first row of a block not yet known complete
            RDB_TIMED(4, {
              if (lane == 0) {
                const int* f = args.flags + dep + (r >> RDB_FLAG_SHIFT);
                RDB_ASSERT(dep >= 0 && dep + (r >> RDB_FLAG_SHIFT) < args.nflags, "lazy dependency flag", dep, r);
                while (ld_relaxed_gpu(f) < args.flag_target) __nanosleep(B200SR_RDB_POLL_NS);
                fence_acq_rel_gpu();
                fence_proxy_async_global();
              }
              __syncwarp();
            });
            ready_blk = r >> RDB_FLAG_SHIFT;
          }
          RDB_TIMED(1, mbar_wait(&bar_empty[stage], phase ^ 1));
          if (lane == 0) {   // (the lane that executed the fences above)
            RDB_ASSERT(r >= -1 && r <= L.H && x0 >= -1 && x0 < L.W && c * L.N + item.n < 3 * L.N, "TMA coordinate", r, x0);
            mbar_arrive_expect_tx(&bar_full[stage], half ? 130 * 64 : 130 * 128);
            const int rr = (RDB_RING > 0 && c > 0 && r >= 0 && r < L.H) ? r % RDB_RING : r;
            tma_load_4d_hint(half ? &amap_h : &amap, &bar_full[stage], sA + stage * RDB_A_STAGE_BYTES, 0, x0, rr,
                             c * L.N + item.n, pol);
          }
          __syncwarp();
          if (++stage == RDB_NSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++wb == RDB_NWBUF) {
          wb = 0;
          wphase ^= 1;
        }
      }
      RDB_STAMP(it, 5);
    }
    if (st_on && lane == 0) {
      long long* o = args.stats + blockIdx.x * 16;
      o[0] = st_acc[4];
      o[1] = st_acc[1];
      o[2] = st_acc[2];
      o[3] = clock64() - st_t0;

    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint64_t adesc0_f = make_smem_desc(smem_u32(sA), 1024, SWZ_128B, 0);
    const uint64_t bdesc0_f = make_smem_desc(smem_u32(sW), 1024, SWZ_128B, 0);
    const uint64_t adesc0_h = make_smem_desc(smem_u32(sA), 512, SWZ_64B, 0);    // 64-byte rows (half chunks)
    const uint64_t bdesc0_h = make_smem_desc(smem_u32(sW), 512, SWZ_64B, 0);
    int stage = 0, phase = 0, wb = 0, wphase = 0;
    uint32_t rempty_par = 0xFFFFu;   // parity to wait for, per 32-column slot (fresh barrier: parity 1 passes)
    int qi = 0, qphase = 0;
    while (true) {
      mbar_wait(&bar_qfull[qi], qphase);
      const int it = s_q[qi];
      const RdbItem item = s_qitem[qi];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_qempty[qi]);
      if (++qi == RDB_QD) {
        qi = 0;
        qphase ^= 1;
      }
      if (it < 0) break;
      RDB_STAMP(it, 9);
      const ConvArgs& L = args.L[item.k];
      const int TH = item.rows;
      const int cout = item.k < 4 ? 32 : 64;
      const int spr = cout >> 5;   // 32-column slots per accumulator row
      const uint32_t wtile = 3u * cout * 128u;
      const uint32_t idesc1 = make_idesc_16(128, cout, false);
      const uint32_t idesc2 = make_idesc_16(128, 2 * cout, false);
      const uint32_t idesc3 = make_idesc_16(128, 3 * cout, false);
      for (int c = 0; c < L.nchunks; ++c) {
        const int ks = (c == L.nchunks - 1) ? L.last_ksteps : 4;
        const bool first_chunk = (c == 0);
        const bool last_chunk = (c == L.nchunks - 1);
        RDB_TIMED(0, mbar_wait(&bar_wfull[wb], wphase));
        const bool half = args.half64 && ks == 2;           // 32-channel chunk held as 64-byte rows (SWIZZLE_64B)
        const uint32_t rb = half ? 64u : 128u;              // bytes per A / B row = stride of the kx tap views
        const uint32_t wt = half ? wtile / 2 : wtile;       // bytes per dx tile of the weight chunk
        const uint64_t adesc0 = half ? adesc0_h : adesc0_f;
        const uint64_t bdesc0 = half ? bdesc0_h : bdesc0_f;
        const uint64_t bdesc_w = bdesc0 + static_cast<uint64_t>((wb * RDB_WBUF_BYTES) >> 4);
        // Two input rows (two pipeline stages) per burst: the tensor pipe buffers only a couple of instructions, so
        // every barrier wait / descriptor computation between bursts is a pipe bubble (probe_umma.cu T6: the gap
        // costs ~150 cycles + ~40-85 per commit regardless of burst length).  24 MMAs per burst halve that cost.
        auto burst = [&](const int y, const int ny) {
          uint32_t dcol[2], idn[2], nblk_[2];
          uint64_t ad0[2], bd0[2];
          bool newr[2];
          int stg[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int yy = y + j;
            const int blk_lo = (yy < 1) ? (1 - yy) : 0;
            const int blk_hi = (TH - yy < 2) ? (TH - yy) : 2;
            const int nblk = blk_hi - blk_lo + 1;
            nblk_[j] = nblk;
            newr[j] = first_chunk && blk_hi == 2 && j < ny;   // accumulator row yy+1 is touched for the first time
            stg[j] = (stage + j) & (RDB_NSTAGES - 1);
            dcol[j] = tmem_base + static_cast<uint32_t>((yy - 1 + blk_lo) * cout);
            ad0[j] = adesc0 + static_cast<uint64_t>((stg[j] * RDB_A_STAGE_BYTES) >> 4);
            bd0[j] = bdesc_w + static_cast<uint64_t>((blk_lo * cout * rb) >> 4);
            idn[j] = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
            if (newr[j]) {
              for (int sl = (yy + 1) * spr; sl < (yy + 2) * spr; ++sl) {
                RDB_TIMED(1, mbar_wait(&bar_rempty[sl], (rempty_par >> sl) & 1u));
                rempty_par ^= 1u << sl;
              }
            }
          }
          RDB_TIMED(2, mbar_wait(&bar_full[stg[0]], phase));
          if (ny == 2) RDB_TIMED(2, mbar_wait(&bar_full[stg[1]], stg[1] < stg[0] ? (phase ^ 1) : phase));
          tc_fence_after();
          const long long t_issue = st_on ? clock64() : 0;
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j < ny) {
                if (newr[j]) {
                  if (nblk_[j] > 1) umma_bf16(dcol[j], ad0[j], bd0[j], nblk_[j] == 3 ? idesc2 : idesc1, 1);
                  umma_bf16(dcol[j] + (nblk_[j] - 1) * cout, ad0[j],
                            bd0[j] + static_cast<uint64_t>(((nblk_[j] - 1) * cout * rb) >> 4), idesc1, 0);
                } else {
                  umma_bf16(dcol[j], ad0[j], bd0[j], idn[j], 1);
                }
                if (ks == 4) {
#pragma unroll
                  for (int i = 1; i < 12; ++i) {
                    const int dx = i >> 2, k = i & 3;
                    umma_bf16(dcol[j], ad0[j] + static_cast<uint64_t>((dx * 128 + k * 32) >> 4),
                              bd0[j] + static_cast<uint64_t>((dx * wtile + k * 32) >> 4), idn[j], 1);
                  }
                } else {
#pragma unroll
                  for (int i = 1; i < 6; ++i) {
                    const int dx = i >> 1, k = i & 1;
                    umma_bf16(dcol[j], ad0[j] + static_cast<uint64_t>((dx * rb + k * 32) >> 4),
                              bd0[j] + static_cast<uint64_t>((dx * wt + k * 32) >> 4), idn[j], 1);
                  }
                }
                umma_commit(&bar_empty[stg[j]]);
                const int yy = y + j;
                if (last_chunk && yy >= 1) {
                  umma_commit(&bar_rfull[(yy - 1) * spr]);   // output row yy-1 is complete
                  RDB_STAMP2(it, yy - 1, 0);
                }
              }
            }
            if (y + ny - 1 == TH) umma_commit(&bar_wempty[wb]);
          }
          __syncwarp();
          if (st_on) {
            RDB_COUNT(3, clock64() - t_issue);
            RDB_COUNT(4, ny);
          }
          stage += ny;
          if (stage >= RDB_NSTAGES) {
            stage -= RDB_NSTAGES;
            phase ^= 1;
          }
        };
        // Row schedule of one chunk: (-1, 0) general; interior pairs through the lean path (rows 1 .. TH-2 feed three
        // full accumulator rows each: N = 3*cout, no edge cases); the last rows through the general path again.
        burst(-1, 2);
        int y = 1;
        // Interior pairs.  The tensor pipe queues only ~2 instructions, so whatever the issuing thread does between
        // the last MMA of one burst and the first of the next is a pipe bubble unless it fits into the ~100-190
        // cycles those two instructions take.  The barrier waits of the NEXT pair (activation stages, and in the
        // first chunk the TMEM slots it touches first) are therefore PROBED (non-blocking) by the issuing thread just
        // before the last two MMAs of the current pair: if everything is already there, the next pair starts without
        // any wait; if not, it waits as usual.  (Blocking there instead was measured slower: it delays this pair's
        // commits, i.e. the stage hand-back to the producer and the rows' hand-over to the epilogue.)
        bool waited = false;   // this pair's barriers were already observed by the issuing thread
        for (; y + 1 <= TH - 2; y += 2) {
          const int s0 = stage, s1 = (stage + 1) & (RDB_NSTAGES - 1);
          const uint32_t p0 = phase, p1 = s1 < s0 ? (phase ^ 1) : phase;
          if (first_chunk) {   // accumulator rows y+1 and y+2 are touched for the first time
            for (int sl = (y + 1) * spr; sl < (y + 3) * spr; ++sl) {
              if (!waited) RDB_TIMED(1, mbar_wait(&bar_rempty[sl], (rempty_par >> sl) & 1u));
              rempty_par ^= 1u << sl;
            }
          }
          if (!waited) {
            RDB_TIMED(2, mbar_wait(&bar_full[s0], p0));
            RDB_TIMED(2, mbar_wait(&bar_full[s1], p1));
          }
          tc_fence_after();
          // the next interior pair (if any): its stages / parities / first-touch slots
          const bool has_next = (y + 3) <= TH - 2;
          const int nst = (stage + 2) & (RDB_NSTAGES - 1);
          const uint32_t nph = (stage + 2 >= RDB_NSTAGES) ? (phase ^ 1) : phase;
          const int n0 = nst, n1 = (nst + 1) & (RDB_NSTAGES - 1);
          const uint32_t np0 = nph, np1 = n1 < n0 ? (nph ^ 1) : nph;
          const uint32_t dc0 = tmem_base + static_cast<uint32_t>((y - 1) * cout);
          const uint64_t a0 = adesc0 + static_cast<uint64_t>((s0 * RDB_A_STAGE_BYTES) >> 4);
          const uint64_t a1 = adesc0 + static_cast<uint64_t>((s1 * RDB_A_STAGE_BYTES) >> 4);
          bool next_ready = false;   // (issuing thread) the next pair's barriers were all found complete
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint32_t dc = dc0 + j * cout;
              const uint64_t ad = j ? a1 : a0;
              if (first_chunk) {
                umma_bf16(dc, ad, bdesc_w, idesc2, 1);
                umma_bf16(dc + 2 * cout, ad, bdesc_w + static_cast<uint64_t>((2 * cout * 128) >> 4), idesc1, 0);
              } else {
                umma_bf16(dc, ad, bdesc_w, idesc3, 1);
              }
              if (ks == 4) {
#pragma unroll
                for (int i = 1; i < 12; ++i) {
                  if (j == 1 && i == 10 && has_next) {
                    bool ok = mbar_test_wait(&bar_full[n0], np0) && mbar_test_wait(&bar_full[n1], np1);
                    if (first_chunk)
                      for (int sl = (y + 3) * spr; sl < (y + 5) * spr; ++sl)
                        ok = ok && mbar_test_wait(&bar_rempty[sl], (rempty_par >> sl) & 1u);
                    next_ready = ok;
                  }
                  const int dx = i >> 2, k = i & 3;
                  umma_bf16(dc, ad + static_cast<uint64_t>((dx * 128 + k * 32) >> 4),
                            bdesc_w + static_cast<uint64_t>((dx * wtile + k * 32) >> 4), idesc3, 1);
                }
              } else {
#pragma unroll
                for (int i = 1; i < 6; ++i) {
                  if (j == 1 && i == 4 && has_next) {
                    bool ok = mbar_test_wait(&bar_full[n0], np0) && mbar_test_wait(&bar_full[n1], np1);
                    if (first_chunk)
                      for (int sl = (y + 3) * spr; sl < (y + 5) * spr; ++sl)
                        ok = ok && mbar_test_wait(&bar_rempty[sl], (rempty_par >> sl) & 1u);
                    next_ready = ok;
                  }
                  const int dx = i >> 1, k = i & 1;
                  umma_bf16(dc, ad + static_cast<uint64_t>((dx * rb + k * 32) >> 4),
                            bdesc_w + static_cast<uint64_t>((dx * wt + k * 32) >> 4), idesc3, 1);
                }
              }
              umma_commit(&bar_empty[j ? s1 : s0]);
              if (last_chunk) {
                umma_commit(&bar_rfull[(y + j - 1) * spr]);   // output row y+j-1 is complete
                RDB_STAMP2(it, y + j - 1, 0);
              }
            }
          }
          waited = __any_sync(0xffffffffu, next_ready);
          stage += 2;
          if (stage >= RDB_NSTAGES) {
            stage -= RDB_NSTAGES;
            phase ^= 1;
          }
        }
        for (; y <= TH; y += 2) burst(y, (y + 1 <= TH) ? 2 : 1);
        if (++wb == RDB_NWBUF) {
          wb = 0;
          wphase ^= 1;
        }
      }
      RDB_STAMP(it, 6);
    }
    if (st_on && lane == 0) {
      long long* o = args.stats + blockIdx.x * 16;
      o[4] = st_acc[0];
      o[5] = st_acc[1];
      o[6] = st_acc[2];
      o[7] = st_acc[3];
      o[8] = st_acc[4];
      o[9] = clock64() - st_t0;
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9)
    const int eg = (warp - 2) >> 2;          // row-parity group
    const int q = warp & 3;                  // TMEM lane quarter
    const int m = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t rfull_par = 0;                  // parity to wait for, per slot
    int qi = 0, qphase = 0;
    while (true) {
      mbar_wait(&bar_qfull[qi], qphase);
      const int it = s_q[qi];
      const RdbItem item = s_qitem[qi];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_qempty[qi]);
      if (++qi == RDB_QD) {
        qi = 0;
        qphase ^= 1;
      }
      if (it < 0) break;
      const ConvArgs& L = args.L[item.k];
      const int n = item.n;
      const int x = item.tx * 128 + m;
      const bool xp = B200SR_EPI_XPOSE != 0 && item.tx * 128 + 128 <= L.W;   // pair-coalesced stores: full tiles only
      if (item.k == 4) {
        // pull the residual (lo) rows of this item towards L2 while the MMAs run (2 groups x 1 KB per warp-row); the
        // x.hi rows arrive with the chunk-0 TMA loads.  At the RRDB end also x0.lo of the RRDB input pair (last
        // touched three launches ago; prefetching x0.hi as well measured ~1 % slower, more early DRAM traffic).
        for (int Y = eg; Y < item.rows; Y += RDB_NGRP) {
          const size_t o = lo_off(n, item.y0 + Y, item.tx * 128 + q * 32, L.H, L.W) + (lane & 7) * 128 +
                           static_cast<size_t>((lane >> 3) & 1) * LO_GSTRIDE;
          if (lane < 16) {
            if constexpr (LO_IN) prefetch_l2(L.lo_in + o);
            if constexpr (RRDB_END) prefetch_l2(L.xb_lo + o);
          }
        }
        for (int Y = 0; Y < item.rows; ++Y) {
          const int sl = 2 * Y;
          if ((Y % RDB_NGRP) == eg) {
            // Request one of this pixel's pairs now, while the MMAs of the row are still in flight: the RDB input
            // pair x, or at the RRDB end the RRDB input pair x0 -- x0 was last touched three launches ago and
            // comes from DRAM, x is L2-warm (chunk-0 TMA loads, lo prefetch) and is fetched after the accumulators.
            uint32_t ph[4][8], pl[2][8];
            const size_t pix = (static_cast<size_t>(n) * L.H + item.y0 + Y) * L.W + x;
            const size_t loff = lo_off(n, item.y0 + Y, x < L.W ? x : 0, L.H, L.W);
            if (x < L.W) {
              if constexpr (RRDB_END)
                load_trunk_pair<true>(L.xb_hi + pix * L.out_pitch, L.xb_lo + loff, ph, pl);
              else
                load_trunk_pair<LO_IN>(L.hi_in + pix * L.out_pitch, L.lo_in + loff, ph, pl);
            }
            RDB_TIMED(0, mbar_wait(&bar_rfull[sl], (rfull_par >> sl) & 1u));
            if (q == 2 && lane == 0) RDB_STAMP2(it, Y, 1);
            tc_fence_after();
            float acc[64];
            RDB_TIMED(1, load_acc_row<64>(tlane + static_cast<uint32_t>(Y * 64), acc));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&bar_rempty[sl]);
              mbar_arrive(&bar_rempty[sl + 1]);
            }
            RDB_TIMED(2, {
              if (x < L.W && !(B200SR_ABL_NOEPI & 2)) {
                if constexpr (RRDB_END) {
                  uint32_t xh[4][8], xl[2][8];
                  load_trunk_pair<LO_IN>(L.hi_in + pix * L.out_pitch, L.lo_in + loff, xh, xl);
                  trunk_pixel<true, LO_IN, LO_OUT>(L, s_bias[4], acc, xh, xl, ph, pl, n, item.y0 + Y, x, xp);
                } else {
                  trunk_pixel<false, LO_IN, LO_OUT>(L, s_bias[4], acc, ph, pl, ph, pl, n, item.y0 + Y, x, xp);
                }
              }
            });
            RDB_COUNT(3, 1);
            if (q == 2 && lane == 0) RDB_STAMP2(it, Y, 2);
          }
          rfull_par ^= 1u << sl;   // every epilogue warp tracks every slot's phase
        }
#if B200SR_RDB_DISCARD
        // this group drained the item's last row: every MMA of the item (hence every TMA read it made) is complete
        if (((item.rows - 1) % RDB_NGRP) == eg && m >= 1 && m <= 126 && x + 1 < L.W) {
          const size_t plane = static_cast<size_t>(L.N) * L.H * L.W * 64;
          for (int Y = 1; Y <= item.rows - 2; ++Y) {
            const __nv_bfloat16* px = L.hi_in + ((static_cast<size_t>(n) * L.H + item.y0 + Y) * L.W + x) * 64;
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(px + plane) : "memory");
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(px + 2 * plane) : "memory");
          }
        }
#endif
      } else {
        for (int Y = 0; Y < item.rows; ++Y) {
          if ((Y % RDB_NGRP) == eg) {
            RDB_TIMED(0, mbar_wait(&bar_rfull[Y], (rfull_par >> Y) & 1u));
            if (q == 2 && lane == 0) RDB_STAMP2(it, Y, 1);
            tc_fence_after();
            float acc[32];
            RDB_TIMED(4, load_acc_row<32>(tlane + static_cast<uint32_t>(Y * 32), acc));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_rempty[Y]);
            RDB_TIMED(5, {
              if (x < L.W && !(B200SR_ABL_NOEPI & 1))
                epilogue_pixel<32, EPI_ACT_BF16>(L, s_bias[item.k], s_bias[item.k], acc, n,
                                                 RDB_RING > 0 ? (item.y0 + Y) % RDB_RING : item.y0 + Y, x, xp);
            });
            RDB_COUNT(6, 1);
            if (q == 2 && lane == 0) RDB_STAMP2(it, Y, 2);
          }
          rfull_par ^= 1u << Y;
          // publish an block once this warp has stored its rows of it (item.y0 is a multiple of 8); the
          // block is complete when every epilogue warp of every column tile has added 1
          if ((Y & (RDB_FLAG_ROWS - 1)) == RDB_FLAG_ROWS - 1 || Y == item.rows - 1) {
            RDB_TIMED(7, {
              __syncwarp();   // orders every lane's stores before lane 0's release (cumulative at gpu scope)
              RDB_ASSERT(item.flag_base + ((item.y0 + Y) >> RDB_FLAG_SHIFT) < args.nflags, "release flag index",
                         item.flag_base, item.y0 + Y);
              if (lane == 0) red_release_gpu_add(args.flags + item.flag_base + ((item.y0 + Y) >> RDB_FLAG_SHIFT), 1);
            });
          }
        }
      }
      if (warp == 2) RDB_STAMP(it, 7);
    }
  }

  if (st_on && warp == 2 && lane == 0) {
    long long* o = args.stats + blockIdx.x * 16;
    o[10] = st_acc[0];
    o[11] = clock64() - st_t0;
    // epilogue detail reuses the producer's per-conv slots of CTAs' row 12..15 is taken; print via printf-free path:
    o[12] = st_acc[1];   // conv5: tcgen05.ld
    o[13] = st_acc[2];   // conv5: math + global
    o[14] = st_acc[4] + st_acc[5];   // conv1-4: ld + math/global
    o[15] = st_acc[7] * 1000000 + st_acc[3] * 1000 + st_acc[6];   // release cycles | conv5 rows | conv1-4 rows
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, RDB_TMEM_COLS);
}

}  // namespace b200sr
