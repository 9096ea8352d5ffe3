// libb200sr.so -- engine + C ABI (include/b200sr.h).
// Host-side walker of the RRDBNet / SRVGGNetCompact forward graphs over the sm_100a kernels in
// conv3x3_tc.cuh / pointwise.cuh, plus the RealESRGANer pre/post/tile logic
// (reference call site: /root/reference/src/framewright/processors/pytorch_realesrgan.py:160-170, 223).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/b200sr.h"
#include "conv3x3_tc.cuh"
#if B200SR_CTAS_PER_SM == 1   // (the 2-CTAs-per-SM dev build of the per-row kernel has no room for these)
#include "conv3x3_sc.cuh"
#include "hr_last_fused.cuh"
#define B200SR_HAVE_SC 1
#else
#define B200SR_HAVE_SC 0
#endif
#include "pointwise.cuh"
#include "rdb_fused.cuh"
#include "tmap.h"

using namespace b200sr;

namespace {

struct Layer {
  int cin = 0, cout = 0;   // true channel counts
  int coutp = 0;           // tensor-core N (16 / 32 / 48 / 64)
  bool set = false;
  bool fp16 = false;            // 16-bit format of this layer's inputs and weights (false = bf16)
  std::vector<float> w, b;      // host fp32 OIHW / bias
  uint8_t* d_wpack = nullptr;   // packed bf16 image (tensor-core layers)
  uint8_t* d_wsub[4] = {nullptr, nullptr, nullptr, nullptr};   // conv_up1 / conv_up2: one image per sub-pixel phase (a, b)
  uint8_t* d_wrdb = nullptr;    // RDB conv2 / conv4 (Cin % 64 == 32): image whose last chunk has 64-byte rows (SWIZZLE_64B)
  uint8_t* d_wlast9 = nullptr;  // conv_last: kx taps stacked on N (conv3x3_sc.cuh, EPI_LAST9_U8)
  float* d_wfirst = nullptr;    // [9][cin][64] fp32 (first layer, CUDA-core kernel)
  std::vector<float> wfirst;    // the same on the host (3-channel first layer: passed as a kernel parameter)
  float* d_bias = nullptr;      // coutp floats
};

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline uint16_t f2bf(float f) {  // round-to-nearest-even, matches __float2bfloat16_rn for finite values
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

inline uint16_t f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

}  // namespace

// One lane = everything a forward pass needs exclusively: a stream, the activation workspace, the fused-RDB
// work list + completion counters, and (host-buffer calls) device + pinned staging.  Host-buffer calls take a free
// lane each, so concurrent callers (the reference's parallel_frames threads, restorer.py:1894) and the chunks of one
// large call overlap their H2D / D2H copies and launch tails with each other's kernels.
struct Lane {
  int id = 0;
  bool busy = false;
  cudaStream_t stream = nullptr;
  uint8_t* ws = nullptr;          // workspace (grow-only)
  size_t ws_bytes = 0;
  // fused-RDB item tables (one per (N, H, W): tile mode walks six region shapes per frame) and their counters
  struct RdbTable {
    int n = 0, h = 0, w = 0, gen = -1, nitems = 0, nflags = 0;
    RdbItem* d_items = nullptr;
    int* d_flags = nullptr;
  };
  std::vector<RdbTable> rdb_tables;
  int rdb_nitems = 0, rdb_nflags = 0;      // of the table selected by the last ensure_rdb_table
  RdbItem* d_rdb_items = nullptr;
  int* d_rdb_flags = nullptr;
  int rdb_launch_idx = 0;
  int launches = 0;
  // staging of the host-buffer entry point
  uint8_t* dev_in = nullptr;
  uint8_t* dev_out = nullptr;
  uint8_t* pin_in = nullptr;
  uint8_t* pin_out = nullptr;
  size_t dev_in_bytes = 0, dev_out_bytes = 0, pin_in_bytes = 0, pin_out_bytes = 0;
};

struct b200sr_engine {
  b200sr_model_desc desc{};
  int device = 0;
  int num_sms = 148;
  std::vector<Layer> layers;
  std::vector<std::vector<float>> prelu_host;
  std::vector<float*> prelu_dev;
  bool finalized = false;
  std::string err;
  // lanes: `dev_lane` serves the device-pointer entry points (caller's stream), `lanes` the host-buffer ones
  std::mutex mu;                   // lane acquisition, err, prof
  std::condition_variable cv;
  Lane dev_lane;                                   // device-pointer calls on any stream beyond the first few
  std::vector<std::pair<cudaStream_t, std::unique_ptr<Lane>>> dev_lanes;   // one lane per caller stream (<= 4)
  std::vector<std::unique_ptr<Lane>> lanes;
  int opt_lanes = 2;
  int opt_host_chunk = 0;          // frames per lane job of a host-buffer call (0 = auto)
  long long opt_ws_limit_mb = 0;   // > 0: refuse workspaces above this size (tests: forces the OOM path)
  std::atomic<int> last_launches{0};
  int opt_force_th = 0;     // 0 = auto
  int opt_max_ctas = 0;     // 0 = one per SM
  // optional per-kernel-class timing (CUDA events around every launch; option "profile")
  int opt_fused_rdb = 1;    // run each RDB as one persistent kernel (L2-resident intermediates)
  int opt_first_v1 = 0;     // input stage: 0 = constant-bank kernel (3 ch) / tiled kernel (12 ch); 1 = round-1 per-pixel
                            // kernel; 2 = tiled kernel for 3 ch too (all bit-identical; tests / A-B timing)
  int opt_trunk_lo = 0;     // where the residual stream's e5m2 lo part is used: 0 = in the RRDB-level skip only (written
                            // at every RRDB end, read at the next one); 1 = also in the first RDB's own residual add;
                            // 2 = the pair after EVERY RDB (rounds 1-2a)
  int opt_fuse_tail = 1;    // conv_hr + conv_last as one kernel (hr_last_fused.cuh; needs opt_pair and opt_last9)
  int opt_half64 = 1;       // fused RDB: the 32-channel last chunk of conv2 / conv4 as a 32-channel SWIZZLE_64B box
  int opt_pair = 1;         // single-chunk convs through conv3x3_sc_kernel (resident weights, row-pair stages)
  int opt_last9 = 1;        // conv_last with the kx taps stacked on N (needs opt_pair)
  int opt_abl = 0;          // dev: timing ablations of the per-conv kernel (ConvArgs::abl); results are wrong when set
  int opt_w_resident = 1;   // single-chunk convs keep their weights resident in shared memory (more activation stages)
  int opt_fold_up = 1;      // conv_up1/up2: 0 materialised upsampling, 1 duplicated-pixel TMA view, 2 sub-pixel phases
  int opt_rdb_stats = 0;    // dev: collect per-CTA cycle counters of the k-th fused launch of a forward pass (1-based)
  int rdb_gen = 0;          // bumped when a schedule option changes: lanes rebuild their work lists
  int stats_nitems = 0;
  long long* d_rdb_stats = nullptr;
  long long* d_rdb_trace = nullptr;
  long long* d_rdb_trace2 = nullptr;
  int opt_profile = 0;
  struct ProfRec {
    int cls;
    double flops;
    cudaEvent_t e0, e1;
  };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
};

enum ProfClass {
  PC_CONV32_ACT = 0, PC_CONV64_ACT, PC_CONV64_PRELU, PC_CONV64_RDB5, PC_CONV64_RDB5_RRDB, PC_CONV64_ADD,
  PC_CONV16_LAST, PC_CONV48_SRVGG_LAST, PC_FIRST, PC_UPSAMPLE, PC_RDB_FUSED, PC_HR_LAST, PC_COUNT
};

namespace {

thread_local std::string tls_err;   // message of the last failed call on this thread (engines are shared by threads)

int fail(b200sr_engine* e, int code, const std::string& msg) {
  tls_err = msg;
  if (e) {
    std::lock_guard<std::mutex> g(e->mu);
    e->err = msg;
  }
  return code;
}

cudaEvent_t prof_event(b200sr_engine* e) {
  if (!e->ev_pool.empty()) {
    cudaEvent_t ev = e->ev_pool.back();
    e->ev_pool.pop_back();
    return ev;
  }
  cudaEvent_t ev = nullptr;
  cudaEventCreate(&ev);
  return ev;
}
struct ProfScope {
  b200sr_engine* e;
  cudaStream_t st;
  size_t idx = 0;
  bool on;
  ProfScope(b200sr_engine* e_, int cls, double flops, cudaStream_t st_) : e(e_), st(st_), on(e_->opt_profile != 0) {
    if (!on) return;
    std::lock_guard<std::mutex> g(e->mu);
    b200sr_engine::ProfRec r{cls, flops, prof_event(e), prof_event(e)};
    cudaEventRecord(r.e0, st);
    idx = e->prof.size();
    e->prof.push_back(r);
  }
  ~ProfScope() {
    if (!on) return;
    std::lock_guard<std::mutex> g(e->mu);
    cudaEventRecord(e->prof[idx].e1, st);
  }
};

#define CUDA_TRY(e, expr)                                                                          \
  do {                                                                                             \
    cudaError_t err__ = (expr);                                                                    \
    if (err__ != cudaSuccess) {                                                                    \
      cudaGetLastError(); /* reset the sticky-until-read error so the next launch check is clean */ \
      int code__ = (err__ == cudaErrorMemoryAllocation) ? B200SR_ERR_OOM : B200SR_ERR_CUDA;        \
      return fail(e, code__, std::string(#expr) + ": " + cudaGetErrorString(err__));               \
    }                                                                                              \
  } while (0)

int coutp_for(int cout) {
  if (cout <= 16) return 16;
  if (cout <= 32) return 32;
  if (cout <= 48) return 48;
  return 64;
}

// ---- architecture tables (execution order == framewright_b200/archs.py::conv_layers) ----
void build_layers(b200sr_engine* e) {
  const auto& d = e->desc;
  const int nf = d.num_feat, gc = d.num_grow_ch;
  auto add = [&](int cin, int cout) {
    Layer l;
    l.cin = cin;
    l.cout = cout;
    l.coutp = coutp_for(cout);
    e->layers.push_back(std::move(l));
  };
  if (d.arch == B200SR_ARCH_RRDB) {
    add(d.scale == 2 ? 12 : 3, nf);
    for (int b = 0; b < d.num_block; ++b)
      for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 4; ++k) add(nf + k * gc, gc);
        add(nf + 4 * gc, nf);
      }
    add(nf, nf);  // conv_body (bf16 in, fp16 out)
    // HR tail in fp16: its roundings land straight in the [0,1] image (DESIGN.md "16-bit formats")
    add(nf, nf);  // conv_up1
    add(nf, nf);  // conv_up2
    add(nf, nf);  // conv_hr
    add(nf, 3);   // conv_last
    for (size_t i = e->layers.size() - 4; i < e->layers.size(); ++i) e->layers[i].fp16 = true;
  } else {
    // SRVGG has no fp32 residual trunk: all 16-bit tensors are fp16
    add(3, nf);
    for (int i = 0; i < d.num_block; ++i) add(nf, nf);
    add(nf, 3 * d.scale * d.scale);
    for (auto& l : e->layers) l.fp16 = true;
    e->prelu_host.resize(d.num_block + 1);
    e->prelu_dev.assign(d.num_block + 1, nullptr);
  }
}

// Packed tensor-core weight image: [chunk][dx][blk*COUTP + co][64 ch], 128-byte rows, SWIZZLE_128B,
// blk 0/1/2 <-> ky 2/1/0 (input row y feeds output rows y-1, y, y+1), dx <-> kx.
std::vector<uint8_t> pack_weights(const Layer& l) {
  const int nchunks = (l.cin + 63) / 64;
  const int COUTP = l.coutp;
  const size_t tile_bytes = static_cast<size_t>(3) * COUTP * 128;
  std::vector<uint8_t> img(static_cast<size_t>(nchunks) * 3 * tile_bytes, 0);
  for (int c = 0; c < nchunks; ++c)
    for (int dx = 0; dx < 3; ++dx) {
      uint8_t* tile = img.data() + (static_cast<size_t>(c) * 3 + dx) * tile_bytes;
      for (int blk = 0; blk < 3; ++blk) {
        const int ky = 2 - blk;
        for (int co = 0; co < l.cout; ++co) {
          const int r = blk * COUTP + co;
          for (int j = 0; j < 64; ++j) {
            const int ci = c * 64 + j;
            if (ci >= l.cin) break;
            const float v = l.w[((static_cast<size_t>(co) * l.cin + ci) * 3 + ky) * 3 + dx];
            uint32_t off = static_cast<uint32_t>(r) * 128 + (j / 8) * 16 + (j % 8) * 2;
            off ^= ((off >> 7) & 7u) << 4;
            const uint16_t h = l.fp16 ? f2h(v) : f2bf(v);
            memcpy(tile + off, &h, 2);
          }
        }
      }
    }
  return img;
}

// Fused-RDB image of a layer with Cin % 64 == 32 (conv2, conv4): full chunks as in pack_weights, the last chunk
// (32 channels) with 64-byte rows, SWIZZLE_64B, at the same chunk offset -- it is copied as 3 x (3*COUTP x 64 B).
std::vector<uint8_t> pack_weights_rdb_half(const Layer& l) {
  std::vector<uint8_t> img = pack_weights(l);
  const int nchunks = (l.cin + 63) / 64;
  const int COUTP = l.coutp;
  const size_t tile_bytes = static_cast<size_t>(3) * COUTP * 128;
  uint8_t* chunk = img.data() + static_cast<size_t>(nchunks - 1) * 3 * tile_bytes;
  memset(chunk, 0, 3 * tile_bytes);
  const size_t htile = tile_bytes / 2;
  for (int dx = 0; dx < 3; ++dx)
    for (int blk = 0; blk < 3; ++blk) {
      const int ky = 2 - blk;
      for (int co = 0; co < l.cout; ++co) {
        const int r = blk * COUTP + co;
        for (int j = 0; j < 32; ++j) {
          const int ci = (nchunks - 1) * 64 + j;
          const float v = l.w[((static_cast<size_t>(co) * l.cin + ci) * 3 + ky) * 3 + dx];
          uint32_t off = static_cast<uint32_t>(r) * 64 + (j / 8) * 16 + (j % 8) * 2;
          off ^= ((off >> 7) & 3u) << 4;
          const uint16_t h = l.fp16 ? f2h(v) : f2bf(v);
          memcpy(chunk + dx * htile + off, &h, 2);
        }
      }
    }
  return img;
}

// "nearest-2x upsample, then 3x3 conv" == four 2x2 convs on the low-resolution grid, one per output phase (a, b):
// HR row 2y+a reads source rows {y-1, y, y} (a = 0) or {y, y, y+1} (a = 1) for ky = 0, 1, 2, so the taps that read the
// same source row are summed (same for columns).  The 2x2 kernel is embedded in a 3x3 one -- (ky', kx') = position
// relative to the centre -- whose unused row / column is zero and is never issued (ConvArgs::sub_*).
std::vector<uint8_t> pack_subpixel_weights(const Layer& l, int a, int b) {
  Layer t;
  t.cin = l.cin;
  t.cout = l.cout;
  t.coutp = l.coutp;
  t.fp16 = l.fp16;
  t.w.assign(l.w.size(), 0.f);
  const int vmap[2][3] = {{0, 1, 1}, {1, 1, 2}};   // ky (or kx) -> ky' for phase 0 / 1
  for (int co = 0; co < l.cout; ++co)
    for (int ci = 0; ci < l.cin; ++ci) {
      const size_t base = (static_cast<size_t>(co) * l.cin + ci) * 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) t.w[base + vmap[a][ky] * 3 + vmap[b][kx]] += l.w[base + ky * 3 + kx];
    }
  return pack_weights(t);
}

// conv_last (64 -> 3) with the kx taps stacked on N: one dx tile whose ky block b holds rows [kx * 8 + co] (co < 3),
// 32 rows per block (conv3x3_sc.cuh, EPI_LAST9_U8).
std::vector<uint8_t> pack_weights_last9(const Layer& l) {
  const size_t tile_bytes = static_cast<size_t>(3) * 32 * 128;
  std::vector<uint8_t> img(tile_bytes, 0);
  for (int blk = 0; blk < 3; ++blk) {
    const int ky = 2 - blk;
    for (int kx = 0; kx < 3; ++kx)
      for (int co = 0; co < l.cout; ++co) {
        const int r = blk * 32 + kx * 8 + co;
        for (int j = 0; j < 64 && j < l.cin; ++j) {
          const float v = l.w[((static_cast<size_t>(co) * l.cin + j) * 3 + ky) * 3 + kx];
          uint32_t off = static_cast<uint32_t>(r) * 128 + (j / 8) * 16 + (j % 8) * 2;
          off ^= ((off >> 7) & 7u) << 4;
          const uint16_t h = l.fp16 ? f2h(v) : f2bf(v);
          memcpy(img.data() + off, &h, 2);
        }
      }
  }
  return img;
}

#if B200SR_HAVE_SC
template <int COUT, int EPI>
int launch_sc_inst(b200sr_engine* e, Lane* lane, const CUtensorMap& amap, const ConvArgs& a, cudaStream_t st,
                   int pcls, double flops) {
  using Cfg = ScCfg<COUT, EPI>;
  ProfScope prof_scope(e, pcls, flops, st);
  static bool attr_done[16] = {};
  auto kern = conv3x3_sc_kernel<COUT, EPI>;
  if (!attr_done[e->device & 15]) {
    CUDA_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[e->device & 15] = true;
  }
  int grid = std::min(a.ntiles, e->opt_max_ctas > 0 ? e->opt_max_ctas : e->num_sms * ConvCfg<COUT>::CTAS_PER_SM);
  kern<<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>>(amap, a);
  CUDA_TRY(e, cudaGetLastError());
  lane->launches++;
  return B200SR_OK;
}

#endif

template <int COUT, int EPI>
int launch_conv_inst(b200sr_engine* e, Lane* lane, const CUtensorMap& amap, const ConvArgs& a, cudaStream_t st,
                     int pcls, double flops) {
  using Cfg = ConvCfg<COUT>;
  ProfScope prof_scope(e, pcls, flops, st);
  static bool attr_done[16] = {};
  auto kern = conv3x3_tc_kernel<COUT, EPI>;
  if (!attr_done[e->device & 15]) {
    CUDA_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[e->device & 15] = true;
  }
  int grid = std::min(a.ntiles, e->opt_max_ctas > 0 ? e->opt_max_ctas : e->num_sms * Cfg::CTAS_PER_SM);
  kern<<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>>(amap, a);
  CUDA_TRY(e, cudaGetLastError());
  lane->launches++;
  return B200SR_OK;
}

// Rows per tile: fill the 512 TMEM columns unless a smaller TH balances the waves better.
int choose_th(const b200sr_engine* e, int coutp, int N, int H, int W, int xpitch = 128) {
  const int maxth = (512 / B200SR_CTAS_PER_SM) / coutp;
  if (e->opt_force_th > 0) return std::min(e->opt_force_th, maxth);
  const int xt = (W + xpitch - 1) / xpitch;
  const int slots = e->opt_max_ctas > 0 ? e->opt_max_ctas : e->num_sms * B200SR_CTAS_PER_SM;
  double best = 1e30;
  int best_th = maxth;
  for (int th = maxth; th >= 1; --th) {
    const long tiles = static_cast<long>(xt) * ((H + th - 1) / th) * N;
    const long waves = (tiles + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (th + 1.0);  // two halo rows cost about one full row
    if (cost < best - 1e-9) {
      best = cost;
      best_th = th;
    }
  }
  return best_th;
}

struct ConvIO {
  const void* in;   // bf16 NHWC input tensor
  int in_pitch;     // channels per pixel
  int N, H, W;
  int planes = 0;   // > 0: chunk-planar input (ConvArgs::in_planes): `planes` tensors [N][H][W][64] back to back
  int up2 = 0;      // `in` is [N][H/2][W/2][pitch]; the conv reads its nearest-2x upsampling (ConvArgs::in_up2)
};

int launch_conv(b200sr_engine* e, Lane* lane, const Layer& l, int epi, const ConvIO& io, ConvArgs a, cudaStream_t st) {
  CUtensorMap amap;
  const bool map_ok = io.up2 ? tmap_encode_act_up2(&amap, io.in, io.N, io.H / 2, io.W / 2, io.in_pitch, 64, 66, 128)
                             : tmap_encode_act(&amap, io.in, io.planes ? io.planes * io.N : io.N, io.H, io.W, io.in_pitch,
                                               64, ConvCfg<64>::A_ROWS, 128);
  if (!map_ok) return fail(e, B200SR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  a.N = io.N;
  a.H = io.H;
  a.W = io.W;
  a.in_planes = io.planes;
  a.in_up2 = io.up2;
  a.nchunks = (l.cin + 63) / 64;
  a.last_ksteps = (l.cin % 64 == 32) ? 2 : 4;
  a.w_resident = (a.nchunks == 1 && e->opt_w_resident) ? 1 : 0;
  a.abl = e->opt_abl;
  a.TH = choose_th(e, l.coutp, io.N, io.H, io.W);
  a.xtiles = (io.W + 127) / 128;
  a.ytiles = (io.H + a.TH - 1) / a.TH;
  a.ntiles = a.xtiles * a.ytiles * io.N;
  a.wpack = a.sub ? l.d_wsub[a.sub_a * 2 + a.sub_b] : l.d_wpack;
  a.bias = l.d_bias;
  a.in_fp16 = l.fp16 ? 1 : 0;
  // algorithmic FLOPs of this launch: true channel counts, every output pixel, 9 taps (a sub-pixel phase launch
  // produces N x H x W of the 4 N H W output pixels of the upsample + conv it implements, with 4 issued taps each)
  const double fl = 2.0 * 9.0 * l.cin * l.cout * static_cast<double>(io.N) * io.H * io.W;
#if B200SR_HAVE_SC
  if (a.nchunks == 1 && !a.sub && e->opt_pair && l.cin == 64) {
    // product path of every single-chunk conv: resident weights, row-pair stages (conv3x3_sc.cuh)
    if (epi == EPI_LAST_U8 && e->opt_last9 && l.d_wlast9) {
      a.wpack = l.d_wlast9;
      a.TH = choose_th(e, 32, io.N, io.H, io.W, 126);
      a.xtiles = (io.W + 125) / 126;
      a.ytiles = (io.H + a.TH - 1) / a.TH;
      a.ntiles = a.xtiles * a.ytiles * io.N;
      return launch_sc_inst<32, EPI_LAST9_U8>(e, lane, amap, a, st, PC_CONV16_LAST, fl);
    }
    switch (l.coutp * 16 + epi) {
      case 64 * 16 + EPI_ACT_BF16: return launch_sc_inst<64, EPI_ACT_BF16>(e, lane, amap, a, st, PC_CONV64_ACT, fl);
      case 64 * 16 + EPI_PRELU_BF16: return launch_sc_inst<64, EPI_PRELU_BF16>(e, lane, amap, a, st, PC_CONV64_PRELU, fl);
      case 64 * 16 + EPI_ADD_F32: return launch_sc_inst<64, EPI_ADD_F32>(e, lane, amap, a, st, PC_CONV64_ADD, fl);
      case 16 * 16 + EPI_LAST_U8: return launch_sc_inst<16, EPI_LAST_U8>(e, lane, amap, a, st, PC_CONV16_LAST, fl);
      case 48 * 16 + EPI_SRVGG_LAST: return launch_sc_inst<48, EPI_SRVGG_LAST>(e, lane, amap, a, st, PC_CONV48_SRVGG_LAST, fl);
      default: break;   // (Cout = 32 single-chunk convs: RDB conv1 through the per-conv path, tests only)
    }
  }
#endif
  switch (l.coutp * 16 + epi) {
    case 32 * 16 + EPI_ACT_BF16: return launch_conv_inst<32, EPI_ACT_BF16>(e, lane, amap, a, st, PC_CONV32_ACT, fl);
    case 64 * 16 + EPI_ACT_BF16: return launch_conv_inst<64, EPI_ACT_BF16>(e, lane, amap, a, st, PC_CONV64_ACT, fl);
    case 64 * 16 + EPI_PRELU_BF16: return launch_conv_inst<64, EPI_PRELU_BF16>(e, lane, amap, a, st, PC_CONV64_PRELU, fl);
    case 64 * 16 + EPI_RDB5: return launch_conv_inst<64, EPI_RDB5>(e, lane, amap, a, st, PC_CONV64_RDB5, fl);
    case 64 * 16 + EPI_RDB5_RRDB: return launch_conv_inst<64, EPI_RDB5_RRDB>(e, lane, amap, a, st, PC_CONV64_RDB5_RRDB, fl);
    case 64 * 16 + EPI_ADD_F32: return launch_conv_inst<64, EPI_ADD_F32>(e, lane, amap, a, st, PC_CONV64_ADD, fl);
    case 16 * 16 + EPI_LAST_U8: return launch_conv_inst<16, EPI_LAST_U8>(e, lane, amap, a, st, PC_CONV16_LAST, fl);
    case 48 * 16 + EPI_SRVGG_LAST: return launch_conv_inst<48, EPI_SRVGG_LAST>(e, lane, amap, a, st, PC_CONV48_SRVGG_LAST, fl);
    default: return fail(e, B200SR_ERR_INVALID, "no kernel instance for this (Cout, epilogue)");
  }
}

#if B200SR_HAVE_SC
// conv_hr + conv_last as one rolling kernel (hr_last_fused.cuh).  `in` = conv_up2's output [N][H][W][64] fp16.
int launch_hr_last_fused(b200sr_engine* e, Lane* lane, const Layer& l_hr, const Layer& l_last, const void* in, int N, int H,
                         int W, const ConvArgs& base, cudaStream_t st) {
  CUtensorMap amap;
  if (!tmap_encode_act(&amap, in, N, H, W, 64, 64, 130, 128)) return fail(e, B200SR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  HrLastArgs a{};
  a.N = N;
  a.H = H;
  a.W = W;
  a.xtiles = (W + 125) / 126;
  // strips per column tile: balance the waves (a unit costs its rows + 4 halo rows)
  const int slots = e->opt_max_ctas > 0 ? e->opt_max_ctas : e->num_sms;
  double best = 1e30;
  int best_rows = (H + 1) / 2 * 2;
  for (int S = 1; S <= std::max(1, H / 16); ++S) {
    const int rows = ((H + S - 1) / S + 1) / 2 * 2;
    const long units = static_cast<long>(N) * a.xtiles * ((H + rows - 1) / rows);
    const double cost = static_cast<double>((units + slots - 1) / slots) * (rows + 4.0);
    if (cost < best - 1e-9) {
      best = cost;
      best_rows = rows;
    }
  }
  if (e->opt_force_th > 0) best_rows = std::max(2, (e->opt_force_th + 1) / 2 * 2);   // tests: ragged / tiny strips
  a.strip_rows = best_rows;
  a.strips = (H + best_rows - 1) / best_rows;
  a.nunits = N * a.xtiles * a.strips;
  a.w_hr = l_hr.d_wpack;
  a.w_last = l_last.d_wlast9;
  a.bias_hr = l_hr.d_bias;
  a.bias_last = l_last.d_bias;
  a.slope = 0.2f;
  a.dst = base.dst;
  a.dst16 = base.dst16;
  a.dst_h = base.dst_h;
  a.dst_w = base.dst_w;
  a.crop_y0 = base.crop_y0;
  a.crop_x0 = base.crop_x0;
  a.crop_h = base.crop_h;
  a.crop_w = base.crop_w;
  a.dst_y0 = base.dst_y0;
  a.dst_x0 = base.dst_x0;
  const double px = static_cast<double>(N) * H * W;
  ProfScope prof_scope(e, PC_HR_LAST, 2.0 * 9.0 * (64.0 * 64.0 + 64.0 * 3.0) * px, st);
  static bool attr_done[16] = {};
  if (!attr_done[e->device & 15]) {
    CUDA_TRY(e, cudaFuncSetAttribute(hr_last_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_BYTES));
    attr_done[e->device & 15] = true;
  }
  const int grid = std::min(a.nunits, slots);
  hr_last_fused_kernel<<<grid, HL_NTHREADS, HL_SMEM_BYTES, st>>>(amap, a);
  CUDA_TRY(e, cudaGetLastError());
  lane->launches++;
  return B200SR_OK;
}
#endif

// ---- workspace layout for one region of conv-domain size N x H x W ----
struct Region {
  // source frame + padding
  const uint8_t* src;
  uint8_t* dst;
  int sample16 = 0;            // frames hold uint16 samples (upstream's 16-bit image branch) instead of uint8
  int n, Hs, Ws, pre_pad, H1, W1;
  int oy, ox, rh, rw;          // region in padded-image coordinates
  int crop_y0, crop_x0, crop_h, crop_w, dst_y0, dst_x0, dst_h, dst_w;  // in network-output pixels
};

// The HR tail (conv_up1 .. conv_last) runs in groups of `tail_group` frames so that its 2x / 4x tensors (75 % of the
// activation bytes) are sized for ~one 720p frame instead of the whole batch; frames are independent, so the bytes
// produced do not depend on the grouping.
int tail_group(int N, int H, int W) {
  const long budget = 1280L * 720L;
  const long px = static_cast<long>(H) * W;
  return static_cast<int>(std::max(1L, std::min(static_cast<long>(N), budget / std::max(px, 1L))));
}

size_t region_ws_bytes(const b200sr_engine* e, int N, int H, int W) {
  const size_t px = static_cast<size_t>(N) * H * W;
  const size_t pxt = static_cast<size_t>(N) * H * ((W + 127) / 128) * 128;   // fp32 trunk: width padded to 128
  const size_t tpx = static_cast<size_t>(tail_group(N, H, W)) * H * W;
  size_t total = 0;
  auto add = [&](size_t b) { total += align_up(b, 1024); };
  if (e->desc.arch == B200SR_ARCH_RRDB) {
    add(px * 192 * 2);
    add(px * 192 * 2);
    add(px * 192 * 2);            // dense-block tensors P, Q, R (one per RDB of an RRDB)
    add(pxt * 64);
    add(pxt * 64);                // residual-stream lo bytes A, B
    add(pxt * 64 * 4);            // f0 (tile-interleaved fp32)
    add(px * 64 * 2);             // conv_body out
    add(tpx * 4 * 64 * 2);
    add(tpx * 4 * 64 * 2);        // 2x: upsampled (fold_up = 0 only), conv_up1 out
    add(tpx * 16 * 64 * 2);
    add(tpx * 16 * 64 * 2);       // 4x ping-pong
  } else {
    add(px * 64 * 2);
    add(px * 64 * 2);
    add(px * 4 * 4);
  }
  return total;
}

int ensure_ws(b200sr_engine* e, Lane* lane, size_t bytes, cudaStream_t st) {
  if (bytes <= lane->ws_bytes) return B200SR_OK;
  if (e->opt_ws_limit_mb > 0 && bytes > (static_cast<size_t>(e->opt_ws_limit_mb) << 20))
    return fail(e, B200SR_ERR_OOM, std::string("out of memory: ") + std::to_string(bytes >> 20) +
                                       " MiB workspace exceeds the configured limit (option ws_limit_mb)");
  if (lane->ws) {
    CUDA_TRY(e, cudaStreamSynchronize(st));
    if (lane->id <= 0) CUDA_TRY(e, cudaDeviceSynchronize());   // device lanes: earlier enqueues may sit on other streams
    cudaFree(lane->ws);
    lane->ws = nullptr;
    lane->ws_bytes = 0;
  }
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, bytes);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(e, B200SR_ERR_OOM, std::string("out of memory allocating ") + std::to_string(bytes >> 20) + " MiB workspace");
  }
  lane->ws = static_cast<uint8_t*>(p);
  lane->ws_bytes = bytes;
  return B200SR_OK;
}

// ---- fused RDB: item table (skewed row strips) ------------------------------------------------------
constexpr int RDB_STRIP = RDB_STRIP_ROWS;   // rows per strip (= rows per conv1..4 item; conv5 items have half)
int RDB_ORDER[5] = {0, 1, 2, 3, 4};      // order of the convs inside one step of the work list (option rdb_order)
int RDB_STEP_OFF[5] = {0, 1, 2, 3, 5};
int RDB_INTERLEAVE = -1;                 // option rdb_interleave (build_rdb_items)   // step in which conv k reaches strip s: s + RDB_STEP_OFF[k] (option rdb_off)

// Strip s of conv k covers rows [16 s - 8 k, 16 s + 16 - 8 k): every conv is shifted up by 8 rows relative to
// its predecessor, so the rows an item reads (its own +-1) of a lower conv belong to items earlier in the list.
void build_rdb_items(int N, int H, int W, std::vector<RdbItem>& items, int* nflags) {
  const int S = (H + 4 * RDB_SHIFT_ROWS + RDB_STRIP - 1) / RDB_STRIP;
  const int xt = (W + 127) / 128;
  const int nblk = (H + RDB_FLAG_ROWS - 1) / RDB_FLAG_ROWS;   // completion counters per conv and frame
  *nflags = N * 4 * nblk;
  items.clear();
  // Frames advance through the list in groups of G: inside a group all frames take step t together (G times more
  // independent items per step, i.e. G times more time between a producer item and its consumers), groups follow
  // each other.  The L2 working set grows with G x frame width, so G is chosen to keep ~64 items per step: 1 at
  // 720p (10 column tiles; interleaving measured 8 % slower there), 5 for 256-pixel-wide frames (+33 %), 2 for
  // 512-pixel tiles (+12 %).  Option rdb_interleave: -1 auto, 0 off, G > 0 explicit.
  const int G = RDB_INTERLEAVE < 0 ? std::max(1, std::min(N, (64 + 3 * xt) / (6 * xt)))
                                   : (RDB_INTERLEAVE == 0 ? 1 : std::min(N, RDB_INTERLEAVE));
  for (int g0 = 0; g0 < N; g0 += G)
    for (int t = 0; t < S + RDB_STEP_OFF[4]; ++t)
      for (int n = g0; n < std::min(N, g0 + G); ++n)
      for (int kk = 0; kk < 5; ++kk) {
        // within a step all groups are independent; the natural order conv1..conv5 measured best (rdb_sweep.py)
        const int k = RDB_ORDER[kk];
        // in step t conv k works on strip t - RDB_STEP_OFF[k]: its producer ran one or two steps (~60 items each)
        // earlier.  conv5 trails conv4 by two steps: conv4 items are the longest and finish their rows last,
        // conv5 items are short and reach their dependent chunk early.
        const int s = t - RDB_STEP_OFF[k];
        if (s < 0 || s >= S) continue;
        const int lo = std::max(0, s * RDB_STRIP - RDB_SHIFT_ROWS * k), hi = std::min(H, (s + 1) * RDB_STRIP - RDB_SHIFT_ROWS * k);
        if (lo >= hi) continue;
        const int th = k < 4 ? RDB_TH4 : RDB_TH5;
        for (int y0 = lo; y0 < hi; y0 += th)
          for (int tx = 0; tx < xt; ++tx) {
            RdbItem it{};
            it.k = k;
            it.n = n;
            it.y0 = y0;
            it.rows = std::min(th, hi - y0);
            it.tx = tx;
            it.flag_base = k < 4 ? (n * 4 + k) * nblk : -1;
            for (int c = 1; c <= 2; ++c) {
              // chunk c holds channels [64c, 64c+64): x1,x2 (c = 1) or x3,x4 (c = 2); newest producer below k
              const int dep = std::min(2 * c - 1, k - 1);
              const bool used = (k >= 1) && (c == 1 || k >= 3);
              it.dep_base[c - 1] = used ? (n * 4 + dep) * nblk : -1;
            }
            items.push_back(it);
          }
      }
}

int ensure_rdb_table(b200sr_engine* e, Lane* lane, int N, int H, int W, cudaStream_t st) {
  for (auto& t : lane->rdb_tables)
    if (t.n == N && t.h == H && t.w == W && t.gen == e->rdb_gen) {
      lane->d_rdb_items = t.d_items;
      lane->d_rdb_flags = t.d_flags;
      lane->rdb_nitems = t.nitems;
      lane->rdb_nflags = t.nflags;
      return B200SR_OK;
    }
  if (lane->rdb_tables.size() >= 24) {   // shapes come and go (variable frame sizes): drop the oldest table
    CUDA_TRY(e, cudaStreamSynchronize(st));
    cudaFree(lane->rdb_tables.front().d_items);
    cudaFree(lane->rdb_tables.front().d_flags);
    lane->rdb_tables.erase(lane->rdb_tables.begin());
  }
  std::vector<RdbItem> items;
  int nflags = 0;
  build_rdb_items(N, H, W, items, &nflags);
  Lane::RdbTable t;
  CUDA_TRY(e, cudaMalloc(&t.d_items, items.size() * sizeof(RdbItem)));
  cudaError_t err = cudaMalloc(&t.d_flags, static_cast<size_t>(nflags + 1) * sizeof(int));   // + item counter
  if (err != cudaSuccess) {
    cudaFree(t.d_items);
    CUDA_TRY(e, err);
  }
  // pageable source: the copy is staged before the call returns, so `items` may go out of scope
  err = cudaMemcpyAsync(t.d_items, items.data(), items.size() * sizeof(RdbItem), cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    cudaFree(t.d_items);
    cudaFree(t.d_flags);
    CUDA_TRY(e, err);
  }
  t.n = N;
  t.h = H;
  t.w = W;
  t.gen = e->rdb_gen;
  t.nitems = static_cast<int>(items.size());
  t.nflags = nflags;
  lane->rdb_tables.push_back(t);
  lane->d_rdb_items = t.d_items;
  lane->d_rdb_flags = t.d_flags;
  lane->rdb_nitems = t.nitems;
  lane->rdb_nflags = t.nflags;
  return B200SR_OK;
}

// One RDB (layers li .. li+4) as one persistent kernel.
struct TrunkIO {   // residual-stream operands of one RDB's conv5 (TrunkLo, conv3x3_tc.cuh)
  const uint8_t* lo_in;
  uint8_t* lo_out;
  const __nv_bfloat16* xb_hi;
  const uint8_t* xb_lo;
};

int launch_rdb_fused(b200sr_engine* e, Lane* lane, int li, bool rrdb_end, __nv_bfloat16* Dcur, __nv_bfloat16* Dnext,
                     const TrunkIO& tio, int N, int H, int W, cudaStream_t st) {
  int rc = ensure_rdb_table(e, lane, N, H, W, st);
  if (rc) return rc;
  CUtensorMap amap;
  if (!tmap_encode_act(&amap, Dcur, 3 * N, H, W, 64, 64, 130, 128)) return fail(e, B200SR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  CUtensorMap amap_h;   // the first 32 channels of a plane as 64-byte rows (half chunks of conv2 / conv4)
  if (!tmap_encode_act(&amap_h, Dcur, 3 * N, H, W, 64, 32, 130, 64)) return fail(e, B200SR_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  const size_t plane = static_cast<size_t>(N) * H * W * 64;   // elements per chunk plane
  RdbArgs a{};
  double flops = 0;
  for (int k = 0; k < 5; ++k) {
    const Layer& l = e->layers[li + k];
    ConvArgs& c = a.L[k];
    c.N = N;
    c.H = H;
    c.W = W;
    c.nchunks = (l.cin + 63) / 64;
    c.last_ksteps = (l.cin % 64 == 32) ? 2 : 4;
    c.wpack = (e->opt_half64 && l.d_wrdb) ? l.d_wrdb : l.d_wpack;
    c.bias = l.d_bias;
    c.slope = 0.2f;
    c.in_planes = 3;
    if (k < 4) {
      c.out = Dcur + (1 + k / 2) * plane;
      c.out_pitch = 64;
      c.out_choff = 32 * (k & 1);
    } else {
      c.out = Dnext;
      c.out_pitch = 64;
      c.out_choff = 0;
      c.hi_in = Dcur;
      c.lo_in = tio.lo_in;
      c.lo_out = tio.lo_out;
      c.xb_hi = tio.xb_hi;
      c.xb_lo = tio.xb_lo;
    }
    flops += 2.0 * 9.0 * l.cin * l.cout * static_cast<double>(N) * H * W;
  }
  a.items = lane->d_rdb_items;
  a.nitems = lane->rdb_nitems;
  a.flags = lane->d_rdb_flags;
  a.counter = lane->d_rdb_flags + lane->rdb_nflags;
  a.nflags = lane->rdb_nflags;
  a.flag_target = ((W + 127) / 128) * RDB_NEPI_WARPS;
  a.rrdb_end = rrdb_end ? 1 : 0;
  a.half64 = e->opt_half64 ? 1 : 0;
  ++lane->rdb_launch_idx;
  if (lane->rdb_launch_idx == -e->opt_rdb_stats) {   // negative: cycle counters only (no per-item / per-row stamps)
    if (!e->d_rdb_stats) CUDA_TRY(e, cudaMalloc(&e->d_rdb_stats, 296 * 16 * sizeof(long long)));
    a.stats = e->d_rdb_stats;
  }
  if (e->opt_rdb_stats > 0 && lane->rdb_launch_idx == e->opt_rdb_stats) {   // dev tool, single-threaded use only
    if (!e->d_rdb_stats) CUDA_TRY(e, cudaMalloc(&e->d_rdb_stats, 296 * 16 * sizeof(long long)));
    a.stats = e->d_rdb_stats;
    if (e->stats_nitems != lane->rdb_nitems) {
      if (e->d_rdb_trace) cudaFree(e->d_rdb_trace);
      if (e->d_rdb_trace2) cudaFree(e->d_rdb_trace2);
      e->d_rdb_trace = e->d_rdb_trace2 = nullptr;
      e->stats_nitems = lane->rdb_nitems;
    }
    if (!e->d_rdb_trace) CUDA_TRY(e, cudaMalloc(&e->d_rdb_trace, static_cast<size_t>(lane->rdb_nitems) * 10 * sizeof(long long)));
    CUDA_TRY(e, cudaMemsetAsync(e->d_rdb_trace, 0, static_cast<size_t>(lane->rdb_nitems) * 10 * sizeof(long long), st));
    a.trace = e->d_rdb_trace;
    if (!e->d_rdb_trace2) CUDA_TRY(e, cudaMalloc(&e->d_rdb_trace2, static_cast<size_t>(lane->rdb_nitems) * 48 * sizeof(long long)));
    CUDA_TRY(e, cudaMemsetAsync(e->d_rdb_trace2, 0, static_cast<size_t>(lane->rdb_nitems) * 48 * sizeof(long long), st));
    a.trace2 = e->d_rdb_trace2;
  }
  // conv5's epilogue is specialised at compile time (rdb_fused.cuh, MODE): RRDB end | lo part read | lo part written
  const int mode = (rrdb_end ? RDB_MODE_RRDB_END : 0) | (a.L[4].lo_in != nullptr ? RDB_MODE_LO_IN : 0) |
                   (a.L[4].lo_out != nullptr ? RDB_MODE_LO_OUT : 0);
  using RdbKernel = void (*)(const CUtensorMap, const CUtensorMap, const RdbArgs);
  RdbKernel kern = nullptr;
  switch (mode) {   // the combinations option trunk_lo = 0 / 1 / 2 produces
    case 0: kern = rdb_fused_kernel<0>; break;
    case 2: kern = rdb_fused_kernel<2>; break;
    case 5: kern = rdb_fused_kernel<5>; break;
    case 6: kern = rdb_fused_kernel<6>; break;
    case 7: kern = rdb_fused_kernel<7>; break;
    default: return fail(e, B200SR_ERR_STATE, "fused RDB: no kernel for residual-stream mode " + std::to_string(mode));
  }
  static bool attr_done[16][8] = {};
  if (!attr_done[e->device & 15][mode]) {
    CUDA_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RDB_SMEM_BYTES));
    attr_done[e->device & 15][mode] = true;
  }
  ProfScope prof_scope(e, PC_RDB_FUSED, flops, st);
  CUDA_TRY(e, cudaMemsetAsync(lane->d_rdb_flags, 0, static_cast<size_t>(lane->rdb_nflags + 1) * sizeof(int), st));
  const int grid = std::min(a.nitems, e->opt_max_ctas > 0 ? e->opt_max_ctas : e->num_sms * RDB_CTAS);
  kern<<<grid, RDB_NTHREADS, RDB_SMEM_BYTES, st>>>(amap, amap_h, a);
  CUDA_TRY(e, cudaGetLastError());
  lane->launches++;
  return B200SR_OK;
}

int run_first(b200sr_engine* e, Lane* lane, const Region& R, int s, int H, int W, __nv_bfloat16* out, int out_pitch, int out_fp16,
              uint8_t* lo, float* f0, float* inrgb, const float* prelu, cudaStream_t st) {
  const Layer& l = e->layers[0];
  FirstArgs a{};
  a.src = R.src;
  a.src16 = R.sample16;
  a.N = R.n;
  a.Hs = R.Hs;
  a.Ws = R.Ws;
  a.H1 = R.H1;
  a.W1 = R.W1;
  a.oy = R.oy;
  a.ox = R.ox;
  a.s = s;
  a.H = H;
  a.W = W;
  a.cin = l.cin;
  a.w = l.d_wfirst;
  a.bias = l.d_bias;
  a.prelu = prelu;
  a.out = out;
  a.out_pitch = out_pitch;
  a.out_fp16 = out_fp16;
  a.lo = lo;
  a.f0 = f0;
  a.inrgb = inrgb;
  ProfScope prof_scope(e, PC_FIRST, 2.0 * 9.0 * l.cin * 64 * static_cast<double>(R.n) * H * W, st);
  if (e->opt_first_v1 == 1) {   // one thread per pixel, weights in shared memory (round-1 kernel; bit-equality test)
    dim3 grid((W + 127) / 128, H, R.n);
    if (l.cin == 3) {
      const size_t sm = 9 * 3 * 64 * sizeof(float);
      first_conv_kernel<3><<<grid, 128, sm, st>>>(a);
    } else {
      const size_t sm = 9 * 12 * 64 * sizeof(float);
      first_conv_kernel<12><<<grid, 128, sm, st>>>(a);
    }
  } else if (l.cin == 3 && e->opt_first_v1 == 0) {
    // weights, bias and PReLU slopes as a kernel parameter: every FFMA reads its weight from the constant bank
    FirstWeights3 cw;
    memcpy(cw.w, l.wfirst.data(), sizeof(cw.w));
    for (int c = 0; c < 64; ++c) {
      cw.b[c] = l.b[c];
      cw.p[c] = prelu ? e->prelu_host[0][c] : 1.f;
    }
    cw.has_prelu = prelu ? 1 : 0;
    first_conv3_const_kernel<<<dim3((W + 127) / 128, H, R.n), 128, 0, st>>>(a, cw);   // one pixel per thread
  } else if (l.cin == 3) {
    using T = FirstTiled<3>;
    dim3 grid((W + 127) / 128, (H + T::ROWS - 1) / T::ROWS, R.n);
    first_conv_tiled_kernel<3><<<grid, 256, T::SMEM_BYTES, st>>>(a);
  } else {
    using T = FirstTiled<12>;
    static bool attr_done[16] = {};
    if (!attr_done[e->device & 15]) {
      CUDA_TRY(e, cudaFuncSetAttribute(first_conv_tiled_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
      attr_done[e->device & 15] = true;
    }
    dim3 grid((W + 127) / 128, (H + T::ROWS - 1) / T::ROWS, R.n);
    first_conv_tiled_kernel<12><<<grid, 256, T::SMEM_BYTES, st>>>(a);
  }
  CUDA_TRY(e, cudaGetLastError());
  lane->launches++;
  return B200SR_OK;
}

int run_upsample(b200sr_engine* e, Lane* lane, const void* in, void* out, int N, int H, int W, cudaStream_t st) {
  const size_t total = static_cast<size_t>(N) * 2 * H * 2 * W * 8;
  const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(e->num_sms) * 16));
  ProfScope prof_scope(e, PC_UPSAMPLE, 0.0, st);
  upsample2x_kernel<<<grid, 256, 0, st>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), N, H, W);
  CUDA_TRY(e, cudaGetLastError());
  lane->launches++;
  return B200SR_OK;
}

// Regions of one `enhance` call: the whole padded frame, or upstream RealESRGANer.tile_process's tiles
// over the padded image.  Geometry only (src/dst/n are filled in by the caller).
std::vector<Region> plan_regions(int arch, int scale, int h, int w, int tile, int tile_pad, int pre_pad) {
  std::vector<Region> out;
  const int mod = (arch == B200SR_ARCH_RRDB && scale == 2) ? 2 : 1;
  const int H1 = h + pre_pad, W1 = w + pre_pad;
  const int Hp = (H1 + mod - 1) / mod * mod, Wp = (W1 + mod - 1) / mod * mod;
  Region R{};
  R.Hs = h;
  R.Ws = w;
  R.pre_pad = pre_pad;
  R.H1 = H1;
  R.W1 = W1;
  R.dst_h = h * scale;
  R.dst_w = w * scale;
  if (tile <= 0) {
    R.oy = 0;
    R.ox = 0;
    R.rh = Hp;
    R.rw = Wp;
    R.crop_y0 = 0;
    R.crop_x0 = 0;
    R.crop_h = h * scale;
    R.crop_w = w * scale;
    R.dst_y0 = 0;
    R.dst_x0 = 0;
    out.push_back(R);
    return out;
  }
  const int tiles_x = (Wp + tile - 1) / tile, tiles_y = (Hp + tile - 1) / tile;
  for (int ty = 0; ty < tiles_y; ++ty)
    for (int tx = 0; tx < tiles_x; ++tx) {
      const int in_x0 = tx * tile, in_y0 = ty * tile;
      const int in_x1 = std::min(in_x0 + tile, Wp), in_y1 = std::min(in_y0 + tile, Hp);
      const int pad_x0 = std::max(in_x0 - tile_pad, 0), pad_x1 = std::min(in_x1 + tile_pad, Wp);
      const int pad_y0 = std::max(in_y0 - tile_pad, 0), pad_y1 = std::min(in_y1 + tile_pad, Hp);
      R.oy = pad_y0;
      R.ox = pad_x0;
      R.rh = pad_y1 - pad_y0;
      R.rw = pad_x1 - pad_x0;
      R.crop_y0 = (in_y0 - pad_y0) * scale;
      R.crop_x0 = (in_x0 - pad_x0) * scale;
      R.dst_y0 = in_y0 * scale;
      R.dst_x0 = in_x0 * scale;
      R.crop_h = std::min((in_y1 - in_y0) * scale, R.dst_h - R.dst_y0);  // post_process crops the pads
      R.crop_w = std::min((in_x1 - in_x0) * scale, R.dst_w - R.dst_x0);
      if (R.crop_h <= 0 || R.crop_w <= 0) continue;
      out.push_back(R);
    }
  return out;
}

// One forward pass over one region (whole padded frame, or one padded tile).
int run_region(b200sr_engine* e, Lane* lane, const Region& R, cudaStream_t st) {
  const auto& d = e->desc;
  const int s = (d.arch == B200SR_ARCH_RRDB && d.scale == 2) ? 2 : 1;
  if (R.rh % s || R.rw % s) return fail(e, B200SR_ERR_INVALID, "region size not divisible by the unshuffle factor");
  const int N = R.n, H = R.rh / s, W = R.rw / s;
  const size_t px = static_cast<size_t>(N) * H * W;
  int rc = ensure_ws(e, lane, region_ws_bytes(e, N, H, W), st);
  if (rc) return rc;
  uint8_t* cur = lane->ws;
  auto take = [&](size_t b) {
    uint8_t* p = cur;
    cur += align_up(b, 1024);
    return p;
  };
  ConvArgs base{};
  base.slope = 1.f;
  base.dst = R.dst;
  base.dst16 = R.sample16;
  base.dst_h = R.dst_h;
  base.dst_w = R.dst_w;
  base.crop_y0 = R.crop_y0;
  base.crop_x0 = R.crop_x0;
  base.crop_h = R.crop_h;
  base.crop_w = R.crop_w;
  base.dst_y0 = R.dst_y0;
  base.dst_x0 = R.dst_x0;

  if (d.arch == B200SR_ARCH_RRDB) {
    // Dense-block tensors, chunk-planar: plane 0 = x.hi (64 ch), plane 1 = conv1|conv2 outputs, plane 2 =
    // conv3|conv4 outputs, each [N][H][W][64] (128 B per pixel, pixels contiguous: a TMA row box is one contiguous
    // 16.6 KB read and DRAM pages are used whole; with one 384 B-pitch NHWC tensor every box row was a separate
    // 128 B access).  RDB r of every RRDB reads D[r] and writes the next x.hi into plane 0 of D[(r + 1) % 3]; D[0]
    // therefore still holds the RRDB input x0.hi when the third RDB needs it, and is overwritten pixel by pixel
    // by the thread that has just read it.
    const int TG = tail_group(N, H, W);
    const size_t fpx = static_cast<size_t>(H) * W;          // pixels per frame
    const size_t tpx = static_cast<size_t>(TG) * fpx;
    __nv_bfloat16* D[3];
    for (int i = 0; i < 3; ++i) D[i] = reinterpret_cast<__nv_bfloat16*>(take(px * 192 * 2));
    const size_t pxt = static_cast<size_t>(N) * H * ((W + 127) / 128) * 128;
    uint8_t* loA = take(pxt * 64);   // x.lo inside an RRDB
    uint8_t* loB = take(pxt * 64);   // x0.lo (RRDB input), rewritten in place at the RRDB end
    float* f0 = reinterpret_cast<float*>(take(pxt * 64 * 4));
    __nv_bfloat16* U0 = reinterpret_cast<__nv_bfloat16*>(take(px * 64 * 2));
    __nv_bfloat16* U1 = reinterpret_cast<__nv_bfloat16*>(take(tpx * 4 * 64 * 2));
    __nv_bfloat16* U2 = reinterpret_cast<__nv_bfloat16*>(take(tpx * 4 * 64 * 2));
    __nv_bfloat16* U3 = reinterpret_cast<__nv_bfloat16*>(take(tpx * 16 * 64 * 2));
    __nv_bfloat16* U4 = reinterpret_cast<__nv_bfloat16*>(take(tpx * 16 * 64 * 2));

    const size_t plane = px * 64;
    rc = run_first(e, lane, R, s, H, W, D[0], 64, 0, loB, f0, nullptr, nullptr, st);
    if (rc) return rc;
    int li = 1;
    const int cur_d = 0;   // the trunk ends where it started
    for (int b = 0; b < d.num_block; ++b)
      for (int r = 0; r < 3; ++r) {
        __nv_bfloat16* Din = D[r];
        __nv_bfloat16* Dout = D[(r + 1) % 3];
        TrunkIO tio;
        // the pair lives at the RRDB boundaries (loB); inside an RRDB the stream is hi alone unless trunk_lo = 2
        tio.lo_in = r == 0 ? (e->opt_trunk_lo >= 1 ? loB : nullptr) : (e->opt_trunk_lo >= 2 ? loA : nullptr);
        tio.lo_out = r == 2 ? loB : (e->opt_trunk_lo >= 2 ? loA : nullptr);
        tio.xb_hi = D[0];
        tio.xb_lo = loB;
        if (e->opt_fused_rdb) {
          rc = launch_rdb_fused(e, lane, li, r == 2, Din, Dout, tio, N, H, W, st);
          if (rc) return rc;
          li += 5;
        } else {
          ConvIO io{Din, 64, N, H, W, 3};
          for (int k = 0; k < 4; ++k) {
            ConvArgs a = base;
            a.slope = 0.2f;
            a.out = Din + (1 + k / 2) * plane;
            a.out_pitch = 64;
            a.out_choff = 32 * (k & 1);
            rc = launch_conv(e, lane, e->layers[li++], EPI_ACT_BF16, io, a, st);
            if (rc) return rc;
          }
          ConvArgs a = base;
          a.out = Dout;
          a.out_pitch = 64;
          a.out_choff = 0;
          a.hi_in = Din;
          a.lo_in = tio.lo_in;
          a.lo_out = tio.lo_out;
          a.xb_hi = tio.xb_hi;
          a.xb_lo = tio.xb_lo;
          rc = launch_conv(e, lane, e->layers[li++], r == 2 ? EPI_RDB5_RRDB : EPI_RDB5, io, a, st);
          if (rc) return rc;
        }
      }
    {  // conv_body: feat + conv_body(body(feat))
      ConvIO io{D[cur_d], 64, N, H, W};
      ConvArgs a = base;
      a.out = U0;
      a.out_pitch = 64;
      a.out_fp16 = 1;
      a.fadd = f0;
      rc = launch_conv(e, lane, e->layers[li++], EPI_ADD_F32, io, a, st);
      if (rc) return rc;
    }
    // HR tail, TG frames at a time.  conv_up1/conv_up2 read F.interpolate(x, 2, 'nearest') of their input straight
    // from the low-resolution tensor through the duplicated-pixel TMA view (option fold_up = 0: materialise it
    // with upsample2x_kernel).
    const Layer& l_up1 = e->layers[li];
    const Layer& l_up2 = e->layers[li + 1];
    const Layer& l_hr = e->layers[li + 2];
    const Layer& l_last = e->layers[li + 3];
    const size_t dst_frame = static_cast<size_t>(R.dst_h) * R.dst_w * 3 * (R.sample16 ? 2 : 1);
    for (int f0i = 0; f0i < N; f0i += TG) {
      const int n = std::min(TG, N - f0i);
      const __nv_bfloat16* U0f = U0 + static_cast<size_t>(f0i) * fpx * 64;
      ConvArgs tb = base;
      tb.dst = R.dst + static_cast<size_t>(f0i) * dst_frame;
      if (!e->opt_fold_up) {
        rc = run_upsample(e, lane, U0f, U1, n, H, W, st);
        if (rc) return rc;
      }
      // conv_up1 / conv_up2 (+ lrelu).  fold_up 2: four sub-pixel phase launches on the low-resolution grid (2.25x
      // fewer MACs: the taps that read the same source pixel are pre-summed); 1: one launch through the
      // duplicated-pixel TMA view; 0: on the materialised upsampling.
      auto conv_up = [&](const Layer& l, const void* lr, const void* up, __nv_bfloat16* dstt, int h, int w) -> int {
        ConvArgs a = tb;
        a.slope = 0.2f;
        a.out = dstt;
        a.out_pitch = 64;
        a.out_fp16 = 1;
        if (e->opt_fold_up == 2) {
          ConvIO io{lr, 64, n, h, w};
          for (int ph = 0; ph < 4; ++ph) {
            a.sub = 1;
            a.sub_a = ph >> 1;
            a.sub_b = ph & 1;
            a.sub_vlo = a.sub_a ? 0 : 1;
            a.sub_vhi = a.sub_a ? 1 : 2;
            a.sub_dlo = a.sub_b ? 1 : 0;
            a.sub_dhi = a.sub_b ? 2 : 1;
            int r = launch_conv(e, lane, l, EPI_ACT_BF16, io, a, st);
            if (r) return r;
          }
          return B200SR_OK;
        }
        ConvIO io{e->opt_fold_up ? lr : up, 64, n, 2 * h, 2 * w, 0, e->opt_fold_up ? 1 : 0};
        return launch_conv(e, lane, l, EPI_ACT_BF16, io, a, st);
      };
      rc = conv_up(l_up1, U0f, U1, U2, H, W);
      if (rc) return rc;
      if (!e->opt_fold_up) {
        rc = run_upsample(e, lane, U2, U3, n, 2 * H, 2 * W, st);
        if (rc) return rc;
      }
      rc = conv_up(l_up2, U2, U3, U4, 2 * H, 2 * W);
      if (rc) return rc;
#if B200SR_HAVE_SC
      if (e->opt_fuse_tail && e->opt_pair && e->opt_last9 && l_last.d_wlast9) {
        // conv_hr + lrelu + conv_last + clamp/round/quantise + crop in one kernel: the 4x tensor between them stays on chip
        rc = launch_hr_last_fused(e, lane, l_hr, l_last, U4, n, 4 * H, 4 * W, tb, st);
        if (rc) return rc;
        continue;
      }
#endif
      {  // conv_hr + lrelu
        ConvIO io{U4, 64, n, 4 * H, 4 * W};
        ConvArgs a = tb;
        a.slope = 0.2f;
        a.out = U3;
        a.out_pitch = 64;
        a.out_fp16 = 1;
        rc = launch_conv(e, lane, l_hr, EPI_ACT_BF16, io, a, st);
        if (rc) return rc;
      }
      {  // conv_last + clamp/round/quantise + crop
        ConvIO io{U3, 64, n, 4 * H, 4 * W};
        rc = launch_conv(e, lane, l_last, EPI_LAST_U8, io, tb, st);
        if (rc) return rc;
      }
    }
  } else {
    __nv_bfloat16* S[2];
    S[0] = reinterpret_cast<__nv_bfloat16*>(take(px * 64 * 2));
    S[1] = reinterpret_cast<__nv_bfloat16*>(take(px * 64 * 2));
    float* inrgb = reinterpret_cast<float*>(take(px * 4 * 4));
    rc = run_first(e, lane, R, 1, H, W, S[0], 64, 1, nullptr, nullptr, inrgb, e->prelu_dev[0], st);
    if (rc) return rc;
    int cur_s = 0;
    for (int i = 0; i < d.num_block; ++i) {
      ConvIO io{S[cur_s], 64, N, H, W};
      ConvArgs a = base;
      a.out = S[cur_s ^ 1];
      a.out_pitch = 64;
      a.out_fp16 = 1;
      a.prelu = e->prelu_dev[i + 1];
      rc = launch_conv(e, lane, e->layers[1 + i], EPI_PRELU_BF16, io, a, st);
      if (rc) return rc;
      cur_s ^= 1;
    }
    ConvIO io{S[cur_s], 64, N, H, W};
    ConvArgs a = base;
    a.fadd = inrgb;
    rc = launch_conv(e, lane, e->layers[1 + d.num_block], EPI_SRVGG_LAST, io, a, st);
    if (rc) return rc;
  }
  return B200SR_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

const char* b200sr_version(void) {
#ifdef B200SR_DEBUG
  return "b200sr 0.2 (sm_100a, tcgen05/TMEM/TMA) DEBUG: bounds traps on";
#else
  return RDB_CTAS == 2 ? "b200sr 0.2 (sm_100a, tcgen05/TMEM/TMA) fused RDB: 2 CTAs per SM"
                       : "b200sr 0.2 (sm_100a, tcgen05/TMEM/TMA)";
#endif
}

// Message of the last failed call made by THIS thread (engines are shared between threads); falls back to the
// engine's most recent message.
const char* b200sr_last_error(const b200sr_engine* e) {
  if (!e) return "null engine";
  if (tls_err.empty()) {
    std::lock_guard<std::mutex> g(const_cast<b200sr_engine*>(e)->mu);
    tls_err = e->err;
  }
  return tls_err.c_str();
}

int b200sr_create(const b200sr_model_desc* desc, int device, b200sr_engine** out) {
  if (!desc || !out) return B200SR_ERR_INVALID;
  *out = nullptr;
  if (desc->num_feat != 64 || desc->num_grow_ch != 32 || desc->num_block < 1) return B200SR_ERR_INVALID;
  if (desc->arch == B200SR_ARCH_RRDB) {
    if (desc->scale != 4 && desc->scale != 2) return B200SR_ERR_INVALID;
  } else if (desc->arch == B200SR_ARCH_SRVGG) {
    if (desc->scale != 4) return B200SR_ERR_INVALID;
  } else {
    return B200SR_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return B200SR_ERR_CUDA;  // no CUDA device: the product path has no CPU fallback
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200SR_ERR_CUDA;
  if (prop.major != 10) return B200SR_ERR_CUDA;  // kernels are sm_100a only
  auto* e = new b200sr_engine();
  e->desc = *desc;
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  build_layers(e);
  if (cudaSetDevice(device) != cudaSuccess || !tmap_init()) {
    delete e;
    return B200SR_ERR_CUDA;
  }
  *out = e;
  return B200SR_OK;
}

static void free_lane(Lane* l) {
  if (l->ws) cudaFree(l->ws);
  for (auto& t : l->rdb_tables) {
    if (t.d_items) cudaFree(t.d_items);
    if (t.d_flags) cudaFree(t.d_flags);
  }
  if (l->dev_in) cudaFree(l->dev_in);
  if (l->dev_out) cudaFree(l->dev_out);
  if (l->pin_in) cudaFreeHost(l->pin_in);
  if (l->pin_out) cudaFreeHost(l->pin_out);
  if (l->stream) cudaStreamDestroy(l->stream);
  *l = Lane{};
}

void b200sr_destroy(b200sr_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& l : e->layers) {
    if (l.d_wpack) cudaFree(l.d_wpack);
    for (auto* p : l.d_wsub)
      if (p) cudaFree(p);
    if (l.d_wfirst) cudaFree(l.d_wfirst);
    if (l.d_wlast9) cudaFree(l.d_wlast9);
    if (l.d_wrdb) cudaFree(l.d_wrdb);
    if (l.d_bias) cudaFree(l.d_bias);
  }
  for (auto* p : e->prelu_dev)
    if (p) cudaFree(p);
  free_lane(&e->dev_lane);
  for (auto& p : e->dev_lanes) {
    p.second->stream = nullptr;   // the caller's stream, not ours
    free_lane(p.second.get());
  }
  for (auto& l : e->lanes) free_lane(l.get());
  if (e->d_rdb_stats) cudaFree(e->d_rdb_stats);
  if (e->d_rdb_trace) cudaFree(e->d_rdb_trace);
  if (e->d_rdb_trace2) cudaFree(e->d_rdb_trace2);
  for (auto& r : e->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  delete e;
}

int b200sr_num_convs(const b200sr_engine* e) { return e ? static_cast<int>(e->layers.size()) : 0; }
int b200sr_num_prelus(const b200sr_engine* e) { return e ? static_cast<int>(e->prelu_host.size()) : 0; }

int b200sr_set_conv(b200sr_engine* e, int layer, const float* weight, const float* bias, int cout, int cin) {
  if (!e || !weight || !bias) return B200SR_ERR_INVALID;
  if (layer < 0 || layer >= static_cast<int>(e->layers.size())) return fail(e, B200SR_ERR_INVALID, "layer index out of range");
  Layer& l = e->layers[layer];
  if (l.cin != cin || l.cout != cout)
    return fail(e, B200SR_ERR_INVALID, "conv " + std::to_string(layer) + ": expected [" + std::to_string(l.cout) + "," +
                                           std::to_string(l.cin) + ",3,3], got [" + std::to_string(cout) + "," +
                                           std::to_string(cin) + ",3,3]");
  l.w.assign(weight, weight + static_cast<size_t>(cout) * cin * 9);
  l.b.assign(bias, bias + cout);
  l.set = true;
  e->finalized = false;
  return B200SR_OK;
}

int b200sr_set_prelu(b200sr_engine* e, int index, const float* slope, int n) {
  if (!e || !slope) return B200SR_ERR_INVALID;
  if (index < 0 || index >= static_cast<int>(e->prelu_host.size()) || n != e->desc.num_feat)
    return fail(e, B200SR_ERR_INVALID, "bad prelu index / size");
  e->prelu_host[index].assign(slope, slope + n);
  e->finalized = false;
  return B200SR_OK;
}

int b200sr_finalize(b200sr_engine* e) {
  if (!e) return B200SR_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  for (size_t i = 0; i < e->layers.size(); ++i) {
    Layer& l = e->layers[i];
    if (!l.set) return fail(e, B200SR_ERR_STATE, "conv " + std::to_string(i) + " has no weights");
    std::vector<float> bias(l.coutp, 0.f);
    std::copy(l.b.begin(), l.b.end(), bias.begin());
    if (!l.d_bias) CUDA_TRY(e, cudaMalloc(&l.d_bias, bias.size() * 4));
    CUDA_TRY(e, cudaMemcpy(l.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    if (i == 0) {
      std::vector<float> wf(static_cast<size_t>(9) * l.cin * 64);
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < l.cin; ++ci)
          for (int t = 0; t < 9; ++t) wf[(static_cast<size_t>(t) * l.cin + ci) * 64 + co] = l.w[(static_cast<size_t>(co) * l.cin + ci) * 9 + t];
      if (!l.d_wfirst) CUDA_TRY(e, cudaMalloc(&l.d_wfirst, wf.size() * 4));
      CUDA_TRY(e, cudaMemcpy(l.d_wfirst, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice));
      l.wfirst = wf;
    } else {
      std::vector<uint8_t> img = pack_weights(l);
      if (!l.d_wpack) CUDA_TRY(e, cudaMalloc(&l.d_wpack, img.size()));
      CUDA_TRY(e, cudaMemcpy(l.d_wpack, img.data(), img.size(), cudaMemcpyHostToDevice));
      const size_t nl = e->layers.size();
      if (e->desc.arch == B200SR_ARCH_RRDB && l.cin % 64 == 32 && l.cin > 64) {   // RDB conv2 / conv4
        std::vector<uint8_t> himg = pack_weights_rdb_half(l);
        if (!l.d_wrdb) CUDA_TRY(e, cudaMalloc(&l.d_wrdb, himg.size()));
        CUDA_TRY(e, cudaMemcpy(l.d_wrdb, himg.data(), himg.size(), cudaMemcpyHostToDevice));
      }
      if (e->desc.arch == B200SR_ARCH_RRDB && i == nl - 1) {   // conv_last: stacked-kx image
        std::vector<uint8_t> simg = pack_weights_last9(l);
        if (!l.d_wlast9) CUDA_TRY(e, cudaMalloc(&l.d_wlast9, simg.size()));
        CUDA_TRY(e, cudaMemcpy(l.d_wlast9, simg.data(), simg.size(), cudaMemcpyHostToDevice));
      }
      if (e->desc.arch == B200SR_ARCH_RRDB && (i == nl - 4 || i == nl - 3)) {   // conv_up1, conv_up2
        for (int ph = 0; ph < 4; ++ph) {
          std::vector<uint8_t> simg = pack_subpixel_weights(l, ph >> 1, ph & 1);
          if (!l.d_wsub[ph]) CUDA_TRY(e, cudaMalloc(&l.d_wsub[ph], simg.size()));
          CUDA_TRY(e, cudaMemcpy(l.d_wsub[ph], simg.data(), simg.size(), cudaMemcpyHostToDevice));
        }
      }
    }
  }
  for (size_t i = 0; i < e->prelu_host.size(); ++i) {
    if (e->prelu_host[i].empty()) return fail(e, B200SR_ERR_STATE, "prelu " + std::to_string(i) + " has no weights");
    if (!e->prelu_dev[i]) CUDA_TRY(e, cudaMalloc(&e->prelu_dev[i], 64 * 4));
    CUDA_TRY(e, cudaMemcpy(e->prelu_dev[i], e->prelu_host[i].data(), 64 * 4, cudaMemcpyHostToDevice));
  }
  e->finalized = true;
  return B200SR_OK;
}

int b200sr_output_dims(const b200sr_engine* e, int h, int w, int* out_h, int* out_w) {
  if (!e || !out_h || !out_w || h <= 0 || w <= 0) return B200SR_ERR_INVALID;
  *out_h = h * e->desc.scale;
  *out_w = w * e->desc.scale;
  return B200SR_OK;
}

static void padded_dims(const b200sr_engine* e, int h, int w, int pre_pad, int* Hp, int* Wp) {
  int H1 = h + pre_pad, W1 = w + pre_pad;
  const int mod = (e->desc.arch == B200SR_ARCH_RRDB && e->desc.scale == 2) ? 2 : 1;
  *Hp = (H1 + mod - 1) / mod * mod;
  *Wp = (W1 + mod - 1) / mod * mod;
}

int b200sr_workspace_bytes(b200sr_engine* e, int n, int h, int w, int tile, int tile_pad, int pre_pad, size_t* bytes) {
  if (!e || !bytes || n <= 0 || h <= 0 || w <= 0 || tile < 0 || tile_pad < 0 || pre_pad < 0) return B200SR_ERR_INVALID;
  int Hp, Wp;
  padded_dims(e, h, w, pre_pad, &Hp, &Wp);
  const int s = (e->desc.arch == B200SR_ARCH_RRDB && e->desc.scale == 2) ? 2 : 1;
  int rh = Hp, rw = Wp;
  if (tile > 0) {
    rh = std::min(Hp, tile + 2 * tile_pad);
    rw = std::min(Wp, tile + 2 * tile_pad);
  }
  *bytes = region_ws_bytes(e, n, (rh + s - 1) / s, (rw + s - 1) / s);
  return B200SR_OK;
}

static int enqueue_impl(b200sr_engine* e, Lane* lane, const void* src_dev_v, void* dst_dev_v, int n, int h, int w,
                        int tile, int tile_pad, int pre_pad, void* cuda_stream, int sample16) {
  const uint8_t* src_dev = static_cast<const uint8_t*>(src_dev_v);
  uint8_t* dst_dev = static_cast<uint8_t*>(dst_dev_v);
  if (!e) return B200SR_ERR_INVALID;
  if (!e->finalized) return fail(e, B200SR_ERR_STATE, "weights not finalised");
  if (!src_dev || !dst_dev || n <= 0 || h <= 0 || w <= 0 || tile < 0 || tile_pad < 0 || pre_pad < 0)
    return fail(e, B200SR_ERR_INVALID, "bad argument");
  if (pre_pad >= h || pre_pad >= w) return fail(e, B200SR_ERR_INVALID, "pre_pad must be smaller than the frame (reflect padding)");
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  lane->launches = 0;
  lane->rdb_launch_idx = 0;
  const int scale = e->desc.scale;
  int Hp, Wp;
  padded_dims(e, h, w, pre_pad, &Hp, &Wp);
  if ((Hp > h + pre_pad && h + pre_pad < 2) || (Wp > w + pre_pad && w + pre_pad < 2))
    return fail(e, B200SR_ERR_INVALID, "frame too small for reflect mod-padding");
  std::vector<Region> regions = plan_regions(e->desc.arch, scale, h, w, tile, tile_pad, pre_pad);
  {  // one workspace allocation for the largest region (tile mode: six shapes per frame)
    const int us = (e->desc.arch == B200SR_ARCH_RRDB && scale == 2) ? 2 : 1;
    size_t need = 0;
    for (const Region& R : regions) need = std::max(need, region_ws_bytes(e, n, R.rh / us, R.rw / us));
    int rc = ensure_ws(e, lane, need, st);
    if (rc) return rc;
  }
  for (Region R : regions) {
    R.src = src_dev;
    R.dst = dst_dev;
    R.sample16 = sample16;
    R.n = n;
    int rc = run_region(e, lane, R, st);
    if (rc) return rc;
  }
  e->last_launches.store(lane->launches);
  return B200SR_OK;
}

// Device-pointer entry points: asynchronous on the caller's stream.  Every caller stream (up to 4) gets its own lane
// (workspace + work list), so work queued on two streams overlaps -- the tail of one stream's persistent kernels
// with the head of the other's -- without sharing scratch memory.  Calls on the SAME stream are ordered by the
// stream; further streams share `dev_lane` and must be ordered with each other by the caller.
static Lane* device_lane_for(b200sr_engine* e, void* cuda_stream) {
  std::lock_guard<std::mutex> g(e->mu);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  for (auto& p : e->dev_lanes)
    if (p.first == st) return p.second.get();
  if (e->dev_lanes.size() < 4) {
    e->dev_lanes.emplace_back(st, std::unique_ptr<Lane>(new Lane()));
    e->dev_lanes.back().second->id = -1 - static_cast<int>(e->dev_lanes.size());
    return e->dev_lanes.back().second.get();
  }
  return &e->dev_lane;
}

int b200sr_enqueue_u8(b200sr_engine* e, const uint8_t* src_dev, uint8_t* dst_dev, int n, int h, int w, int tile,
                      int tile_pad, int pre_pad, void* cuda_stream) {
  if (!e) return B200SR_ERR_INVALID;
  return enqueue_impl(e, device_lane_for(e, cuda_stream), src_dev, dst_dev, n, h, w, tile, tile_pad, pre_pad,
                      cuda_stream, 0);
}

int b200sr_enqueue_u16(b200sr_engine* e, const uint16_t* src_dev, uint16_t* dst_dev, int n, int h, int w, int tile,
                       int tile_pad, int pre_pad, void* cuda_stream) {
  if (!e) return B200SR_ERR_INVALID;
  return enqueue_impl(e, device_lane_for(e, cuda_stream), src_dev, dst_dev, n, h, w, tile, tile_pad, pre_pad,
                      cuda_stream, 1);
}

// ---- host-buffer entry point: pipelined over the engine's lanes ---------------------------------------------------
namespace {

bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

Lane* acquire_lane(b200sr_engine* e, bool block) {
  std::unique_lock<std::mutex> lk(e->mu);
  while (static_cast<int>(e->lanes.size()) < std::max(1, e->opt_lanes)) {
    e->lanes.emplace_back(new Lane());
    e->lanes.back()->id = static_cast<int>(e->lanes.size());
  }
  while (true) {
    for (auto& l : e->lanes)
      if (!l->busy) {
        l->busy = true;
        return l.get();
      }
    if (!block) return nullptr;
    e->cv.wait(lk);
  }
}

void release_lane(b200sr_engine* e, Lane* l) {
  {
    std::lock_guard<std::mutex> g(e->mu);
    l->busy = false;
  }
  e->cv.notify_one();
}

int grow_dev(b200sr_engine* e, uint8_t** p, size_t* have, size_t need) {
  if (need <= *have) return B200SR_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  CUDA_TRY(e, cudaMalloc(p, need));
  *have = need;
  return B200SR_OK;
}

int grow_pinned(b200sr_engine* e, uint8_t** p, size_t* have, size_t need) {
  if (need <= *have) return B200SR_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *have = 0;
  CUDA_TRY(e, cudaHostAlloc(p, need, cudaHostAllocPortable));
  *have = need;
  return B200SR_OK;
}

struct HostJob {
  Lane* lane;
  int f0, cnt;
};

}  // namespace

// n frames from host memory to host memory.  The frames are cut into jobs of `chunk` frames; every job takes a free
// lane and queues H2D copy -> forward pass -> D2H copy on that lane's stream, so job k+1's upload and job k-1's
// download overlap job k's kernels (two lanes = double buffering), and so do the jobs of concurrent callers.
// Pinned (cudaHostAlloc / cudaHostRegister) buffers are copied directly; pageable ones are staged through the
// lane's pinned buffers.  Deadlock-free: a caller blocks for a lane only while it holds none.
static int upscale_host_impl(b200sr_engine* e, const void* src_host, void* dst_host, int n, int h, int w, int tile,
                             int tile_pad, int pre_pad, int sample16) {
  if (!e || !src_host || !dst_host || n <= 0 || h <= 0 || w <= 0) return B200SR_ERR_INVALID;
  if (!e->finalized) return fail(e, B200SR_ERR_STATE, "weights not finalised");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const size_t in_frame = static_cast<size_t>(h) * w * 3 * (sample16 ? 2 : 1);
  const size_t out_frame = in_frame * e->desc.scale * e->desc.scale;
  const bool src_pinned = host_ptr_is_pinned(src_host), dst_pinned = host_ptr_is_pinned(dst_host);
  // frames per job: ~two 720p frames of pixels keeps the persistent kernels' launch tails short relative to the
  // job while two lanes in flight hide them (measured flat from there up)
  int chunk = e->opt_host_chunk;
  if (chunk <= 0) chunk = static_cast<int>(std::max<long>(1, (2L * 1280 * 720) / (static_cast<long>(h) * w)));
  chunk = std::min(chunk, n);
  const uint8_t* src = static_cast<const uint8_t*>(src_host);
  uint8_t* dst = static_cast<uint8_t*>(dst_host);
  std::deque<HostJob> inflight;
  int status = B200SR_OK;
  int launches = 0;

  auto retire = [&](const HostJob& j, bool copy_out) {
    cudaError_t err = cudaStreamSynchronize(j.lane->stream);
    if (err != cudaSuccess && status == B200SR_OK)
      status = fail(e, err == cudaErrorMemoryAllocation ? B200SR_ERR_OOM : B200SR_ERR_CUDA,
                    std::string("cudaStreamSynchronize: ") + cudaGetErrorString(err));
    if (copy_out && status == B200SR_OK && !dst_pinned)
      memcpy(dst + static_cast<size_t>(j.f0) * out_frame, j.lane->pin_out, static_cast<size_t>(j.cnt) * out_frame);
    launches += j.lane->launches;
    release_lane(e, j.lane);
  };
  auto submit = [&](Lane* L, int f0, int cnt) -> int {
    if (!L->stream) CUDA_TRY(e, cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
    const size_t ib = static_cast<size_t>(cnt) * in_frame, ob = static_cast<size_t>(cnt) * out_frame;
    int rc = grow_dev(e, &L->dev_in, &L->dev_in_bytes, ib);
    if (!rc) rc = grow_dev(e, &L->dev_out, &L->dev_out_bytes, ob);
    if (!rc && !src_pinned) rc = grow_pinned(e, &L->pin_in, &L->pin_in_bytes, ib);
    if (!rc && !dst_pinned) rc = grow_pinned(e, &L->pin_out, &L->pin_out_bytes, ob);
    if (rc) return rc;
    const uint8_t* sp = src + static_cast<size_t>(f0) * in_frame;
    if (!src_pinned) {
      memcpy(L->pin_in, sp, ib);
      sp = L->pin_in;
    }
    CUDA_TRY(e, cudaMemcpyAsync(L->dev_in, sp, ib, cudaMemcpyHostToDevice, L->stream));
    rc = enqueue_impl(e, L, L->dev_in, L->dev_out, cnt, h, w, tile, tile_pad, pre_pad, L->stream, sample16);
    if (rc) return rc;
    uint8_t* dp = dst_pinned ? dst + static_cast<size_t>(f0) * out_frame : L->pin_out;
    CUDA_TRY(e, cudaMemcpyAsync(dp, L->dev_out, ob, cudaMemcpyDeviceToHost, L->stream));
    return B200SR_OK;
  };

  for (int f0 = 0; f0 < n && status == B200SR_OK; f0 += chunk) {
    const int cnt = std::min(chunk, n - f0);
    Lane* L = inflight.empty() ? acquire_lane(e, true) : acquire_lane(e, false);
    while (!L) {   // every lane is busy and this call holds some: finish our oldest job first
      retire(inflight.front(), true);
      inflight.pop_front();
      L = inflight.empty() ? acquire_lane(e, true) : acquire_lane(e, false);
    }
    const int rc = submit(L, f0, cnt);
    if (rc) {
      status = rc;
      cudaStreamSynchronize(L->stream);
      release_lane(e, L);
      break;
    }
    inflight.push_back(HostJob{L, f0, cnt});
  }
  while (!inflight.empty()) {
    retire(inflight.front(), true);
    inflight.pop_front();
  }
  if (status == B200SR_OK) e->last_launches.store(launches);
  return status;
}

int b200sr_upscale_host_u8(b200sr_engine* e, const uint8_t* src_host, uint8_t* dst_host, int n, int h, int w, int tile,
                           int tile_pad, int pre_pad) {
  return upscale_host_impl(e, src_host, dst_host, n, h, w, tile, tile_pad, pre_pad, 0);
}

int b200sr_upscale_host_u16(b200sr_engine* e, const uint16_t* src_host, uint16_t* dst_host, int n, int h, int w,
                            int tile, int tile_pad, int pre_pad) {
  return upscale_host_impl(e, src_host, dst_host, n, h, w, tile, tile_pad, pre_pad, 1);
}

// Pinned host memory for callers that want zero-copy staging (the Python layer allocates `enhance`'s result arrays
// here, and the multi-GPU scheduler registers its shared-memory frame rings).
void* b200sr_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void b200sr_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int b200sr_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return B200SR_ERR_INVALID;
  if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
    cudaGetLastError();
    return B200SR_ERR_CUDA;
  }
  return B200SR_OK;
}
int b200sr_host_unregister(void* p) {
  if (!p) return B200SR_ERR_INVALID;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return B200SR_ERR_CUDA;
  }
  return B200SR_OK;
}

int b200sr_last_launch_count(const b200sr_engine* e) { return e ? e->last_launches.load() : 0; }

int b200sr_set_option(b200sr_engine* e, const char* key, int value) {
  if (!e || !key) return B200SR_ERR_INVALID;
  if (!strcmp(key, "force_th")) {
    e->opt_force_th = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "max_ctas")) {
    e->opt_max_ctas = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "rdb_off")) {   // decimal digits-pairs: value = o1 + 100*o2 + 10000*o3 + 1000000*o4
    RDB_STEP_OFF[1] = value % 100;
    RDB_STEP_OFF[2] = (value / 100) % 100;
    RDB_STEP_OFF[3] = (value / 10000) % 100;
    RDB_STEP_OFF[4] = (value / 1000000) % 100;
    e->rdb_gen++;   // lanes rebuild their work lists
    return B200SR_OK;
  }
  if (!strcmp(key, "rdb_interleave")) {
    RDB_INTERLEAVE = value;
    e->rdb_gen++;
    return B200SR_OK;
  }
  if (!strcmp(key, "rdb_order")) {   // five decimal digits, e.g. 32104
    int v = value;
    for (int i = 4; i >= 0; --i) {
      RDB_ORDER[i] = v % 10;
      v /= 10;
    }
    e->rdb_gen++;
    return B200SR_OK;
  }
  if (!strcmp(key, "rdb_stats")) {
    e->opt_rdb_stats = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "fused_rdb")) {
    e->opt_fused_rdb = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "first_v1")) {
    e->opt_first_v1 = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "fold_up")) {
    e->opt_fold_up = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "trunk_lo")) {
    e->opt_trunk_lo = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "rdb_half64")) {
    e->opt_half64 = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "fuse_tail")) {
    e->opt_fuse_tail = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "pair")) {
    e->opt_pair = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "last9")) {
    e->opt_last9 = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "abl")) {
    e->opt_abl = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "w_resident")) {
    e->opt_w_resident = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "lanes")) {   // host-buffer lanes (streams + workspaces); takes effect for lanes not yet created
    if (value < 1 || value > 8) return fail(e, B200SR_ERR_INVALID, "lanes must be 1..8");
    std::lock_guard<std::mutex> g(e->mu);
    for (auto& l : e->lanes)
      if (l->busy) return B200SR_ERR_STATE;
    cudaSetDevice(e->device);
    while (static_cast<int>(e->lanes.size()) > value) {
      free_lane(e->lanes.back().get());
      e->lanes.pop_back();
    }
    e->opt_lanes = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "host_chunk")) {   // frames per lane job of a host-buffer call (0 = auto)
    e->opt_host_chunk = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "ws_limit_mb")) {   // tests: refuse larger workspaces with B200SR_ERR_OOM
    e->opt_ws_limit_mb = value;
    return B200SR_OK;
  }
  if (!strcmp(key, "profile")) {  // 1: time every launch with CUDA events; read back with b200sr_get_profile
    e->opt_profile = value;
    return B200SR_OK;
  }
  return fail(e, B200SR_ERR_INVALID, std::string("unknown option ") + key);
}

int b200sr_get_profile(b200sr_engine* e, int nclass, double* ms, double* flops, int* launches) {
  if (!e || !ms || !flops || !launches || nclass < PC_COUNT) return B200SR_ERR_INVALID;
  for (int i = 0; i < nclass; ++i) {
    ms[i] = 0;
    flops[i] = 0;
    launches[i] = 0;
  }
  for (auto& r : e->prof) {
    cudaEventSynchronize(r.e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms[r.cls] += t;
    flops[r.cls] += r.flops;
    launches[r.cls] += 1;
    e->ev_pool.push_back(r.e0);
    e->ev_pool.push_back(r.e1);
  }
  e->prof.clear();
  return B200SR_OK;
}

// Dev tool: per-CTA cycle counters of the LAST fused-RDB launch (option "rdb_stats" = 1); [grid][16] long long.
int b200sr_debug_rdb_stats(b200sr_engine* e, long long* out, int max_ctas) {
  if (!e || !out || !e->d_rdb_stats) return -1;
  const int n = std::min(max_ctas, std::min(e->num_sms, 148));
  if (cudaMemcpy(out, e->d_rdb_stats, static_cast<size_t>(n) * 16 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

// Dev tool: per-item timestamps of the traced fused-RDB launch; [nitems][10] long long.  Returns nitems.
int b200sr_debug_rdb_trace(b200sr_engine* e, long long* out, int max_items) {
  if (!e || !e->d_rdb_trace) return -1;
  const int n = std::min(std::abs(max_items), e->stats_nitems);
  if (max_items < 0) {   // negative count: fetch the per-row trace ([nitems][16][3]) instead
    if (out && cudaMemcpy(out, e->d_rdb_trace2, static_cast<size_t>(n) * 48 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e->stats_nitems;
  }
  if (out && cudaMemcpy(out, e->d_rdb_trace, static_cast<size_t>(n) * 10 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return e->stats_nitems;
}

// ---- test hooks that need no GPU ----------------------------------------------------------------
int b200sr_debug_plan_regions(int arch, int scale, int h, int w, int tile, int tile_pad, int pre_pad, int* out,
                              int max_regions) {
  if (!out || h <= 0 || w <= 0 || tile < 0 || tile_pad < 0 || pre_pad < 0) return -1;
  std::vector<Region> rs = plan_regions(arch, scale, h, w, tile, tile_pad, pre_pad);
  int n = 0;
  for (const Region& R : rs) {
    if (n >= max_regions) break;
    int* o = out + n * 10;
    o[0] = R.oy; o[1] = R.ox; o[2] = R.rh; o[3] = R.rw; o[4] = R.crop_y0; o[5] = R.crop_x0;
    o[6] = R.crop_h; o[7] = R.crop_w; o[8] = R.dst_y0; o[9] = R.dst_x0;
    ++n;
  }
  return static_cast<int>(rs.size());
}

// Packs one conv layer's weights exactly as b200sr_finalize does; returns the image size in bytes
// (call with out == nullptr to query it).
long long b200sr_debug_pack_weights(const float* weight, int cout, int cin, int fp16, uint8_t* out, long long out_bytes) {
  if (!weight || cout <= 0 || cin <= 0) return -1;
  Layer l;
  l.cin = cin;
  l.cout = cout;
  l.coutp = coutp_for(cout);
  l.fp16 = fp16 != 0;
  l.w.assign(weight, weight + static_cast<size_t>(cout) * cin * 9);
  std::vector<uint8_t> img = pack_weights(l);
  if (out) {
    if (out_bytes < static_cast<long long>(img.size())) return -1;
    memcpy(out, img.data(), img.size());
  }
  return static_cast<long long>(img.size());
}

// The fused-RDB work list for n frames of h x w: 8 ints per item (k n y0 rows tx flag_base dep_base[0] dep_base[1]);
// returns the item count (may exceed max_items).
int b200sr_debug_rdb_flag_rows(void) { return RDB_FLAG_ROWS; }

int b200sr_debug_rdb_items(int n, int h, int w, int* out, int max_items) {
  if (n <= 0 || h <= 0 || w <= 0) return -1;
  std::vector<RdbItem> items;
  int nflags = 0;
  build_rdb_items(n, h, w, items, &nflags);
  for (size_t i = 0; i < items.size() && static_cast<int>(i) < max_items && out; ++i) {
    const RdbItem& it = items[i];
    int* o = out + i * 8;
    o[0] = it.k; o[1] = it.n; o[2] = it.y0; o[3] = it.rows; o[4] = it.tx; o[5] = it.flag_base;
    o[6] = it.dep_base[0]; o[7] = it.dep_base[1];
  }
  return static_cast<int>(items.size());
}

// Rows per tile the launcher would pick (wave balancing) for an output of n x h x w on `num_sms` SMs.
int b200sr_debug_choose_th(int coutp, int n, int h, int w, int num_sms) {
  b200sr_engine e;
  e.num_sms = num_sms;
  return choose_th(&e, coutp, n, h, w);
}

// ---- test hook: one tensor-core conv layer on caller-provided device tensors -------------------
// in: bf16 NHWC [n][h][w][in_pitch]; weight fp32 OIHW host; out: bf16 NHWC [n][h][w][out_pitch].
// epi: 0 = leaky(slope) -> bf16 slice ; 1 = prelu(prelu_host) -> bf16.
int b200sr_debug_conv3x3(int device, const void* in_dev, int n, int h, int w, int in_pitch, int cin,
                         const float* weight, const float* bias, int cout, int epi, float slope,
                         const float* prelu_host, void* out_dev, int out_pitch, int out_choff, int fp16, int force_th,
                         int max_ctas, void* cuda_stream, char* errbuf, int errbuf_len) {
  b200sr_engine e;
  e.device = device;
  auto report = [&](int rc) {
    if (errbuf && errbuf_len > 0) snprintf(errbuf, errbuf_len, "%s", e.err.c_str());
    return rc;
  };
  if (cudaSetDevice(device) != cudaSuccess) return report(fail(&e, B200SR_ERR_CUDA, "cudaSetDevice failed"));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return report(fail(&e, B200SR_ERR_CUDA, "needs an sm_100 device"));
  e.num_sms = prop.multiProcessorCount;
  e.opt_force_th = force_th;
  e.opt_max_ctas = max_ctas;
  if (!tmap_init()) return report(fail(&e, B200SR_ERR_CUDA, "no cuTensorMapEncodeTiled"));
  if (cin % 32 || cin < 32 || cin > in_pitch) return report(fail(&e, B200SR_ERR_INVALID, "cin must be a multiple of 32 and <= in_pitch"));
  Layer l;
  l.cin = cin;
  l.cout = cout;
  l.coutp = coutp_for(cout);
  l.fp16 = fp16 != 0;
  if (!((l.coutp == 32 && epi == 0) || (l.coutp == 64 && (epi == 0 || epi == 1))))
    return report(fail(&e, B200SR_ERR_INVALID, "debug conv supports Cout 32 (leaky) and 64 (leaky / prelu)"));
  l.w.assign(weight, weight + static_cast<size_t>(cout) * cin * 9);
  l.b.assign(bias, bias + cout);
  std::vector<uint8_t> img = pack_weights(l);
  std::vector<float> bp(l.coutp, 0.f);
  std::copy(l.b.begin(), l.b.end(), bp.begin());
  float* d_prelu = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  int rc = B200SR_OK;
  auto body = [&]() -> int {
    CUDA_TRY(&e, cudaMalloc(&l.d_wpack, img.size()));
    CUDA_TRY(&e, cudaMalloc(&l.d_bias, bp.size() * 4));
    CUDA_TRY(&e, cudaMemcpy(l.d_wpack, img.data(), img.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(&e, cudaMemcpy(l.d_bias, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    ConvArgs a{};
    a.slope = slope;
    a.out = static_cast<__nv_bfloat16*>(out_dev);
    a.out_pitch = out_pitch;
    a.out_choff = out_choff;
    a.out_fp16 = fp16 ? 1 : 0;
    if (epi == 1) {
      if (!prelu_host) return fail(&e, B200SR_ERR_INVALID, "prelu slopes missing");
      CUDA_TRY(&e, cudaMalloc(&d_prelu, 64 * 4));
      CUDA_TRY(&e, cudaMemcpy(d_prelu, prelu_host, 64 * 4, cudaMemcpyHostToDevice));
      a.prelu = d_prelu;
    }
    ConvIO io{in_dev, in_pitch, n, h, w};
    Lane lane;
    int r = launch_conv(&e, &lane, l, epi == 1 ? EPI_PRELU_BF16 : EPI_ACT_BF16, io, a, st);
    if (r) return r;
    CUDA_TRY(&e, cudaStreamSynchronize(st));
    return B200SR_OK;
  };
  rc = body();
  if (l.d_wpack) cudaFree(l.d_wpack);
  if (l.d_bias) cudaFree(l.d_bias);
  if (d_prelu) cudaFree(d_prelu);
  return report(rc);
}

}  // extern "C"
