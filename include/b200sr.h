/*
 * b200sr.h -- C ABI of the B200-native Real-ESRGAN upscaling engine (libb200sr.so).
 *
 * The reference (FrameWright) has no FFI on this path: its upscale processor is Python that
 * constructs and calls third-party PyTorch code.  Each entry point below therefore cites the
 * reference *Python* interface whose work it replaces (paths relative to
 * /root/reference/src/framewright/).  The Python host layer (package `framewright_b200`)
 * binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: plain C, no exceptions, no torch types.  Every function returns a status
 * code (B200SR_OK == 0); b200sr_last_error() gives the message of the calling thread's last
 * failed call.  One engine per GPU.  Threading: the host-buffer calls (b200sr_upscale_host_*)
 * are thread-safe -- each takes one of the engine's lanes (stream + workspace + staging), which is
 * how the reference's parallel_frames threads (restorer.py:1894) overlap copies and kernels; the
 * device-pointer calls (b200sr_enqueue_*) are stream-ordered and must be ordered with each other
 * by the caller; weight loading / options / destroy need the engine idle.
 * Frames are uint8 HWC, BGR channel order (what cv2.imread yields, pytorch_realesrgan.py:198).
 */
#ifndef B200SR_H_
#define B200SR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SR_OK 0
#define B200SR_ERR_INVALID 1   /* bad argument / unsupported geometry */
#define B200SR_ERR_CUDA 2      /* CUDA runtime or driver error */
#define B200SR_ERR_OOM 3       /* device allocation failed -> host layer reports "GPU out of memory" */
#define B200SR_ERR_STATE 4     /* weights not finalised, etc. */

#define B200SR_ARCH_RRDB 0     /* basicsr RRDBNet   (processors/pytorch_realesrgan.py:107,112,117) */
#define B200SR_ARCH_SRVGG 1    /* realesrgan SRVGGNetCompact (BASELINE.json north_star) */

typedef struct b200sr_engine b200sr_engine;

/* Architecture descriptor == the RRDBNet(...)/SRVGGNetCompact(...) constructor arguments the
 * reference passes at processors/pytorch_realesrgan.py:103-129 and cli.py:715-723. */
typedef struct b200sr_model_desc {
  int arch;        /* B200SR_ARCH_* */
  int scale;       /* network scale: 4, or 2 (RRDBNet with pixel-unshuffle input) */
  int num_feat;    /* 64 */
  int num_block;   /* RRDB blocks (23 / 6) or SRVGG body convs (32 / 16) */
  int num_grow_ch; /* 32 */
} b200sr_model_desc;

/* Replaces: model construction + `.to(device)` inside RealESRGANer.__init__
 * (called at processors/pytorch_realesrgan.py:160-170). */
int b200sr_create(const b200sr_model_desc* desc, int device, b200sr_engine** out);
void b200sr_destroy(b200sr_engine* e);

/* Number of 3x3 conv layers (execution order: processors/... RRDBNet.forward order, see
 * framewright_b200/archs.py::conv_layers) and of PReLU layers. */
int b200sr_num_convs(const b200sr_engine* e);
int b200sr_num_prelus(const b200sr_engine* e);

/* Replaces: `load_state_dict(strict=True)` + `.half()` in RealESRGANer.__init__.
 * weight: fp32 [cout][cin][3][3] (PyTorch OIHW), bias: fp32 [cout]; host pointers. */
int b200sr_set_conv(b200sr_engine* e, int layer, const float* weight, const float* bias, int cout, int cin);
int b200sr_set_prelu(b200sr_engine* e, int index, const float* slope, int n);
/* Packs weights to the tensor-core layout (bf16, K-major, 128B-swizzled tiles) and uploads. */
int b200sr_finalize(b200sr_engine* e);

/* Output geometry of `enhance` at the network's native scale. */
int b200sr_output_dims(const b200sr_engine* e, int h, int w, int* out_h, int* out_w);

/* Device bytes the engine holds for n frames of h x w (tile > 0: per-tile buffers). */
int b200sr_workspace_bytes(b200sr_engine* e, int n, int h, int w, int tile, int tile_pad, int pre_pad, size_t* bytes);

/* Replaces: RealESRGANer.pre_process / tile_process / process / post_process and the uint8
 * conversion at the end of RealESRGANer.enhance (called at processors/pytorch_realesrgan.py:223,
 * processors/enhancement/super_resolution.py:524, cli.py:769), for the 3-channel uint8 branch:
 *   src: device pointer, n frames [h][w][3] u8 BGR; dst: device pointer, n frames
 *   [h*scale][w*scale][3] u8 BGR.  tile == 0 -> whole frame; tile > 0 -> upstream tile loop
 *   (tile, tile_pad) with zero padding at each padded tile's border (seam-exact).
 *   pre_pad: reflect pad right/bottom before, cropped after.  Asynchronous on `cuda_stream`. */
int b200sr_enqueue_u8(b200sr_engine* e, const uint8_t* src_dev, uint8_t* dst_dev, int n, int h, int w, int tile,
                      int tile_pad, int pre_pad, void* cuda_stream);

/* Same with HOST buffers (the end-to-end call, what RealESRGANer.enhance's `.to(device)` ... `.cpu()` does,
 * called at processors/pytorch_realesrgan.py:223): the n frames are cut into jobs of a few frames, each job
 * queues H2D copy -> forward -> D2H copy on a free lane's stream, two lanes by default, so uploads and
 * downloads overlap the neighbouring job's kernels (and those of concurrent callers).  Pinned buffers
 * (b200sr_host_alloc / b200sr_host_register) are copied directly, pageable ones staged through pinned memory.
 * Returns when dst_host holds the result.  Thread-safe. */
int b200sr_upscale_host_u8(b200sr_engine* e, const uint8_t* src_host, uint8_t* dst_host, int n, int h, int w,
                           int tile, int tile_pad, int pre_pad);

/* The same two calls for frames of uint16 samples (upstream RealESRGANer.enhance's 16-bit branch: img / 65535 in,
 * round(clamp(out, 0, 1) * 65535) out).  [N][H][W][3] uint16 BGR -> [N][sH][sW][3] uint16 BGR. */
int b200sr_enqueue_u16(b200sr_engine* e, const uint16_t* src_dev, uint16_t* dst_dev, int n, int h, int w, int tile,
                       int tile_pad, int pre_pad, void* cuda_stream);
int b200sr_upscale_host_u16(b200sr_engine* e, const uint16_t* src_host, uint16_t* dst_host, int n, int h, int w,
                            int tile, int tile_pad, int pre_pad);

/* Pinned host memory: allocate (cudaHostAlloc, portable) / free, or pin an existing range (shared-memory frame
 * rings of the multi-GPU scheduler) so that the host-buffer calls copy without staging. */
void* b200sr_host_alloc(size_t bytes);
void b200sr_host_free(void* p);
int b200sr_host_register(void* p, size_t bytes);
int b200sr_host_unregister(void* p);

/* Number of kernels the last enqueue launched (bench.py reports it as gpu_launches). */
int b200sr_last_launch_count(const b200sr_engine* e);

/* Debug / measurement hooks (not part of the reference surface).
 * Options: "fused_rdb" (1: one persistent kernel per residual dense block; 0: five per-conv launches, same bytes),
 * "fold_up" (1: conv_up1/up2 read the nearest-2x upsampling through the duplicated-pixel TMA view; 0: materialise it),
 * "trunk_lo" (where the residual stream's e5m2 lo part is used -- 0: in the RRDB-level skip only, 1: also in the first
 * RDB's residual add, 2: the pair after every RDB),
 * "rdb_half64" (1: the 32-channel last chunk of RDB conv2 / conv4 is loaded as a 32-channel SWIZZLE_64B box; 0: a
 * 64-channel box of which half is used; identical bytes),
 * "fuse_tail" (1: conv_hr + conv_last as one rolling kernel, the 4x tensor between them stays on chip; identical bytes),
 * "pair" (1: single-chunk convs run through conv3x3_sc_kernel -- resident weights, row-pair stages; 0: the per-row
 * conv3x3_tc_kernel, identical bytes), "last9" (1: conv_last with the kx taps stacked on N; 0: per-tap form, within 1 LSB),
 * "w_resident" (per-row kernel only: convs with Cin <= 64 keep their weights resident), "abl" (timing ablations of the
 * per-conv kernels -- 1 no epilogue stores, 2 no MMAs, 4 no TMA loads; results are wrong when set),
 * "profile" (1: CUDA events around every launch, read with b200sr_get_profile), "force_th", "max_ctas",
 * "lanes" (host-buffer lanes, default 2), "host_chunk" (frames per lane job, 0 = auto), "ws_limit_mb" (tests: refuse
 * larger workspaces with B200SR_ERR_OOM),
 * "rdb_off" / "rdb_order" (fused work-list schedule), "rdb_stats" (instrumented builds only). */
int b200sr_set_option(b200sr_engine* e, const char* key, int value);

/* Per-kernel-class timing collected while option "profile" is 1 (CUDA events around every launch).
 * Arrays of length nclass >= 12, indexed: 0 conv<32,act> 1 conv<64,act> 2 conv<64,prelu> 3 conv<64,rdb5>
 * 4 conv<64,rdb5+rrdb> 5 conv<64,add> 6 conv<16,last_u8> 7 conv<48,srvgg_last> 8 first_conv 9 upsample2x 10 rdb_fused 11 hr_last_fused.
 * flops = algorithmic FLOPs (true channel counts).  Clears the collected records. */
int b200sr_get_profile(b200sr_engine* e, int nclass, double* ms, double* flops, int* launches);

/* Dev tool: wait / issue cycle counters of the last fused-RDB launch, 16 long long per CTA (needs option
 * "rdb_stats" = 1).  Returns the number of CTAs written. */
int b200sr_debug_rdb_stats(b200sr_engine* e, long long* out, int max_ctas);
/* Dev tool: per-item globaltimer stamps of the traced launch (library built with -DB200SR_RDB_STATS), 10 per item:
 * claim, dep1 begin/end, dep2 begin/end, loads issued, MMAs issued, epilogue done, CTA, unused. */
int b200sr_debug_rdb_trace(b200sr_engine* e, long long* out, int max_items);

/* Host-only test hooks (no GPU needed).
 * plan_regions: the regions one enhance call is split into (tile == 0: one region; else upstream
 *   tile_process geometry); 10 ints per region: oy ox rh rw crop_y0 crop_x0 crop_h crop_w dst_y0 dst_x0.
 *   Returns the region count (may exceed max_regions), -1 on bad arguments.
 * pack_weights: the tensor-core weight image of one layer; returns its size in bytes.
 * choose_th: output rows per CTA tile the launcher picks. */
int b200sr_debug_plan_regions(int arch, int scale, int h, int w, int tile, int tile_pad, int pre_pad, int* out,
                              int max_regions);
long long b200sr_debug_pack_weights(const float* weight, int cout, int cin, int fp16, uint8_t* out, long long out_bytes);
int b200sr_debug_choose_th(int coutp, int n, int h, int w, int num_sms);
/* rdb_items: work list of the fused-RDB kernel; 8 ints per item: k n y0 rows tx flag_base dep_base0 dep_base1
 * (counter of 8-row block b of a conv = base + b).  Returns the item count. */
int b200sr_debug_rdb_items(int n, int h, int w, int* out, int max_items);
/* Output rows per completion counter of the fused-RDB schedule (flag_base / dep_base index blocks of this many rows). */
int b200sr_debug_rdb_flag_rows(void);

/* Test hook: one tensor-core 3x3 conv layer on caller-provided device tensors (bf16 NHWC in/out,
 * fp32 OIHW host weights).  epi 0: leaky(slope) ; epi 1: PReLU(prelu_host[64]); fp16 != 0: tensors and
 * weights are fp16 instead of bf16.  Synchronous. */
int b200sr_debug_conv3x3(int device, const void* in_dev, int n, int h, int w, int in_pitch, int cin,
                         const float* weight, const float* bias, int cout, int epi, float slope,
                         const float* prelu_host, void* out_dev, int out_pitch, int out_choff, int fp16, int force_th,
                         int max_ctas, void* cuda_stream, char* errbuf, int errbuf_len);

const char* b200sr_last_error(const b200sr_engine* e);
const char* b200sr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SR_H_ */
