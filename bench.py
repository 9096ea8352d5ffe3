#!/usr/bin/env python
"""Benchmark of the Real-ESRGAN upscaling hot path on B200 (contract: task prompt, section 4).

A "step" is one pass of the hot path (RealESRGAN_x4plus RRDBNet x4, untiled) over one batch of B synthetic 1280x720
uint8 frames -- the configuration BASELINE.json's metric ("RRDBNet x4 frames/sec @720p input") is quoted on.  With
N > 1 GPUs (torchrun, one rank per GPU) every rank runs the same number of frames (weak scaling, frame sharding, no
collective on the data path; the process group exists only for the timing barrier / max-over-ranks).

  value         frames/s with inputs already resident in HBM (device-pointer C-ABI call, B frames per step)
  e2e           frames/s through the reference-facing plugin call: `get_upsampler(cfg).enhance(img)` -- one pageable
                numpy frame in, one numpy frame out per call, from `parallel_frames` caller threads as the reference
                drives it (restorer.py:1894); H2D + forward + D2H inside.  `e2e.variants` adds the single-thread
                figure, `enhance_frames_batch` (B frames per call) and the raw host-buffer C-ABI call.
  roofline      dominant kernel (the fused residual-dense-block kernel) vs the measured sustained dense bf16 peak;
                `per_class` carries every kernel class with its tensor / HBM fraction
  parity        frame 0 of the TIMED output against the fp32 CPU oracle on the same weights and frame (N = 1)
  cpu_baseline  that oracle run, timed: one full 720p frame on the host cores (no extrapolation)
  configs       BASELINE.json configs 1-4, device-resident (frames/s, TFLOP/s)

`--impl reference` times only the CPU reference path (oracle/; the reference's own dependencies are not installable
here): each step is one full-width 1280x96 band of a 720p frame.
`--workload clip2000` pushes a 2000-frame synthetic clip through the product scheduler
(`MultiGPUDistributor.distribute_frames`, strong scaling over the visible GPUs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = "RealESRGAN_x4plus"
H, W = 720, 1280
METRIC = "frames_per_sec_rrdbnet_x4_720p"
UNIT = "frames/s"
BAND = 96      # rows of a 720p frame per step of the CPU reference arm
STREAMS = 1    # set from --streams


def workload_config(B: int, world: int) -> dict:
    return {"workload": f"{MODEL} x4 on {B} synthetic 1280x720 uint8 frames per step per GPU, untiled, "
                        "random-init weights (seed 0)",
            "frames_per_step_per_gpu": B, "streams": STREAMS,
            "l2": "per-step working set (>5 GB of activations) exceeds the 126 MB L2",
            "parallelism": f"frame-sharded x{world}, no collective on the data path (process group: timing barrier only)"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": float(p["bf16_tflops"]), "bf16_tflops_sustained": float(p["bf16_tflops_sustained"]),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = max(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def synthetic_frames(n: int, seed: int, h: int = H, w: int = W) -> np.ndarray:
    """n synthetic uint8 BGR frames, generated from seed + frame index (no disk)."""
    from oracle.oracle import synthetic_frame  # data generator only (shared with the tests)

    return np.stack([synthetic_frame(h, w, seed=seed * 100003 + i, kind="mixed") for i in range(n)])


# --------------------------------------------------------------------------------------------------
def _cpu_oracle(model=MODEL):
    import torch

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_synthetic_state_dict(model, 0)
    return oracle.make_upsampler(model, sd, tile=0, pre_pad=0), torch.get_num_threads()


def run_reference(args):
    """The reference's CPU path (fp32 torch restatement, all host threads).  One step = one full-width 1280 x BAND
    band of a 720p frame (the network is fully convolutional: cost is proportional to pixels)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle

    up, cores = _cpu_oracle()
    frame = oracle.synthetic_frame(H, W, seed=4 * 100003, kind="mixed")
    steps, warmup = max(args.steps, 1), max(min(args.warmup, 2), 1)
    bands = [np.ascontiguousarray(frame[(i * BAND) % (H - BAND + 1):][:BAND]) for i in range(steps + warmup)]
    for b in bands[:warmup]:
        up.enhance(b, outscale=4)
    t0 = time.perf_counter()
    for b in bands[warmup:]:
        up.enhance(b, outscale=4)
    dt = (time.perf_counter() - t0) / steps
    fps = (BAND / H) / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, world),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"one 1280x{BAND} full-width band of the 720p frame per step ({dt:.2f} s/step), "
                                   f"fp32 torch oracle on {cores} threads; frames/s = ({BAND}/720) / s_per_step"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
def _time_device(eng, dev_in, dev_out, steps, torch, **kw):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        eng.upscale_device(dev_in, out=dev_out, **kw)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)


def bench_configs(torch, peaks):
    """BASELINE.json configs 1-4, device-resident: frames/s, ms/step, algorithmic TFLOP/s."""
    from framewright_b200.archs import MODEL_ARCHS, make_synthetic_state_dict
    from framewright_b200.engine import B200Engine

    cases = [
        ("cfg1_x4plus_256x256_batch8", "RealESRGAN_x4plus", 8, 256, 256, {}, 1.0),
        ("cfg2_general_x4v3_640x480_batch64", "realesr-general-x4v3", 64, 480, 640, {}, 1.0),
        ("cfg3_x2plus_1080p", "RealESRGAN_x2plus", 2, 1080, 1920, {}, 0.25),          # MACs are per unshuffled pixel
        ("cfg4_x4plus_720p_tile512_pad10", "RealESRGAN_x4plus", 4, 720, 1280, {"tile": 512, "tile_pad": 10}, 976800 / 921600),
        # the CLI's variant (cli.py:742-750: pre_pad = 10): tiles over the 1290x730 padded image, 997 500 px processed
        ("cfg4_cli_prepad10", "RealESRGAN_x4plus", 4, 720, 1280, {"tile": 512, "tile_pad": 10, "pre_pad": 10}, 997500 / 921600),
    ]
    out = {}
    for key, name, n, h, w, kw, px_factor in cases:
        eng = B200Engine(name, make_synthetic_state_dict(name, 0), gpu_id=torch.cuda.current_device())
        s = MODEL_ARCHS[name].scale
        frames = torch.from_numpy(synthetic_frames(min(n, 4), seed=11, h=h, w=w)).cuda()
        if frames.shape[0] < n:
            frames = frames.repeat((n + frames.shape[0] - 1) // frames.shape[0], 1, 1, 1)[:n].contiguous()
        dst = torch.empty((n, h * s, w * s, 3), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            eng.upscale_device(frames, out=dst, **kw)
        torch.cuda.synchronize()
        steps = 5
        ms = _time_device(eng, frames, dst, steps, torch, **kw) / steps
        fps = n / (ms * 1e-3)
        flops = 2.0 * MODEL_ARCHS[name].macs_per_input_pixel() * h * w * px_factor
        out[key] = {"frames_per_sec": fps, "ms_per_step": ms, "frames_per_step": n,
                    "tflops": fps * flops / 1e12, "frac_of_burst_peak": fps * flops / 1e12 / peaks["bf16_tflops"],
                    "launches_per_step": eng.last_launch_count}
        eng.close()
        del frames, dst
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    import framewright_b200  # noqa: F401
    from framewright_b200 import _native
    from framewright_b200 import pytorch_realesrgan as plugin
    from framewright_b200.archs import MODEL_ARCHS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    _native.load()
    # random-init weights of the named architecture (there is no network for checkpoints): explicit opt-in, the
    # plugin never falls back to them silently
    os.environ["B200SR_SYNTHETIC_WEIGHTS"] = "0"
    B = args.batch
    cfg = plugin.PyTorchESRGANConfig(model_name=MODEL, scale_factor=4, tile_size=0, gpu_id=local_rank)
    upsampler = plugin.get_upsampler(cfg)        # the object the reference's callers hold
    assert upsampler.tile_size == 0
    eng = upsampler.engine
    frames = synthetic_frames(B, seed=4 + rank)
    dev_in = torch.from_numpy(frames).cuda()
    dev_out = torch.empty((B, H * 4, W * 4, 3), dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`): K steps on `--streams` CUDA streams (every stream has its own engine
    # lane; with 2, the tail of one step's persistent kernels overlaps the head of the next step's)
    streams = [torch.cuda.Stream() for _ in range(max(1, args.streams))]
    dev_outs = [dev_out] + [torch.empty_like(dev_out) for _ in streams[1:]]
    main = torch.cuda.current_stream()

    def run_steps(k: int) -> None:
        for s in streams:
            s.wait_stream(main)
        for i in range(k):
            j = i % len(streams)
            with torch.cuda.stream(streams[j]):
                eng.upscale_device(dev_in, out=dev_outs[j])
        for s in streams:
            main.wait_stream(s)

    run_steps(max(args.warmup, len(streams)))
    launches_per_step = eng.last_launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record(main)
        run_steps(args.steps)
        ev1.record(main)
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    fps = world * B * args.steps / (ms_total * 1e-3)
    timed_frame0 = dev_outs[(args.steps - 1) % len(streams)][0].cpu().numpy()   # frame 0 of the last timed step
    for o in dev_outs[1:]:
        assert torch.equal(o, dev_out)
    del dev_outs[1:]
    torch.cuda.empty_cache()

    # ---- device-resident, one frame per call
    one_in, one_out = dev_in[:1].contiguous(), dev_out[:1]
    for _ in range(2):
        eng.upscale_device(one_in, out=one_out)
    barrier()
    n1 = max(4, min(args.steps * B, 24))
    ms1 = max_over_ranks(_time_device(eng, one_in, one_out, n1, torch))
    fps_batch1 = world * n1 / (ms1 * 1e-3)

    # ---- end to end through the plugin call (`e2e`)
    e2e_frames = max(2 * B, min(args.steps * B, 48))
    pageable = [frames[i % B].copy() for i in range(e2e_frames)]     # ordinary numpy arrays, as cv2.imread returns

    def run_enhance(nthreads: int) -> float:
        keep = {}                                 # outputs are consumed and dropped, as a caller that writes them out

        def work(k):
            for i in range(k, e2e_frames, nthreads):
                out = upsampler.enhance(pageable[i], outscale=4)[0]
                if i == 0:
                    keep[0] = out
                del out

        barrier()
        t0 = time.perf_counter()
        if nthreads == 1:
            work(0)
        else:
            ts = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        run_enhance.last = keep[0]
        return world * e2e_frames / (ms * 1e-3)

    run_enhance(4)                                # warm-up: lanes, pinned result pool, workspaces
    e2e_1t = run_enhance(1)
    e2e_2t = run_enhance(2)
    e2e_4t = run_enhance(4)
    e2e_frame0 = run_enhance.last
    for _ in range(2):                            # warm-up (result buffers of this size enter the pinned pool)
        batch_out = plugin.enhance_frames_batch(frames, cfg)
        del batch_out
    barrier()
    nb = max(2, min(args.steps, 6))
    t0 = time.perf_counter()
    for it in range(nb):
        batch_out = plugin.enhance_frames_batch(frames, cfg)
        if it < nb - 1:
            del batch_out
    e2e_batch = world * B * nb / (max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3)
    big = np.concatenate([frames] * 4)            # 16 frames in one call: the engine pipelines its own chunks
    big_out = plugin.enhance_frames_batch(big, cfg)
    del big_out
    barrier()
    t0 = time.perf_counter()
    big_out = plugin.enhance_frames_batch(big, cfg)
    e2e_big = world * big.shape[0] / (max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3)
    same_bytes = bool(np.array_equal(e2e_frame0, timed_frame0) and np.array_equal(batch_out[0], timed_frame0)
                      and np.array_equal(big_out[4], timed_frame0))
    del big_out, batch_out

    # ---- per-kernel-class timing (CUDA events around every launch, separate pass so it cannot perturb `value`)
    # Sustained: the profiled steps follow warm steps back to back (a single pass after an idle gap would run at
    # boost clocks and flatter every kernel); the per-class figures are averages over `psteps` steps.
    psteps = max(2, min(args.steps, 5))
    for _ in range(2):
        eng.upscale_device(dev_in, out=dev_out)
    eng.set_option("profile", 1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(psteps):
        eng.upscale_device(dev_in, out=dev_out)
    pe1.record()
    torch.cuda.synchronize()
    prof_step_ms = pe0.elapsed_time(pe1) / psteps
    prof = eng.get_profile()
    eng.set_option("profile", 0)
    for v in prof.values():
        v["ms"] /= psteps
        v["flops"] /= psteps
        v["launches"] //= psteps
    peaks = _peaks()
    dom = max((k for k in prof if prof[k]["flops"] > 0), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
    conv_ms = sum(v["ms"] for v in prof.values())
    # DRAM traffic of the dominant kernel, per launch, from the committed `ncu --set full` capture of this command's
    # kernel at the same frames-per-launch (profiles/r0x_rdb_fused_traffic.json; null if the batch differs)
    traffic = None
    for tname in ("r02d_rdb_fused_traffic.json", "r02b_rdb_fused_traffic.json", "r02_rdb_fused_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if dom == "rdb_fused" and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if int(tj.get("frames_per_launch", -1)) == B:
                traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
                break
    # HBM-bound stages: algorithmic bytes per launch
    lr_px, hr_px = float(B) * H * W, float(B) * H * W * 16
    hbm_bytes = {"first_conv": lr_px * (3 + 128 + 64 + 256), "conv16_last_u8": hr_px * (128 + 3)}
    # conv_hr + conv_last fused: reads conv_up2's output once, writes the u8 frame; the tensor between them stays on chip
    info_bytes = {"hr_last_fused": hr_px * (128 + 3)}
    per_class = {}
    for k, v in prof.items():
        ent = {"ms": round(v["ms"], 3), "launches": v["launches"],
               "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0}
        if k in hbm_bytes and v["ms"] > 0:
            gbs = hbm_bytes[k] / (v["ms"] * 1e-3) / 1e9
            ent.update({"bound": "hbm", "algorithmic_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"]})
        if k in info_bytes and v["ms"] > 0:
            gbs = info_bytes[k] / (v["ms"] * 1e-3) / 1e9
            ent.update({"bound": "tensor", "algorithmic_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"]})
        per_class[k] = ent
    # the kernel is timed inside a long step (69 back-to-back launches under the power cap): sustained peak
    peak = peaks["bf16_tflops_sustained"]
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
        "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu)",
        "frac_vs_burst_peak": ach / peaks["bf16_tflops"],
        "algorithmic_flops_per_launch": d["flops"] / d["launches"],
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}; kernel timed inside a long step; "
                       f"burst figure {peaks['bf16_tflops']})",
        "launches": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
        "share_of_step": d["ms"] / conv_ms,
        "kernel_ms_per_step": conv_ms, "profiled_step_ms": prof_step_ms,
        "idle_between_kernels_ms": prof_step_ms - conv_ms,
        "per_class": per_class,
    }
    flops_frame = 2.0 * MODEL_ARCHS[MODEL].macs_per_input_pixel() * H * W
    whole = fps / world * flops_frame / 1e12

    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "dtype_detail": "bf16 RRDB trunk (residual stream bf16, plus an e5m2 lo part at the RRDB boundaries), fp16 HR tail, fp32 accumulate",
        "data": "synthetic",
        "config": workload_config(B, world),
        "value_batch1": fps_batch1,
        "e2e": {"value": e2e_2t, "unit": UNIT,
                "call": "framewright_b200.pytorch_realesrgan.get_upsampler(cfg).enhance(img, outscale=4): one pageable "
                        "1280x720 numpy frame in, one 5120x2880 numpy frame out per call, 2 caller threads "
                        "(parallel_frames = 2, restorer.py:1894)",
                "h2d_bytes_per_step": int(frames.nbytes), "d2h_bytes_per_step": int(frames.nbytes) * 16,
                "frames": e2e_frames,
                "variants": {"enhance_1_thread": e2e_1t, "enhance_2_threads": e2e_2t, "enhance_4_threads": e2e_4t,
                             f"enhance_frames_batch_{B}": e2e_batch, "enhance_frames_batch_16": e2e_big},
                "outputs_equal_device_path": same_bytes},
        "gpu_launches": launches_per_step * args.steps,
        "output_megapixels_per_sec": fps * (H * 4 * W * 4) / 1e6,
        "tensor_tflops_whole_step_per_gpu": whole,
        "tensor_frac_whole_step": whole / peaks["bf16_tflops"],
        "tensor_frac_whole_step_vs_sustained": whole / peaks["bf16_tflops_sustained"],
        "roofline": roofline,
        "clocks": clk.summary(),
    }
    if rank == 0 and world == 1 and not args.no_configs:
        line["configs"] = bench_configs(torch, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the fp32 oracle on frame 0 of the timed batch: parity of the benchmarked output + the CPU baseline
        from oracle import oracle

        up, cores = _cpu_oracle()
        up.enhance(np.ascontiguousarray(frames[0][:32, :256]), outscale=4)      # thread-pool warm-up
        t0 = time.perf_counter()
        ref0 = up.enhance(frames[0], outscale=4)[0]
        cdt = time.perf_counter() - t0
        rep = oracle.parity_report(ref0, timed_frame0)
        line["parity"] = {"frac_within_1lsb": rep["frac_within_1lsb"], "psnr_db": rep["psnr_db"],
                          "max_abs": rep["max_abs"], "gate": "frac >= 0.999 and psnr >= 45 dB",
                          "pass": bool(rep["frac_within_1lsb"] >= oracle.GATE_FRAC_WITHIN_1LSB
                                       and rep["psnr_db"] >= oracle.GATE_PSNR_DB),
                          "what": "frame 0 of the timed device-resident output vs the fp32 CPU oracle (5120x2880x3)"}
        try:
            # the CUDA path against a vector made by the REFERENCE'S OWN RRDB code (oracle/ref_pin.py; committed fixture)
            from framewright_b200.archs import make_synthetic_state_dict
            from framewright_b200.engine import B200Engine
            from oracle import ref_pin

            z = np.load(os.path.join(ROOT, "tests", "golden", "reference_made", "RealESRGAN_x4plus_24x28_mixed31.npz"))
            e2 = B200Engine(MODEL, make_synthetic_state_dict(MODEL, 0), gpu_id=local_rank)
            got = e2.upscale_host(np.ascontiguousarray(z["input"]))
            e2.close()
            r2 = oracle.parity_report(ref_pin.quantise(z["net_out"]), got)
            line["parity"]["reference_made_vector"] = {
                "frac_within_1lsb": r2["frac_within_1lsb"], "psnr_db": min(float(r2["psnr_db"]), 999.0),
                "max_abs": r2["max_abs"],
                "what": "CUDA path vs the output of the reference's in-tree ESRGAN generator (aesrgan_face.py:171-268) "
                        "on the same weights, 24x28 frame, quantised per upstream's post-process"}
        except Exception as e:   # never at the cost of the line
            line["parity"]["reference_made_vector"] = {"error": f"{type(e).__name__}: {e}"}
        line["cpu_baseline"] = {"value": 1.0 / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"one full 1280x720 frame (frame 0 of the timed batch), one pass of the fp32 "
                                          f"torch oracle: {cdt:.1f} s; no extrapolation"}
    # ---- the PRODUCT scheduler over the same N GPUs (strong scaling of a fixed clip): rank 0 drives
    # MultiGPUDistributor.distribute_frames with one persistent worker process per GPU; the other ranks have released
    # their engines and wait.  (`--workload clip2000` runs the full 2000-frame clip the same way.)
    plugin.clear_upsampler_cache()
    del dev_in, dev_out
    torch.cuda.empty_cache()
    barrier()
    sched_hung = False
    if rank == 0 and args.sched_frames > 0:
        # the contract line must not depend on this leg: an exception becomes an `error` entry, and a leg that does not
        # come back within --sched-timeout seconds is abandoned (the line is printed, then the process leaves hard)
        import threading

        box = {}

        def leg():
            try:
                from tools.clip_bench import run_clip_job

                box["res"] = run_clip_job(world, args.sched_frames, batch=2, threads=3)
            except BaseException as e:
                box["res"] = {"error": f"{type(e).__name__}: {e}"}

        th = threading.Thread(target=leg, name="product-scheduler-leg", daemon=True)
        th.start()
        th.join(timeout=args.sched_timeout)
        sched_hung = th.is_alive()
        line["product_scheduler"] = {"error": f"no result within {args.sched_timeout} s"} if sched_hung else box["res"]
    if world > 1:
        # host-side rendezvous (an NCCL barrier would park a spinning kernel on the GPUs the scheduler is using)
        from datetime import timedelta

        from torch.distributed import distributed_c10d as c10d

        store = c10d._get_default_store()
        if rank == 0:
            store.set("b200sr_sched_done", "1")
        else:
            store.wait(["b200sr_sched_done"], timedelta(seconds=900))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if sched_hung:
        os._exit(0)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4, help="720p frames per step per GPU")
    ap.add_argument("--streams", type=int, default=1, help="CUDA streams the timed steps alternate between "
                    "(2 overlaps the launch tails of consecutive steps: measured +0.3 %%)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="step", choices=["step", "clip2000"])
    ap.add_argument("--clip-frames", type=int, default=2000)
    ap.add_argument("--sched-frames", type=int, default=-1,
                    help="frames of the clip pushed through the product scheduler after the main measurement "
                         "(default: 96 per GPU; 0 = skip)")
    ap.add_argument("--sched-timeout", type=float, default=420.0, help="seconds the product-scheduler leg may take")
    ap.add_argument("--clip-batch", type=int, default=2, help="frames per engine call in the scheduler's workers")
    ap.add_argument("--clip-threads", type=int, default=3, help="runner threads per worker process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    global STREAMS
    STREAMS = args.streams
    if args.sched_frames < 0:
        args.sched_frames = 96 * max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "clip2000":
        from tools.clip_bench import run_clip

        return run_clip(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
