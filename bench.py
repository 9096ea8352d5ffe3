#!/usr/bin/env python
"""Benchmark of the Real-ESRGAN upscaling hot path on B200 (contract: task prompt, section 4).

A "step" is one pass of the hot path (RealESRGAN_x4plus RRDBNet x4, untiled) over one batch of
synthetic 1280x720 uint8 frames -- the configuration BASELINE.json's metric ("RRDBNet x4 frames/sec @720p
input") is quoted on.  With N > 1 GPUs (torchrun, one rank per GPU) every rank runs the same number of
frames (weak scaling, frame sharding, no collective on the data path).

  value      frames/s with inputs already resident in HBM (device-pointer C-ABI call)
  e2e        frames/s through the host-buffer C-ABI call (pinned host in, H2D + forward + D2H inside)
  roofline   dominant kernel (the fused residual-dense-block kernel) vs the measured sustained dense bf16 peak
  cpu_baseline  the fp32 oracle (the reference's CPU path restated) on a bounded crop, rank 0 / N=1 only

`--impl reference` times only the CPU reference path (oracle/; the reference's own dependencies are not
installable here).
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


def synthetic_frames(n: int, seed: int) -> np.ndarray:
    """n synthetic 720p uint8 BGR frames, generated from seed + frame index (no disk)."""
    from oracle.oracle import synthetic_frame  # data generator only (shared with the tests)

    return np.stack([synthetic_frame(H, W, seed=seed * 100003 + i, kind="mixed") for i in range(n)])


# --------------------------------------------------------------------------------------------------
def cpu_reference_fps(sample_hw: int, steps: int, warmup: int):
    """The reference's CPU path (fp32 torch, all host threads) on a bounded crop of the 720p workload."""
    import torch

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_synthetic_state_dict(MODEL, 0)
    up = oracle.make_upsampler(MODEL, sd, tile=0, pre_pad=0)
    crop = oracle.synthetic_frame(sample_hw, sample_hw, seed=4, kind="mixed")
    for _ in range(warmup):
        up.enhance(crop, outscale=4)
    t0 = time.perf_counter()
    for _ in range(steps):
        up.enhance(crop, outscale=4)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    fps_720p = (sample_hw * sample_hw) / (H * W) / dt
    return fps_720p, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 160
    fps, dt, cores = cpu_reference_fps(sample, steps=max(args.steps, 1), warmup=max(min(args.warmup, 2), 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} x4, 1280x720 uint8 frames, untiled (CPU: {sample}x{sample} crop per step, "
                               "frames/s scaled by pixel count)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample}x{sample} crop of a 720p frame per step, fp32 torch oracle, "
                                   f"{dt:.2f} s/step; frames/s = crop_px / frame_px / s"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import framewright_b200  # noqa: F401
    from framewright_b200 import _native
    from framewright_b200.archs import MODEL_ARCHS, make_synthetic_state_dict
    from framewright_b200.engine import B200Engine

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
    B = args.batch
    sd = make_synthetic_state_dict(MODEL, 0)
    eng = B200Engine(MODEL, sd, gpu_id=local_rank)
    frames = synthetic_frames(B, seed=4 + rank)
    dev_in = torch.from_numpy(frames).cuda()
    dev_out = torch.empty((B, H * 4, W * 4, 3), dtype=torch.uint8, device="cuda")
    pin_in = torch.from_numpy(frames).pin_memory()
    pin_out = torch.empty((B, H * 4, W * 4, 3), dtype=torch.uint8).pin_memory()

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

    # ---- device-resident throughput (`value`)
    for _ in range(args.warmup):
        eng.upscale_device(dev_in, out=dev_out)
    launches_per_step = eng.last_launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        for _ in range(args.steps):
            eng.upscale_device(dev_in, out=dev_out)
        ev1.record()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    fps = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer call (`e2e`)
    e2e_steps = max(2, min(args.steps, 5))
    eng.upscale_host(pin_in.numpy(), out=pin_out.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.upscale_host(pin_in.numpy(), out=pin_out.numpy())
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_fps = world * B * e2e_steps / (e2e_ms * 1e-3)
    checksum = int(pin_out.numpy()[0, ::97, ::89].astype(np.int64).sum())

    # ---- per-kernel-class timing (CUDA events around every launch, separate pass so it cannot perturb `value`)
    eng.set_option("profile", 1)
    eng.upscale_device(dev_in, out=dev_out)
    torch.cuda.synchronize()
    prof = eng.get_profile()
    eng.set_option("profile", 0)
    peaks = _peaks()
    dom = max((k for k in prof if prof[k]["flops"] > 0), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
    conv_ms = sum(v["ms"] for v in prof.values())
    # DRAM traffic of the dominant kernel, per launch, from the committed `ncu --set full` capture of this command's
    # kernel at the same frames-per-launch (profiles/r01_rdb_fused_traffic.json; null if the batch differs)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_rdb_fused_traffic.json")
    if dom == "rdb_fused" and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if int(tj.get("frames_per_launch", -1)) == B:
            traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
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
        "per_class": {k: {"ms": round(v["ms"], 3), "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0,
                          "launches": v["launches"]} for k, v in prof.items()},
    }
    flops_frame = 2.0 * MODEL_ARCHS[MODEL].macs_per_input_pixel() * H * W
    whole = fps / world * flops_frame / 1e12

    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "dtype_detail": "bf16 RRDB trunk (residual stream bf16 hi + e5m2 lo), fp16 HR tail, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"{MODEL} x4 on {B} synthetic 1280x720 uint8 frames per step per GPU, untiled, "
                               "random-init weights (seed 0)",
                   "frames_per_step_per_gpu": B, "l2": "per-step working set (>5 GB of activations) exceeds the 126 MB L2",
                   "parallelism": f"frame-sharded x{world}, no collective"},
        "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
                "d2h_bytes_per_step": int(pin_out.numpy().nbytes), "steps": e2e_steps, "output_checksum": checksum},
        "gpu_launches": launches_per_step * args.steps,
        "output_megapixels_per_sec": fps * (H * 4 * W * 4) / 1e6,
        "tensor_tflops_whole_step_per_gpu": whole,
        "tensor_frac_whole_step": whole / peaks["bf16_tflops"],
        "tensor_frac_whole_step_vs_sustained": whole / peaks["bf16_tflops_sustained"],
        "roofline": roofline,
        "clocks": clk.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cfps, cdt, cores = cpu_reference_fps(128, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": cfps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"128x128 crop of a 720p frame, 2 timed passes of the fp32 torch oracle "
                                          f"({cdt:.2f} s each); frames/s = crop_px / frame_px / s"}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4, help="720p frames per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
