#!/usr/bin/env python
"""`bench.py --workload clip2000`: BASELINE config 5 as written -- a 2000-frame synthetic 1280x720 clip pushed through
the PRODUCT scheduler (`MultiGPUDistributor.distribute_frames`, no `process_fn`): persistent worker process per GPU,
contiguous shards + tail stealing, frames generated in the workers from seed + frame index (no disk), results copied
back to host memory and reduced to a checksum by the sink.  Strong scaling: the clip is fixed, `--gpus N` varies.

Timing: wall clock in the parent from job submission to the last frame's completion message (what a caller of the
scheduler observes; includes the per-frame messaging).  The workers are warm (a short job runs first) so that
process start-up, engine construction and workspace allocation are not part of the figure."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class SyntheticClip:
    """Frame i of a synthetic clip, generated where it is needed from (seed, i): one of `pool` smooth-gradient +
    texture + noise base frames, rolled vertically by i // pool rows (cheap enough to run at hundreds of frames/s)."""

    def __init__(self, n: int, h: int = 720, w: int = 1280, seed: int = 4, pool: int = 8):
        self.n, self.h, self.w, self.seed, self.pool = n, h, w, seed, pool
        self._frames = None

    def __len__(self) -> int:
        return self.n

    def name(self, i: int) -> str:
        return f"frame_{i + 1:08d}.png"

    _CACHE: dict = {}     # per process: (h, w, seed, pool) -> base frames (a worker serves several jobs)

    def open(self) -> None:
        if self._frames is not None:
            return
        key = (self.h, self.w, self.seed, self.pool)
        if key in SyntheticClip._CACHE:
            self._frames = SyntheticClip._CACHE[key]
            return
        yy, xx = np.mgrid[0:self.h, 0:self.w].astype(np.float32)
        frames = []
        for k in range(self.pool):
            rng = np.random.default_rng(self.seed * 100003 + k)
            img = np.empty((self.h, self.w, 3), np.float32)
            for c in range(3):
                fx, fy = rng.uniform(0.5, 3.0, 2)
                ph = rng.uniform(0, 6.28, 2)
                img[:, :, c] = 0.5 + 0.35 * np.sin(fx * 6.28 * xx / self.w + ph[0]) * np.cos(fy * 6.28 * yy / self.h + ph[1])
            img += 0.08 * np.sin(xx[..., None] * 0.9) * np.sin(yy[..., None] * 1.1)
            img += rng.normal(0, 0.04, size=img.shape).astype(np.float32)
            frames.append(np.clip(img * 255.0, 0, 255).round().astype(np.uint8))
        self._frames = SyntheticClip._CACHE[key] = frames

    def close(self) -> None:
        pass

    def load(self, i: int) -> np.ndarray:
        if self._frames is None:
            self.open()
        return np.roll(self._frames[i % self.pool], (i // self.pool) % self.h, axis=0)

    def __getstate__(self):
        return {"n": self.n, "h": self.h, "w": self.w, "seed": self.seed, "pool": self.pool}

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._frames = None


def run_clip_job(n_gpus: int, frames: int, batch: int = 2, threads: int = 3, model: str = "RealESRGAN_x4plus") -> dict:
    """One clip of `frames` synthetic 720p frames through MultiGPUDistributor.distribute_frames on the first `n_gpus`
    GPUs (workers warmed by a short job first).  Returns the measurement as a dict."""
    import torch

    import framewright_b200  # noqa: F401
    from framewright_b200 import multi_gpu as mg
    from framewright_b200.scheduler import ChecksumSink

    os.environ["B200SR_SYNTHETIC_WEIGHTS"] = "0"      # inherited by the worker processes (explicit opt-in)
    n_gpus = max(1, min(n_gpus, torch.cuda.device_count()))
    gpus = mg.query_gpus()[:n_gpus]
    h, w = 720, 1280
    d = mg.MultiGPUDistributor(gpus=gpus, strategy=mg.LoadBalanceStrategy.ROUND_ROBIN, workers_per_gpu=threads - 1,
                               batch=batch, model_name=model, scale=4, tile=0)
    t_start = time.time()
    try:
        # warm-up job on the same clip (same seed): engines, workspaces, pinned pools and the workers' base frames
        warm = d.distribute_frames(SyntheticClip(8 * n_gpus, h, w, seed=4), None, None, sink=ChecksumSink())
        assert not warm.errors, warm.errors
        startup_s = time.time() - t_start
        t0 = time.time()
        res = d.distribute_frames(SyntheticClip(frames, h, w, seed=4), None, None, sink=ChecksumSink())
        dt = time.time() - t0
        rr = d.last_run
    finally:
        d.close()
    ok = res.total_frames
    return {"value": ok / dt, "unit": "frames/s", "n_gpus": n_gpus, "frames": frames, "frames_ok": ok,
            "frames_failed": len(res.errors), "frames_per_gpu": {str(g): len(v) for g, v in res.frames_per_gpu.items()},
            "stolen": rr.stolen, "retried": len(rr.retried), "wall_s": dt, "startup_s": startup_s,
            "batch": batch, "threads_per_worker": threads,
            "output_checksum": int(sum(v[1] for v in rr.ok.values()) % (1 << 61)),
            "what": f"{model} x4 over a {frames}-frame synthetic 1280x720 clip through MultiGPUDistributor.distribute_frames "
                    "(persistent worker process per GPU, contiguous shards + tail stealing, frames generated in the "
                    "workers from seed + index, results copied to host memory)",
            "timing": "parent wall clock, job submission -> last completion message; workers warm"}


def run_clip(args) -> int:
    if int(os.environ.get("RANK", "0")) != 0:     # launched under torchrun: one scheduler drives all GPUs
        return 0
    m = run_clip_job(args.gpus, args.clip_frames, batch=args.clip_batch, threads=args.clip_threads)
    h, w = 720, 1280
    line = {
        "metric": "frames_per_sec_rrdbnet_x4_720p", "value": m["value"], "unit": "frames/s", "n_gpus": m["n_gpus"],
        "steps": 1, "warmup": 1, "ms_per_step": m["wall_s"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": m["what"], "frames": args.clip_frames, "batch": args.clip_batch,
                   "threads_per_worker": args.clip_threads, "parallelism": f"frame-sharded x{m['n_gpus']}, no collective"},
        "e2e": {"value": m["value"], "unit": "frames/s", "h2d_bytes_per_step": args.clip_frames * h * w * 3,
                "d2h_bytes_per_step": args.clip_frames * h * w * 3 * 16},
    }
    line.update({k: m[k] for k in ("frames_ok", "frames_failed", "frames_per_gpu", "stolen", "retried", "wall_s",
                                   "startup_s", "timing", "output_checksum")})
    print(json.dumps(line))
    return 0 if m["frames_ok"] == args.clip_frames else 1
