#!/usr/bin/env python
"""Device-resident throughput of the BASELINE.json configurations 1-4 (dev tool; bench.py measures config 5's
per-GPU workload).  python tools/configs_bench.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import MODEL_ARCHS, make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402

CFGS = [
    ("cfg1", "RealESRGAN_x4plus", 8, 256, 256, dict()),
    ("cfg2", "realesr-general-x4v3", 64, 480, 640, dict()),
    ("cfg3", "RealESRGAN_x2plus", 2, 1080, 1920, dict()),
    ("cfg4", "RealESRGAN_x4plus", 2, 720, 1280, dict(tile=512, tile_pad=10)),
    ("cfg4b", "RealESRGAN_x4plus", 2, 720, 1280, dict(tile=512, tile_pad=10, pre_pad=10)),
    ("anime6B", "RealESRGAN_x4plus_anime_6B", 4, 720, 1280, dict()),
    ("animevideov3", "realesr-animevideov3", 16, 720, 1280, dict()),
]
rng = np.random.default_rng(0)
INTER = int(os.environ.get("INTERLEAVE", "-1"))
for tag, model, N, H, W, kw in CFGS:
    if os.environ.get("ONLY") and tag not in os.environ["ONLY"].split(","):
        continue
    eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
    if INTER >= 0:
        eng.set_option("rdb_interleave", INTER)
    x = torch.from_numpy(rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)).cuda()
    for _ in range(2):
        y = eng.upscale_device(x, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        y = eng.upscale_device(x, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    arch = MODEL_ARCHS[model]
    px = N * H * W / (4 if arch.scale == 2 and arch.kind == "rrdb" else 1)
    fl = 2.0 * arch.macs_per_input_pixel() * px
    print(f"{tag:13s} {model:28s} N={N:3d} {H}x{W} {kw}: {ms:9.2f} ms/step {N / ms * 1e3:8.2f} frames/s "
          f"{fl / ms / 1e9:7.0f} TFLOP/s (algorithmic, untiled count) out {tuple(y.shape)}", flush=True)
    eng.close()
    del x, y
    torch.cuda.empty_cache()
