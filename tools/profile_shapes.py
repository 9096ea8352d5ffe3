#!/usr/bin/env python
"""Per-kernel-class throughput for several frame shapes (dev tool): python tools/profile_shapes.py N,H,W [N,H,W ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import MODEL_ARCHS, make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402

model = os.environ.get("MODEL", "RealESRGAN_x4plus")
eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    eng.set_option(k, int(v))
rng = np.random.default_rng(0)
for spec in sys.argv[1:]:
    N, H, W = [int(v) for v in spec.split(",")]
    x = torch.from_numpy(rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)).cuda()
    for _ in range(2):
        y = eng.upscale_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        y = eng.upscale_device(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * MODEL_ARCHS[model].macs_per_input_pixel() * N * H * W
    eng.set_option("profile", 1)
    eng.upscale_device(x)
    torch.cuda.synchronize()
    prof = eng.get_profile()
    eng.set_option("profile", 0)
    print(f"== {model} N={N} {H}x{W}: {ms:.2f} ms/step, {N / ms * 1e3:.2f} frames/s, {fl / ms / 1e9:.0f} TFLOP/s whole step")
    for k, v in prof.items():
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
        print(f"   {k:18s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {tf:7.1f} TFLOP/s  {v['ms'] / v['launches'] * 1e3:7.1f} us/launch")
    del x, y
eng.close()
