#!/usr/bin/env python
"""Sweep fused-RDB step offsets and batch size on the full x4plus model at 720p (dev tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa
from framewright_b200.archs import make_synthetic_state_dict
from framewright_b200.engine import B200Engine
model = "RealESRGAN_x4plus"
eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
rng = np.random.default_rng(0)
def run(N, offs, reps=2, inter=0):
    eng.set_option("rdb_interleave", inter)
    x = torch.from_numpy(rng.integers(0, 256, size=(N, 720, 1280, 3), dtype=np.uint8)).cuda()
    eng.set_option("rdb_off", offs[0] + 100*offs[1] + 10000*offs[2] + 1000000*offs[3])
    for _ in range(2): y = eng.upscale_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): y = eng.upscale_device(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"N={N} off={offs} interleave={inter}: {ms:8.2f} ms/step  {N/ms*1e3:6.2f} frames/s", flush=True)
for order in (1234, 40123, 43210, 4123, 1234, 40123):   # digits = conv order within a step (leading 0 dropped)
    eng.set_option("rdb_order", order)
    print("order", f"{order:05d}", end=" ")
    run(4, (1,2,3,5), inter=-1)
eng.close()
