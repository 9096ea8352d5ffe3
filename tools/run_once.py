#!/usr/bin/env python
"""One forward pass of the hot path (for ncu / compute-sanitizer captures): python tools/run_once.py [model] [H] [W] [N]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "RealESRGAN_x4plus"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 720
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
N = int(sys.argv[4]) if len(sys.argv) > 4 else 1
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:   # e.g. OPTS=rdb_interleave=0,trunk_lo=2
    eng.set_option(k, int(v))
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)).cuda()
for _ in range(reps):
    y = eng.upscale_device(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), "launches", eng.last_launch_count, "checksum", int(y[0, ::37, ::41].sum()))
eng.close()
