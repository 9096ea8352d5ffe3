#!/usr/bin/env python
"""Race hunt for the fused-RDB schedule (dev tool): the per-conv path is the reference, the fused path is repeated
many times on full-size and ragged inputs; any dependency race shows up as a byte mismatch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(7)
bad = 0
for model, shapes in (("RealESRGAN_x4plus", [(4, 720, 1280), (1, 720, 1280)]),
                      ("RealESRGAN_x4plus_anime_6B", [(8, 256, 256), (2, 522, 532), (3, 218, 266), (2, 333, 517), (5, 40, 1000)]),
                      ("RealESRGAN_x2plus", [(1, 1080, 1920)])):
    eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
    for n, h, w in shapes:
        x = torch.from_numpy(rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)).cuda()
        eng.set_option("fused_rdb", 0)
        ref = eng.upscale_device(x).clone()
        eng.set_option("fused_rdb", 1)
        mism = 0
        for _ in range(reps):
            got = eng.upscale_device(x)
            if not torch.equal(got, ref):
                mism += 1
        bad += mism
        print(f"{model} {n}x{h}x{w}: {reps} fused runs, {mism} mismatching", flush=True)
        del x, ref
    eng.close()
    torch.cuda.empty_cache()
print("STRESS", "FAILED" if bad else "ok")
sys.exit(1 if bad else 0)
