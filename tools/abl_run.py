#!/usr/bin/env python
"""Dev tool: sustained step time, SM clock and board power of one library build (B200SR_LIB=...): used with the
timing-ablation builds (-DB200SR_ABL_*) to see which component of the fused RDB kernel the power-capped step pays for.
  B200SR_LIB=libb200sr_ablne1.so python tools/abl_run.py [N H W steps]"""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402

N, H, W, steps = [int(v) for v in (sys.argv[1:5] + ["4", "720", "1280", "20"][len(sys.argv) - 1:])]
model = os.environ.get("MODEL", "RealESRGAN_x4plus")
eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    eng.set_option(k, int(v))
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)).cuda()
samples, stop = [], False


def sampler():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.05)


for _ in range(5):
    eng.upscale_device(x)
torch.cuda.synchronize()
th = threading.Thread(target=sampler)
th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.upscale_device(x)
e1.record()
torch.cuda.synchronize()
stop = True
th.join()
ms = e0.elapsed_time(e1) / steps
eng.set_option("profile", 1)
y = eng.upscale_device(x)
torch.cuda.synchronize()
prof = eng.get_profile()
import hashlib
digest = hashlib.md5(y.cpu().numpy().tobytes()).hexdigest()[:12]
mid = samples[len(samples) // 4:] or samples
clk = sorted(s[0] for s in mid)[len(mid) // 2]
pw = sorted(s[1] for s in mid)[len(mid) // 2]
rdb = prof.get("rdb_fused", {"ms": 0, "launches": 1})
print(f"{os.environ.get('B200SR_LIB', 'libb200sr.so'):24s} {ms:8.2f} ms/step  {N / ms * 1e3:6.2f} frames/s  sm {clk} MHz  {pw:6.0f} W  "
      f"rdb_fused {rdb['ms']:.2f} ms / {rdb['launches']} launches  md5 {digest}")
eng.close()
