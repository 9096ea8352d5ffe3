#!/usr/bin/env python
"""Where the fused-RDB kernel's warps spend their cycles (dev tool): python tools/rdb_stats.py [N H W]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402

N, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 720, 1280)
model = os.environ.get("MODEL", "RealESRGAN_x4plus_anime_6B")   # LAUNCH < 0: counters only, no trace stamps
eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)).cuda()
eng.upscale_device(x)
eng.set_option("rdb_stats", int(os.environ.get("LAUNCH", "2")))
eng.upscale_device(x)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (148 * 16))()
n = eng._lib.b200sr_debug_rdb_stats(eng._h, buf, 148)
a = np.frombuffer(buf, dtype=np.int64, count=n * 16).reshape(n, 16).astype(np.float64)
names = ["P dep-wait", "P empty-wait", "P wempty-wait", "P total", "M wfull-wait", "M rempty-wait", "M full-wait",
         "M issue(blocked in MMA/commit)", "M stages", "M total", "E(warp2) rfull-wait", "E total",
         "E conv5 tcgen05.ld", "E conv5 math+global", "E conv1-4 ld+math+global", "E packed(release|rows5|rows14)"]
tot = a[:, 9].mean()
print(f"last fused RDB launch, {n} CTAs, mean MMA-warp lifetime {tot:.0f} cycles")
for i, nm in enumerate(names):
    print(f"  {nm:32s} mean {a[:, i].mean():12.0f}  min {a[:, i].min():12.0f}  max {a[:, i].max():12.0f}   {a[:, i].mean() / tot * 100:5.1f} % of MMA lifetime")
print(f"  issue cycles per stage: {a[:, 7].sum() / a[:, 8].sum():.0f}")
pk = a[:, 15]; rel = pk // 1000000; r5 = (pk // 1000) % 1000; r14 = pk % 1000
print(f"  epilogue warp 2: conv5 rows {r5.mean():.0f}: ld {a[:,12].sum()/r5.sum():.0f} cyc/row, math+global {a[:,13].sum()/r5.sum():.0f} cyc/row; conv1-4 rows {r14.mean():.0f}: {a[:,14].sum()/r14.sum():.0f} cyc/row; release total {rel.mean():.0f} cyc")
eng.close()
