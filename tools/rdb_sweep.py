#!/usr/bin/env python
"""Sweep the fused-RDB work-list step offsets (dev tool)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa
from framewright_b200.archs import make_synthetic_state_dict
from framewright_b200.engine import B200Engine
model = "RealESRGAN_x4plus_anime_6B"
eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(1, 720, 1280, 3), dtype=np.uint8)).cuda()
ref = None
for order, offs in [(1234, (1,2,3,5)), (32104, (1,2,3,5)), (43210, (1,2,3,5)), (32104, (1,2,3,4)), (32104, (1,2,4,6)), (1234, (1,2,3,4)), (32104,(1,1,2,3)), (30214,(1,2,3,5))]:
    eng.set_option("rdb_order", order)
    eng.set_option("rdb_off", offs[0] + 100*offs[1] + 10000*offs[2] + 1000000*offs[3])
    eng.set_option("rdb_stats", 0)
    for _ in range(2): y = eng.upscale_device(x)
    torch.cuda.synchronize()
    if ref is None: ref = y.clone()
    same = bool(torch.equal(ref, y))
    eng.set_option("profile", 1); eng.upscale_device(x); torch.cuda.synchronize(); pr = eng.get_profile(); eng.set_option("profile", 0)
    eng.set_option("rdb_stats", 1); eng.upscale_device(x); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (148 * 16))(); n = eng._lib.b200sr_debug_rdb_stats(eng._h, buf, 148)
    a = np.frombuffer(buf, dtype=np.int64, count=n * 16).reshape(n, 16).astype(np.float64)
    r = pr["rdb_fused"]
    print(f"order={order:05d} off={offs} identical={same} rdb_fused {r['ms']/r['launches']*1e3:7.1f} us/launch {r['flops']/r['ms']/1e9:6.0f} TF/s | last launch: dep-wait {a[:,0].mean()/a[:,9].mean()*100:4.1f}% full-wait {a[:,6].mean()/a[:,9].mean()*100:4.1f}% rempty {a[:,5].mean()/a[:,9].mean()*100:4.1f}% issue {a[:,7].mean()/a[:,9].mean()*100:4.1f}%", flush=True)
eng.close()
