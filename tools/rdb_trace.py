#!/usr/bin/env python
"""Per-item timeline of one fused-RDB launch (dev tool; needs a library built with -DB200SR_RDB_STATS)."""
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
eng.upscale_device(x)
eng.set_option("rdb_stats", int(os.environ.get("LAUNCH", "2")))
eng.upscale_device(x); torch.cuda.synchronize()
lib = eng._lib
n = lib.b200sr_debug_rdb_trace(eng._h, None, 0)
buf = (ctypes.c_longlong * (n * 10))()
lib.b200sr_debug_rdb_trace(eng._h, buf, n)
ti = np.frombuffer(buf, dtype=np.int64, count=n * 10).reshape(n, 10).copy()
cta_col = ti[:, 8].copy()
t0i = ti[:, 0].min()
t = np.where(ti > 0, ti - t0i, 0).astype(np.float64)   # ns since the first claim (0 = not stamped)
t[:, 8] = cta_col
ibuf = (ctypes.c_int * (n * 8))()
lib.b200sr_debug_rdb_items(1, 720, 1280, ibuf, n)
items = np.frombuffer(ibuf, dtype=np.int32, count=n * 8).reshape(n, 8)
t0 = 0.0
us = lambda a: a / 1e3
print(f"{n} items, launch span {us(t[:, 7].max() - t0):.1f} us")
claim = t[:, 0] - t0
order_ok = np.all(np.diff(claim) >= -2000)
print("claims monotonic (within 2 us):", bool(order_ok), " max backward step us:", us(-np.diff(claim).min()))
for k in range(5):
    m = items[:, 0] == k
    dur = us(t[m, 7] - t[m, 0]); prod = us(t[m, 5] - t[m, 0]); mma = us(t[m, 6] - t[m, 0])
    d1 = us(np.where(t[m, 1] > 0, t[m, 2] - t[m, 1], 0)); d2 = us(np.where(t[m, 3] > 0, t[m, 4] - t[m, 3], 0))
    s1 = us(np.where(t[m, 1] > 0, t[m, 1] - t[m, 0], 0)); s2 = us(np.where(t[m, 3] > 0, t[m, 3] - t[m, 0], 0))
    print(f"conv{k+1}: {m.sum():4d} items | claim->epilogue done {dur.mean():6.1f} us (p90 {np.percentile(dur,90):6.1f}) | loads issued at {prod.mean():6.1f} | MMAs issued at {mma.mean():6.1f} | "
          f"dep1 reached at {s1.mean():5.1f} waits {d1.mean():5.1f} (p90 {np.percentile(d1,90):5.1f}) | dep2 reached at {s2.mean():5.1f} waits {d2.mean():5.1f} (p90 {np.percentile(d2,90):5.1f})")
# per-CTA utilisation: sum of item durations vs span
cta = t[:, 8].astype(int)
busy = np.zeros(cta.max() + 1); 
for c in range(cta.max() + 1):
    m = cta == c
    busy[c] = (t[m, 6] - t[m, 0]).sum()
print(f"per-CTA items {np.bincount(cta).mean():.1f}; sum(claim->MMA issued) / span: {busy.mean() / (t[:, 7].max() - t0) * 100:.1f} %")
# claim rate over time
edges = np.linspace(0, claim.max(), 11)
print("items claimed per decile of the launch:", np.histogram(claim, edges)[0].tolist())
np.save(os.path.join(ROOT, "gpurun_out", "rdb_trace.npy"), np.concatenate([items.astype(np.float64), t], axis=1))
for c in (0, 77):
    idx = np.where(cta == c)[0]
    print("CTA", c)
    prev = 0.0
    for i in idx[:19]:
        k, fn, y0, rows, tx = items[i, :5]
        tt = t[i] / 1e3
        d1 = (tt[1] - tt[0], tt[2] - tt[1]) if tt[1] > 0 else (0, 0)
        d2 = (tt[3] - tt[0], tt[4] - tt[3]) if tt[3] > 0 else (0, 0)
        print(f"  item {i:5d} conv{k+1} y0={y0:3d} tx={tx} claim {tt[0]:7.1f} (prev mma end {prev - tt[0]:+6.1f}) | mma got item +{tt[9]-tt[0]:5.1f} | dep1 at +{d1[0]:5.1f} wait {d1[1]:5.1f} | dep2 at +{d2[0]:5.1f} wait {d2[1]:5.1f} | loads done +{tt[5]-tt[0]:5.1f} | mma done +{tt[6]-tt[0]:5.1f} | epi done +{tt[7]-tt[0]:5.1f}")
        prev = tt[6]
buf2 = (ctypes.c_longlong * (n * 48))()
lib.b200sr_debug_rdb_trace(eng._h, buf2, -n)
r = np.frombuffer(buf2, dtype=np.int64, count=n * 48).reshape(n, 16, 3).copy()
for i in [int(v) for v in np.where(cta == 0)[0][:6]]:
    k, fn, y0, rows, tx = items[i, :5]
    print(f"item {i} conv{k+1}: claim {t[i,0]/1e3:.1f} mma done {t[i,6]/1e3:.1f} epi done {t[i,7]/1e3:.1f}")
    for Y in range(rows):
        c, w, d = [(v - t0i) / 1e3 if v > 0 else float('nan') for v in r[i, Y]]
        print(f"    row {Y:2d}: mma commit {c:8.1f}  epi(q2) wait passed {w:8.1f}  stored {d:8.1f}")
eng.close()
