#!/usr/bin/env python
"""Adversarial test of the fused-RDB cross-CTA protocol (compute-sanitizer is closed on this pool).

  python tools/stress_fused.py [--iters 60] [--seed 7] [--big]

Every iteration draws a random shape (ragged heights / widths around the strip, flag-block and column-tile
boundaries), a random batch and a random `max_ctas` between 1 and 148 (fewer resident CTAs than work items, down to
ONE CTA that must run the whole dependency graph by itself in claim order), runs the per-conv schedule (and the
per-row conv kernel) as the reference, the fused schedule + the row-pair single-chunk kernel twice, and compares bytes;
the stacked-kx conv_last (another fp32 summation order) must stay within 1 LSB.  Run it with B200SR_LIB=libb200sr_debug.so to execute
the build with bounds traps on every flag index, item field and TMA coordinate (`-DB200SR_DEBUG`).
`--big` adds the full-size shapes (4 x 720p, 1080p x2, the 720p tile regions)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200 import _native  # noqa: E402
from framewright_b200.archs import make_synthetic_state_dict  # noqa: E402
from framewright_b200.engine import B200Engine  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    _native.build()
    print("library:", _native.load().b200sr_version().decode(), flush=True)
    rng = np.random.default_rng(args.seed)
    bad = 0
    model = "RealESRGAN_x4plus_anime_6B"      # 18 fused launches per forward: the protocol, not the depth, is under test
    eng = B200Engine(model, make_synthetic_state_dict(model, 0), gpu_id=0)
    edges = [7, 8, 9, 15, 16, 17, 23, 24, 25, 31, 32, 33, 47, 48, 49, 63, 64, 65]
    wedges = [8, 31, 127, 128, 129, 255, 256, 257, 300, 383, 385]
    for it in range(args.iters):
        n = int(rng.integers(1, 5))
        h = int(rng.choice(edges)) if rng.random() < 0.5 else int(rng.integers(8, 200))
        w = int(rng.choice(wedges)) if rng.random() < 0.5 else int(rng.integers(8, 700))
        ctas = int(rng.choice([1, 2, 3, 5, 8, 17, 37, 74, 147, 148])) if rng.random() < 0.7 else 0
        x = torch.from_numpy(rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)).cuda()
        # reference: five per-conv launches per RDB and the per-row conv kernel everywhere (pair = 0); under test: the
        # fused RDB kernel + the row-pair single-chunk kernel (bit-identical), then the stacked-kx conv_last on top
        # (different fp32 summation order: within 1 LSB)
        eng.set_option("max_ctas", 0)
        eng.set_option("fused_rdb", 0)
        eng.set_option("pair", 0)
        eng.set_option("last9", 0)
        ref = eng.upscale_device(x).clone()
        eng.set_option("fused_rdb", 1)
        eng.set_option("pair", 1)
        eng.set_option("max_ctas", ctas)
        ok = all(torch.equal(eng.upscale_device(x), ref) for _ in range(2))
        eng.set_option("last9", 1)
        eng.set_option("fuse_tail", 0)
        y9 = eng.upscale_device(x).clone()
        d9 = (y9.to(torch.int16) - ref.to(torch.int16)).abs().max().item()
        eng.set_option("fuse_tail", 1)                      # conv_hr + conv_last as one rolling kernel: same bytes
        ok = ok and d9 <= 1 and torch.equal(eng.upscale_device(x), y9)
        torch.cuda.synchronize()
        bad += 0 if ok else 1
        print(f"[{it:3d}] {n}x{h}x{w} max_ctas={ctas or 148}: {'ok' if ok else 'MISMATCH'}", flush=True)
        del x, ref
    eng.set_option("max_ctas", 0)
    eng.close()
    if args.big:
        for mdl, shapes in (("RealESRGAN_x4plus", [(4, 720, 1280)]),
                            ("RealESRGAN_x4plus_anime_6B", [(8, 256, 256), (2, 522, 532), (3, 218, 266), (5, 40, 1000)]),
                            ("RealESRGAN_x2plus", [(1, 1080, 1920)])):
            eng = B200Engine(mdl, make_synthetic_state_dict(mdl, 0), gpu_id=0)
            for n, h, w in shapes:
                x = torch.from_numpy(rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)).cuda()
                eng.set_option("fused_rdb", 0)
                ref = eng.upscale_device(x).clone()
                eng.set_option("fused_rdb", 1)
                mism = sum(0 if torch.equal(eng.upscale_device(x), ref) else 1 for _ in range(5))
                bad += mism
                print(f"{mdl} {n}x{h}x{w}: 5 fused runs, {mism} mismatching", flush=True)
                del x, ref
            eng.close()
            torch.cuda.empty_cache()
    print("STRESS", "FAILED" if bad else "ok", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
