#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/oracle.py) -- TEST INFRASTRUCTURE.

These vectors are produced by the ORACLE ITSELF (regression vectors: the reference's own dependencies are not
installable here and its tests hold no vectors for this path, see oracle/oracle.py); they pin the oracle and
the synthetic-weight recipe against drift and give the GPU tests a CPU-free comparison target.  The REFERENCE-made
vectors (the RRDBNet network through the reference's in-tree ESRGAN generator) are oracle/ref_pin.py's,
under tests/golden/reference_made/.

    python oracle/gen_golden.py          # rewrites tests/golden/
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200.archs import make_synthetic_state_dict  # noqa: E402
from oracle import oracle  # noqa: E402

# name, h, w, frame kind, frame seed, tile, tile_pad, pre_pad
CASES = [
    ("RealESRGAN_x4plus", 40, 56, "mixed", 11, 0, 10, 0),
    ("RealESRGAN_x4plus", 72, 88, "noise", 12, 32, 6, 0),       # tile mode, 3x3 ragged tiles
    ("RealESRGAN_x4plus", 36, 44, "mixed", 13, 0, 10, 10),      # pre_pad (cli.py:742-750 variant)
    ("RealESRGAN_x4plus_anime_6B", 48, 48, "mixed", 14, 0, 10, 0),
    ("RealESRGAN_x2plus", 45, 63, "mixed", 15, 0, 10, 0),       # odd size -> reflect mod-pad
    ("RealESRGAN_x2plus", 64, 80, "noise", 16, 32, 4, 0),
    ("realesr-general-x4v3", 40, 52, "mixed", 17, 0, 10, 0),
    ("realesr-animevideov3", 50, 38, "noise", 18, 24, 5, 3),
]


def case_name(c):
    name, h, w, kind, seed, tile, tile_pad, pre_pad = c
    return f"{name}_{h}x{w}_{kind}{seed}_t{tile}p{tile_pad}pp{pre_pad}"


def main():
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for c in CASES:
        name, h, w, kind, seed, tile, tile_pad, pre_pad = c
        sd = make_synthetic_state_dict(name, 0)
        img = oracle.synthetic_frame(h, w, seed=seed, kind=kind)
        out, mode = oracle.make_upsampler(name, sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(img)
        assert mode == "RGB"
        path = os.path.join(out_dir, case_name(c) + ".npz")
        np.savez_compressed(path, input=img, output=out, meta=np.array([h, w, seed, tile, tile_pad, pre_pad]))
        print(path, out.shape, os.path.getsize(path))


if __name__ == "__main__":
    main()
