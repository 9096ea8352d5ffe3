"""The CPU oracle against the committed golden fixtures (tests/golden/, made by oracle/gen_golden.py).

These fixtures are the oracle's own outputs: they pin the oracle + synthetic-weight recipe against drift (whole path:
pre / tile / post-processing, every model family).  The reference-made vectors that pin the RRDBNet arithmetic to the
reference's own code are tests/golden/reference_made/ (tests/test_reference_pin.py).  The
oracle is fp32 torch on CPU: results are allowed to differ from the fixture by 1 LSB on <= 0.1 % of pixels
(different oneDNN kernels / thread counts change the fp32 summation order).
"""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _parse(path):
    base = os.path.basename(path)[:-4]
    name = base.split("_" + base.split("_")[-3])[0] if False else None
    return base


def test_fixtures_exist():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    import framewright_b200  # noqa: F401
    from framewright_b200.archs import MODEL_ARCHS, make_synthetic_state_dict
    from oracle import oracle

    base = os.path.basename(path)[:-4]
    name = next(n for n in sorted(MODEL_ARCHS, key=len, reverse=True) if base.startswith(n + "_"))
    z = np.load(path)
    h, w, seed, tile, tile_pad, pre_pad = [int(v) for v in z["meta"]]
    kind = "noise" if "_noise" in base else "mixed"
    img = oracle.synthetic_frame(h, w, seed=seed, kind=kind)
    assert np.array_equal(img, z["input"]), "synthetic frame generator drifted"
    sd = make_synthetic_state_dict(name, 0)
    out, mode = oracle.make_upsampler(name, sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(img)
    s = MODEL_ARCHS[name].scale
    assert mode == "RGB" and out.shape == (h * s, w * s, 3) and out.dtype == np.uint8
    rep = oracle.parity_report(z["output"], out)
    assert rep["max_abs"] <= 1 and rep["frac_exact"] >= 0.999, rep


def test_oracle_tile_geometry_matches_survey():
    """720p, tile=512, tile_pad=10 -> six tiles (522,522) (532,522) (266,522) (522,218) (532,218) (266,218)."""
    import math

    H, W, tile, pad = 720, 1280, 512, 10
    sizes = []
    for y in range(math.ceil(H / tile)):
        for x in range(math.ceil(W / tile)):
            x0, y0 = x * tile, y * tile
            x1, y1 = min(x0 + tile, W), min(y0 + tile, H)
            sizes.append((min(x1 + pad, W) - max(x0 - pad, 0), min(y1 + pad, H) - max(y0 - pad, 0)))
    assert sizes == [(522, 522), (532, 522), (266, 522), (522, 218), (532, 218), (266, 218)]
    assert sum(a * b for a, b in sizes) == 976800


def test_pixel_unshuffle_matches_torch():
    import torch
    import torch.nn.functional as F

    from oracle.oracle import pixel_unshuffle

    x = torch.arange(2 * 3 * 8 * 12, dtype=torch.float32).view(2, 3, 8, 12)
    assert torch.equal(pixel_unshuffle(x, 2), F.pixel_unshuffle(x, 2))
    assert torch.equal(pixel_unshuffle(x, 4), F.pixel_unshuffle(x, 4))


def test_macs_per_pixel_match_survey():
    from framewright_b200.archs import MODEL_ARCHS

    want = {"RealESRGAN_x4plus": 17926848, "RealESRGAN_x2plus": 17932032, "RealESRGAN_x4plus_anime_6B": 5706432,
            "realesr-general-x4v3": 1209024, "realesr-animevideov3": 619200}
    for k, v in want.items():
        assert MODEL_ARCHS[k].macs_per_input_pixel() == v


def test_oracle_gray_rgba_outscale_branches():
    """Upstream enhance(): gray -> 'L', 4-channel -> 'RGBA', outscale != netscale -> LANCZOS4 resize."""
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    sd = make_synthetic_state_dict("realesr-animevideov3", 0)
    up = oracle.make_upsampler("realesr-animevideov3", sd)
    img = oracle.synthetic_frame(20, 24, seed=1)
    g, mode = up.enhance(img[:, :, 0])
    assert mode == "L" and g.shape == (80, 96)
    rgba = np.dstack([img, img[:, :, 1]])
    o, mode = up.enhance(rgba)
    assert mode == "RGBA" and o.shape == (80, 96, 4)
    o, mode = up.enhance(img, outscale=2)
    assert o.shape == (40, 48, 3)
