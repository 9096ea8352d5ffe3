"""BASELINE.json configurations 2-5 at their FULL sizes: the CUDA path against the fp32 CPU oracle computed in-test.

Every other oracle comparison in the suite is <= 256x256; these cases cover what only full size exercises -- offsets
beyond 2^31 bytes, `choose_th` at H = 2880, the six-region 522/532/266 x 522/218 tile plan on the device, all 148
CTAs with ~6000 work items -- under the BASELINE gate (>= 99.9 % of uint8 pixels within 1 LSB, PSNR >= 45 dB).
The oracle costs ~20-30 s per 720p frame on the GPU box's host cores.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(name, seed=0):
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import B200Engine

    sd = make_synthetic_state_dict(name, seed)
    return B200Engine(name, sd, gpu_id=0), sd


def _check(ref, got, label):
    from oracle import oracle

    rep = oracle.parity_report(ref, got)
    print(label, rep)
    assert ref.shape == got.shape
    assert rep["frac_within_1lsb"] >= oracle.GATE_FRAC_WITHIN_1LSB, (label, rep)
    assert rep["psnr_db"] >= oracle.GATE_PSNR_DB, (label, rep)
    return rep


def _oracle(name, sd, img, **kw):
    from oracle import oracle

    torch.set_num_threads(torch.get_num_threads())
    with torch.no_grad():
        return oracle.make_upsampler(name, sd, **kw).enhance(img)[0]


def test_cfg5_720p_untiled_against_oracle(native_lib):
    """Config 5's per-frame workload: RealESRGAN_x4plus, one 1280x720 frame, untiled -> 5120x2880."""
    from oracle import oracle

    eng, sd = _engine("RealESRGAN_x4plus")
    img = oracle.synthetic_frame(720, 1280, seed=4 * 100003, kind="mixed")   # frame 0 of bench.py's clip (seed 4)
    got = eng.upscale_host(img)
    assert got.shape == (2880, 5120, 3)
    ref = _oracle("RealESRGAN_x4plus", sd, img, tile=0, pre_pad=0)
    _check(ref, got, "cfg5 720p untiled")
    # the same frame as member 1 of a 2-frame batch through the device-pointer call
    batch = torch.from_numpy(np.stack([img[::-1].copy(), img])).cuda()
    out = eng.upscale_device(batch)
    torch.cuda.synchronize()
    assert np.array_equal(out[1].cpu().numpy(), got)
    eng.close()


@pytest.mark.parametrize("pre_pad", [0, 10])
def test_cfg4_720p_tile512_against_tiled_oracle_and_seam_exact(native_lib, pre_pad):
    """Config 4: 1280x720, tile=512, tile_pad=10 (pre_pad 0 = PyTorchESRGANConfig default, 10 = cli.py:742-750)
    against the oracle run in tile mode with the same geometry; and "seam-exact": each of the six regions equals a
    standalone run of that padded tile, bit for bit (upstream zero-pads every conv at the padded-tile border)."""
    from framewright_b200 import _native
    from oracle import oracle
    import ctypes

    name = "RealESRGAN_x4plus"
    eng, sd = _engine(name)
    img = oracle.synthetic_frame(720, 1280, seed=3, kind="mixed")
    got = eng.upscale_host(img, tile=512, tile_pad=10, pre_pad=pre_pad)
    assert got.shape == (2880, 5120, 3)
    ref = _oracle(name, sd, img, tile=512, tile_pad=10, pre_pad=pre_pad)
    _check(ref, got, f"cfg4 720p tile512/10 pre_pad {pre_pad}")

    # region plan of the engine (the C ABI's host-only hook) == six regions
    lib = _native.load()
    buf = (ctypes.c_int * (16 * 10))()
    nreg = lib.b200sr_debug_plan_regions(_native.ARCH_RRDB, 4, 720, 1280, 512, 10, pre_pad, buf, 16)
    assert nreg == 6
    padded = np.pad(img, ((0, pre_pad), (0, pre_pad), (0, 0)), mode="reflect") if pre_pad else img
    for r in range(nreg):
        oy, ox, rh, rw, cy0, cx0, ch, cw, dy0, dx0 = [buf[r * 10 + i] for i in range(10)]
        tile_in = np.ascontiguousarray(padded[oy:oy + rh, ox:ox + rw])
        alone = eng.upscale_host(tile_in)                       # standalone, untiled, no pre_pad
        want = alone[cy0:cy0 + ch, cx0:cx0 + cw]
        assert np.array_equal(got[dy0:dy0 + ch, dx0:dx0 + cw], want), f"region {r} differs from its standalone run"
    eng.close()


def test_cfg3_x2plus_1080p_against_oracle(native_lib):
    """Config 3: RealESRGAN_x2plus (pixel-unshuffle input), one 1920x1080 frame -> 3840x2160."""
    from oracle import oracle

    eng, sd = _engine("RealESRGAN_x2plus")
    img = oracle.synthetic_frame(1080, 1920, seed=2, kind="mixed")
    got = eng.upscale_host(img)
    assert got.shape == (2160, 3840, 3)
    ref = _oracle("RealESRGAN_x2plus", sd, img, tile=0, pre_pad=0)
    _check(ref, got, "cfg3 x2plus 1080p")
    eng.close()


def test_cfg2_srvgg_batch64_members_against_oracle(native_lib):
    """Config 2: realesr-general-x4v3 on a batch of 64 640x480 frames; members 0 / 37 / 63 against the oracle."""
    from oracle import oracle

    name = "realesr-general-x4v3"
    eng, sd = _engine(name)
    frames = np.stack([oracle.synthetic_frame(480, 640, seed=1000 + i, kind="mixed" if i % 2 else "noise")
                       for i in range(64)])
    out = eng.upscale_host(frames)
    assert out.shape == (64, 1920, 2560, 3)
    for i in (0, 37, 63):
        ref = _oracle(name, sd, frames[i], tile=0, pre_pad=0)
        _check(ref, out[i], f"cfg2 general-x4v3 batch member {i}")
    eng.close()
