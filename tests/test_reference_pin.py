"""The RRDBNet arithmetic pinned against the REFERENCE'S OWN in-tree ESRGAN generator (oracle/ref_pin.py).

`/root/reference/src/framewright/processors/aesrgan_face.py:171-268` defines `ResidualDenseBlock`, `RRDB` and the
`AESRGAN` trunk + upsampling tail; with its attention gate at the constructed value (gamma = 0: identity) it is
`RRDBNet(scale=4)` by the reference's own lines.  `tests/golden/reference_made/*.npz` hold what THAT code computed
(float32 network output) for seeded frames and the synthetic checkpoints -- reference-made vectors, not oracle-made.

  * CPU, always:  the oracle's network forward and its whole `enhance` against the committed reference-made vectors;
  * CPU, where /root/reference is mounted:  the reference module run live, bit-compared with the oracle's `RRDBNet`,
    and the committed vectors re-derived (the fixtures are what the generator produces);
  * GPU:  the CUDA path (C ABI) against the reference-made vectors under the BASELINE gate.

  * the frame path: the oracle's `enhance` against the reference's own pre- / post-processing around the network
    (`_enhance_face` :516-542), which has upstream's conventions except that it truncates where upstream rounds.

What this does not pin (no reference code restates it): upstream's final `round`, `pre_pad` / mod-pad, the tile loop, the
gray / alpha / 16-bit branches, and `SRVGGNetCompact`.
"""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "reference_made", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]


def _case(path):
    import framewright_b200  # noqa: F401
    from framewright_b200.archs import MODEL_ARCHS

    base = os.path.basename(path)[:-4]
    name = next(n for n in sorted(MODEL_ARCHS, key=len, reverse=True) if base.startswith(n + "_"))
    z = np.load(path)
    nb, cin, h, w, seed = [int(v) for v in z["meta"]]
    return name, nb, cin, z["input"], z["net_out"]


def test_fixtures_exist_and_cover_the_rrdb_family():
    from oracle import ref_pin

    assert len(FIXTURES) == len(ref_pin.CASES) >= 4
    names = {_case(p)[0] for p in FIXTURES}
    assert names == {"RealESRGAN_x4plus", "RealESRGAN_x4plus_anime_6B", "RealESRGAN_x2plus"}


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_oracle_network_reproduces_the_reference_made_vector(path):
    """oracle.RRDBNet (incl. its own pixel-unshuffle for the x2 model) vs what the reference's module computed.
    fp32 on CPU: another host's oneDNN kernels may order the sums differently, hence 2e-5 instead of equality."""
    import torch

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle, ref_pin

    name, nb, cin, img, want = _case(path)
    model, scale = oracle.build_model(name)
    model.load_state_dict(make_synthetic_state_dict(name, 0), strict=True)
    x = ref_pin.network_input(img, 3)                       # the oracle's forward does the unshuffle itself
    with torch.no_grad():
        got = model.eval()(x).squeeze(0).numpy()
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= 2e-5


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_oracle_enhance_reproduces_the_quantised_reference_made_vector(path):
    """The oracle's whole `RealESRGANer.enhance` (pre-process, network, post-process, uint8) against the reference-made
    network output quantised per upstream's post-process (clamp, RGB -> BGR, round(x * 255))."""
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle, ref_pin

    name, nb, cin, img, net_out = _case(path)
    out, mode = oracle.make_upsampler(name, make_synthetic_state_dict(name, 0)).enhance(img)
    want = ref_pin.quantise(net_out)
    assert mode == "RGB" and out.shape == want.shape and out.dtype == np.uint8
    rep = oracle.parity_report(want, out)
    assert rep["max_abs"] <= 1 and rep["frac_exact"] >= 0.9995, rep      # a value within 1e-5 of x.5 may round the other way


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/framewright/processors/aesrgan_face.py"),
                    reason="reference tree not mounted")
def test_reference_module_live_equals_oracle_and_fixtures():
    """The unmodified reference file, imported from where it lies, in this process: the oracle's RRDBNet gives the
    same bits (same op order through the same torch kernels), and the committed fixtures are what the generator makes."""
    import torch

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle, ref_pin

    mod = ref_pin.load_reference_module()
    for c in ref_pin.CASES:
        name, nb, cin = c[:3]
        img, ref_out, ref_frame = ref_pin.run_case(mod, c)
        model, _ = oracle.build_model(name)
        model.load_state_dict(make_synthetic_state_dict(name, 0), strict=True)
        with torch.no_grad():
            ours = model.eval()(ref_pin.network_input(img, 3)).squeeze(0).numpy()
        assert np.array_equal(ours, ref_out), (name, float(np.abs(ours - ref_out).max()))
        z = np.load(os.path.join(ref_pin.OUT_DIR, ref_pin.case_name(c) + ".npz"))
        assert np.array_equal(z["input"], img)
        assert float(np.abs(z["net_out"] - ref_out).max()) <= 2e-5
        if ref_frame is not None:                       # the reference's own frame path, live
            assert np.array_equal(z["frame_out_truncated"], ref_frame)


@pytest.mark.parametrize("path", [p for p in FIXTURES if "x2plus" not in p], ids=[i for i in IDS if "x2plus" not in i])
def test_oracle_frame_path_against_the_reference_frame_path(path):
    """Frame in -> frame out.  `frame_out_truncated` is what the reference's own pre- / post-processing around the
    network produced (`AESRGANFaceRestorer._enhance_face`, aesrgan_face.py:516-542: BGR -> RGB, / 255, NCHW, model, HWC,
    `np.clip(x * 255, 0, 255).astype(uint8)`, RGB -> BGR).  `RealESRGANer.enhance` has the same conventions but ROUNDS
    where the reference truncates: the oracle's frame must be the reference's, or the reference's + 1 exactly where
    the fraction of x * 255 is >= 0.5 -- which pins channel order, normalisation and layout on both sides of the net."""
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    name, nb, cin, img, net_out = _case(path)
    ref_frame = np.load(path)["frame_out_truncated"]
    ours, mode = oracle.make_upsampler(name, make_synthetic_state_dict(name, 0)).enhance(img)
    assert mode == "RGB" and ours.shape == ref_frame.shape == (img.shape[0] * 4, img.shape[1] * 4, 3)
    d = ours.astype(np.int16) - ref_frame.astype(np.int16)
    assert d.min() >= 0 and d.max() <= 1
    v = np.clip(net_out, 0.0, 1.0)[[2, 1, 0]].transpose(1, 2, 0).astype(np.float64) * 255.0     # BGR, HWC
    frac = v - np.floor(v)
    sure = np.abs(frac - 0.5) > 1e-3                    # away from the rounding boundary (fp32 summation order)
    assert np.array_equal(d[sure] == 1, frac[sure] > 0.5)
    assert 0.3 < float((d == 1).mean()) < 0.7           # about half the pixels round up: the relation is not vacuous
    # a channel-order or layout slip on either side would break it: the same check with R and B swapped fails
    assert (np.abs(ours[:, :, ::-1].astype(np.int16) - ref_frame.astype(np.int16)) > 1).mean() > 0.05


def test_emulated_engine_rounding_passes_the_gate_on_the_reference_made_vectors():
    """No GPU here: the CPU emulation of the engine's rounding points (tests/emulate.py) predicts that the CUDA path
    clears the gate against the reference-made vectors (the GPU test below is the real check)."""
    import sys

    import torch

    sys.path.insert(0, HERE)
    from emulate import emulate_rrdb

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle, ref_pin

    for path in FIXTURES:
        name, nb, cin, img, net_out = _case(path)
        got = emulate_rrdb(make_synthetic_state_dict(name, 0), img, scale=2 if cin == 12 else 4, num_block=nb,
                           tail_dtype=torch.float16, tail_w_dtype=torch.float16)
        rep = oracle.parity_report(ref_pin.quantise(net_out), got)
        assert rep["frac_within_1lsb"] >= oracle.GATE_FRAC_WITHIN_1LSB and rep["psnr_db"] >= oracle.GATE_PSNR_DB, (name, rep)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_gpu_against_reference_made_vector(native_lib, path):
    """The CUDA path against vectors the reference's own RRDB code produced (gate: >= 99.9 % within 1 LSB, >= 45 dB)."""
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import B200Engine
    from oracle import oracle, ref_pin

    name, nb, cin, img, net_out = _case(path)
    eng = B200Engine(name, make_synthetic_state_dict(name, 0), gpu_id=0)
    got = eng.upscale_host(np.ascontiguousarray(img))
    eng.close()
    want = ref_pin.quantise(net_out)
    rep = oracle.parity_report(want, got)
    print(os.path.basename(path), rep)
    assert got.shape == want.shape
    assert rep["frac_within_1lsb"] >= oracle.GATE_FRAC_WITHIN_1LSB, rep
    assert rep["psnr_db"] >= oracle.GATE_PSNR_DB, rep


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/framewright/metrics.py"), reason="reference tree not mounted")
def test_gate_metric_is_the_references_psnr():
    """The gate's PSNR (`oracle.calculate_psnr`, used by `parity_report`) against the reference's own
    `calculate_psnr` (`metrics.py:433-458`, the module loaded unmodified): same number on the same pair of images."""
    import importlib.util
    import sys

    from oracle import oracle

    spec = importlib.util.spec_from_file_location("_ref_metrics", "/root/reference/src/framewright/metrics.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    dont, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = dont
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    for noise in (1, 2, 7, 40):
        b = np.clip(a.astype(np.int16) + rng.integers(-noise, noise + 1, a.shape), 0, 255).astype(np.uint8)
        assert abs(oracle.calculate_psnr(a, b) - mod.calculate_psnr(a, b)) < 1e-9
        assert abs(oracle.parity_report(a, b)["psnr_db"] - mod.calculate_psnr(a, b)) < 1e-9
    assert oracle.calculate_psnr(a, a) == mod.calculate_psnr(a, a) == float("inf")
