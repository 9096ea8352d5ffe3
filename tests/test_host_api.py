"""Host-side mirror of the reference upscale processor API (no GPU).  The first block restates, against the
B200 module, the live tests of /root/reference/tests/test_processors/test_pytorch_realesrgan.py (:59-116, :230-237)
and re-implements the ones the reference cannot run (fixtures `mock_torch`/`mock_cv2` are never registered there)."""
from pathlib import Path

import numpy as np
import pytest

import framewright_b200  # noqa: F401
from framewright_b200 import multi_gpu as mg
from framewright_b200 import pytorch_realesrgan as pr
from framewright_b200 import super_resolution as srm


# ---- PyTorchESRGANConfig (reference tests :59-116) -------------------------------------------------
def test_config_defaults():
    c = pr.PyTorchESRGANConfig()
    assert (c.model_name, c.scale_factor, c.tile_size, c.tile_pad, c.pre_pad, c.half_precision, c.gpu_id) == (
        "RealESRGAN_x4plus", 4, 0, 10, 0, True, 0)


@pytest.mark.parametrize("name", ["RealESRGAN_x4plus", "RealESRGAN_x4plus_anime_6B", "RealESRGAN_x2plus",
                                  "realesr-animevideov3", "realesr-general-x4v3"])
def test_config_valid_models(name):
    pr.PyTorchESRGANConfig(model_name=name).validate()


def test_config_invalid_model():
    with pytest.raises(ValueError, match="Invalid model"):
        pr.PyTorchESRGANConfig(model_name="nope").validate()


@pytest.mark.parametrize("scale", [1, 3, 8])
def test_config_invalid_scale(scale):
    with pytest.raises(ValueError, match="Scale factor must be 2 or 4"):
        pr.PyTorchESRGANConfig(scale_factor=scale).validate()


# ---- name map (reference tests :230-237) -------------------------------------------------------------
def test_ncnn_name_map():
    assert pr.convert_ncnn_model_name("realesrgan-x4plus") == "RealESRGAN_x4plus"
    assert pr.convert_ncnn_model_name("realesrgan-x4plus-anime") == "RealESRGAN_x4plus_anime_6B"
    assert pr.convert_ncnn_model_name("realesr-animevideov3") == "realesr-animevideov3"
    assert pr.convert_ncnn_model_name("realesrnet-x4plus") == "realesr-general-x4v3"
    assert pr.convert_ncnn_model_name("realesrgan-x2plus") == "RealESRGAN_x2plus"
    assert pr.convert_ncnn_model_name("unknown-model") == "RealESRGAN_x4plus"
    assert set(pr.NCNN_TO_PYTORCH_MODEL.values()) == set(pr.VALID_MODELS)


# ---- availability / error convention ---------------------------------------------------------------
def test_availability_is_memoised(monkeypatch):
    monkeypatch.setattr(pr, "_PYTORCH_ESRGAN_AVAILABLE", True)
    assert pr.is_pytorch_esrgan_available() is True
    monkeypatch.setattr(pr, "_PYTORCH_ESRGAN_AVAILABLE", False)
    assert pr.is_pytorch_esrgan_available() is False


def test_get_upsampler_raises_without_engine(monkeypatch):
    monkeypatch.setattr(pr, "_PYTORCH_ESRGAN_AVAILABLE", False)
    with pytest.raises(RuntimeError, match="not available"):
        pr.get_upsampler(pr.PyTorchESRGANConfig())


def test_enhance_frame_never_raises_unreadable_input(tmp_path):
    ok, err = pr.enhance_frame_pytorch(tmp_path / "missing.png", tmp_path / "out.png", pr.PyTorchESRGANConfig())
    assert ok is False and err.startswith("Failed to read image:")


def test_enhance_frame_reports_invalid_config(tmp_path):
    ok, err = pr.enhance_frame_pytorch(tmp_path / "x.png", tmp_path / "o.png", pr.PyTorchESRGANConfig(model_name="bad"))
    assert ok is False and "Invalid model" in err


class _FakeUpsampler:
    def __init__(self, exc=None):
        self.exc = exc
        self.calls = 0

    def enhance(self, img, outscale=None):
        self.calls += 1
        if self.exc:
            raise self.exc
        return np.zeros((img.shape[0] * 4, img.shape[1] * 4, 3), np.uint8), "RGB"

    def close(self):
        pass


def _write_png(path, h=8, w=8):
    import cv2

    cv2.imwrite(str(path), np.full((h, w, 3), 127, np.uint8))


def test_enhance_frame_success_writes_output(tmp_path, monkeypatch):
    up = _FakeUpsampler()
    monkeypatch.setattr(pr, "get_upsampler", lambda cfg: up)
    monkeypatch.setattr(pr, "_available_vram_mb", lambda gpu: 50000.0)
    _write_png(tmp_path / "in.png")
    ok, err = pr.enhance_frame_pytorch(tmp_path / "in.png", tmp_path / "out.png", pr.PyTorchESRGANConfig())
    assert (ok, err) == (True, None) and (tmp_path / "out.png").exists() and up.calls == 1


def test_enhance_frame_oom_message(tmp_path, monkeypatch):
    """OOM -> (False, 'GPU out of memory: ...Try: 1) Reduce tile_size...'): callers grep for 'memory' to shrink
    tiles (reference restorer.py:1746; intended behaviour of reference test :182-201)."""
    from framewright_b200.engine import EngineOutOfMemory

    monkeypatch.setattr(pr, "get_upsampler", lambda cfg: _FakeUpsampler(EngineOutOfMemory("GPU out of memory: 4 GiB")))
    monkeypatch.setattr(pr, "_available_vram_mb", lambda gpu: 50000.0)
    _write_png(tmp_path / "in.png")
    ok, err = pr.enhance_frame_pytorch(tmp_path / "in.png", tmp_path / "out.png", pr.PyTorchESRGANConfig())
    assert ok is False and err.startswith("GPU out of memory") and "Reduce tile_size" in err and "memory" in err


@pytest.mark.parametrize("avail,want", [(9000.0, 0), (5000.0, 512), (3000.0, 384), (1000.0, 256)])
def test_auto_tile_mutates_config(tmp_path, monkeypatch, avail, want):
    """Auto mode (tile_size == 0) picks the tile from free VRAM and mutates the config like the reference
    (:203-218; intended behaviour of reference tests :256-306: > 8000 MB -> 0, 2000-4000 MB -> 384)."""
    monkeypatch.setattr(pr, "get_upsampler", lambda cfg: _FakeUpsampler())
    monkeypatch.setattr(pr, "_available_vram_mb", lambda gpu: avail)
    _write_png(tmp_path / "in.png")
    cfg = pr.PyTorchESRGANConfig(tile_size=0)
    ok, _ = pr.enhance_frame_pytorch(tmp_path / "in.png", tmp_path / "out.png", cfg)
    assert ok and cfg.tile_size == want


def test_clear_cache_is_idempotent():
    pr.clear_upsampler_cache()
    pr.clear_upsampler_cache()
    assert pr._UPSAMPLER is None


# ---- SRBackend mirror ------------------------------------------------------------------------------
def test_sr_backend_names_scales_and_quirk():
    b = srm.B200RealESRGANBackend(srm.SRConfig(scale=4), None, "x4plus")
    assert b.name == "realesrgan_x4plus" and b.supported_scales == [4]
    b2 = srm.B200RealESRGANBackend(srm.SRConfig(scale=2), None, "x2plus")
    assert b2.supported_scales == [2] and b2._get_model_name() == "RealESRGAN_x2plus"
    # reference quirk: variant "x2" is not in the map -> silently RealESRGAN_x4plus
    assert srm.B200RealESRGANBackend(model_variant="x2")._get_model_name() == "RealESRGAN_x4plus"
    assert srm.B200RealESRGANBackend(model_variant="general")._get_model_name() == "realesr-general-x4v3"


def test_sr_backend_vram_estimate_formula():
    b = srm.B200RealESRGANBackend()
    assert b.estimate_vram_usage(1280, 720, 4) == 2000 + (1280 * 720 * 3 * 4 * 17) // (1024 * 1024)
    assert srm.B200RealESRGANBackend(model_variant="anime").estimate_vram_usage(64, 64, 4) == 1500


def test_sr_config_validation():
    with pytest.raises(ValueError):
        srm.SRConfig(scale=3)
    with pytest.raises(ValueError):
        srm.SRConfig(quality_preset="ultra")


class _BatchUpsampler:
    """Stand-in for the B200 upsampler: x4 nearest, counts calls, can fail on a marker pixel value."""
    scale = 4

    def __init__(self):
        self.batches, self.singles = [], 0

    def enhance_batch(self, frames, out=None):
        self.batches.append(len(frames))
        if (frames[:, 0, 0, 0] == 66).any():
            raise RuntimeError("boom")
        return np.repeat(np.repeat(frames, 4, axis=1), 4, axis=2)

    def enhance(self, img, outscale=None):
        self.singles += 1
        if img.ndim == 3 and img[0, 0, 0] == 66:
            raise RuntimeError("boom")
        out = np.repeat(np.repeat(img, 4, axis=0), 4, axis=1)
        return out, "RGB"

    def close(self):
        pass


def test_sr_backend_dir_api_batches_and_collects_failures(tmp_path, monkeypatch):
    """frames-dir in/out: sorted *.png, same names, consecutive same-size frames in one engine call, unreadable /
    failing frames -> warnings 'Frame <name>: <err>', progress per frame, never raises."""
    import cv2

    ind = tmp_path / "in"
    ind.mkdir()
    for i in range(9):
        cv2.imwrite(str(ind / f"frame_{i:08d}.png"), np.full((8, 8, 3), 10 * i, np.uint8))
    cv2.imwrite(str(ind / "frame_00000009.png"), np.full((6, 12, 3), 99, np.uint8))      # another size: own batch
    (ind / "frame_00000010.png").write_bytes(b"not a png")
    cv2.imwrite(str(ind / "frame_00000011.png"), np.full((6, 12, 3), 66, np.uint8))      # the engine fails on this one
    up = _BatchUpsampler()
    monkeypatch.setattr(srm, "get_upsampler", lambda cfg: up)
    prog = []
    res = srm.B200RealESRGANBackend().upscale_frames(ind, tmp_path / "out", 4, prog.append)
    assert (res.frames_processed, res.frames_failed) == (10, 2)
    assert sorted(res.warnings) == sorted([f"Frame frame_00000010.png: Failed to read image: {ind / 'frame_00000010.png'}",
                                           "Frame frame_00000011.png: boom"])
    assert up.batches == [4, 4, 2] and up.singles == 3      # frame 8 alone; the failed pair (9, 11) retried one by one
    assert len(prog) == 12 and prog == sorted(prog) and prog[-1] == 1.0
    assert res.output_dir == tmp_path / "out" and res.backend_used == "realesrgan_x4plus" and res.avg_fps > 0
    for i in range(9):
        got = cv2.imread(str(tmp_path / "out" / f"frame_{i:08d}.png"))
        assert got.shape == (32, 32, 3) and int(got[0, 0, 0]) == 10 * i
    assert not (tmp_path / "out" / "frame_00000011.png").exists()
    empty = srm.B200RealESRGANBackend().upscale_frames(tmp_path / "none", tmp_path / "out2", 4)
    assert empty.warnings == ["No frames found"]


def test_sr_backend_dir_api_recovers_after_out_of_memory(tmp_path, monkeypatch):
    """One frame runs the device out of memory: it fails with the reference's 'GPU out of memory ... Try: 1) ...' text,
    the upsampler cache is cleared (the engine is destroyed), and the FOLLOWING frames get a fresh upsampler instead of
    failing on the closed one."""
    import cv2

    from framewright_b200.engine import EngineOutOfMemory

    ind = tmp_path / "in"
    ind.mkdir()
    for i in range(6):
        cv2.imwrite(str(ind / f"frame_{i:08d}.png"), np.full((8, 8, 3), 66 if i == 1 else 10 * i, np.uint8))

    class _Up(_BatchUpsampler):
        closed = False

        def _check(self, frames):
            if self.closed:
                raise RuntimeError("upsampler is closed")
            if (np.asarray(frames)[..., 0, 0, 0] == 66).any():
                raise EngineOutOfMemory("GPU out of memory: workspace")

        def enhance_batch(self, frames, out=None):
            self._check(frames)
            return np.repeat(np.repeat(frames, 4, axis=1), 4, axis=2)

        def enhance(self, img, outscale=None):
            self._check(img[None])
            return np.repeat(np.repeat(img, 4, axis=0), 4, axis=1), "RGB"

    made = []

    def get(cfg):
        if not made or made[-1].closed:
            made.append(_Up())
        return made[-1]

    def clear():
        made[-1].closed = True

    monkeypatch.setattr(srm, "get_upsampler", get)
    monkeypatch.setattr(srm, "clear_upsampler_cache", clear)
    res = srm.B200RealESRGANBackend().upscale_frames(ind, tmp_path / "out", 4)
    assert (res.frames_processed, res.frames_failed) == (5, 1) and len(made) == 2
    assert len(res.warnings) == 1 and res.warnings[0].startswith("Frame frame_00000001.png: GPU out of memory")
    assert "Try: 1) Reduce tile_size" in res.warnings[0]
    assert sorted(p.name for p in (tmp_path / "out").glob("*.png")) == [f"frame_{i:08d}.png" for i in (0, 2, 3, 4, 5)]


def test_super_resolution_facade_selects_falls_back_and_processes_lists(monkeypatch):
    """`SuperResolution` (reference :1194-1527): auto selection, fallback from a backend that does not exist here,
    `process(frames: List)` batching runs of same-size frames, `upscale_frame`, factories."""
    up = _BatchUpsampler()
    monkeypatch.setattr(srm, "get_upsampler", lambda cfg: up)
    monkeypatch.setattr(srm, "is_pytorch_esrgan_available", lambda: True)
    sr = srm.SuperResolution(srm.SRConfig(scale=4, backend="auto"))
    assert sr.backend.name == "realesrgan_x4" and sr.get_backend_info()["supported_scales"] == [4]
    assert srm.SuperResolution(srm.SRConfig(scale=4, backend="hat_large", quality_preset="quality")).backend.name \
        == "realesrgan_x4"                                       # unavailable backend -> the preset's fallback chain
    assert srm.SuperResolution(srm.SRConfig(backend="realesrgan_anime")).backend._get_model_name() \
        == "RealESRGAN_x4plus_anime_6B"
    assert set(sr.get_available_backends()) == {"realesrgan_x2", "realesrgan_x4", "realesrgan_anime",
                                                "realesrgan_general", "realesrgan_animevideo"}
    frames = [np.full((4, 6, 3), i, np.uint8) for i in range(5)] + [np.full((3, 3, 3), 7, np.uint8)]
    outs = sr.process(frames, scale=4)
    assert [o.shape for o in outs] == [(16, 24, 3)] * 5 + [(12, 12, 3)] and [int(o[0, 0, 0]) for o in outs] == [0, 1, 2, 3, 4, 7]
    assert up.batches == [5, 1]
    assert sr.upscale_frame(frames[0]).shape == (16, 24, 3)
    assert sr.estimate_vram_usage(1280, 720) == 2000 + (1280 * 720 * 3 * 4 * 17) // (1024 * 1024)
    assert srm.create_super_resolution(scale=2).backend.supported_scales == [2]
    assert srm.get_recommended_backend() == "realesrgan_x4" and "realesrgan_x4" in srm.list_available_backends()
    monkeypatch.setattr(srm, "is_pytorch_esrgan_available", lambda: False)
    with pytest.raises(RuntimeError, match="No super-resolution backend available"):
        srm.SuperResolution(srm.SRConfig())


# ---- frame scheduler (reference tests/test_multi_gpu.py:510-554 restated) -------------------------------
def _gpus(n, free=None, util=None):
    return [mg.GPUInfo(i, f"GPU{i}", 180000, (free or [180000] * n)[i], (util or [0.0] * n)[i]) for i in range(n)]


def test_assign_round_robin_and_least_loaded():
    frames = [Path(f"f{i}.png") for i in range(10)]
    a = mg.assign_frames(frames, _gpus(2), mg.LoadBalanceStrategy.ROUND_ROBIN)
    assert len(a[0]) == 5 and len(a[1]) == 5 and a[0][0] == frames[0] and a[1][0] == frames[1]
    a = mg.assign_frames(frames, _gpus(2, util=[90.0, 10.0]), mg.LoadBalanceStrategy.LEAST_LOADED)
    assert a[1][0] == frames[0]


def test_assign_vram_aware_and_weighted_cover_all_frames():
    frames = [Path(f"f{i}.png") for i in range(11)]
    for strat in (mg.LoadBalanceStrategy.VRAM_AWARE, mg.LoadBalanceStrategy.WEIGHTED):
        a = mg.assign_frames(frames, _gpus(3, free=[90000, 45000, 45000]), strat)
        got = sorted(f for v in a.values() for f in v)
        assert got == sorted(frames)
        assert len(a[0]) >= len(a[1])


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 2000, 2001):
        for g in (1, 2, 4, 8):
            spans = [mg.shard_range(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert mg.shard_range(2000, 8, 3) == (750, 1000)


def test_gpu_health_and_result_summary():
    assert not mg.GPUInfo(0, "x", 1, 1, 0.0, temperature_c=95.0).is_healthy
    assert mg.GPUInfo(0, "x", 1, 1, 0.0, temperature_c=90.0).is_healthy
    r = mg.DistributionResult(frames_per_gpu={0: [Path("a")], 1: []}, errors={"b": "x"})
    assert r.total_frames == 1 and r.success_rate == 50.0 and "GPU0: 1" in r.summary()


def test_distributor_empty_and_no_gpus():
    d = mg.MultiGPUDistributor(gpus=[])
    assert d.distribute_frames([], None, Path(".")).total_frames == 0
    res = d.distribute_frames([Path("a.png")], lambda *a: (None, True, None), Path("."))
    assert res.errors == {"a.png": "No GPUs available"}


def test_distributor_process_fn_retry_on_other_gpu(tmp_path):
    frames = [tmp_path / f"f{i}.png" for i in range(6)]
    seen = []

    def fn(inp, outdir, gpu):
        seen.append((inp.name, gpu))
        if gpu == 0 and inp.name == "f2.png":
            return None, False, "gpu0 failed"
        return outdir / inp.name, True, None

    d = mg.MultiGPUDistributor(gpus=_gpus(2), strategy=mg.LoadBalanceStrategy.ROUND_ROBIN)
    res = d.distribute_frames(frames, fn, tmp_path / "out")
    assert res.total_frames == 6 and not res.errors
    assert ("f2.png", 0) in seen and ("f2.png", 1) in seen and frames[2] in res.retried_frames
    assert res.speedup_factor > 1.0


def test_dni_interpolates_state_dicts():
    """RealESRGANer.dni (upstream: blend of two checkpoints by dni_weight), on plain and file-style wrapped dicts."""
    import torch

    from framewright_b200.upsampler import RealESRGANer

    a = {"w": torch.ones(2, 3), "b": torch.zeros(3)}
    b = {"w": torch.full((2, 3), 3.0), "b": torch.ones(3)}
    out = RealESRGANer.dni(a, {"params": b}, [0.25, 0.75])
    assert torch.allclose(out["w"], torch.full((2, 3), 2.5)) and torch.allclose(out["b"], torch.full((3,), 0.75))
    import pytest
    from framewright_b200.engine import EngineError
    with pytest.raises(EngineError):
        RealESRGANer.dni(a, {"w": torch.ones(2, 3)}, [0.5, 0.5])


def test_tile_size_helpers_keep_the_reference_rules(monkeypatch):
    """The expectations of the reference's own tests for `calculate_optimal_tile_size` / `get_adaptive_tile_sequence`
    (`/root/reference/tests/test_utils_gpu.py:126-207`), restated against `tile_sizing` (whose numbers come from the
    engine's real footprint instead of the reference's cuDNN-era coefficients)."""
    from framewright_b200 import tile_sizing as ts

    monkeypatch.setattr(ts, "_free_vram_mb", lambda: 24000)
    assert ts.calculate_optimal_tile_size(frame_resolution=(1920, 1080), scale_factor=4) == 0      # test_no_tiling_needed
    monkeypatch.setattr(ts, "_free_vram_mb", lambda: 2000)
    t = ts.calculate_optimal_tile_size(frame_resolution=(3840, 2160), scale_factor=4)              # test_tiling_needed_low_vram
    assert t > 0 and t % 32 == 0
    assert ts.calculate_optimal_tile_size((3840, 2160), 4, available_vram_mb=4000) > 0             # test_with_explicit_vram
    assert ts.calculate_optimal_tile_size((7680, 4320), 4, available_vram_mb=1000) >= 128          # test_minimum_tile_size
    seq = ts.get_adaptive_tile_sequence(frame_resolution=(1920, 1080), scale_factor=4, starting_tile_size=512)
    assert all(a > b for a, b in zip(seq, seq[1:])) and all(x % 32 == 0 for x in seq)              # test_generate_sequence
    assert 128 in ts.get_adaptive_tile_sequence((1920, 1080), 4, min_tile_size=128)                # test_minimum_in_sequence
    # what is specific to this engine: the workspace a tile needs shrinks with the tile, and a B200 never tiles 720p
    assert ts.engine_workspace_mb("RealESRGAN_x4plus", 1280, 720, tile=256) < ts.engine_workspace_mb("RealESRGAN_x4plus", 1280, 720)
    assert ts.calculate_optimal_tile_size((1280, 720), 4, available_vram_mb=180000) == 0
    # no device and nothing given: the reference's conservative 2048 MB default
    monkeypatch.setattr(ts, "_free_vram_mb", lambda: None)
    assert ts.calculate_optimal_tile_size((1920, 1080), 4) >= 128
