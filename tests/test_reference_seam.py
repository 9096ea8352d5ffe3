"""The UNMODIFIED reference files drive the B200 host layer through the `realesrgan` / `basicsr` shims.

Runs only where /root/reference is mounted (this container).  The reference package cannot be imported as a
whole (its infrastructure/models/ directory is missing, SURVEY.md finding 4), so the hot-path modules are
loaded with stub parent packages.  There is no GPU here: the engine behind the shim is
replaced, for this test only, by the CPU oracle, which checks the seam (constructor arguments, `enhance`
signature and return type, file in / file out) and not the CUDA path.
"""
import importlib
import os
import sys
import types

from pathlib import Path

import numpy as np
import pytest

REF = "/root/reference/src/framewright"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


@pytest.fixture()
def ref_modules(monkeypatch):
    import framewright_b200  # noqa: F401
    from framewright_b200 import shims

    saved = {k: v for k, v in sys.modules.items() if k == "framewright" or k.startswith("framewright.")}
    for name, path in [("framewright", REF), ("framewright.processors", REF + "/processors"),
                       ("framewright.processors.enhancement", REF + "/processors/enhancement"),
                       ("framewright.utils", REF + "/utils")]:
        m = types.ModuleType(name)
        m.__path__ = [path]
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    shims.install()
    pr = importlib.import_module("framewright.processors.pytorch_realesrgan")
    yield pr
    for k in [k for k in sys.modules if k == "framewright" or k.startswith("framewright.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_reference_processor_calls_shim(ref_modules, tmp_path, monkeypatch):
    import cv2

    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    pr = ref_modules
    created = {}

    class OracleEngine:  # stands in for B200Engine (no GPU in this container)
        def __init__(self, arch, state_dict, gpu_id=0):
            created["arch"] = arch
            name = next(k for k, v in up_mod.MODEL_ARCHS.items() if v == arch)
            self.name, self.sd = name, state_dict

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            created["tile"] = (tile, tile_pad, pre_pad)
            return oracle.make_upsampler(self.name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(frames)[0]

        def close(self):
            pass

    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    # the reference hands RealESRGANer the release URL of the checkpoint; offline, the shim resolves it to the file of
    # that name in the weights directory (and raises if there is none -- never random weights)
    import torch

    wdir = tmp_path / "weights"
    wdir.mkdir()
    torch.save({"params_ema": make_synthetic_state_dict("RealESRGAN_x4plus_anime_6B", 0)},
               str(wdir / "RealESRGAN_x4plus_anime_6B.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    # the reference probes `import torch; from realesrgan import RealESRGANer; from basicsr... import RRDBNet`
    assert pr.is_pytorch_esrgan_available() is True
    img = oracle.synthetic_frame(24, 28, seed=2, kind="mixed")
    cv2.imwrite(str(tmp_path / "frame_00000001.png"), img)
    cfg = pr.PyTorchESRGANConfig(model_name="RealESRGAN_x4plus_anime_6B", scale_factor=4, tile_size=16, tile_pad=4)
    ok, err = pr.enhance_frame_pytorch(tmp_path / "frame_00000001.png", tmp_path / "out.png", cfg)
    assert (ok, err) == (True, None), err
    got = cv2.imread(str(tmp_path / "out.png"), cv2.IMREAD_UNCHANGED)
    assert created["arch"].num_block == 6 and created["tile"] == (16, 4, 0)
    sd = make_synthetic_state_dict("RealESRGAN_x4plus_anime_6B", 0)
    want = oracle.make_upsampler("RealESRGAN_x4plus_anime_6B", sd, tile=16, tile_pad=4, pre_pad=0).enhance(img)[0]
    assert got.shape == (96, 112, 3) and np.array_equal(got, want)
    pr.clear_upsampler_cache()


def test_reference_and_mirror_expose_same_surface(ref_modules):
    from framewright_b200 import pytorch_realesrgan as mine

    pr = ref_modules
    for name in ("PyTorchESRGANConfig", "is_pytorch_esrgan_available", "get_upsampler", "enhance_frame_pytorch",
                 "clear_upsampler_cache", "convert_ncnn_model_name", "NCNN_TO_PYTORCH_MODEL"):
        assert hasattr(pr, name) and hasattr(mine, name)
    assert pr.NCNN_TO_PYTORCH_MODEL == mine.NCNN_TO_PYTORCH_MODEL
    import dataclasses

    ref_fields = [(f.name, f.default) for f in dataclasses.fields(pr.PyTorchESRGANConfig)]
    my_fields = [(f.name, f.default) for f in dataclasses.fields(mine.PyTorchESRGANConfig)]
    assert ref_fields == my_fields


def _reference_method(path, cls, name):
    """Source of one method of the reference, verbatim, as a plain function (the module itself cannot be imported:
    `framewright.restorer` pulls in `framewright.infrastructure.models`, which does not exist in the reference tree)."""
    import ast
    import textwrap

    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    return textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
    raise AssertionError(f"{cls}.{name} not found in {path}")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("ref_file,ref_cls", [("restorer.py", "VideoRestorer"),
                                              ("mixins/frame_processing.py", "FrameProcessingMixin")])
def test_video_restorer_enhance_stage_runs_verbatim_against_the_mirror(tmp_path, monkeypatch, ref_file, ref_cls):
    """(Both copies of the caller: `VideoRestorer`'s own method and the `FrameProcessingMixin` one, which imports the
    processor module and the validator relatively inside the method -- resolved here to the mirrors.)
    SURVEY 8 a10: `VideoRestorer._enhance_single_frame_pytorch` (restorer.py:1420-1460) -- the reference's caller,
    its source text unmodified -- executed against this repo's module functions: ncnn model name -> config ->
    `enhance_frame_pytorch` -> output validation, returning `(output_path, ok, err)`; and the "memory" error string
    its retry ladder greps for (:1746)."""
    import types
    from typing import Optional, Tuple

    import cv2
    import torch

    from framewright_b200 import pytorch_realesrgan as mine
    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import EngineOutOfMemory
    from framewright_b200.restorer_adapter import validate_frame_integrity
    from oracle import oracle

    fail = {"oom": False}

    class OracleEngine:
        def __init__(self, arch, state_dict, gpu_id=0):
            self.name = next(k for k, v in up_mod.MODEL_ARCHS.items() if v == arch)
            self.sd = state_dict

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            if fail["oom"]:
                raise EngineOutOfMemory("GPU out of memory: out of memory allocating 6300 MiB workspace")
            return oracle.make_upsampler(self.name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(frames)[0]

        def close(self):
            pass

    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    monkeypatch.setattr(mine, "is_pytorch_esrgan_available", lambda: True)
    monkeypatch.setattr(mine, "_auto_tile", lambda gpu: 0)
    monkeypatch.setattr(mine, "_available_vram_mb", lambda gpu: 50000.0)
    wdir = tmp_path / "weights"
    wdir.mkdir()
    torch.save({"params_ema": make_synthetic_state_dict("RealESRGAN_x4plus_anime_6B", 0)},
               str(wdir / "RealESRGAN_x4plus_anime_6B.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    mine.clear_upsampler_cache()

    ns = {"Path": Path, "Tuple": Tuple, "Optional": Optional,
          "is_pytorch_esrgan_available": mine.is_pytorch_esrgan_available,
          "convert_ncnn_model_name": mine.convert_ncnn_model_name, "PyTorchESRGANConfig": mine.PyTorchESRGANConfig,
          "enhance_frame_pytorch": mine.enhance_frame_pytorch, "validate_frame_integrity": validate_frame_integrity}
    if ref_cls == "FrameProcessingMixin":      # its `from ..processors.pytorch_realesrgan import ...` / `from ..validators import ...`
        ns.update(__package__="framewright.mixins", __name__="framewright.mixins.frame_processing")
        for name in ("framewright", "framewright.mixins", "framewright.processors", "framewright.validators"):
            m = types.ModuleType(name)
            m.__path__ = []
            monkeypatch.setitem(sys.modules, name, m)
        monkeypatch.setitem(sys.modules, "framewright.processors.pytorch_realesrgan", mine)
        sys.modules["framewright.validators"].validate_frame_integrity = validate_frame_integrity
    exec(_reference_method(os.path.join(REF, ref_file), ref_cls, "_enhance_single_frame_pytorch"), ns)
    restorer = types.SimpleNamespace(config=types.SimpleNamespace(model_name="realesrgan-x4plus-anime", scale_factor=4,
                                                                  gpu_id=None))
    img = oracle.synthetic_frame(20, 24, seed=5, kind="mixed")
    src, dst = tmp_path / "frame_00000001.png", tmp_path / "enhanced_frame_00000001.png"
    cv2.imwrite(str(src), img)
    out_path, ok, err = ns["_enhance_single_frame_pytorch"](restorer, src, dst, 0)
    assert (out_path, ok, err) == (dst, True, None)
    got = cv2.imread(str(dst), cv2.IMREAD_UNCHANGED)
    want = oracle.make_upsampler("RealESRGAN_x4plus_anime_6B", make_synthetic_state_dict("RealESRGAN_x4plus_anime_6B", 0),
                                 tile=0, pre_pad=0).enhance(img)[0]
    assert np.array_equal(got, want)
    v = validate_frame_integrity(dst)
    assert v.is_valid and (v.width, v.height) == (96, 80)
    fail["oom"] = True
    _, ok, err = ns["_enhance_single_frame_pytorch"](restorer, src, tmp_path / "o2.png", 0)
    assert ok is False and ("vram" in err.lower() or "memory" in err.lower())         # restorer.py:1746's test
    assert not validate_frame_integrity(tmp_path / "o2.png").is_valid
    mine.clear_upsampler_cache()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_ensemble_member_runs_through_the_mirror(tmp_path, monkeypatch):
    """`EnsembleSR`'s Real-ESRGAN member (`processors/ensemble_sr.py:132-223`, the reference file loaded
    unmodified) -- `ModelProcessor("realesrgan").process_frame(frame)` builds a `PyTorchESRGANConfig`, calls
    `get_upsampler` and then `enhance_frame_pytorch(frame, upsampler, config)` with an ndARRAY: through the mirror that
    returns the upscaled frame (checked against the oracle)."""
    import torch

    from framewright_b200 import pytorch_realesrgan as mine
    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    class OracleEngine:
        def __init__(self, arch, state_dict, gpu_id=0):
            self.name = next(k for k, v in up_mod.MODEL_ARCHS.items() if v == arch)
            self.sd = state_dict

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            return oracle.make_upsampler(self.name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(frames)[0]

        def close(self):
            pass

    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    monkeypatch.setattr(mine, "is_pytorch_esrgan_available", lambda: True)
    monkeypatch.setattr(mine, "_auto_tile", lambda gpu: 0)
    monkeypatch.setattr(mine, "MODEL_SCALES", dict(mine.MODEL_SCALES))
    wdir = tmp_path / "weights"
    wdir.mkdir()
    # (the member asks for RealESRGAN_x4plus: a 1-block file under that name would not load strictly, so the 23-block
    # synthetic network it is -- on a 12 x 16 frame the oracle takes a few seconds)
    sd = make_synthetic_state_dict("RealESRGAN_x4plus", 0)
    torch.save({"params_ema": sd}, str(wdir / "RealESRGAN_x4plus.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    mine.clear_upsampler_cache()

    saved = {k: v for k, v in sys.modules.items() if k == "framewright" or k.startswith("framewright.")}
    try:
        for name, path in [("framewright", REF), ("framewright.processors", REF + "/processors")]:
            m = types.ModuleType(name)
            m.__path__ = [path]
            monkeypatch.setitem(sys.modules, name, m)
        monkeypatch.setitem(sys.modules, "framewright.processors.pytorch_realesrgan", mine)
        monkeypatch.setattr(sys, "dont_write_bytecode", True)
        ens = importlib.import_module("framewright.processors.ensemble_sr")
        member = ens.ModelProcessor("realesrgan", gpu_id=0)
        assert member.load()
        frame = oracle.synthetic_frame(12, 16, seed=9, kind="mixed")
        out = member.process_frame(frame)
        assert isinstance(out, np.ndarray) and out.shape == (48, 64, 3) and out.dtype == np.uint8
        ref, _ = oracle.make_upsampler("RealESRGAN_x4plus", sd, tile=0, pre_pad=0).enhance(frame)
        assert np.array_equal(out, ref)
        assert member.process_frame(np.zeros((4, 4, 5), np.uint8)) is None      # engine error -> None, as the caller expects
    finally:
        mine.clear_upsampler_cache()
        for k in [k for k in sys.modules if k == "framewright" or k.startswith("framewright.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_sr_facade_with_the_b200_operator_registered(tmp_path, monkeypatch):
    """SURVEY 8 a9 / INTEGRATION.md layer 2, on the UNMODIFIED reference `processors/enhancement/super_resolution.py`:
    (1) its own `RealESRGANBackend` runs over the module mirror (`..pytorch_realesrgan` -> this repo's functions);
    (2) with `B200RealESRGANBackend` registered the way INTEGRATION.md shows (module global + `BACKENDS` entries) the
    reference's `SuperResolution` facade selects it, and `upscale_frame` / `process` / `upscale` (frames dir) give the
    oracle's pixels."""
    import cv2
    import torch

    from framewright_b200 import pytorch_realesrgan as mine
    from framewright_b200 import super_resolution as my_sr
    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    class OracleEngine:
        def __init__(self, arch, state_dict, gpu_id=0):
            self.name = next(k for k, v in up_mod.MODEL_ARCHS.items() if v == arch)
            self.sd = state_dict

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            up = oracle.make_upsampler(self.name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad)
            if frames.ndim == 4:      # the frames-dir path batches same-size frames
                return np.stack([up.enhance(f)[0] for f in frames])
            return up.enhance(frames)[0]

        def close(self):
            pass

    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    for mod in (mine, my_sr):
        monkeypatch.setattr(mod, "is_pytorch_esrgan_available", lambda: True)
    monkeypatch.setattr(mine, "_auto_tile", lambda gpu: 0)
    monkeypatch.setattr(mine, "_available_vram_mb", lambda gpu: 50000.0)
    wdir = tmp_path / "weights"
    wdir.mkdir()
    name = "RealESRGAN_x4plus_anime_6B"
    sd = make_synthetic_state_dict(name, 0)
    torch.save({"params_ema": sd}, str(wdir / (name + ".pth")))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    mine.clear_upsampler_cache()
    frame = oracle.synthetic_frame(16, 20, seed=11, kind="mixed")
    want = oracle.make_upsampler(name, sd, tile=0, pre_pad=0).enhance(frame)[0]

    saved = {k: v for k, v in sys.modules.items() if k == "framewright" or k.startswith("framewright.")}
    try:
        for mname, path in [("framewright", REF), ("framewright.processors", REF + "/processors"),
                            ("framewright.processors.enhancement", REF + "/processors/enhancement")]:
            m = types.ModuleType(mname)
            m.__path__ = [path]
            monkeypatch.setitem(sys.modules, mname, m)
        monkeypatch.setitem(sys.modules, "framewright.processors.pytorch_realesrgan", mine)
        monkeypatch.setattr(sys, "dont_write_bytecode", True)
        sr = importlib.import_module("framewright.processors.enhancement.super_resolution")
        hw = types.SimpleNamespace(tier=None, vram_free_mb=180000, gpu_vendor=None)

        # (1) the reference's own operator over the module mirror
        ref_backend = sr.RealESRGANBackend(sr.SRConfig(scale=4), hw, "anime")
        assert ref_backend.is_available()
        assert np.array_equal(ref_backend.upscale_frame(frame, 4), want)
        ind = tmp_path / "in"
        ind.mkdir()
        for i in range(3):
            cv2.imwrite(str(ind / f"frame_{i + 1:08d}.png"), frame)
        res = ref_backend.upscale_frames(ind, tmp_path / "out_ref", 4)
        assert (res.frames_processed, res.frames_failed) == (3, 0)
        assert np.array_equal(cv2.imread(str(tmp_path / "out_ref" / "frame_00000002.png")), want)

        # (2) INTEGRATION.md layer 2: the B200 operator registered in the reference facade
        monkeypatch.setattr(sr, "RealESRGANBackend", my_sr.B200RealESRGANBackend)
        for key in ("realesrgan_x2", "realesrgan_x4", "realesrgan_anime"):
            monkeypatch.setitem(sr.SuperResolution.BACKENDS, key, my_sr.B200RealESRGANBackend)
        facade = sr.SuperResolution(sr.SRConfig(scale=4, backend="realesrgan_anime"), hw)
        assert isinstance(facade.backend, my_sr.B200RealESRGANBackend) and facade.backend.name == "realesrgan_anime"
        assert np.array_equal(facade.upscale_frame(frame), want)
        outs = facade.process([frame, frame])
        assert len(outs) == 2 and all(np.array_equal(o, want) for o in outs)
        res = facade.upscale(ind, tmp_path / "out_b200")
        assert (res.frames_processed, res.frames_failed, res.backend_used) == (3, 0, "realesrgan_anime")
        assert np.array_equal(cv2.imread(str(tmp_path / "out_b200" / "frame_00000003.png")), want)
    finally:
        mine.clear_upsampler_cache()
        for k in [k for k in sys.modules if k == "framewright" or k.startswith("framewright.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_streaming_pipeline_with_the_b200_enhancer(tmp_path, monkeypatch):
    """the UNMODIFIED `StreamingPipeline` (`processors/streaming.py:815-1175`: extract / enhance / write
    threads, bounded buffers, batches of `batch_size` `PipelineFrame`s) with `make_streaming_enhancer(...)` as its
    `set_enhancer` function: every frame comes out processed, in the files the write stage names, with the stand-in
    engine's pixels; an unreadable frame is skipped by the reference's write stage, not fatal."""
    import cv2

    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import make_streaming_enhancer
    from sched_helpers import fake_engine

    saved = {k: v for k, v in sys.modules.items() if k == "framewright" or k.startswith("framewright.")}
    try:
        for mname, path in [("framewright", REF), ("framewright.processors", REF + "/processors"),
                            ("framewright.utils", REF + "/utils")]:
            m = types.ModuleType(mname)
            m.__path__ = [path]
            monkeypatch.setitem(sys.modules, mname, m)
        monkeypatch.setattr(sys, "dont_write_bytecode", True)
        st = importlib.import_module("framewright.processors.streaming")
        ind, outd = tmp_path / "in", tmp_path / "out"
        ind.mkdir()
        n = 11
        for i in range(n):
            cv2.imwrite(str(ind / f"frame_{i:08d}.png"), np.full((6, 8, 3), 10 * i, np.uint8))
        (ind / f"frame_{n:08d}.png").write_bytes(b"not a png")
        batches, prog = [], []
        pipe = st.StreamingPipeline(st.StreamingConfig(batch_size=4, max_buffer_size=6, cleanup_between_chunks=False))
        pipe.set_enhancer(make_streaming_enhancer(PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2),
                                                  upsampler=fake_engine({"gpu_id": 1}), output_dir=outd))
        pipe.set_batch_callback(lambda r: batches.append(list(r.frame_indices)))
        paths = pipe.process(ind, outd, progress_callback=lambda f, m: prog.append(f))
        assert paths == [outd / f"frame_{i:08d}.png" for i in range(n)]            # the unreadable frame is skipped
        assert batches == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]] and prog[-1] == 1.0
        for i in range(n):
            got = cv2.imread(str(paths[i]))
            assert got.shape == (12, 16, 3) and int(got[0, 0, 0]) == 10 * i
        assert not (outd / f"frame_{n:08d}.png").exists()
    finally:
        for k in [k for k in sys.modules if k == "framewright" or k.startswith("framewright.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_multi_gpu_processor_calls_the_b200_process_func(monkeypatch):
    """`MultiGPUProcessor.process_frames` + `_process_single_frame`
    (`infrastructure/gpu/distributor.py:687-786`, source text unmodified; the module itself needs a GPU backend to
    initialise) with `make_process_func(config)` as the `process_func(frame, device_id)` operator: per-GPU executor
    threads, results sorted by frame index, an engine error becomes `ProcessingResult(success=False)`."""
    import logging
    import threading
    import time
    from concurrent.futures import ThreadPoolExecutor, as_completed
    from typing import Callable, List, Optional

    from framewright_b200 import pytorch_realesrgan as mine
    from framewright_b200.restorer_adapter import ProcessingResult, make_process_func
    from sched_helpers import FakeUpsampler

    made = {}

    class Up(FakeUpsampler):
        def enhance(self, img, outscale=None):
            if img.ndim not in (2, 3):
                raise ValueError("img must be an HxW or HxWxC ndarray")      # what RealESRGANer.enhance raises
            return super().enhance(img, outscale)

    def fake_get_upsampler(cfg):
        return made.setdefault(cfg.gpu_id, Up({"gpu_id": cfg.gpu_id}))

    monkeypatch.setattr(mine, "get_upsampler", fake_get_upsampler)
    ns = {"np": np, "time": time, "logger": logging.getLogger("ref"), "ProcessingResult": ProcessingResult,
          "as_completed": as_completed, "List": List, "Optional": Optional, "Callable": Callable}
    path = os.path.join(REF, "infrastructure/gpu/distributor.py")
    exec(_reference_method(path, "MultiGPUProcessor", "process_frames"), ns)
    exec(_reference_method(path, "MultiGPUProcessor", "_process_single_frame"), ns)
    n = 9
    plan = types.SimpleNamespace(frame_assignments={i: i % 2 for i in range(n)})
    stats = types.SimpleNamespace(errors=0, times=[], update_timing=lambda t: None)
    self = types.SimpleNamespace(
        _initialized=True, _executors={0: ThreadPoolExecutor(1), 1: ThreadPoolExecutor(1)},
        distributor=types.SimpleNamespace(get_optimal_distribution=lambda k: plan, _stats={0: stats, 1: stats}))
    self._process_single_frame = lambda *a: ns["_process_single_frame"](self, *a)
    frames = [np.full((4, 6, 3), i, np.uint8) for i in range(n - 1)] + [np.zeros((4, 6, 5, 1, 1), np.uint8)]
    seen = []
    fn = make_process_func(mine.PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2))
    results = ns["process_frames"](self, frames, fn, callback=lambda r: seen.append(r.frame_index))
    for ex in self._executors.values():
        ex.shutdown()
    assert [r.frame_index for r in results] == list(range(n)) and sorted(seen) == list(range(n))
    assert [r.device_id for r in results] == [i % 2 for i in range(n)] and sorted(made) == [0, 1]
    assert all(r.success and r.output.shape == (8, 12, 3) and int(r.output[0, 0, 0]) == r.frame_index for r in results[:-1])
    assert results[-1].success is False and results[-1].error and stats.errors == 1


# ---- the remaining call sites of the upsampler duck type, their source text unmodified ------
def _reference_function(path, name):
    """Source of one module-level function of a reference file, verbatim."""
    import ast

    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
    raise AssertionError(f"{name} not found in {path}")


@pytest.fixture()
def shimmed_oracle_engine(monkeypatch, tmp_path):
    """`from realesrgan import RealESRGANer` / `from basicsr.archs.rrdbnet_arch import RRDBNet` resolve to the shims;
    the engine behind them is the CPU oracle (no GPU here); synthetic checkpoints under the names upstream ships."""
    import torch

    import framewright_b200  # noqa: F401
    from framewright_b200 import shims
    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    calls = []

    class OracleEngine:
        def __init__(self, arch, state_dict, gpu_id=0):
            self.name = next(k for k, v in up_mod.MODEL_ARCHS.items() if v == arch)
            self.sd = state_dict

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            calls.append((self.name, tile, tile_pad, pre_pad))
            return oracle.make_upsampler(self.name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(frames)[0]

        def close(self):
            pass

    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("realesrgan", "basicsr")}
    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    shims.install()
    models = tmp_path / "home" / ".framewright" / "models"
    models.mkdir(parents=True)
    for name in ("RealESRGAN_x4plus_anime_6B", "realesr-animevideov3"):
        torch.save({"params_ema": make_synthetic_state_dict(name, 0)}, str(models / f"{name}.pth"))
    # (a 6-block stand-in under the x4plus FILE name keeps the CPU oracle fast: the architecture comes from the
    #  RRDBNet(...) object the caller passes, which is exactly what is being tested)
    yield calls, models
    for k in [k for k in sys.modules if k.split(".")[0] in ("realesrgan", "basicsr")]:
        del sys.modules[k]
    sys.modules.update(saved)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_cli_enhance_frames_runs_verbatim_over_the_shims(shimmed_oracle_engine, tmp_path, monkeypatch, capsys):
    """`framewright enhance-frames` (`cli.py:699-775`, `_enhance_with_realesrgan`, source unmodified): RRDBNet(...) object +
    `RealESRGANer(scale, model_path, model, tile=512, tile_pad=10, pre_pad=10, half=True)` + `enhance(img, outscale)` per
    frame file.  The 'anime' branch names realesr-animevideov3.pth and passes a 6-block RRDBNet: the checkpoint's name
    selects the SRVGG network upstream ships under it (SURVEY finding 3)."""
    import types

    import cv2
    import tqdm

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    calls, models = shimmed_oracle_engine
    monkeypatch.setenv("HOME", str(tmp_path / "home"))
    ns = {"Path": Path, "sys": sys, "tqdm": tqdm.tqdm,
          "print_colored": lambda msg, color=None: print(msg),
          "Colors": types.SimpleNamespace(OKCYAN="", WARNING="", FAIL="", OKGREEN="", OKBLUE="")}
    exec(_reference_function(os.path.join(REF, "cli.py"), "_enhance_with_realesrgan"), ns)
    ind, outd = tmp_path / "frames", tmp_path / "enhanced"
    ind.mkdir()
    outd.mkdir()
    frames = []
    for i in range(3):
        img = oracle.synthetic_frame(18, 22, seed=80 + i, kind="mixed")
        cv2.imwrite(str(ind / f"frame_{i + 1:08d}.png"), img)
        frames.append(img)
    (ind / "frame_00000004.png").write_bytes(b"not a png")                     # unreadable: counted as failed, no raise
    files = sorted(ind.glob("*.png"))
    ns["_enhance_with_realesrgan"](types.SimpleNamespace(model="realesr-animevideov3-anime"), ind, outd, 4, files)
    assert calls and set(calls) == {("realesr-animevideov3", 512, 10, 10)}
    want_up = oracle.make_upsampler("realesr-animevideov3", make_synthetic_state_dict("realesr-animevideov3", 0),
                                    tile=512, tile_pad=10, pre_pad=10)
    for i, img in enumerate(frames):
        got = cv2.imread(str(outd / f"frame_{i + 1:08d}.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(got, want_up.enhance(img)[0])
    out = capsys.readouterr().out
    assert "Enhanced 3/4 frames" in out and "1 frames failed" in out
    # no checkpoint of the requested model and no fallback file: the reference prints and exits 1 -- never random weights
    for f in models.glob("*.pth"):
        f.unlink()
    with pytest.raises(SystemExit):
        ns["_enhance_with_realesrgan"](types.SimpleNamespace(model="realesrgan-x4plus"), ind, outd, 4, files)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_gfpgan_background_upsampler_runs_verbatim_over_the_shims(shimmed_oracle_engine, tmp_path, monkeypatch):
    """`FaceRestorer._get_bg_upsampler` (`processors/face_restore.py:379-401`, source unmodified): RRDBNet(23 blocks) object
    + `RealESRGANer(scale=4, model_path='RealESRGAN_x4plus.pth', model=model, tile=400, tile_pad=10, pre_pad=0, half=True)`
    -- a BARE file name, resolved in the weights directory."""
    import shutil
    import types

    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    calls, models = shimmed_oracle_engine
    ns = {}
    exec(_reference_method(os.path.join(REF, "processors/face_restore.py"), "FaceRestorer", "_get_bg_upsampler"), ns)
    me = types.SimpleNamespace(bg_upsampler="realesrgan")
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(tmp_path / "empty"))
    assert ns["_get_bg_upsampler"](me) is None               # no such checkpoint: the reference swallows it -> no upsampler
    assert ns["_get_bg_upsampler"](types.SimpleNamespace(bg_upsampler="none")) is None
    import torch

    wdir = tmp_path / "weights"
    wdir.mkdir()
    torch.save({"params": make_synthetic_state_dict("RealESRGAN_x4plus", 0)}, str(wdir / "RealESRGAN_x4plus.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    up = ns["_get_bg_upsampler"](me)
    assert up is not None and (up.scale, up.tile_size, up.tile_pad, up.pre_pad) == (4, 400, 10, 0)
    img = oracle.synthetic_frame(12, 14, seed=9, kind="mixed")
    out, mode = up.enhance(img, outscale=2)                  # GFPGAN asks for its own upscale factor: resized result
    assert mode == "RGB" and out.shape == (24, 28, 3) and calls[-1] == ("RealESRGAN_x4plus", 400, 10, 0)
    up.close()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_dependency_check_finds_the_shims(shimmed_oracle_engine, monkeypatch):
    """`utils/dependencies.py::check_realesrgan` (:263-345, the module loaded unmodified): with no ncnn binary on the
    machine it verifies `import torch; from realesrgan import RealESRGANer; from basicsr.archs.rrdbnet_arch import
    RRDBNet` and reads `realesrgan.__version__` -- which is what lets `VideoRestorer.__init__` pass
    `_verify_dependencies` (restorer.py:393-406)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_ref_dependencies", os.path.join(REF, "utils", "dependencies.py"))
    dep = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "_ref_dependencies", dep)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    spec.loader.exec_module(dep)
    monkeypatch.setattr(dep.shutil, "which", lambda cmd: None)
    monkeypatch.setattr(dep.Path, "home", classmethod(lambda cls: Path("/nonexistent-home")))
    info = dep.check_realesrgan()
    assert info.installed is True and info.additional_info["backend"] == "pytorch"
    assert info.version.endswith("+b200sr") and info.path == "realesrgan (Python package)"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_video_restorer_thread_parallel_caller_runs_verbatim_against_the_mirror(tmp_path, monkeypatch):
    """SURVEY 8 a10 / (b) threading: `VideoRestorer._enhance_frames_parallel` (restorer.py:1823-1973), `_enhance_single_frame`
    (:1386-1418) and `_enhance_single_frame_pytorch` (:1420-1460) -- source text unmodified -- drive the mirror's
    `enhance_frame_pytorch` from `parallel_frames` = 4 threads: every frame written and checkpointed, progress to 1.0;
    an out-of-memory answer walks the caller's tile ladder ("memory" in the message, :1746 / :1878) and the frames
    still complete on the smaller tile."""
    import shutil
    import threading
    import time
    import types
    from concurrent.futures import ThreadPoolExecutor, as_completed
    from typing import List, Optional, Tuple

    import cv2
    import torch

    from framewright_b200 import pytorch_realesrgan as mine
    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import EngineOutOfMemory
    from framewright_b200.restorer_adapter import EnhancementError, validate_frame_integrity
    from oracle import oracle

    name = "RealESRGAN_x4plus_anime_6B"
    state = {"oom_above_tile": None, "threads": set(), "tiles": []}

    class OracleEngine:
        def __init__(self, arch, state_dict, gpu_id=0):
            self.sd = state_dict

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            state["threads"].add(threading.get_ident())
            lim = state["oom_above_tile"]
            if lim is not None and (tile == 0 or tile > lim):
                raise EngineOutOfMemory("GPU out of memory: out of memory allocating workspace")
            state["tiles"].append((tile, hash(np.asarray(frames).tobytes())))
            return oracle.make_upsampler(name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad).enhance(frames)[0]

        def close(self):
            pass

    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    monkeypatch.setattr(mine, "is_pytorch_esrgan_available", lambda: True)
    monkeypatch.setattr(mine, "_auto_tile", lambda gpu: 0)
    monkeypatch.setattr(mine, "_available_vram_mb", lambda gpu: 50000.0)
    wdir = tmp_path / "weights"
    wdir.mkdir()
    torch.save({"params_ema": make_synthetic_state_dict(name, 0)}, str(wdir / f"{name}.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    mine.clear_upsampler_cache()

    class Report:
        def __init__(self):
            self.ok, self.errors = 0, []

        def add_success(self):
            self.ok += 1

        def add_error(self, frame, exc):
            self.errors.append((frame, str(exc)))

    logger = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, error=lambda *a, **k: None,
                                   debug=lambda *a, **k: None)
    ns = {"Path": Path, "Tuple": Tuple, "Optional": Optional, "List": List, "ErrorReport": Report, "time": time,
          "shutil": shutil, "logger": logger, "ThreadPoolExecutor": ThreadPoolExecutor, "as_completed": as_completed,
          "EnhancementError": EnhancementError, "is_pytorch_esrgan_available": mine.is_pytorch_esrgan_available,
          "convert_ncnn_model_name": mine.convert_ncnn_model_name, "PyTorchESRGANConfig": mine.PyTorchESRGANConfig,
          "enhance_frame_pytorch": mine.enhance_frame_pytorch, "validate_frame_integrity": validate_frame_integrity}
    for meth in ("_enhance_frames_parallel", "_enhance_frames_sequential", "_enhance_single_frame",
                 "_enhance_single_frame_pytorch"):
        exec(_reference_method(os.path.join(REF, "restorer.py"), "VideoRestorer", meth), ns)

    ind, outd = tmp_path / "frames", tmp_path / "enhanced"
    ind.mkdir()
    outd.mkdir()
    imgs, frames = [], []
    for i in range(8):
        img = oracle.synthetic_frame(40, 44, seed=120 + i, kind="mixed")
        p = ind / f"frame_{i + 1:08d}.png"
        cv2.imwrite(str(p), img)
        imgs.append(img)
        frames.append(p)
    checkpointed, progress = [], []
    me = types.SimpleNamespace(
        config=types.SimpleNamespace(parallel_frames=4, enhanced_dir=outd, max_retries=2, retry_delay=0.0,
                                     continue_on_error=False, model_name="realesrgan-x4plus-anime", scale_factor=4,
                                     gpu_id=None),
        checkpoint_manager=types.SimpleNamespace(update_frame=lambda **k: checkpointed.append(k["frame_number"])),
        _vram_monitor=None, _reset_stage_timing=lambda stage: None, _record_frame_time=lambda t: None,
        _check_disk_space=lambda: None, _get_enhancement_backend=lambda: "pytorch",
        _update_progress=lambda **k: progress.append(k["progress"]))
    for meth in ("_enhance_single_frame", "_enhance_single_frame_pytorch"):
        setattr(me, meth, types.MethodType(ns[meth], me))

    rep = Report()
    n = ns["_enhance_frames_parallel"](me, frames, 0, [32, 24, 16], rep)
    assert n == 8 and rep.ok == 8 and not rep.errors and sorted(checkpointed) == list(range(1, 9))
    assert progress[-1] == 1.0 and len(state["threads"]) > 1                     # really called from several threads
    want_up = oracle.make_upsampler(name, make_synthetic_state_dict(name, 0), tile=0, pre_pad=0)
    for i, img in enumerate(imgs):
        assert np.array_equal(cv2.imread(str(outd / frames[i].name), cv2.IMREAD_UNCHANGED), want_up.enhance(img)[0])

    # the sequential caller (:1707-1821, parallel_frames = 1), the same three frames; then with a device on which only
    # tiles <= 24 fit: its ladder (one step per failure, no race) ends on 24 and the failed frame is retried there
    shutil.rmtree(outd)
    outd.mkdir()
    checkpointed.clear()
    rep = Report()
    assert ns["_enhance_frames_sequential"](me, frames[:3], 0, [32, 24, 16], rep) == 3 and rep.ok == 3
    assert sorted(checkpointed) == [1, 2, 3]
    assert np.array_equal(cv2.imread(str(outd / frames[1].name), cv2.IMREAD_UNCHANGED), want_up.enhance(imgs[1])[0])
    state.update(oom_above_tile=24, tiles=[])
    mine.clear_upsampler_cache()
    rep = Report()
    assert ns["_enhance_frames_sequential"](me, frames[:3], 48, [48, 32, 24, 16], rep) == 3 and rep.ok == 3
    assert {t for t, _ in state["tiles"]} == {24}
    state.update(oom_above_tile=None, tiles=[])

    # the device "fills up": nothing above a 24-pixel tile fits any more.  The caller starts from its configured tile
    # and every failing thread steps the shared tile size down once (the reference's own logic, racy by design: the
    # ladder may overshoot), so the frames end on SOME tile <= 24 -- each output must be the tiled result for the tile
    # it was computed with.
    shutil.rmtree(outd)
    outd.mkdir()
    state.update(oom_above_tile=24, tiles=[])
    me.config.max_retries = 6       # (one thread alone may need five steps to reach 24; at most 8 failures in all < 9 rungs)
    mine.clear_upsampler_cache()
    rep = Report()
    n = ns["_enhance_frames_parallel"](me, frames, 48, [40, 36, 32, 28, 24, 20, 16, 12, 8], rep)
    assert n == 8 and rep.ok == 8 and not rep.errors
    used = {h: t for t, h in state["tiles"]}
    assert len(used) == 8 and all(0 < t <= 24 for t in used.values())
    sd = make_synthetic_state_dict(name, 0)
    for i, img in enumerate(imgs):
        t = used[hash(img.tobytes())]
        want = oracle.make_upsampler(name, sd, tile=t, tile_pad=10, pre_pad=0).enhance(img)[0]
        assert np.array_equal(cv2.imread(str(outd / frames[i].name), cv2.IMREAD_UNCHANGED), want), (i, t)
    mine.clear_upsampler_cache()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_the_documented_enhance_frames_patch_applies_and_runs(tmp_path, monkeypatch):
    """SURVEY 8 f2, the binding a maintainer would add: the `+` lines of INTEGRATION.md's `enhance_frames` diff are taken
    from that file, inserted into the reference's `VideoRestorer.enhance_frames` source (restorer.py:1604-1705) at the
    documented place, and the patched method runs -- frames directory in, `enhanced_dir` out through the sharded
    scheduler (two worker processes, stand-in engine), the reference's own `CheckpointManager` fed per frame,
    `_update_progress` per frame, resume skipping what the checkpoint holds."""
    import functools
    import types

    import cv2

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sched_helpers import fake_engine
    from test_adapters import _checkpoint_module

    import framewright_b200  # noqa: F401
    from framewright_b200 import multi_gpu as mg
    from framewright_b200 import pytorch_realesrgan as mine
    from framewright_b200 import restorer_adapter as ra
    from framewright_b200.tile_sizing import get_adaptive_tile_sequence

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = doc[doc.index("--- a/src/framewright/restorer.py   (enhance_frames"):]
    block = block[:block.index("```")]
    plus = [ln[3:] for ln in block.splitlines() if ln.startswith("  +")]          # "  +<code>" -> "<code>"
    assert len(plus) >= 10 and plus[0].strip().startswith("if self._get_enhancement_backend()")
    src = _reference_method(os.path.join(REF, "restorer.py"), "VideoRestorer", "enhance_frames").splitlines()
    at = next(i for i, ln in enumerate(src) if 'logger.info(f"Enhancing {total_frames} frames using' in ln)
    indent_ref = len(src[at]) - len(src[at].lstrip())
    indent_doc = len(plus[0]) - len(plus[0].lstrip())
    patched = src[:at + 1] + [" " * indent_ref + ln[indent_doc:] for ln in plus] + src[at + 1:]

    class Report:
        def __init__(self, total_operations=0):
            self.total, self.ok, self.errors = total_operations, 0, []

        def add_success(self):
            self.ok += 1

        def add_error(self, name, exc):
            self.errors.append(name)

        def summary(self):
            return f"{self.ok} ok"

    quiet = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, error=lambda *a, **k: None,
                                  debug=lambda *a, **k: None)
    ns = {"logger": quiet, "EnhancementError": ra.EnhancementError, "ErrorReport": Report,
          "get_adaptive_tile_sequence": get_adaptive_tile_sequence, "PyTorchESRGANConfig": mine.PyTorchESRGANConfig,
          "convert_ncnn_model_name": mine.convert_ncnn_model_name}
    exec("\n".join(patched), ns)

    d = mg.MultiGPUDistributor(gpus=[mg.GPUInfo(0, "GPU0", 1, 1, 0.0), mg.GPUInfo(1, "GPU1", 1, 1, 0.0)],
                               strategy=mg.LoadBalanceStrategy.ROUND_ROBIN, workers_per_gpu=1,
                               model_name="RealESRGAN_x2plus", scale=2)
    # (no GPU here: the adapter the patch imports is given the CPU pool and the stand-in engine; nothing else differs)
    monkeypatch.setattr(ra, "enhance_frames_batched",
                        functools.partial(ra.enhance_frames_batched, distributor=d, engine_factory=fake_engine))
    try:
        frames_dir, enhanced = tmp_path / "frames", tmp_path / "enhanced"
        frames_dir.mkdir()
        for i in range(9):
            cv2.imwrite(str(frames_dir / f"frame_{i + 1:08d}.png"), np.full((6, 8, 3), 11 * i, np.uint8))
        ck = _checkpoint_module()
        cm = ck.CheckpointManager(tmp_path, checkpoint_interval=2)
        cm.create_checkpoint(stage="extract", total_frames=9, source_path="clip.mp4")
        progress = []
        me = types.SimpleNamespace(
            _dedup_result=None, checkpoint_manager=cm, metadata={"width": 8, "height": 6},
            config=types.SimpleNamespace(frames_dir=frames_dir, enhanced_dir=enhanced, unique_frames_dir=tmp_path / "u",
                                         model_name="realesrgan-x2plus", scale_factor=2, gpu_id=None,
                                         continue_on_error=False, parallel_frames=1),
            _get_enhancement_backend=lambda: "pytorch", _get_tile_size=lambda: 0,
            _update_progress=lambda **k: progress.append(k))
        assert ns["enhance_frames"](me) == 9
        assert sorted(p.name for p in enhanced.glob("*.png")) == [f"frame_{i + 1:08d}.png" for i in range(9)]
        got = cv2.imread(str(enhanced / "frame_00000004.png"), cv2.IMREAD_UNCHANGED)
        assert got.shape == (12, 16, 3) and int(got[0, 0, 0]) == 33
        loaded = cm.load_checkpoint()
        assert loaded.stage == "enhance" and sorted(f.frame_number for f in loaded.frames if f.processed) == list(range(1, 10))
        done = [p["frames_completed"] for p in progress if "eta_seconds" not in p]
        assert done[0] == 0 and sorted(done[1:]) == list(range(1, 10)) and progress[-1]["progress"] == 1.0
        assert me._error_report.ok == 9 and not me._error_report.errors
        # a second run resumes: the checkpoint holds every frame, nothing is enhanced again
        progress.clear()
        assert ns["enhance_frames"](me) == 9 and not progress
        # the ncnn backend is untouched by the patch: the reference's own path continues below it
        me._get_enhancement_backend = lambda: "ncnn"
        me._enhance_frames_sequential = lambda frames, tile, seq, rep: 0
        me.checkpoint_manager = None
        assert ns["enhance_frames"](me) == 9       # (its final count is the directory listing)
    finally:
        d.close()
