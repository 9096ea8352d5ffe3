"""The optional operator interfaces of SURVEY.md section 8 (b) item 3, driven by the UNMODIFIED reference files:
`plugins/manager.py` loads `plugins_contrib/b200_realesrgan.py` from a plugin directory and hands out the processor;
`engine/pipeline.py`'s `Pipeline` runs a stage whose processor is `B200VideoProcessor` / `B200FrameProcessor`;
`infrastructure/gpu/backends/base.py`'s registry constructs `B200Backend` after `register_backend`.

No GPU here: as in tests/test_reference_seam.py the engine behind `RealESRGANer` is replaced, for these tests only, by
the CPU oracle -- they check the seams (registration, settings, batching, sizes, error behaviour), not the CUDA path
(tests/test_gpu_boundary.py::test_plugin_adapters_on_the_device does that).
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/src/framewright"
HAVE_REF = os.path.isdir(REF)
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
MODEL = "RealESRGAN_x4plus_anime_6B"


@pytest.fixture()
def oracle_engine(monkeypatch, tmp_path):
    """get_upsampler works without a GPU: oracle-backed engine, synthetic checkpoints in a weights directory."""
    import torch

    import framewright_b200  # noqa: F401
    from framewright_b200 import pytorch_realesrgan as pr
    from framewright_b200 import upsampler as up_mod
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    calls = []

    class OracleEngine:
        def __init__(self, arch, state_dict, gpu_id=0):
            self.name = next(k for k, v in up_mod.MODEL_ARCHS.items() if v == arch)
            self.sd, self.gpu_id = state_dict, gpu_id

        def upscale_host(self, frames, out=None, tile=0, tile_pad=10, pre_pad=0):
            calls.append((self.name, tuple(np.shape(frames)), tile))
            up = oracle.make_upsampler(self.name, self.sd, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad)
            if np.ndim(frames) == 4:
                return np.stack([up.enhance(f)[0] for f in frames])
            return up.enhance(frames)[0]

        def close(self):
            pass

    wdir = tmp_path / "weights"
    wdir.mkdir()
    for name in (MODEL, "RealESRGAN_x2plus"):
        torch.save({"params_ema": make_synthetic_state_dict(name, 0)}, str(wdir / f"{name}.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    monkeypatch.setattr(up_mod, "B200Engine", OracleEngine)
    monkeypatch.setattr(pr, "is_pytorch_esrgan_available", lambda: True)
    monkeypatch.setattr(pr, "_auto_tile", lambda gpu_id: 0)
    from framewright_b200 import plugin_adapters as pa

    monkeypatch.setattr(pa, "is_pytorch_esrgan_available", lambda: True)
    pr.clear_upsampler_cache()
    yield calls, wdir
    pr.clear_upsampler_cache()


@pytest.fixture()
def ref_packages(monkeypatch):
    """Stub parent packages (the reference package cannot be imported as a whole, SURVEY.md finding 4); the modules
    under test are the reference's own files."""
    saved = {k: v for k, v in sys.modules.items() if k == "framewright" or k.startswith("framewright.")}
    for k in saved:
        del sys.modules[k]
    for name, path in [("framewright", REF), ("framewright.engine", REF + "/engine"),
                       ("framewright.infrastructure", REF + "/infrastructure"),
                       ("framewright.infrastructure.gpu", REF + "/infrastructure/gpu"),
                       ("framewright.infrastructure.gpu.backends", REF + "/infrastructure/gpu/backends")]:
        m = types.ModuleType(name)
        m.__path__ = [path]
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    yield
    for k in [k for k in sys.modules if k == "framewright" or k.startswith("framewright.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _frames(n, h=20, w=24, seed=50):
    from oracle import oracle

    return [oracle.synthetic_frame(h, w, seed=seed + i, kind="mixed") for i in range(n)]


def _want(frames, name=MODEL):
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    up = oracle.make_upsampler(name, make_synthetic_state_dict(name, 0))
    return [up.enhance(f)[0] for f in frames]


# ---- the session all adapters share (no reference needed) ------------------------------------------------------------
def test_session_settings_devices_and_sizes(oracle_engine):
    from framewright_b200 import plugin_adapters as pa

    assert pa.gpu_id_of("cuda") == 0 and pa.gpu_id_of("cuda:3") == 3 and pa.gpu_id_of(2) == 2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pa.gpu_id_of("cpu")
    s = pa.UpscaleSession("cuda:0", {"model": "realesrgan-x2plus", "tile": 32, "unknown_key": 1})
    assert s.settings["model_name"] == "RealESRGAN_x2plus" and s.net_scale == 2 and s.scale == 2.0
    assert s.config().tile_size == 32 and s.output_size((45, 63)) == (90, 126)
    s.update({"scale": 3})                              # an outscale that is not the network's: resized result
    assert s.output_size((10, 20)) == (30, 60)
    f = _frames(1)[0]
    assert s.frame(f).shape == (60, 72, 3)
    assert pa.UpscaleSession("cuda:0", {"model": "no-such-model"}).settings["model_name"] == "RealESRGAN_x4plus"   # :263-275


def test_frame_processor_batches_same_size_runs(oracle_engine):
    from framewright_b200 import plugin_adapters as pa

    calls, _ = oracle_engine
    proc = pa.B200FrameProcessor(device="cuda:0", model_name=MODEL, max_batch=3)
    frames = _frames(5) + _frames(2, h=16, w=16, seed=70)
    want = _want(frames)
    got = proc.process_frames(frames)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert [c[1][0] if len(c[1]) == 4 else 1 for c in calls] == [3, 2, 2]     # 5 same-size frames: 3 + 2; then 2
    assert np.array_equal(proc.process_frame(frames[0]), want[0]) and np.array_equal(proc(frames[1]), want[1])
    out = proc.process_frame(frames[0], scale=2)                               # per-call stage params
    assert out.shape == (40, 48, 3)
    proc.close()


def test_video_processor_round_trip(oracle_engine, tmp_path):
    """Video file in, upscaled video file out (cv2 codecs): every frame, scaled size, progress to 1.0; failures raise."""
    import cv2

    from framewright_b200 import plugin_adapters as pa

    src, dst = tmp_path / "in.avi", tmp_path / "out.avi"
    frames = _frames(7, h=16, w=24)
    w = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (24, 16))
    assert w.isOpened()
    for f in frames:
        w.write(f)
    w.release()
    prog = []
    vp = pa.B200VideoProcessor(device="cuda:0", fourcc="MJPG", model_name=MODEL, max_batch=3)
    assert vp.process_video(src, dst, progress_callback=prog.append) is True
    assert vp.frames_processed == 7 and prog[-1] == 1.0 and prog == sorted(prog)
    cap = cv2.VideoCapture(str(dst))
    assert int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == 96 and int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)) == 64
    n = 0
    while cap.read()[0]:
        n += 1
    cap.release()
    assert n == 7
    with pytest.raises(ValueError, match="Cannot open video"):
        vp.process_video(tmp_path / "missing.avi", dst)
    vp.close()


# ---- the unmodified reference files ----------------------------------------------------------------------------------
@needs_ref
def test_reference_plugin_manager_loads_and_runs_the_plugin_file(oracle_engine, ref_packages, tmp_path, monkeypatch):
    import shutil

    import framewright_b200
    from framewright_b200 import plugin_adapters as pa

    monkeypatch.setenv("HOME", str(tmp_path / "home"))            # PluginManager also scans ~/.framewright/plugins
    plugins = importlib.import_module("framewright.plugins")     # the reference's package: base, manager, hooks
    pdir = tmp_path / "plugins"
    pdir.mkdir()
    shutil.copy(os.path.join(os.path.dirname(framewright_b200.__file__), "plugins_contrib", "b200_realesrgan.py"), pdir)
    mgr = plugins.PluginManager(plugin_dirs=[pdir])               # auto_load -> PluginLoader.load_from_file
    infos = mgr.find_plugins_for_capability(plugins.PluginCapability.UPSCALE)
    assert [i.metadata.name for i in infos] == [pa.PLUGIN_NAME]
    listed = mgr.list_plugins(plugin_type="processor")
    assert listed[0]["name"] == pa.PLUGIN_NAME and listed[0]["supports_cpu"] is False and listed[0]["capabilities"] == ["UPSCALE"]
    # the manager's default device is "cpu": the plugin refuses it, get_plugin logs and returns None (manager.py:318-321)
    assert mgr.get_processor(pa.PLUGIN_NAME) is None
    mgr.set_device("cuda:0")
    up = mgr.get_processor(pa.PLUGIN_NAME, {"model_name": MODEL, "max_batch": 4})
    assert isinstance(up, plugins.ProcessorPlugin) and up.is_initialized and up.supports_batch()
    assert up.validate_requirements() == [] or all("VRAM" in i for i in up.validate_requirements())
    frames = _frames(3)
    want = _want(frames)
    assert np.array_equal(up.process_frame(frames[0], 0), want[0])
    got = up.process_batch(frames, start_frame=0)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert up.estimate_output_size((20, 24)) == (80, 96) and up.get_temporal_radius() == 0
    again = mgr.get_processor(pa.PLUGIN_NAME, {"model": "realesrgan-x2plus"})    # same instance, settings updated
    assert again is up and up.estimate_output_size((20, 24)) == (40, 48)
    assert up.process_frame(frames[0], 0).shape == (40, 48, 3)
    mgr.release_all()
    assert not up.is_initialized
    with pytest.raises(RuntimeError, match="not initialized"):
        up.process_frame(frames[0], 0)
    # registration without a file: register_plugin(manager)
    mgr2 = plugins.PluginManager(plugin_dirs=[], auto_load=False)
    cls = pa.register_plugin(mgr2)
    assert mgr2.registry.get(pa.PLUGIN_NAME).plugin_class is cls and issubclass(cls, plugins.ProcessorPlugin)


@needs_ref
def test_reference_pipeline_runs_an_upscale_stage(oracle_engine, ref_packages, tmp_path):
    import cv2

    from framewright_b200 import plugin_adapters as pa

    pl = importlib.import_module("framewright.engine.pipeline")
    src, dst = tmp_path / "in.avi", tmp_path / "out.avi"
    w = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (24, 16))
    for f in _frames(4, h=16, w=24):
        w.write(f)
    w.release()
    pipe = pl.Pipeline(name="upscale-only")
    vp = pa.B200VideoProcessor(device="cuda:0", fourcc="MJPG", model_name=MODEL)
    pipe.add_stage(vp, config=pl.StageConfig(params={"max_batch": 2}), name="upscale")
    events = []
    pipe.on_event(None, lambda e: events.append(e.event_type.name))
    res = pipe.run(src, dst, resume_from_checkpoint=False)
    assert res.succeeded, res.error
    assert res.get_stage_result("upscale").output is True and vp.frames_processed == 4
    assert "STAGE_COMPLETED" in events and "PROGRESS_UPDATE" in events
    cap = cv2.VideoCapture(str(dst))
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (96, 64)
    cap.release()
    # a failing stage is a raised exception there: the pipeline reports FAILED
    res = pipe.run(tmp_path / "missing.avi", dst, resume_from_checkpoint=False)
    assert not res.succeeded and "Cannot open video" in str(res.get_stage_result("upscale").error)
    # the frame-processor protocol object is recognised too (hasattr(processor, "process_frame"), pipeline.py:1174)
    assert hasattr(pa.B200FrameProcessor, "process_frame") and not hasattr(pa.B200FrameProcessor, "process_video")


@needs_ref
def test_reference_backend_registry_constructs_the_b200_backend(oracle_engine, ref_packages):
    from framewright_b200 import plugin_adapters as pa

    calls, wdir = oracle_engine
    base = importlib.import_module("framewright.infrastructure.gpu.backends.base")
    det = importlib.import_module("framewright.infrastructure.gpu.detector")
    cls = pa.register_compute_backend()
    assert issubclass(cls, base.Backend)
    be = base.get_backend(det.BackendType.CUDA, device_id=0, force_new=True)      # the reference's factory
    assert type(be) is cls and be.backend_type == det.BackendType.CUDA and "B200" in be.name
    with be:                                                                        # initialize / cleanup
        assert be.is_initialized
        caps = be.get_capabilities()
        assert MODEL in caps.supported_models and caps.supports_batching and caps.to_dict()["backend_type"] == "cuda"
        assert set(be.get_memory_info()) >= {"total_mb", "used_mb", "free_mb"}
        assert be.load_model(MODEL) is True
        assert be.load_model("hat_l") is False                                      # not a model of this path
        assert be.load_model("x2", model_path=wdir / "RealESRGAN_x2plus.pth", model="RealESRGAN_x2plus") is True
        frames = _frames(3)
        want = _want(frames)
        assert np.array_equal(be.run_inference(MODEL, frames[0]), want[0])
        stack = be.run_inference(MODEL, np.stack(frames))
        assert stack.shape == (3, 80, 96, 3) and all(np.array_equal(a, b) for a, b in zip(stack, want))
        assert be.run_inference("x2", frames[0]).shape == (40, 48, 3)
        with pytest.raises(ValueError, match="not loaded"):
            be.run_inference("RealESRGAN_x4plus", frames[0])
        be.unload_model("x2")
        be.unload_model(MODEL)
    assert not be.is_initialized
