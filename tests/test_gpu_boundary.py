"""Every layer of the drop-in boundary ON THE DEVICE: module functions (`get_upsampler`, `enhance_frame_pytorch`,
`clear_upsampler_cache`), the `RealESRGANer` duck type through the `realesrgan` / `basicsr` shim imports exactly as
the reference's call sites construct it (`cli.py:715-750`, `face_restore.py:379-401`), the operator backend
(`upscale_frame` / `upscale_frames` / `upscale_array`), error behaviour (OOM -> (False, "GPU out of memory ..."),
missing weights -> error, never random weights) and thread-parallel callers (`restorer.py:1894`)."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODELS = ["RealESRGAN_x4plus", "RealESRGAN_x4plus_anime_6B", "RealESRGAN_x2plus", "realesr-general-x4v3",
          "realesr-animevideov3"]


@pytest.fixture()
def weights_dir(tmp_path, monkeypatch):
    """Checkpoint files as upstream ships them ({'params_ema': state_dict}), synthetic weights inside."""
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200 import pytorch_realesrgan as pr

    d = tmp_path / "weights"
    d.mkdir()
    for i, name in enumerate(MODELS):
        key = "params_ema" if i % 2 == 0 else "params"
        torch.save({key: make_synthetic_state_dict(name, 0)}, str(d / f"{name}.pth"))
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(d))
    pr.clear_upsampler_cache()
    yield d
    pr.clear_upsampler_cache()


def _gate(ref, got, label):
    from oracle import oracle

    rep = oracle.parity_report(ref, got)
    print(label, rep)
    assert ref.shape == got.shape
    assert rep["frac_within_1lsb"] >= oracle.GATE_FRAC_WITHIN_1LSB and rep["psnr_db"] >= oracle.GATE_PSNR_DB, (label, rep)


def _oracle(name, img, **kw):
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    return oracle.make_upsampler(name, make_synthetic_state_dict(name, 0), **kw).enhance(img)[0]


def test_get_upsampler_auto_tile_and_enhance(native_lib, weights_dir):
    from framewright_b200 import pytorch_realesrgan as pr
    from oracle import oracle

    assert pr.is_pytorch_esrgan_available() is True
    cfg = pr.PyTorchESRGANConfig(model_name="RealESRGAN_x4plus_anime_6B")
    up = pr.get_upsampler(cfg)
    assert up.tile_size == 0                      # reference ladder (:136-151): >= 24 GB -> no tiling; a B200 has 180 GB
    assert pr.get_upsampler(cfg) is up            # cached
    img = oracle.synthetic_frame(60, 90, seed=1, kind="mixed")
    out, mode = up.enhance(img, outscale=4)
    assert mode == "RGB" and out.shape == (240, 360, 3)
    _gate(_oracle("RealESRGAN_x4plus_anime_6B", img), out, "get_upsampler.enhance")
    other = pr.get_upsampler(pr.PyTorchESRGANConfig(model_name="RealESRGAN_x4plus_anime_6B", tile_size=48))
    assert other is not up and other.tile_size == 48   # the cache is keyed by the config (reference :153-156 is not)


@pytest.mark.parametrize("name,scale", [("RealESRGAN_x4plus", 4), ("RealESRGAN_x2plus", 2), ("realesr-general-x4v3", 4)])
def test_enhance_frame_pytorch_file_to_file(native_lib, weights_dir, tmp_path, name, scale):
    import cv2

    from framewright_b200 import pytorch_realesrgan as pr
    from oracle import oracle

    img = oracle.synthetic_frame(54, 76, seed=8, kind="mixed")
    src, dst = tmp_path / "frame_00000001.png", tmp_path / "enhanced" / "frame_00000001.png"
    dst.parent.mkdir()
    cv2.imwrite(str(src), img)
    cfg = pr.PyTorchESRGANConfig(model_name=name, scale_factor=scale)
    ok, err = pr.enhance_frame_pytorch(src, dst, cfg)
    assert (ok, err) == (True, None), err
    assert cfg.tile_size == 0                     # auto mode: > 8000 MB free -> 0 (reference :208-218 mutates the config)
    got = cv2.imread(str(dst), cv2.IMREAD_UNCHANGED)
    _gate(_oracle(name, img), got, f"enhance_frame_pytorch {name}")
    ok, err = pr.enhance_frame_pytorch(tmp_path / "missing.png", dst, cfg)
    assert ok is False and err.startswith("Failed to read image")


def test_oom_is_reported_not_raised(native_lib, weights_dir, tmp_path):
    """A workspace the device cannot hold -> (False, 'GPU out of memory ...'), the cache is dropped (reference
    :237-244), and the next call with a smaller tile succeeds (the caller's retry ladder, restorer.py:1746)."""
    import cv2

    from framewright_b200 import pytorch_realesrgan as pr
    from oracle import oracle

    img = oracle.synthetic_frame(96, 128, seed=3, kind="mixed")
    src, dst = tmp_path / "in.png", tmp_path / "out.png"
    cv2.imwrite(str(src), img)
    cfg = pr.PyTorchESRGANConfig(model_name="RealESRGAN_x4plus_anime_6B", tile_size=0)
    up = pr.get_upsampler(cfg)
    up.engine.set_option("ws_limit_mb", 8)        # stands in for a full device
    ok, err = pr.enhance_frame_pytorch(src, dst, cfg)
    assert ok is False and err.startswith("GPU out of memory") and "memory" in err.lower()
    assert not dst.exists()
    assert pr._UPSAMPLERS == {}                    # cache cleared
    cfg2 = pr.PyTorchESRGANConfig(model_name="RealESRGAN_x4plus_anime_6B", tile_size=32)
    ok, err = pr.enhance_frame_pytorch(src, dst, cfg2)
    assert (ok, err) == (True, None)
    _gate(_oracle("RealESRGAN_x4plus_anime_6B", img, tile=32, tile_pad=10), cv2.imread(str(dst)), "after OOM, tiled")


def test_missing_weights_is_an_error_not_random_weights(native_lib, tmp_path, monkeypatch):
    import cv2

    from framewright_b200 import pytorch_realesrgan as pr
    from framewright_b200.upsampler import RealESRGANer
    from oracle import oracle

    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(tmp_path / "empty"))
    monkeypatch.setenv("B200SR_DOWNLOAD_TIMEOUT", "2")
    monkeypatch.delenv("B200SR_SYNTHETIC_WEIGHTS", raising=False)
    pr.clear_upsampler_cache()
    with pytest.raises(FileNotFoundError):
        RealESRGANer(scale=4, model_path="https://example.invalid/RealESRGAN_x4plus.pth", model_name="RealESRGAN_x4plus")
    src = tmp_path / "in.png"
    cv2.imwrite(str(src), oracle.synthetic_frame(32, 32, seed=1))
    ok, err = pr.enhance_frame_pytorch(src, tmp_path / "out.png", pr.PyTorchESRGANConfig())
    assert ok is False and "not available" in err and not (tmp_path / "out.png").exists()


def test_shim_imports_as_cli_and_face_restore_construct_it(native_lib, weights_dir, monkeypatch):
    """`from realesrgan import RealESRGANer; from basicsr.archs.rrdbnet_arch import RRDBNet` + the exact constructor
    calls of cli.py:715-750 (tile=512, tile_pad=10, pre_pad=10, absolute model_path) and face_restore.py:388-399
    (model_path='RealESRGAN_x4plus.pth', tile=400, pre_pad=0)."""
    from framewright_b200 import shims
    from oracle import oracle

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("realesrgan", "basicsr")}
    shims.install()
    try:
        from basicsr.archs.rrdbnet_arch import RRDBNet
        from realesrgan import RealESRGANer
        import realesrgan

        assert isinstance(realesrgan.__version__, str)
        img = oracle.synthetic_frame(70, 110, seed=12, kind="mixed")
        # cli.py
        model_arch = RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=4)
        up = RealESRGANer(scale=4, model_path=str(weights_dir / "RealESRGAN_x4plus.pth"), model=model_arch, tile=512,
                          tile_pad=10, pre_pad=10, half=True)
        out, _ = up.enhance(img, outscale=4)
        _gate(_oracle("RealESRGAN_x4plus", img, tile=512, tile_pad=10, pre_pad=10), out, "cli.py pattern")
        up.close()
        # face_restore.py (GFPGAN background upsampler): bare file name, tile=400; called on BGR crops, outscale 2
        model = RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=4)
        bg = RealESRGANer(scale=4, model_path="RealESRGAN_x4plus.pth", model=model, tile=400, tile_pad=10, pre_pad=0,
                          half=True)
        big = oracle.synthetic_frame(450, 420, seed=13, kind="mixed")      # > one 400-pixel tile in both directions
        out, _ = bg.enhance(big, outscale=4)
        _gate(_oracle("RealESRGAN_x4plus", big, tile=400, tile_pad=10, pre_pad=0), out, "face_restore.py pattern")
        out2, _ = bg.enhance(img, outscale=2)      # GFPGANer asks for its own upscale factor: LANCZOS4 resize after x4
        assert out2.shape == (140, 220, 3)
        bg.close()
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("realesrgan", "basicsr")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_backend_frame_dir_and_array_paths(native_lib, weights_dir, tmp_path):
    import cv2

    from framewright_b200.super_resolution import B200RealESRGANBackend, SRConfig
    from oracle import oracle

    be = B200RealESRGANBackend(SRConfig(scale=4), None, "anime")
    assert be.is_available() and be.name == "realesrgan_anime" and be.supported_scales == [4]
    frames = [oracle.synthetic_frame(40, 64, seed=20 + i, kind="mixed") for i in range(5)]
    one = be.upscale_frame(frames[0], scale=4)
    _gate(_oracle("RealESRGAN_x4plus_anime_6B", frames[0]), one, "backend.upscale_frame")
    arr = be.upscale_array(np.stack(frames))
    assert arr.shape == (5, 160, 256, 3) and np.array_equal(arr[0], one)
    ind, outd = tmp_path / "frames", tmp_path / "enhanced"
    ind.mkdir()
    for i, f in enumerate(frames):
        cv2.imwrite(str(ind / f"frame_{i + 1:08d}.png"), f)
    (ind / "frame_00000099.png").write_bytes(b"not a png")            # unreadable frame -> warning, not an exception
    seen = []
    res = be.upscale_frames(ind, outd, scale=4, progress_callback=seen.append)
    assert res.frames_processed == 5 and res.frames_failed == 1 and res.output_dir == outd
    assert any(w.startswith("Frame frame_00000099.png:") for w in res.warnings)
    assert seen and abs(seen[-1] - 1.0) < 1e-9 and seen == sorted(seen)
    for i in range(5):
        got = cv2.imread(str(outd / f"frame_{i + 1:08d}.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(got, arr[i])
    be.clear_cache()


def test_thread_parallel_callers_share_one_upsampler(native_lib, weights_dir):
    """restorer.py:1894 runs enhance_frame_pytorch from `parallel_frames` threads: concurrent `enhance` calls on the
    one cached upsampler (engine lanes) return what sequential calls return, for mixed sizes."""
    from framewright_b200 import pytorch_realesrgan as pr
    from oracle import oracle

    up = pr.get_upsampler(pr.PyTorchESRGANConfig(model_name="RealESRGAN_x4plus_anime_6B"))
    imgs = [oracle.synthetic_frame(40 + 8 * (i % 3), 64 + 16 * (i % 2), seed=50 + i, kind="mixed") for i in range(12)]
    want = [up.enhance(im)[0].copy() for im in imgs]
    got = [None] * len(imgs)
    errs = []

    def work(k):
        try:
            for i in range(k, len(imgs), 4):
                got[i] = up.enhance(imgs[i])[0]
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    # a call racing with clear_upsampler_cache() either completes or raises a clean EngineError -- never a crash
    pr.clear_upsampler_cache()
    from framewright_b200.engine import EngineError
    with pytest.raises(EngineError):
        up.enhance(imgs[0])


def test_host_path_chunks_lanes_and_pinned_buffers(native_lib):
    """The host-buffer call with more frames than one lane job: pageable and pinned inputs, 1 and 3 lanes, explicit
    chunk sizes -- identical bytes to frame-by-frame calls."""
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import PINNED_POOL, B200Engine
    from oracle import oracle

    name = "realesr-animevideov3"
    eng = B200Engine(name, make_synthetic_state_dict(name, 0), gpu_id=0)
    frames = np.stack([oracle.synthetic_frame(72, 100, seed=70 + i, kind="mixed") for i in range(7)])
    want = np.stack([eng.upscale_host(f) for f in frames])
    for lanes, chunk in ((1, 2), (2, 1), (3, 2), (2, 0)):
        eng.set_option("lanes", lanes)
        eng.set_option("host_chunk", chunk)
        assert np.array_equal(eng.upscale_host(frames), want), (lanes, chunk)
        pin = PINNED_POOL.empty(frames.shape, np.uint8)
        pin[...] = frames
        out = np.empty_like(want)                                  # pageable destination
        assert np.array_equal(eng.upscale_host(pin, out=out), want), (lanes, chunk, "pinned in / pageable out")
    assert eng.last_launch_count > 0
    eng.close()


def test_engine_load_is_strict(native_lib):
    """upstream `load_state_dict(strict=True)`: a missing or unexpected key, or a wrong shape, is an error."""
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import B200Engine, EngineError

    name = "realesr-animevideov3"
    sd = make_synthetic_state_dict(name, 0)
    missing = {k: v for k, v in sd.items() if k != "body.4.bias"}
    with pytest.raises(EngineError, match="missing"):
        B200Engine(name, missing, gpu_id=0)
    extra = dict(sd, **{"body.99.weight": torch.zeros(1)})
    with pytest.raises(EngineError, match="unexpected"):
        B200Engine(name, extra, gpu_id=0)
    bad = dict(sd, **{"body.2.weight": torch.zeros(64, 32, 3, 3)})
    with pytest.raises(EngineError, match="shape"):
        B200Engine(name, bad, gpu_id=0)
    B200Engine(name, sd, gpu_id=0).close()


def test_engine_workspace_formula_matches_the_c_abi(native_lib):
    """`tile_sizing.engine_workspace_bytes` (pure Python, used for tile-size decisions without touching the device)
    equals `b200sr_workspace_bytes` for every architecture, whole frames, tiles, pads and batches."""
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import B200Engine
    from framewright_b200.tile_sizing import calculate_optimal_tile_size, engine_workspace_bytes

    for name in ("RealESRGAN_x4plus_anime_6B", "RealESRGAN_x2plus", "realesr-animevideov3"):
        eng = B200Engine(name, make_synthetic_state_dict(name, 0), gpu_id=0)
        for (n, h, w, tile, pad, pre) in [(1, 720, 1280, 0, 10, 0), (4, 720, 1280, 0, 10, 0), (2, 721, 1279, 512, 10, 10),
                                          (8, 256, 256, 0, 10, 0), (1, 1080, 1920, 400, 10, 0), (3, 45, 63, 32, 4, 3)]:
            assert eng.workspace_bytes(n, h, w, tile, pad, pre) == engine_workspace_bytes(name, w, h, n, tile, pad, pre), \
                (name, n, h, w, tile, pad, pre)
        eng.close()
    free_mb = torch.cuda.mem_get_info(0)[0] >> 20
    assert calculate_optimal_tile_size((1280, 720), 4, available_vram_mb=free_mb) == 0      # a 720p frame fits whole
    assert calculate_optimal_tile_size((1280, 720), 4, available_vram_mb=2000) >= 128


def test_plugin_adapters_on_the_device(native_lib, weights_dir, tmp_path):
    """The adapters behind the reference's optional operator interfaces (plugin_adapters.py) with the REAL engine:
    frame processor (batched runs of same-size frames == single calls, gate vs the oracle), a session that owns a
    checkpoint named by path, and the video processor (cv2 file in -> upscaled cv2 file out)."""
    import cv2

    from framewright_b200 import plugin_adapters as pa
    from framewright_b200 import pytorch_realesrgan as pr
    from oracle import oracle

    name = "RealESRGAN_x4plus_anime_6B"
    frames = [oracle.synthetic_frame(40, 56, seed=60 + i, kind="mixed") for i in range(5)]
    frames.append(oracle.synthetic_frame(32, 48, seed=70, kind="noise"))
    proc = pa.B200FrameProcessor(device="cuda:0", model_name=name, max_batch=4)
    got = proc.process_frames(frames)
    up = pr.get_upsampler(pr.PyTorchESRGANConfig(model_name=name, gpu_id=0))
    for f, g in zip(frames, got):
        assert np.array_equal(g, up.enhance(f)[0])
    _gate(_oracle(name, frames[0]), got[0], "B200FrameProcessor")
    assert np.array_equal(proc.process_frame(frames[5]), got[5])
    own = pa.UpscaleSession("cuda:0", {"model_name": name}, model_path=weights_dir / f"{name}.pth")
    assert own.open() is not up and np.array_equal(own.frame(frames[1]), got[1])
    own.close()
    src, dst = tmp_path / "in.avi", tmp_path / "out.avi"
    w = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (56, 40))
    assert w.isOpened()
    for f in frames[:5]:
        w.write(f)
    w.release()
    prog = []
    vp = pa.B200VideoProcessor(device="cuda:0", fourcc="MJPG", model_name=name, max_batch=2)
    assert vp.process_video(src, dst, progress_callback=prog.append) is True
    assert vp.frames_processed == 5 and prog[-1] == 1.0
    cap = cv2.VideoCapture(str(dst))
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (224, 160)
    n = 0
    while cap.read()[0]:
        n += 1
    cap.release()
    assert n == 5
    assert proc.process_frame(frames[0], scale=2).shape == (80, 112, 3)     # outscale != net scale: resized result
    proc.close()
