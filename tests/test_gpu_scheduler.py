"""The product scheduler with REAL engines: one worker process per visible GPU (2+ when the box has them), files,
shared-memory arrays and the ordered ring; results must equal what one engine produces in this process."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAME = "RealESRGAN_x4plus_anime_6B"


@pytest.fixture(scope="module")
def setup(native_lib):
    from framewright_b200 import multi_gpu as mg
    from framewright_b200.archs import make_synthetic_state_dict
    from framewright_b200.engine import B200Engine
    from oracle import oracle

    os.environ["B200SR_SYNTHETIC_WEIGHTS"] = "0"      # the workers inherit it (explicit opt-in to synthetic weights)
    gpus = mg.query_gpus()[:4]
    frames = np.stack([oracle.synthetic_frame(48, 80, seed=200 + i, kind="mixed") for i in range(23)])
    eng = B200Engine(NAME, make_synthetic_state_dict(NAME, 0), gpu_id=0)
    want = eng.upscale_host(frames)
    eng.close()
    d = mg.MultiGPUDistributor(gpus=gpus, strategy=mg.LoadBalanceStrategy.ROUND_ROBIN, workers_per_gpu=2, batch=2,
                               model_name=NAME, scale=4)
    yield d, gpus, frames, want
    d.close()
    os.environ.pop("B200SR_SYNTHETIC_WEIGHTS", None)


def test_frames_dir_through_the_distributor(setup, tmp_path):
    import cv2

    d, gpus, frames, want = setup
    ind = tmp_path / "frames"
    ind.mkdir()
    paths = []
    for i, f in enumerate(frames):
        p = ind / f"frame_{i + 1:08d}.png"
        cv2.imwrite(str(p), f)
        paths.append(p)
    prog = []
    res = d.distribute_frames(paths, None, tmp_path / "enhanced", progress_callback=lambda f, m: prog.append(f))
    assert res.total_frames == len(frames) and not res.errors, res.errors
    assert len(prog) == len(frames) and abs(prog[-1] - 1.0) < 1e-9
    for i in range(len(frames)):
        got = cv2.imread(str(tmp_path / "enhanced" / f"frame_{i + 1:08d}.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(got, want[i]), i
    if len(gpus) > 1:
        assert all(len(v) > 0 for v in res.frames_per_gpu.values())      # every GPU took part


def test_frame_arrays_and_ordered_stream_through_the_pool(setup):
    from framewright_b200.scheduler import ArraySink, ArraySource, SharedArray

    d, gpus, frames, want = setup
    pool = d._get_pool([g.id for g in gpus])
    n, h, w = frames.shape[:3]
    sin, sout = SharedArray(frames.shape), SharedArray(want.shape)
    try:
        sin.array[...] = frames
        res = pool.run(ArraySource(sin), ArraySink(sout), d._esr_config(), batch=3)
        assert sorted(res.ok) == list(range(n)) and not res.errors
        assert np.array_equal(sout.array, want)
    finally:
        sin.release()
        sout.release()
    got = {}
    order = []

    def emit(i, out):
        order.append(i)
        got[i] = out.copy()

    res = pool.stream(iter(frames), d._esr_config(), emit, num_frames=n, frame_shape=(h, w), scale=4, batch=2, window=8)
    assert not res.errors and order == list(range(n))
    assert all(np.array_equal(got[i], want[i]) for i in range(n))


def test_batch_path_facade_and_process_frames_on_the_device(setup, tmp_path):
    """The callers either side of the path with REAL engines: `enhance_frames_batched` (checkpoint + progress per
    frame), `SuperResolution(...).upscale(dir, dir)` / `.process(list)`, `multi_gpu_process_frames`."""
    import types

    import cv2

    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import enhance_frames_batched, multi_gpu_process_frames
    from framewright_b200.super_resolution import SRConfig, SuperResolution

    d, gpus, frames, want = setup
    ind = tmp_path / "frames"
    ind.mkdir()
    paths = []
    for i, f in enumerate(frames[:9]):
        p = ind / f"frame_{i + 1:08d}.png"
        cv2.imwrite(str(p), f)
        paths.append(p)
    # f2: the batch path behind VideoRestorer.enhance_frames
    updates, prog = [], []
    cm = types.SimpleNamespace(load_checkpoint=lambda: None, update_stage=lambda s: updates.append(("stage", s)),
                               update_frame=lambda **k: updates.append(("frame", k["frame_number"], str(k["output_path"]))),
                               get_unprocessed_frames=lambda fs: fs, force_save=lambda: None)
    cfg = PyTorchESRGANConfig(model_name=NAME, scale_factor=4)
    n = enhance_frames_batched(paths, tmp_path / "enhanced", cfg, checkpoint_manager=cm, distributor=d,
                               update_progress=lambda **k: prog.append(k["frames_completed"]))
    assert n == 9 and updates[0] == ("stage", "enhance") and sorted(u[1] for u in updates[1:]) == list(range(1, 10))
    assert prog[0] == 0 and sorted(prog[1:-1]) == list(range(1, 10))
    for i in range(9):
        assert np.array_equal(cv2.imread(str(tmp_path / "enhanced" / f"frame_{i + 1:08d}.png"), cv2.IMREAD_UNCHANGED), want[i])
    # a9: the facade, frames directory and frame list
    sr = SuperResolution(SRConfig(scale=4, backend="realesrgan_anime"))
    res = sr.upscale(ind, tmp_path / "sr_out")
    assert res.frames_processed == 9 and res.frames_failed == 0 and res.backend_used == "realesrgan_anime"
    assert np.array_equal(cv2.imread(str(tmp_path / "sr_out" / "frame_00000004.png"), cv2.IMREAD_UNCHANGED), want[3])
    outs = sr.process([f for f in frames[:5]], scale=4)
    assert all(np.array_equal(o, w) for o, w in zip(outs, want[:5]))
    sr.clear_cache()
    # f4: MultiGPUProcessor.process_frames shape over shared memory
    pool = d._get_pool([g.id for g in gpus])
    results = multi_gpu_process_frames([f for f in frames[:7]], cfg, pool=pool)
    assert [r.frame_index for r in results] == list(range(7)) and all(r.success for r in results)
    assert all(np.array_equal(r.output, want[r.frame_index]) for r in results)


def test_raw_video_pipe_through_the_gpus(setup):
    """f1: raw bgr24 frames from a pipe -> ordered shared-memory ring over the GPUs -> raw bgr24 frames to a pipe."""
    import io

    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import upscale_raw_stream

    d, gpus, frames, want = setup
    pool = d._get_pool([g.id for g in gpus])
    n, h, w = frames.shape[:3]
    src, dst = io.BytesIO(frames.tobytes()), io.BytesIO()
    cfg = PyTorchESRGANConfig(model_name=NAME, scale_factor=4)
    assert upscale_raw_stream(src, dst, w, h, cfg, num_frames=n, pool=pool) == n
    assert np.array_equal(np.frombuffer(dst.getvalue(), np.uint8).reshape(want.shape), want)
    # single engine in this process (no pool): same bytes
    src, dst = io.BytesIO(frames[:6].tobytes()), io.BytesIO()
    assert upscale_raw_stream(src, dst, w, h, cfg, num_frames=6, batch=4) == 6
    assert np.array_equal(np.frombuffer(dst.getvalue(), np.uint8).reshape(want[:6].shape), want[:6])
