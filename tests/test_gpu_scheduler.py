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
