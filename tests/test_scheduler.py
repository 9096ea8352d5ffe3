"""The product multi-GPU scheduler on CPU: real worker processes (spawn), a stand-in engine.  Covers sharding +
stealing, per-frame callbacks, retry on another GPU (<= 3 attempts), a worker that dies, the frame-array and the
ordered ring (stream) transports, and the file path."""
import os
import threading

import numpy as np
import pytest

from sched_helpers import CountingSource, fake_engine

CFG = {"model_name": "RealESRGAN_x2plus", "scale_factor": 2, "tile_size": 0, "tile_pad": 0, "pre_pad": 0}


@pytest.fixture(scope="module")
def pool():
    import framewright_b200  # noqa: F401
    from framewright_b200.scheduler import SchedulerPool

    p = SchedulerPool([0, 1, 2], workers_per_gpu=2, start_timeout=120)
    yield p
    p.close()


def test_claim_table_shards_and_steals_from_the_back():
    import multiprocessing as mp

    from framewright_b200.scheduler import ClaimTable

    t = ClaimTable(mp.get_context("spawn"), 2)
    t.reset([(0, 5), (5, 10)], 10)
    assert t.claim(0, 2) == [0, 1] and t.claim(0, 2) == [2, 3] and t.claim(0, 2) == [4]
    assert t.claim(0, 2) == [8, 9]                 # own shard empty: the BACK of the other one
    assert t.claim(1, 4) == [5, 6, 7] and t.claim(1, 1) == [] and t.remaining() == 0
    t.reset([(0, 10), (0, 0)], 3)                  # ring mode: only indices below `avail` may be taken
    assert t.claim(0, 4) == [0, 1, 2] and t.claim(0, 4) == []
    t.set_avail(5)
    assert t.claim(1, 4) == [] and t.claim(0, 4) == [3, 4]
    t.set_avail(9)                                 # front = 0: shard 1's worker draws from the FRONT of shard 0 ...
    assert t.claim(1, 2, steal=False, front=0) == [5, 6] and t.claim(0, 2) == [7, 8]
    assert t.held_by(1) == [5, 6]                  # ... and holds what it took in its own row
    assert t.claim(1, 2, steal=False, front=0) == []


def test_every_frame_once_with_per_frame_callbacks(pool):
    from framewright_b200.scheduler import ChecksumSink

    n = 50
    seen, prog = [], []
    res = pool.run(CountingSource(n), ChecksumSink(), CFG, batch=2, engine_factory=fake_engine,
                   frame_callback=lambda i, name, ok, err, gpu: seen.append((i, name, ok, gpu)),
                   progress_callback=lambda f, m: prog.append((f, m)))
    assert sorted(res.ok) == list(range(n)) and not res.errors and not res.retried
    assert sorted(i for i, *_ in seen) == list(range(n))              # one callback per frame, as it completes
    assert [p[0] for p in prog] == sorted(p[0] for p in prog) and abs(prog[-1][0] - 1.0) < 1e-12
    assert prog[-1][1] == f"Processed {n}/{n} frames"
    assert sum(len(v) for v in res.frames_per_gpu.values()) == n
    assert all(seen_name == f"frame_{i + 1:08d}.png" for i, seen_name, _, _ in seen)
    assert len(res.ok[7]) == 2 and isinstance(res.ok[7][1], int)       # sink info travels back


def test_slow_gpu_loses_its_tail_to_the_others(pool):
    from framewright_b200.scheduler import ChecksumSink

    n = 60
    cfg = dict(CFG, tile_pad=30)                    # stand-in: GPU 0 takes 30 ms per frame, the others ~0
    res = pool.run(CountingSource(n), ChecksumSink(), cfg, batch=2, engine_factory=fake_engine)
    assert sorted(res.ok) == list(range(n))
    assert len(res.frames_per_gpu[0]) < n // 3 and res.stolen > 0      # its contiguous shard was [0, 20)
    assert res.total_time < 0.030 * 20 * 0.9 + 2.0


def test_failed_frames_are_retried_on_another_gpu(pool):
    from framewright_b200.scheduler import ChecksumSink

    n = 24
    # (steal=False: GPU 1's shard can only be claimed by GPU 1, so it certainly fails frames -- with stealing the
    # other two may empty its shard before its runner threads have started)
    res = pool.run(CountingSource(n), ChecksumSink(), CFG, batch=2, engine_factory=fake_engine, fail_on_gpus=(1,),
                   steal=False)
    assert sorted(res.ok) == list(range(n)) and not res.errors
    assert res.retried and not res.frames_per_gpu[1]
    res = pool.run(CountingSource(6), ChecksumSink(), CFG, batch=1, engine_factory=fake_engine,
                   fail_on_gpus=(0, 1, 2))
    assert not res.ok and sorted(res.errors) == list(range(6))
    assert all("injected failure" in e for e in res.errors.values())


def test_frame_array_in_frame_array_out_through_shared_memory(pool):
    from framewright_b200.scheduler import ArraySink, ArraySource, SharedArray

    n, h, w = 17, 10, 12
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    sin, sout = SharedArray((n, h, w, 3)), SharedArray((n, 2 * h, 2 * w, 3))
    try:
        sin.array[...] = frames
        res = pool.run(ArraySource(sin), ArraySink(sout), CFG, batch=3, engine_factory=fake_engine)
        assert sorted(res.ok) == list(range(n))
        assert np.array_equal(sout.array, np.repeat(np.repeat(frames, 2, axis=1), 2, axis=2))
    finally:
        sin.release()
        sout.release()


def test_stream_is_ordered_and_bounded(pool):
    """Ring mode: an iterator of frames in, results emitted strictly in order, 8 slots of shared memory in flight."""
    n, h, w = 41, 8, 10
    frames = [np.full((h, w, 3), i, np.uint8) for i in range(n)]
    got = []
    res = pool.stream(iter(frames), CFG, lambda i, out: got.append((i, int(out[0, 0, 0]), out.shape)),
                      num_frames=n, frame_shape=(h, w), scale=2, batch=2, window=8, engine_factory=fake_engine)
    assert sorted(res.ok) == list(range(n))
    assert [g[0] for g in got] == list(range(n))                      # emitted in order
    assert all(v == i and shp == (2 * h, 2 * w, 3) for i, v, shp in got)


def test_stream_is_shared_by_all_gpus(pool):
    """Ring mode is self-scheduling: every GPU takes the next chunk from the front of the one ordered stream, so a
    slow GPU 0 (the stand-in engine sleeps 4 ms per frame there) leaves most of the stream to the other two."""
    n, h, w = 120, 8, 10
    frames = [np.full((h, w, 3), i % 251, np.uint8) for i in range(n)]
    got = []
    cfg = dict(CFG, tile_pad=4)
    res = pool.stream(iter(frames), cfg, lambda i, out: got.append((i, int(out[0, 0, 0]))), num_frames=n,
                      frame_shape=(h, w), scale=2, batch=2, window=12, engine_factory=fake_engine)
    assert sorted(res.ok) == list(range(n)) and not res.errors
    assert got == [(i, i % 251) for i in range(n)]                    # still strictly in order
    per_gpu = {g: len(v) for g, v in res.frames_per_gpu.items()}
    assert per_gpu[1] + per_gpu[2] > n // 2, per_gpu
    assert per_gpu[1] > 0 and per_gpu[2] > 0, per_gpu


def test_stream_that_ends_early_does_not_hang(pool):
    """A pipe that delivers fewer frames than announced (container metadata is often off by a few): the frames that
    arrived are emitted in order, the rest is reported as failed, `stream` returns."""
    n, have, h, w = 25, 10, 8, 10
    frames = [np.full((h, w, 3), i, np.uint8) for i in range(have)]
    got = []
    res = pool.stream(iter(frames), CFG, lambda i, out: got.append(i), num_frames=n, frame_shape=(h, w), scale=2,
                      batch=2, window=8, engine_factory=fake_engine)
    assert got == list(range(have)) and sorted(res.ok) == list(range(have))
    assert sorted(res.errors) == list(range(have, n)) and "ended after 10 of 25" in res.errors[n - 1]
    bad = [np.zeros((h, w, 3), np.uint8), np.zeros((h + 1, w, 3), np.uint8)]          # a frame of the wrong size
    res = pool.stream(iter(bad), CFG, lambda i, out: None, num_frames=2, frame_shape=(h, w), scale=2, batch=1,
                      window=4, engine_factory=fake_engine)
    assert sorted(res.ok) == [0] and "failed at frame 1" in res.errors[1]


def test_exception_in_a_callback_ends_the_job_and_leaves_the_pool_usable(pool):
    from framewright_b200.scheduler import ChecksumSink

    def boom(i, name, ok, err, gpu):
        raise BrokenPipeError("encoder went away")

    with pytest.raises(BrokenPipeError):
        pool.run(CountingSource(30), ChecksumSink(), CFG, batch=2, engine_factory=fake_engine, frame_callback=boom)
    res = pool.run(CountingSource(20), ChecksumSink(), CFG, batch=2, engine_factory=fake_engine)
    assert sorted(res.ok) == list(range(20)) and not res.errors


def test_files_in_files_out_and_unreadable_frame(pool, tmp_path):
    import cv2

    from framewright_b200.scheduler import PathSource, PngSink

    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir()
    paths = []
    for i in range(7):
        p = ind / f"frame_{i + 1:08d}.png"
        cv2.imwrite(str(p), np.full((6, 7, 3), 10 * i, np.uint8))
        paths.append(p)
    bad = ind / "frame_00000008.png"
    bad.write_bytes(b"nope")
    paths.append(bad)
    res = pool.run(PathSource(paths), PngSink(outd), CFG, batch=2, engine_factory=fake_engine)
    assert sorted(res.ok) == list(range(7)) and list(res.errors) == [7]
    assert "Failed to read image" in res.errors[7]
    for i in range(7):
        img = cv2.imread(str(outd / f"frame_{i + 1:08d}.png"))
        assert img.shape == (12, 14, 3) and int(img[0, 0, 0]) == 10 * i
    assert res.ok[3][1] == str(outd / "frame_00000004.png")


def test_a_dead_worker_does_not_hang_the_job():
    """A worker killed mid-job (native crash, OOM killer): its claimed frames are retried elsewhere, its shard is
    stolen, `run` returns; the next job runs on the survivors."""
    import framewright_b200  # noqa: F401
    from framewright_b200.scheduler import ChecksumSink, SchedulerPool

    p = SchedulerPool([0, 1], workers_per_gpu=1, start_timeout=120)
    try:
        n = 20
        done = {}

        def run():
            # (GPU 0 is made slow -- 25 ms per frame -- so that GPU 1 certainly claims work before it dies)
            done["res"] = p.run(CountingSource(n), ChecksumSink(), dict(CFG, tile_pad=25), batch=2,
                                engine_factory=fake_engine, crash_on_gpus=(1,))

        t = threading.Thread(target=run, daemon=True)
        t.start()
        t.join(timeout=60)
        assert not t.is_alive(), "run() hung on a dead worker"
        res = done["res"]
        assert sorted(res.ok) == list(range(n)) and not res.errors
        assert res.dead_gpus == [1] and res.retried and set(res.frames_per_gpu[0]) == set(range(n))
        assert p.alive_gpus() == [0]
        res = p.run(CountingSource(5), ChecksumSink(), CFG, batch=2, engine_factory=fake_engine)
        assert sorted(res.ok) == list(range(5))
    finally:
        p.close()


def test_distributor_product_path_over_the_pool(tmp_path):
    """`MultiGPUDistributor.distribute_frames(frames, process_fn=None, output_dir)`: persistent per-GPU workers, PNG
    in / PNG out, strategy-sized contiguous shards, per-frame progress, result in the reference's `DistributionResult`."""
    import cv2

    import framewright_b200  # noqa: F401
    from framewright_b200 import multi_gpu as mg

    ind = tmp_path / "frames"
    ind.mkdir()
    frames = []
    for i in range(13):
        p = ind / f"frame_{i + 1:08d}.png"
        cv2.imwrite(str(p), np.full((6, 8, 3), 5 * i, np.uint8))
        frames.append(p)
    gpus = [mg.GPUInfo(i, f"GPU{i}", 180000, f, 0.0) for i, f in enumerate((120000, 60000))]
    d = mg.MultiGPUDistributor(gpus=gpus, strategy=mg.LoadBalanceStrategy.VRAM_AWARE, workers_per_gpu=1,
                               model_name="RealESRGAN_x2plus", scale=2)
    assert d._shard_ranges(13, gpus) == [(0, 9), (9, 13)]      # VRAM_AWARE sizes (8 + 4, remainder round-robin), contiguous
    prog, per_frame = [], []
    try:
        res = d.distribute_frames(frames, None, tmp_path / "out", progress_callback=lambda f, m: prog.append(f),
                                  frame_callback=lambda i, name, ok, err, gpu: per_frame.append((i, ok)),
                                  engine_factory=fake_engine)
        assert isinstance(res, mg.DistributionResult) and res.total_frames == 13 and not res.errors
        assert len(prog) == 13 and prog == sorted(prog) and sorted(i for i, _ in per_frame) == list(range(13))
        assert sorted(p.name for v in res.frames_per_gpu.values() for p in v) == sorted(f.name for f in frames)
        assert cv2.imread(str(tmp_path / "out" / "frame_00000003.png")).shape == (12, 16, 3)
        assert res.success_rate == 100.0 and "Processed 13 frames across 2 GPUs" in res.summary()
        # the workers persist: a second job reuses them (no respawn)
        pids = {g: p.pid for g, p in d._pool._procs.items()}
        res2 = d.distribute_frames(frames[:4], None, tmp_path / "out2", engine_factory=fake_engine)
        assert res2.total_frames == 4 and {g: p.pid for g, p in d._pool._procs.items()} == pids
    finally:
        d.close()


def test_a_hung_worker_is_terminated_after_the_stall_timeout():
    """A worker that is alive but never answers (hung GPU): after `stall_timeout` it is terminated, its frames are
    retried on the other GPU, the job completes."""
    import framewright_b200  # noqa: F401
    from framewright_b200.scheduler import ChecksumSink, SchedulerPool

    from sched_helpers import hanging_engine

    p = SchedulerPool([0, 1], workers_per_gpu=1, start_timeout=120, stall_timeout=3.0)
    try:
        res = p.run(CountingSource(12), ChecksumSink(), dict(CFG, tile_pad=25), batch=2, engine_factory=hanging_engine)
        assert sorted(res.ok) == list(range(12)) and not res.errors
        assert res.dead_gpus == [1] and res.retried and p.alive_gpus() == [0]
    finally:
        p.close()
