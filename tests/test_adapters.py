"""Adapters for the callers either side of the hot path (restorer_adapter.py): the batch path behind
`VideoRestorer.enhance_frames` with the reference's own `CheckpointManager`, the `StreamingPipeline` enhancer, the
`MultiGPUProcessor.process_frames` shape and the raw-video pipes -- on CPU with a stand-in engine."""
import importlib
import io
import os
import sys
import types

import numpy as np
import pytest

from sched_helpers import fake_engine

REF = "/root/reference/src/framewright"


def _checkpoint_module():
    """The reference's checkpoint.py, unmodified, when the reference tree is present."""
    if not os.path.isdir(REF):
        return None
    saved = {k: sys.modules.get(k) for k in ("framewright", "framewright.checkpoint")}
    m = types.ModuleType("framewright")
    m.__path__ = [REF]
    sys.modules["framewright"] = m
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        return importlib.import_module("framewright.checkpoint")
    finally:
        sys.dont_write_bytecode = old
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


class _MiniCheckpoint:
    """API-compatible stand-in (update_frame / update_stage / load_checkpoint / get_unprocessed_frames)."""

    def __init__(self):
        self.stage, self.frames = None, {}

    def load_checkpoint(self):
        return types.SimpleNamespace(stage=self.stage) if self.stage else None

    def update_stage(self, stage):
        self.stage = stage

    def update_frame(self, frame_number, input_path, output_path=None, checksum=None):
        self.frames[frame_number] = (str(input_path), str(output_path))

    def get_unprocessed_frames(self, all_frames):
        return [f for f in all_frames if int(f.stem.split("_")[-1]) not in self.frames]

    def force_save(self):
        pass


def _write_frames(d, n, h=6, w=8):
    import cv2

    d.mkdir(parents=True, exist_ok=True)
    paths = []
    for i in range(n):
        p = d / f"frame_{i + 1:08d}.png"
        cv2.imwrite(str(p), np.full((h, w, 3), 7 * i % 250, np.uint8))
        paths.append(p)
    return paths


@pytest.fixture()
def distributor():
    import framewright_b200  # noqa: F401
    from framewright_b200 import multi_gpu as mg

    d = mg.MultiGPUDistributor(gpus=[mg.GPUInfo(0, "GPU0", 1, 1, 0.0), mg.GPUInfo(1, "GPU1", 1, 1, 0.0)],
                               strategy=mg.LoadBalanceStrategy.ROUND_ROBIN, workers_per_gpu=1,
                               model_name="RealESRGAN_x2plus", scale=2)
    yield d
    d.close()


@pytest.mark.parametrize("real", [False, True])
def test_enhance_frames_batched_feeds_checkpoint_and_progress_and_resumes(tmp_path, distributor, real):
    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import enhance_frames_batched

    frames = _write_frames(tmp_path / "frames", 11)
    enhanced = tmp_path / "enhanced"
    if real:
        ck = _checkpoint_module()
        if ck is None:
            pytest.skip("reference tree not present")
        cm = ck.CheckpointManager(tmp_path, checkpoint_interval=3)
        cm.create_checkpoint(stage="extract", total_frames=11, source_path="clip.mp4")
    else:
        cm = _MiniCheckpoint()
    cfg = PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2)
    prog = []
    n = enhance_frames_batched(frames[:6], enhanced, cfg, checkpoint_manager=cm, distributor=distributor,
                               update_progress=lambda **k: prog.append(k), engine_factory=fake_engine)
    assert n == 6 and len(list(enhanced.glob("*.png"))) == 6
    per_frame = [p for p in prog if 0.0 < p["progress"] < 1.0 or p.get("frames_completed") not in (0, None)]
    assert [p["frames_completed"] for p in prog if "eta_seconds" not in p][1:] == [1, 2, 3, 4, 5, 6]
    assert prog[0]["progress"] == 0.0 and prog[-1]["progress"] == 1.0 and prog[-1]["eta_seconds"] == 0.0 and per_frame
    if real:
        loaded = cm.load_checkpoint()
        assert loaded.stage == "enhance" and sorted(f.frame_number for f in loaded.frames if f.processed) == [1, 2, 3, 4, 5, 6]
        assert loaded.frames[0].output_path.endswith(".png")
    # resume: only the unprocessed frames run again
    seen = []
    n = enhance_frames_batched(frames, enhanced, cfg, checkpoint_manager=cm, distributor=distributor,
                               update_progress=lambda **k: seen.append(k), engine_factory=fake_engine)
    assert n == 11 and len(list(enhanced.glob("*.png"))) == 11
    assert max(k["frames_total"] for k in seen) == 5                  # 6 of 11 were skipped
    assert enhance_frames_batched(frames, enhanced, cfg, checkpoint_manager=cm, distributor=distributor) == 11


def test_enhance_frames_batched_failure_policy(tmp_path, distributor):
    """A frame every GPU fails: continue_on_error copies the original (the video still assembles) and records the
    error; without it the call raises `EnhancementError` like `_enhance_frames_sequential`."""
    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import EnhancementError, enhance_frames_batched

    frames = _write_frames(tmp_path / "frames", 4)
    frames[2].write_bytes(b"garbage")
    cfg = PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2)
    report = types.SimpleNamespace(ok=0, errs=[])
    report.add_success = lambda: setattr(report, "ok", report.ok + 1)
    report.add_error = lambda name, exc: report.errs.append((name, str(exc)))
    n = enhance_frames_batched(frames, tmp_path / "out", cfg, distributor=distributor, error_report=report,
                               continue_on_error=True, engine_factory=fake_engine)
    assert n == 4 and report.ok == 3 and report.errs[0][0] == "frame_00000003.png" and "Failed to read" in report.errs[0][1]
    assert (tmp_path / "out" / "frame_00000003.png").read_bytes() == b"garbage"
    with pytest.raises(EnhancementError, match="Failed to enhance frame frame_00000003.png"):
        enhance_frames_batched(frames, tmp_path / "out2", cfg, distributor=distributor, continue_on_error=False,
                               engine_factory=fake_engine)


def test_streaming_enhancer_batches_pipeline_frames(tmp_path):
    from dataclasses import dataclass, field
    from pathlib import Path
    from typing import Any, Dict, Optional

    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import make_streaming_enhancer

    @dataclass
    class PipelineFrame:            # processors/streaming.py:805-812
        index: int
        path: Path
        data: Optional[Any] = None
        metadata: Dict[str, Any] = field(default_factory=dict)
        processed: bool = False
        error: Optional[str] = None

    paths = _write_frames(tmp_path / "f", 5)
    frames = [PipelineFrame(i, p) for i, p in enumerate(paths)]
    frames.append(PipelineFrame(5, tmp_path / "missing.png"))
    frames.append(PipelineFrame(6, paths[0], data=np.full((3, 3, 3), 9, np.uint8)))      # in-memory frame, other size
    up = fake_engine({"gpu_id": 1})
    fn = make_streaming_enhancer(PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2), upsampler=up)
    out = fn(frames)
    assert out is frames and [f.processed for f in out] == [True] * 5 + [False, True]
    assert out[0].data.shape == (12, 16, 3) and out[6].data.shape == (6, 6, 3) and "Failed to read" in out[5].error


def test_multi_gpu_process_frames_shape(tmp_path):
    import framewright_b200  # noqa: F401
    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import multi_gpu_process_frames
    from framewright_b200.scheduler import SchedulerPool

    frames = [np.full((5, 7, 3), i, np.uint8) for i in range(9)]
    seen = []
    with SchedulerPool([0, 1], workers_per_gpu=1, start_timeout=120) as pool:
        res = multi_gpu_process_frames(frames, PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2),
                                       pool=pool, callback=seen.append, engine_factory=fake_engine)
    assert [r.frame_index for r in res] == list(range(9)) and all(r.success for r in res) and len(seen) == 9
    assert all(r.output.shape == (10, 14, 3) and int(r.output[0, 0, 0]) == r.frame_index for r in res)
    assert {r.device_id for r in res} <= {0, 1}


def test_raw_video_pipes_round_trip():
    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import RawVideoReader, upscale_raw_stream

    h, w, n = 6, 10, 7
    rng = np.random.default_rng(0)
    clip = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    src = io.BytesIO(clip.tobytes() + b"\x00" * 5)                    # a truncated trailing frame is dropped
    dst = io.BytesIO()
    cfg = PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2)
    written = upscale_raw_stream(src, dst, w, h, cfg, num_frames=n, batch=3, upsampler=fake_engine({"gpu_id": 1}))
    assert written == n
    out = np.frombuffer(dst.getvalue(), np.uint8).reshape(n, 2 * h, 2 * w, 3)
    assert np.array_equal(out, np.repeat(np.repeat(clip, 2, axis=1), 2, axis=2))
    assert len(list(RawVideoReader(io.BytesIO(clip.tobytes()), w, h))) == n


_FAKE_FFMPEG = '''#!/usr/bin/env python3
"""Stand-in for ffmpeg in the pipe test: `-i <file> ... pipe:1` writes the file's bytes to stdout (decoder);
`-i pipe:0 ... -y <out>` copies stdin to <out> and records its argv next to it (encoder)."""
import json, os, sys
a = sys.argv[1:]
src = a[a.index("-i") + 1]
if src == "pipe:0":
    out = a[-1]
    if os.environ.get("FAKE_FFMPEG_ENCODER_DIES"):
        sys.stdin.buffer.read(1000)
        sys.stderr.write("Unknown encoder 'libx265'\\n")
        sys.exit(1)
    with open(out, "wb") as f:
        while True:
            b = sys.stdin.buffer.read(1 << 16)
            if not b:
                break
            f.write(b)
    json.dump(a, open(out + ".argv.json", "w"))
else:
    if not os.path.exists(src):
        sys.stderr.write(src + ": No such file or directory\\n")
        sys.exit(1)
    sys.stdout.buffer.write(open(src, "rb").read())
'''


def test_video_file_through_ffmpeg_pipes(tmp_path, monkeypatch):
    """f1 end to end: decoder process -> raw frames -> engine -> raw frames -> encoder process, no frame files; the
    encoder gets the reference's codec arguments (restorer.py:3001-3027) and the SCALED frame size; a failing ffmpeg on
    either side is an `EnhancementError` carrying its stderr, not a hang."""
    import json
    import stat

    from framewright_b200.pytorch_realesrgan import PyTorchESRGANConfig
    from framewright_b200.restorer_adapter import (EnhancementError, ffmpeg_decode_command, ffmpeg_encode_command,
                                                   upscale_video_ffmpeg)

    exe = tmp_path / "ffmpeg"
    exe.write_text(_FAKE_FFMPEG)
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    h, w, n = 6, 10, 9
    clip = np.random.default_rng(1).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    src, dst, audio = tmp_path / "in.raw", tmp_path / "out.bin", tmp_path / "audio.flac"
    src.write_bytes(clip.tobytes())
    audio.write_bytes(b"x")
    cfg = PyTorchESRGANConfig(model_name="RealESRGAN_x2plus", scale_factor=2)
    prog = []
    written = upscale_video_ffmpeg(src, dst, cfg, width=w, height=h, framerate=24, num_frames=n, audio_path=audio,
                                   crf=16, preset="slow", ffmpeg=str(exe), batch=4,
                                   upsampler=fake_engine({"gpu_id": 1}), progress_callback=lambda f, m: prog.append(f))
    assert written == n and prog[-1] == 1.0
    out = np.frombuffer(dst.read_bytes(), np.uint8).reshape(n, 2 * h, 2 * w, 3)
    assert np.array_equal(out, np.repeat(np.repeat(clip, 2, axis=1), 2, axis=2))
    argv = json.load(open(str(dst) + ".argv.json"))
    assert argv[argv.index("-s") + 1] == f"{2 * w}x{2 * h}" and argv[argv.index("-framerate") + 1] == "24"
    assert argv[argv.index("-c:v") + 1] == "libx265" and argv[argv.index("-crf") + 1] == "16"
    assert argv[argv.index("-preset") + 1] == "slow" and argv[argv.index("-pix_fmt", argv.index("-c:v")) + 1] == "yuv420p10le"
    assert argv[argv.index("-c:a") + 1] == "flac" and str(audio) in argv and argv[-2:] == ["-y", str(dst)]
    assert ffmpeg_decode_command("a.mp4")[-5:] == ["-f", "rawvideo", "-pix_fmt", "bgr24", "pipe:1"]
    assert "-c:a" not in ffmpeg_encode_command("o.mp4", 8, 8, 30)                       # no audio track given
    with pytest.raises(EnhancementError, match="decoder failed.*No such file"):
        upscale_video_ffmpeg(tmp_path / "missing.raw", dst, cfg, width=w, height=h, ffmpeg=str(exe),
                             upsampler=fake_engine({"gpu_id": 1}))
    monkeypatch.setenv("FAKE_FFMPEG_ENCODER_DIES", "1")
    big = np.zeros((40, 64, 64, 3), np.uint8)                      # enough bytes to fill the pipe of a dead encoder
    (tmp_path / "big.raw").write_bytes(big.tobytes())
    with pytest.raises(EnhancementError, match="encoder failed.*Unknown encoder"):
        upscale_video_ffmpeg(tmp_path / "big.raw", dst, cfg, width=64, height=64, ffmpeg=str(exe),
                             upsampler=fake_engine({"gpu_id": 1}))
    with pytest.raises(EnhancementError, match="cannot start the decoder"):
        upscale_video_ffmpeg(src, dst, cfg, width=w, height=h, ffmpeg=str(tmp_path / "no-such-ffmpeg"),
                             upsampler=fake_engine({"gpu_id": 1}))
