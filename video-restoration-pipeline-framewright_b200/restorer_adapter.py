"""Adapters for the callers either side of the hot path.

f2  `enhance_frames_batched`: what `VideoRestorer.enhance_frames` / `_enhance_frames_sequential` / `_parallel`
    (`/root/reference/src/framewright/restorer.py:1604-1973`) do around one `enhance_frame_pytorch` call per frame,
    but handing the WHOLE frame list to the sharded scheduler: resume from the checkpoint, `update_stage("enhance")`,
    a `CheckpointManager.update_frame(frame_number, input_path, output_path)` and an `_update_progress(...)` per
    frame AS IT COMPLETES, the out-of-memory tile ladder ("memory" / "vram" in the error -> next smaller tile of
    `get_adaptive_tile_sequence`), `continue_on_error` (copy the original frame so the video still assembles) and the
    `ErrorReport` bookkeeping.
f4  `make_streaming_enhancer` -> the `enhance_fn(List[PipelineFrame]) -> List[PipelineFrame]` that
    `StreamingPipeline.set_enhancer` takes (`processors/streaming.py:876-885`); `make_process_func` -> the
    `process_func(frame, device_id)` the reference's own `MultiGPUProcessor.process_frames` calls
    (`infrastructure/gpu/distributor.py:687-752`); `multi_gpu_process_frames` -> the same call shape over the
    scheduler's worker processes and shared-memory transport.
f1  `RawVideoReader` / `RawVideoWriter`: raw bgr24 frame pipes (what `ffmpeg -f rawvideo -pix_fmt bgr24 -` produces /
    consumes) and `upscale_raw_stream`, the decode -> engine -> encode path with no PNG round trip and no per-frame
    `ffprobe` (`restorer.py:1109-1118, 3001-3027`, `validators.py:165-181`); `upscale_video_ffmpeg` runs it between
    an ffmpeg decoder and an ffmpeg encoder process (the reference's own codec arguments).
"""
from __future__ import annotations

import logging
import shutil
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any, BinaryIO, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .pytorch_realesrgan import PyTorchESRGANConfig
from .tile_sizing import get_adaptive_tile_sequence

logger = logging.getLogger(__name__)


class EnhancementError(RuntimeError):
    """Mirror of `framewright.errors.EnhancementError` for callers that do not import the reference package."""


def _is_vram_error(msg: Optional[str]) -> bool:
    return bool(msg) and ("vram" in msg.lower() or "memory" in msg.lower())   # restorer.py:1746


def _frame_number(path: Path) -> int:
    return int(Path(path).stem.split("_")[-1])                                  # restorer.py:1774


def enhance_frames_batched(frames: Sequence[Path], enhanced_dir: Path, config: PyTorchESRGANConfig, *,
                           checkpoint_manager: Any = None, update_progress: Optional[Callable[..., None]] = None,
                           error_report: Any = None, continue_on_error: bool = True,
                           frame_resolution: Optional[Tuple[int, int]] = None, tile_sequence: Optional[List[int]] = None,
                           distributor: Any = None, gpu_ids: Optional[List[int]] = None,
                           error_cls: type = EnhancementError, **hooks) -> int:
    """Enhance `frames` (frame_XXXXXXXX.png paths) into `enhanced_dir/<same name>` on all healthy GPUs.
    Returns the number of frames enhanced (frames replaced by their original under `continue_on_error` count, as in
    the reference).  `checkpoint_manager`, `update_progress(stage=, progress=, frames_completed=, frames_total=)` and
    `error_report` (`add_success()` / `add_error(name, exc)`) are the reference's own objects, passed in by the caller."""
    from .multi_gpu import GPUInfo, LoadBalanceStrategy, MultiGPUDistributor, query_gpus

    frames = [Path(f) for f in frames]
    if not frames:
        raise error_cls("No frames found to enhance")
    total_all = len(frames)
    if checkpoint_manager is not None:
        checkpoint = checkpoint_manager.load_checkpoint()
        if checkpoint and checkpoint.stage == "enhance":
            frames = checkpoint_manager.get_unprocessed_frames(frames)
            logger.info(f"Resuming enhancement: {len(frames)} frames remaining")
    if not frames:
        logger.info("All frames already enhanced")
        return total_all
    enhanced_dir = Path(enhanced_dir)
    enhanced_dir.mkdir(parents=True, exist_ok=True)
    if checkpoint_manager is not None:
        checkpoint_manager.update_stage("enhance")
    if tile_sequence is None:
        tile_sequence = get_adaptive_tile_sequence(frame_resolution or (1920, 1080), config.scale_factor,
                                                   starting_tile_size=config.tile_size or None)
    total = len(frames)
    if update_progress:
        update_progress(stage="enhance_frames", progress=0.0, frames_completed=0, frames_total=total)
    own = distributor is None
    if own:
        gpus = None
        if gpu_ids is not None:
            gpus = [g for g in query_gpus() if g.id in gpu_ids] or [GPUInfo(i, f"GPU{i}", 1, 1, 0.0) for i in gpu_ids]
        distributor = MultiGPUDistributor(gpus=gpus, strategy=LoadBalanceStrategy.ROUND_ROBIN,
                                          model_name=config.model_name, scale=config.scale_factor,
                                          tile=config.tile_size, tile_pad=config.tile_pad, pre_pad=config.pre_pad)
    state = {"completed": 0, "enhanced": 0}
    try:
        todo = list(frames)
        tile_index = 0
        while todo:
            vram_failed: List[Path] = []
            last_err: Dict[str, str] = {}

            def on_frame(i: int, name: str, ok: bool, err: Optional[str], gpu: int, _todo=todo) -> None:
                src = _todo[i]
                if not ok and _is_vram_error(err) and tile_index + 1 < len(tile_sequence):
                    vram_failed.append(src)                 # retried below with the next smaller tile
                    last_err[src.name] = err or ""
                    return
                out = enhanced_dir / src.name
                if ok:
                    state["enhanced"] += 1
                    if error_report is not None:
                        error_report.add_success()
                    if checkpoint_manager is not None:
                        checkpoint_manager.update_frame(frame_number=_frame_number(src), input_path=src, output_path=out)
                else:
                    if error_report is not None:
                        error_report.add_error(src.name, error_cls(err or "Unknown error"))
                    if not continue_on_error:
                        state["fatal"] = f"Failed to enhance frame {src.name}: {err}"
                    else:
                        try:   # keep the video assemblable: the original frame stands in (restorer.py:1790-1800)
                            shutil.copy2(src, out)
                            logger.warning(f"Frame {src.name} enhancement failed, using original. Error: {err}")
                            state["enhanced"] += 1
                        except Exception as copy_err:
                            logger.error(f"Could not copy original frame: {copy_err}")
                state["completed"] += 1
                if update_progress:
                    update_progress(stage="enhance_frames", progress=state["completed"] / total,
                                    frames_completed=state["completed"], frames_total=total)

            distributor.distribute_frames(todo, None, enhanced_dir, frame_callback=on_frame, **hooks)
            if state.get("fatal"):
                raise error_cls(state["fatal"])
            if not vram_failed:
                break
            tile_index += 1
            new_tile = tile_sequence[tile_index]
            logger.info(f"VRAM error, reducing tile size to {new_tile}")
            config.tile_size = new_tile                      # the reference mutates the running tile size the same way
            distributor._engine_kwargs.update(tile=new_tile)
            todo = sorted(vram_failed)
    finally:
        if own:
            distributor.close()
        if checkpoint_manager is not None and hasattr(checkpoint_manager, "force_save"):
            try:
                checkpoint_manager.force_save()
            except Exception:  # pragma: no cover
                pass
    if update_progress:
        update_progress(stage="enhance_frames", progress=1.0, frames_completed=state["completed"],
                        frames_total=total, eta_seconds=0.0)
    return state["enhanced"] + (total_all - total)


# ------------------------------------------------------------------------------------------------ f1: no ffprobe fork
@dataclass
class FrameValidation:
    """Mirror of `framewright.validators.FrameValidation` (:47-55)."""
    frame_path: Path
    is_valid: bool
    width: int = 0
    height: int = 0
    file_size: int = 0
    error_message: Optional[str] = None
    quality: Any = None


def validate_frame_integrity(frame_path: Path) -> FrameValidation:
    """`validators.validate_frame_integrity` (:143-200) without forking `ffprobe` for every frame: the same checks
    (exists, non-empty, a decodable image header with non-zero dimensions) by reading the PNG IHDR / JPEG SOF marker.
    At 27 frames/s per GPU a process fork per output frame is the pipeline bound the survey measured (f1)."""
    frame_path = Path(frame_path)
    result = FrameValidation(frame_path=frame_path, is_valid=False)
    if not frame_path.exists():
        result.error_message = "File does not exist"
        return result
    result.file_size = frame_path.stat().st_size
    if result.file_size == 0:
        result.error_message = "File is empty"
        return result
    try:
        with open(frame_path, "rb") as f:
            head = f.read(32)
            if head[:8] == b"\x89PNG\r\n\x1a\n" and head[12:16] == b"IHDR":
                result.width = int.from_bytes(head[16:20], "big")
                result.height = int.from_bytes(head[20:24], "big")
            elif head[:2] == b"\xff\xd8":                      # JPEG: walk the segments to the first SOFn marker
                f.seek(2)
                while True:
                    b = f.read(4)
                    if len(b) < 4 or b[0] != 0xFF:
                        break
                    marker, seglen = b[1], int.from_bytes(b[2:4], "big")
                    if 0xC0 <= marker <= 0xCF and marker not in (0xC4, 0xC8, 0xCC):
                        d = f.read(5)
                        result.height, result.width = int.from_bytes(d[1:3], "big"), int.from_bytes(d[3:5], "big")
                        break
                    f.seek(seglen - 2, 1)
            else:
                result.error_message = "No video stream found"
                return result
    except OSError as e:
        result.error_message = f"read error: {e}"
        return result
    if result.width == 0 or result.height == 0:
        result.error_message = "Invalid dimensions"
        return result
    result.is_valid = True
    return result


# ------------------------------------------------------------------------------------------------ f4
def make_streaming_enhancer(config: PyTorchESRGANConfig, load: Optional[Callable[[Path], np.ndarray]] = None,
                            upsampler: Any = None, output_dir: Optional[Path] = None) -> Callable[[List[Any]], List[Any]]:
    """`enhance_fn` for `StreamingPipeline.set_enhancer`: takes the pipeline's chunk of `PipelineFrame`s (`.index`,
    `.path`, `.data`, `.processed`, `.error`), runs every run of same-size frames as one engine batch, and returns the
    same objects with `.data` replaced by the upscaled frame and `.processed` set.
    `output_dir`: also write every upscaled frame to `output_dir/frame_<index:08d>.png` -- the paths the reference's
    write stage returns (`streaming.py:1088-1090`) but does not itself fill ("in real implementation, this would save
    enhanced data")."""
    from .pytorch_realesrgan import get_upsampler

    if output_dir is not None:
        output_dir = Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)

    def _load(p: Path) -> np.ndarray:
        import cv2

        img = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
        if img is None:
            raise IOError(f"Failed to read image: {p}")
        return img

    loader = load or _load

    def enhance_fn(frames: List[Any]) -> List[Any]:
        up = upsampler or get_upsampler(config)
        imgs: List[Optional[np.ndarray]] = []
        for f in frames:
            try:
                imgs.append(f.data if isinstance(getattr(f, "data", None), np.ndarray) else loader(f.path))
            except Exception as e:
                imgs.append(None)
                f.error = str(e)
        i = 0
        while i < len(frames):
            if imgs[i] is None:
                i += 1
                continue
            j = i + 1
            plain = imgs[i].ndim == 3 and imgs[i].shape[2] == 3 and imgs[i].dtype == np.uint8
            while plain and j < len(frames) and imgs[j] is not None and imgs[j].shape == imgs[i].shape \
                    and imgs[j].dtype == np.uint8 and j - i < 8:
                j += 1
            try:
                if j - i > 1:
                    outs = up.enhance_batch(np.stack(imgs[i:j]))
                else:
                    outs = [up.enhance(imgs[i], outscale=config.scale_factor)[0]]
                for k in range(i, j):
                    frames[k].data = outs[k - i]
                    frames[k].processed = True
                    frames[k].error = None
                    if output_dir is not None:
                        import cv2

                        dst = output_dir / f"frame_{frames[k].index:08d}.png"
                        if not cv2.imwrite(str(dst), outs[k - i]) or not dst.exists():
                            frames[k].processed = False
                            frames[k].error = "Output file was not created"
            except Exception as e:
                for k in range(i, j):
                    frames[k].error = str(e)
            i = j
        return frames

    return enhance_fn


@dataclass
class ProcessingResult:
    """Mirror of `infrastructure.gpu.distributor.ProcessingResult`."""
    frame_index: int
    device_id: int
    success: bool
    output: Optional[np.ndarray] = None
    error: Optional[str] = None
    elapsed_seconds: float = 0.0


def make_process_func(config: PyTorchESRGANConfig) -> Callable[[np.ndarray, int], np.ndarray]:
    """The `process_func(frame, device_id) -> processed_frame` that the reference's own
    `MultiGPUProcessor.process_frames` / `process_batch` (`infrastructure/gpu/distributor.py:687-752, 788-`) call from
    their per-GPU executor threads: the upscaling operator on GPU `device_id` (one cached upsampler per GPU, thread-safe;
    a [N,H,W,3] stack -- `process_batch` -- runs as one engine batch).  Exceptions propagate: the reference turns them
    into `ProcessingResult(success=False, error=str(e))`."""
    import dataclasses

    from .pytorch_realesrgan import get_upsampler

    config.validate()

    def process_func(frame: np.ndarray, device_id: int) -> np.ndarray:
        up = get_upsampler(dataclasses.replace(config, gpu_id=int(device_id)))
        if isinstance(frame, np.ndarray) and frame.ndim == 4:
            return up.enhance_batch(frame)
        return up.enhance(frame, outscale=config.scale_factor)[0]

    return process_func


def multi_gpu_process_frames(frames: List[np.ndarray], config: PyTorchESRGANConfig, pool: Any = None,
                             callback: Optional[Callable[[ProcessingResult], None]] = None,
                             gpu_ids: Optional[List[int]] = None, **hooks) -> List[ProcessingResult]:
    """`MultiGPUProcessor.process_frames(frames, process_func, callback)` for the upscaling operator: a list of frame
    arrays in, per-frame `ProcessingResult`s out (sorted by frame index), frames sharded over the GPUs through shared
    memory (no files).  The reference's `process_func(frame, device_id)` is the engine itself here."""
    from .multi_gpu import query_gpus
    from .scheduler import ArraySink, ArraySource, SchedulerPool, SharedArray

    if not frames:
        return []
    shape = frames[0].shape
    if any(f.shape != shape or f.dtype != np.uint8 for f in frames) or len(shape) != 3 or shape[2] != 3:
        raise ValueError("multi_gpu_process_frames needs same-size uint8 BGR frames")
    s = int(config.scale_factor)
    own = pool is None
    if own:
        ids = gpu_ids if gpu_ids is not None else [g.id for g in query_gpus()] or [0]
        pool = SchedulerPool(ids, workers_per_gpu=3)
    sin = SharedArray((len(frames),) + tuple(shape))
    sout = SharedArray((len(frames), shape[0] * s, shape[1] * s, 3))
    results: Dict[int, ProcessingResult] = {}
    t0 = time.time()
    try:
        a = sin.array
        for i, f in enumerate(frames):
            a[i] = f
        cfg = {"model_name": config.model_name, "scale_factor": s, "tile_size": config.tile_size,
               "tile_pad": config.tile_pad, "pre_pad": config.pre_pad}

        def on_frame(i: int, name: str, ok: bool, err: Optional[str], gpu: int) -> None:
            r = ProcessingResult(i, gpu, ok, sout.array[i].copy() if ok else None, err, time.time() - t0)
            results[i] = r
            if callback:
                callback(r)

        pool.run(ArraySource(sin), ArraySink(sout), cfg, batch=2, frame_callback=on_frame, **hooks)
    finally:
        sin.release()
        sout.release()
        if own:
            pool.close()
    return [results[i] for i in sorted(results)]


# ------------------------------------------------------------------------------------------------ f1
class RawVideoReader:
    """Frames from a raw bgr24 byte stream (`ffmpeg -i in.mp4 -f rawvideo -pix_fmt bgr24 -`): width x height x 3 bytes
    per frame, no container, no per-frame files."""

    def __init__(self, stream: BinaryIO, width: int, height: int):
        self.stream, self.width, self.height = stream, int(width), int(height)
        self.frame_bytes = self.width * self.height * 3

    def __iter__(self) -> Iterator[np.ndarray]:
        while True:
            buf = bytearray(self.frame_bytes)
            view, got = memoryview(buf), 0
            while got < self.frame_bytes:
                n = self.stream.readinto(view[got:])
                if not n:
                    break
                got += n
            if got < self.frame_bytes:
                if got:
                    logger.warning("raw stream ended inside a frame (%d of %d bytes)", got, self.frame_bytes)
                return
            yield np.frombuffer(buf, np.uint8).reshape(self.height, self.width, 3)


class RawVideoWriter:
    """Frames to a raw bgr24 byte stream (`ffmpeg -f rawvideo -pix_fmt bgr24 -s WxH -r fps -i - out.mp4`)."""

    def __init__(self, stream: BinaryIO):
        self.stream = stream
        self.frames = 0

    def write(self, frame: np.ndarray) -> None:
        self.stream.write(memoryview(np.ascontiguousarray(frame)).cast("B"))
        self.frames += 1


def upscale_raw_stream(src: BinaryIO, dst: BinaryIO, width: int, height: int, config: PyTorchESRGANConfig,
                       num_frames: Optional[int] = None, pool: Any = None, batch: int = 4, upsampler: Any = None,
                       progress_callback: Optional[Callable[[float, str], None]] = None) -> int:
    """decode pipe -> engine -> encode pipe, frames in order, bounded memory.  With a `SchedulerPool` (and a known
    `num_frames`) the frames are spread over its GPUs through the shared-memory ring; otherwise one engine in this
    process batches them.  Returns the number of frames written."""
    from .pytorch_realesrgan import get_upsampler

    reader, writer = RawVideoReader(src, width, height), RawVideoWriter(dst)
    s = int(config.scale_factor)
    if pool is not None and num_frames:
        cfg = {"model_name": config.model_name, "scale_factor": s, "tile_size": config.tile_size,
               "tile_pad": config.tile_pad, "pre_pad": config.pre_pad}
        pool.stream(iter(reader), cfg, lambda i, out: writer.write(out), num_frames=num_frames,
                    frame_shape=(height, width), scale=s, batch=2, progress_callback=progress_callback)
        return writer.frames
    up = upsampler or get_upsampler(config)
    chunk: List[np.ndarray] = []

    def flush() -> None:
        if not chunk:
            return
        outs = up.enhance_batch(np.stack(chunk))
        for o in outs:
            writer.write(o)
        chunk.clear()
        if progress_callback and num_frames:
            progress_callback(min(1.0, writer.frames / num_frames), f"Processed {writer.frames}/{num_frames} frames")

    for frame in reader:
        chunk.append(frame)
        if len(chunk) >= batch:
            flush()
    flush()
    return writer.frames


def ffmpeg_decode_command(video_path: Any, ffmpeg: str = "ffmpeg") -> List[str]:
    """The decoder side of `VideoRestorer.extract_frames` (`restorer.py:1109-1118`) with the PNG sequence replaced by
    raw bgr24 frames on stdout."""
    return [ffmpeg, "-v", "error", "-i", str(video_path), "-f", "rawvideo", "-pix_fmt", "bgr24", "pipe:1"]


def ffmpeg_encode_command(output_path: Any, width: int, height: int, framerate: Any, audio_path: Any = None,
                          codec: str = "libx265", pix_fmt: str = "yuv420p10le", crf: int = 18, preset: str = "medium",
                          ffmpeg: str = "ffmpeg") -> List[str]:
    """The encoder side of `VideoRestorer.reassemble_video` (`restorer.py:3001-3027`: `-framerate`, optional audio as
    FLAC, `-c:v <codec> -crf <crf> -preset <preset> -pix_fmt <fmt> -y out`) with the PNG sequence replaced by raw
    bgr24 frames of `width` x `height` on stdin."""
    cmd = [ffmpeg, "-v", "error", "-f", "rawvideo", "-pix_fmt", "bgr24", "-s", f"{int(width)}x{int(height)}",
           "-framerate", str(framerate), "-i", "pipe:0"]
    if audio_path is not None and Path(audio_path).exists():
        cmd += ["-i", str(audio_path), "-c:a", "flac"]
    cmd += ["-c:v", codec, "-crf", str(crf), "-preset", preset, "-pix_fmt", pix_fmt, "-y", str(output_path)]
    return cmd


def upscale_video_ffmpeg(video_path: Any, output_path: Any, config: PyTorchESRGANConfig, *, width: int, height: int,
                         framerate: Any = 30, num_frames: Optional[int] = None, audio_path: Any = None,
                         codec: str = "libx265", pix_fmt: str = "yuv420p10le", crf: int = 18, preset: str = "medium",
                         ffmpeg: str = "ffmpeg", pool: Any = None, batch: int = 4, upsampler: Any = None,
                         progress_callback: Optional[Callable[[float, str], None]] = None) -> int:
    """video file -> ffmpeg decoder -> raw frames -> engine(s) -> raw frames -> ffmpeg encoder -> video file: the
    reference's extract_frames / enhance_frames / reassemble_video sequence (`restorer.py:1076-1160, 1604-1705,
    2950-3060`) as ONE pass with no frame files in between.  `width`, `height`, `framerate` (and `num_frames`, which lets
    a `SchedulerPool` spread the frames over its GPUs) are what the reference reads into `self.metadata` with ffprobe
    before it starts.  Returns the number of frames written; raises `EnhancementError` when either ffmpeg fails."""
    import subprocess
    import tempfile

    s = int(config.scale_factor)
    dec_cmd = ffmpeg_decode_command(video_path, ffmpeg)
    enc_cmd = ffmpeg_encode_command(output_path, width * s, height * s, framerate, audio_path, codec, pix_fmt, crf,
                                    preset, ffmpeg)
    # stderr goes to files: nobody reads those pipes while the frames flow, and a full pipe would stall ffmpeg
    with tempfile.TemporaryFile() as dec_err, tempfile.TemporaryFile() as enc_err:
        try:
            dec = subprocess.Popen(dec_cmd, stdout=subprocess.PIPE, stderr=dec_err, stdin=subprocess.DEVNULL)
        except OSError as e:
            raise EnhancementError(f"cannot start the decoder ({dec_cmd[0]}): {e}") from e
        try:
            enc = subprocess.Popen(enc_cmd, stdin=subprocess.PIPE, stderr=enc_err, stdout=subprocess.DEVNULL)
        except OSError as e:
            dec.kill()
            dec.wait()
            raise EnhancementError(f"cannot start the encoder ({enc_cmd[0]}): {e}") from e
        written, failure = 0, None
        try:
            written = upscale_raw_stream(dec.stdout, enc.stdin, width, height, config, num_frames=num_frames, pool=pool,
                                         batch=batch, upsampler=upsampler, progress_callback=progress_callback)
        except BrokenPipeError as e:        # the encoder went away: its exit code and stderr say why
            failure = e
        except BaseException:
            for p in (dec, enc):
                p.kill()
            raise
        finally:
            for stream in (enc.stdin, dec.stdout):
                try:
                    stream.close()
                except Exception:
                    pass
            dec_rc, enc_rc = dec.wait(), enc.wait()

        def tail(f) -> str:
            f.seek(0)
            return f.read()[-800:].decode("utf-8", "replace").strip()

        if enc_rc != 0 or failure is not None:
            raise EnhancementError(f"encoder failed (exit code {enc_rc}) after {written} frames: {tail(enc_err)}")
        if dec_rc != 0:
            raise EnhancementError(f"decoder failed (exit code {dec_rc}) after {written} frames: {tail(dec_err)}")
    return written
