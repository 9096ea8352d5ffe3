"""Operator API of the upscaling path: the `SRBackend` interface, its Real-ESRGAN implementation and the
`SuperResolution` facade of `/root/reference/src/framewright/processors/enhancement/super_resolution.py`
(`SRConfig` :164-199, `SRResult` :206-230, `SRBackend` :237-311, `RealESRGANBackend` :441-601, `SuperResolution`
:1194-1527, factories :1534-1613) on the B200 engine.

Same names, arguments, result records and error behaviour -- a frames directory in, a frames directory out, per-frame
failures collected in `SRResult.warnings` as "Frame <name>: <err>", nothing raised -- but the frames-dir path is built
for this engine instead of one `enhance_frame_pytorch` call per file: frames are decoded ahead on a thread pool,
consecutive same-size frames run as one batch through the engine's pipelined host path (`enhance_batch`: H2D / D2H
of one lane job under the kernels of the other), and PNG encoding of batch k overlaps the forward pass of batch k+1.
`process(frames: List)` / `upscale_array` are frame-array in, frame-array out with no files at all.

Only the Real-ESRGAN backends are on this repository's path; the facade accepts the reference's other backend names
(hat_*, vrt, basicvsr_pp, diffusion, ensemble, realesrgan_ncnn) and treats them as unavailable, so its selection /
fallback logic ends on a `realesrgan_*` backend exactly as the reference's does on a machine without those packages.
"""
from __future__ import annotations

import logging
import time
from abc import ABC, abstractmethod
from collections import deque
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Deque, Dict, List, Optional, Tuple, Union

import numpy as np

from .pytorch_realesrgan import (PyTorchESRGANConfig, clear_upsampler_cache, get_upsampler,
                                 is_pytorch_esrgan_available)

logger = logging.getLogger(__name__)


@dataclass
class SRConfig:
    """Unified super-resolution configuration (reference :164-199, same fields and validation)."""

    scale: int = 4
    backend: str = "auto"
    half_precision: bool = True
    tile_size: int = 0  # 0 = auto, None = no tiling
    tile_overlap: int = 32
    temporal_window: int = 7
    gpu_id: int = 0
    quality_preset: str = "balanced"
    fallback_chain: Optional[List[str]] = None

    def __post_init__(self) -> None:
        if self.scale not in (2, 4):
            raise ValueError(f"scale must be 2 or 4, got {self.scale}")
        if self.temporal_window < 1:
            raise ValueError(f"temporal_window must be >= 1, got {self.temporal_window}")
        valid_presets = ["fast", "balanced", "quality", "maximum"]
        if self.quality_preset not in valid_presets:
            raise ValueError(f"quality_preset must be one of {valid_presets}")


@dataclass
class SRResult:
    """Result of super-resolution processing (reference :206-230)."""

    frames_processed: int = 0
    frames_failed: int = 0
    output_dir: Optional[Path] = None
    backend_used: str = "unknown"
    processing_time_seconds: float = 0.0
    avg_fps: float = 0.0
    peak_vram_mb: int = 0
    scale_factor: int = 4
    warnings: List[str] = field(default_factory=list)


class SRBackend(ABC):
    """Abstract operator interface (reference :237-311)."""

    @property
    @abstractmethod
    def name(self) -> str: ...

    @property
    @abstractmethod
    def supported_scales(self) -> List[int]: ...

    @abstractmethod
    def is_available(self) -> bool: ...

    @abstractmethod
    def estimate_vram_usage(self, width: int, height: int, scale: int) -> int: ...

    @abstractmethod
    def upscale_frame(self, frame: np.ndarray, scale: int = 4) -> np.ndarray: ...

    @abstractmethod
    def upscale_frames(self, input_dir: Path, output_dir: Path, scale: int = 4,
                       progress_callback: Optional[Callable[[float], None]] = None) -> SRResult: ...

    def clear_cache(self) -> None:
        pass


def _device_bytes_in_use(gpu_id: int) -> int:
    """Device memory in use (the engine allocates with cudaMalloc, outside torch's allocator statistics)."""
    try:
        import torch

        if not torch.cuda.is_available():
            return 0
        free, total = torch.cuda.mem_get_info(gpu_id)
        return int(total - free)
    except Exception:
        return 0


def upscale_frame_list(up: Any, frames: List[np.ndarray], scale: float = 4, max_batch: int = 16) -> List[np.ndarray]:
    """A list of frames through one upsampler: runs of same-size BGR uint8 frames as ONE engine call each (at most
    `max_batch` frames), anything else (gray, alpha, 16-bit, an `outscale` that is not the network's) frame by frame."""
    out: List[Optional[np.ndarray]] = [None] * len(frames)
    i = 0
    while i < len(frames):
        f = frames[i]
        plain = isinstance(f, np.ndarray) and f.ndim == 3 and f.shape[2] == 3 and f.dtype == np.uint8
        j = i + 1
        if plain and float(scale) == float(up.scale):
            while j < len(frames) and j - i < max_batch and isinstance(frames[j], np.ndarray) \
                    and frames[j].shape == f.shape and frames[j].dtype == np.uint8:
                j += 1
            res = up.enhance_batch(np.stack(frames[i:j]))
            for k in range(i, j):
                out[k] = res[k - i]
        else:
            out[i] = up.enhance(f, outscale=scale)[0]
        i = j
    return out  # type: ignore[return-value]


class B200RealESRGANBackend(SRBackend):
    """The reference's `RealESRGANBackend` contract on the B200 engine."""

    _VARIANTS = {
        "x2plus": "RealESRGAN_x2plus",
        "x4plus": "RealESRGAN_x4plus",
        "anime": "RealESRGAN_x4plus_anime_6B",
        "animevideo": "realesr-animevideov3",
        "general": "realesr-general-x4v3",
    }
    BATCH = 4            # frames per engine call on the frames-dir path
    IO_THREADS = 6       # PNG decode / encode threads
    LOOKAHEAD = 12       # frames decoded ahead of the forward pass

    def __init__(self, config: Optional[SRConfig] = None, hardware: Any = None, model_variant: str = "x4plus"):
        self.config = config or SRConfig()
        self.hardware = hardware
        self.model_variant = model_variant
        self._esrgan_config: Optional[PyTorchESRGANConfig] = None

    @property
    def name(self) -> str:
        return f"realesrgan_{self.model_variant}"

    @property
    def supported_scales(self) -> List[int]:
        return [2] if "x2" in self.model_variant else [4]

    def is_available(self) -> bool:
        return is_pytorch_esrgan_available()

    def _get_model_name(self) -> str:
        # unknown variants fall back to x4plus exactly like the reference (:483-492) -- including the facade's
        # "realesrgan_x2" (variant "x2"): the x4 network runs and `enhance(outscale=2)` resizes its result, as there
        return self._VARIANTS.get(self.model_variant, "RealESRGAN_x4plus")

    def _ensure_config(self) -> PyTorchESRGANConfig:
        if self._esrgan_config is None:
            self._esrgan_config = PyTorchESRGANConfig(
                model_name=self._get_model_name(),
                scale_factor=self.config.scale,
                tile_size=self.config.tile_size if self.config.tile_size is not None else 0,
                half_precision=self.config.half_precision,
                gpu_id=self.config.gpu_id,
            )
        return self._esrgan_config

    def estimate_vram_usage(self, width: int, height: int, scale: int) -> int:
        """Reference formula (:507-512): base 2000 MB (1500 for anime) + 12 B/px x (1 + scale^2).  (What this engine
        really holds for a frame is `tile_sizing.engine_workspace_mb`.)"""
        base_vram = 1500 if "anime" in self.model_variant else 2000
        frame_vram = (width * height * 3 * 4 * (1 + scale * scale)) // (1024 * 1024)
        return base_vram + frame_vram

    # ---- frame-array paths
    def upscale_frame(self, frame: np.ndarray, scale: int = 4) -> np.ndarray:
        output, _ = get_upsampler(self._ensure_config()).enhance(frame, outscale=scale)
        return output

    def upscale_array(self, frames: np.ndarray) -> np.ndarray:
        """[N,H,W,3] uint8 BGR -> [N,sH,sW,3]: one call, pipelined inside the engine."""
        return get_upsampler(self._ensure_config()).enhance_batch(frames)

    def process(self, frames: List[np.ndarray], scale: int = 4) -> List[np.ndarray]:
        """List of frames in, list of frames out (`SuperResolution.process`, :1502-1519): runs of same-size BGR uint8
        frames go through `upscale_array`, anything else (gray, alpha, 16-bit) through `upscale_frame`."""
        return upscale_frame_list(get_upsampler(self._ensure_config()), frames, scale)

    # ---- frames directory in, frames directory out
    def upscale_frames(self, input_dir: Path, output_dir: Path, scale: int = 4,
                       progress_callback: Optional[Callable[[float], None]] = None) -> SRResult:
        import cv2

        from .engine import EngineOutOfMemory

        result = SRResult(backend_used=self.name, scale_factor=scale)
        start_time = time.time()
        cfg = self._ensure_config()
        input_dir, output_dir = Path(input_dir), Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        result.output_dir = output_dir
        frames = sorted(input_dir.glob("*.png")) or sorted(input_dir.glob("*.jpg"))
        if not frames:
            result.warnings.append("No frames found")
            return result
        total = len(frames)
        done = 0
        peak_vram = 0

        def finished(path: Path, err: Optional[str]) -> None:
            nonlocal done
            done += 1
            if err is None:
                result.frames_processed += 1
            else:
                result.frames_failed += 1
                result.warnings.append(f"Frame {path.name}: {err}")
            if progress_callback:
                progress_callback(done / total)

        def read(path: Path):
            return cv2.imread(str(path), cv2.IMREAD_UNCHANGED)

        def write(path: Path, img: np.ndarray) -> Optional[str]:
            try:
                if not cv2.imwrite(str(path), img) or not path.exists():
                    return "Output file was not created"
                return None
            except Exception as e:
                return str(e)

        try:
            get_upsampler(cfg)
        except Exception as e:   # engine / weights unavailable: every frame fails with that message
            for p in frames:
                finished(p, str(e))
            result.processing_time_seconds = time.time() - start_time
            return result

        with ThreadPoolExecutor(max_workers=self.IO_THREADS) as io:
            reads: Deque[Tuple[Path, Future]] = deque()
            writes: Deque[Tuple[Path, Future]] = deque()
            nxt = 0

            def drain_writes(block: bool) -> None:
                while writes and (block or writes[0][1].done()):
                    p, fut = writes.popleft()
                    finished(p, fut.result())

            def run_batch(batch: List[Tuple[Path, np.ndarray]]) -> None:
                nonlocal peak_vram
                if not batch:
                    return
                try:
                    # (looked up per batch: after an out-of-memory failure the cache is cleared -- the engine is gone --
                    # and the next batch gets a fresh one, as the reference's per-frame get_upsampler does)
                    up = get_upsampler(cfg)
                    plain = all(im.ndim == 3 and im.shape[2] == 3 and im.dtype == np.uint8 for _, im in batch) \
                        and float(scale) == float(up.scale)
                    if plain and len(batch) > 1:
                        outs = up.enhance_batch(np.stack([im for _, im in batch]))
                    else:
                        outs = [up.enhance(im, outscale=scale)[0] for _, im in batch]
                except EngineOutOfMemory as e:
                    if len(batch) > 1:                      # smaller launches first, then give up frame by frame
                        for item in batch:
                            run_batch([item])
                        return
                    clear_upsampler_cache()
                    finished(batch[0][0], f"GPU out of memory: {e}\n"
                                          f"Try: 1) Reduce tile_size, 2) Use smaller model, 3) Close other GPU applications")
                    return
                except Exception as e:
                    if len(batch) > 1:                      # find the frame(s) at fault: one by one
                        for item in batch:
                            run_batch([item])
                        return
                    finished(batch[0][0], str(e))
                    return
                peak_vram = max(peak_vram, _device_bytes_in_use(self.config.gpu_id))
                for (p, _), o in zip(batch, outs):
                    writes.append((p, io.submit(write, output_dir / p.name, o)))
                del outs
                while len(writes) > 2 * self.BATCH + self.IO_THREADS:   # bound the encoded-frame backlog (44 MB each)
                    p, fut = writes.popleft()
                    finished(p, fut.result())

            batch: List[Tuple[Path, np.ndarray]] = []
            while nxt < total or reads:
                while nxt < total and len(reads) < self.LOOKAHEAD:
                    reads.append((frames[nxt], io.submit(read, frames[nxt])))
                    nxt += 1
                p, fut = reads.popleft()
                try:
                    img = fut.result()
                except Exception as e:  # pragma: no cover - cv2.imread returns None instead of raising
                    img, err = None, str(e)
                else:
                    err = f"Failed to read image: {p}"
                if img is None:
                    finished(p, err)
                    continue
                if batch and (batch[0][1].shape != img.shape or batch[0][1].dtype != img.dtype):
                    run_batch(batch)
                    batch = []
                batch.append((p, img))
                if len(batch) >= self.BATCH:
                    run_batch(batch)
                    batch = []
                drain_writes(block=False)
            run_batch(batch)
            drain_writes(block=True)

        result.processing_time_seconds = time.time() - start_time
        result.peak_vram_mb = peak_vram // (1024 * 1024)
        if result.processing_time_seconds > 0 and result.frames_processed > 0:
            result.avg_fps = result.frames_processed / result.processing_time_seconds
        return result

    def clear_cache(self) -> None:
        clear_upsampler_cache()


# ---------------------------------------------------------------------------------------------------------------
@dataclass
class _Hardware:
    """What the facade reads from the reference's `HardwareInfo` (`infrastructure/gpu/detector.py`), from torch."""
    tier: Any = "cuda"
    gpu_vendor: Any = "nvidia"
    gpu_name: str = "unknown"
    vram_total_mb: int = 0
    vram_free_mb: int = 0


def detect_hardware(gpu_id: int = 0) -> _Hardware:
    try:
        import torch

        if torch.cuda.is_available():
            p = torch.cuda.get_device_properties(gpu_id)
            free, total = torch.cuda.mem_get_info(gpu_id)
            return _Hardware("cuda", "nvidia", p.name, int(total >> 20), int(free >> 20))
    except Exception:
        pass
    return _Hardware("cpu", "unknown", "none", 0, 0)


class SuperResolution:
    """Backend selection + fallback, `upscale(dir, dir)`, `upscale_frame`, `process(list)` (reference :1194-1527)."""

    BACKENDS: Dict[str, Optional[type]] = {
        "realesrgan_ncnn": None,                       # ncnn-vulkan binary: not this repository's path
        "realesrgan_x2": B200RealESRGANBackend,
        "realesrgan_x4": B200RealESRGANBackend,
        "realesrgan_anime": B200RealESRGANBackend,
        "realesrgan_general": B200RealESRGANBackend,   # realesr-general-x4v3 (SRVGGNetCompact)
        "realesrgan_animevideo": B200RealESRGANBackend,
        "hat_small": None, "hat_base": None, "hat_large": None, "basicvsr_pp": None, "vrt": None,
        "diffusion": None, "ensemble": None,
    }
    FALLBACK_CHAINS: Dict[str, List[str]] = {
        "fast": ["realesrgan_ncnn", "realesrgan_x4", "realesrgan_x2"],
        "balanced": ["hat_base", "realesrgan_x4", "hat_small", "realesrgan_ncnn"],
        "quality": ["hat_large", "hat_base", "vrt", "basicvsr_pp", "realesrgan_x4"],
        "maximum": ["ensemble", "diffusion", "hat_large", "vrt", "hat_base"],
    }

    def __init__(self, config: Optional[SRConfig] = None, hardware: Any = None):
        self.config = config or SRConfig()
        self.hardware = hardware or detect_hardware(self.config.gpu_id)
        self._available_backends: Dict[str, bool] = {}
        self.backend = self._select_backend()

    def _create_backend(self, backend_name: str) -> Optional[SRBackend]:
        if backend_name not in self.BACKENDS:
            logger.warning(f"Unknown backend: {backend_name}")
            return None
        cls = self.BACKENDS[backend_name]
        if cls is None:
            return None
        try:
            return cls(self.config, self.hardware, backend_name.replace("realesrgan_", ""))
        except Exception as e:
            logger.warning(f"Failed to create backend {backend_name}: {e}")
            return None

    def _check_backend_available(self, backend_name: str) -> bool:
        if backend_name not in self._available_backends:
            b = self._create_backend(backend_name)
            self._available_backends[backend_name] = b is not None and b.is_available()
        return self._available_backends[backend_name]

    def _select_optimal_backend_for_hardware(self) -> str:
        # the reference's ladder (:1326-1386) restricted to what exists here: scale picks the network
        order = ["realesrgan_x2", "realesrgan_x4"] if self.config.scale == 2 else ["realesrgan_x4", "realesrgan_x2"]
        for name in order:
            if self._check_backend_available(name):
                return name
        raise RuntimeError("No super-resolution backend available")

    def _select_backend(self) -> SRBackend:
        backend_name = self.config.backend
        if backend_name == "auto":
            backend_name = self._select_optimal_backend_for_hardware()
            logger.info(f"Auto-selected backend: {backend_name}")
        backend = self._create_backend(backend_name)
        if backend and backend.is_available():
            return backend
        chain = self.config.fallback_chain or self.FALLBACK_CHAINS.get(self.config.quality_preset,
                                                                       self.FALLBACK_CHAINS["balanced"])
        logger.warning(f"Backend {backend_name} not available, trying fallbacks")
        for name in list(chain) + ["realesrgan_x4"]:
            if name == backend_name:
                continue
            fb = self._create_backend(name)
            if fb and fb.is_available():
                logger.info(f"Using fallback backend: {name}")
                return fb
        raise RuntimeError("No super-resolution backend available after fallbacks")

    def get_available_backends(self) -> List[str]:
        return [n for n in self.BACKENDS if self._check_backend_available(n)]

    def get_backend_info(self) -> Dict[str, Any]:
        hw = self.hardware
        tier = getattr(hw, "tier", "cuda")
        return {"name": self.backend.name, "supported_scales": self.backend.supported_scales,
                "hardware_tier": getattr(tier, "value", tier), "vram_total_mb": getattr(hw, "vram_total_mb", 0),
                "vram_free_mb": getattr(hw, "vram_free_mb", 0), "gpu_name": getattr(hw, "gpu_name", "unknown")}

    def estimate_vram_usage(self, width: int, height: int) -> int:
        return self.backend.estimate_vram_usage(width, height, self.config.scale)

    def upscale_frame(self, frame: np.ndarray) -> np.ndarray:
        return self.backend.upscale_frame(frame, self.config.scale)

    def upscale(self, input_dir: Union[str, Path], output_dir: Union[str, Path],
                progress_callback: Optional[Callable[[float], None]] = None) -> SRResult:
        logger.info(f"Starting {self.config.scale}x upscale with {self.backend.name} backend")
        return self.backend.upscale_frames(Path(input_dir), Path(output_dir), self.config.scale, progress_callback)

    def process(self, frames: List[np.ndarray], scale: int = 4) -> List[np.ndarray]:
        if hasattr(self.backend, "process"):
            return self.backend.process(frames, scale)
        return [self.backend.upscale_frame(f, scale) for f in frames]

    def clear_cache(self) -> None:
        if self.backend:
            self.backend.clear_cache()


def create_super_resolution(scale: int = 4, backend: str = "auto", quality_preset: str = "balanced", gpu_id: int = 0,
                            **kwargs) -> SuperResolution:
    config = SRConfig(scale=scale, backend=backend, quality_preset=quality_preset, gpu_id=gpu_id, **kwargs)
    return SuperResolution(config, detect_hardware(gpu_id))


def upscale_frames(input_dir: Union[str, Path], output_dir: Union[str, Path], scale: int = 4, backend: str = "auto",
                   progress_callback: Optional[Callable[[float], None]] = None, model_variant: Optional[str] = None,
                   **kwargs) -> SRResult:
    """Convenience function (reference :1565-1587).  `model_variant` ("anime", "general", ...) picks the network
    directly."""
    if model_variant is not None:
        backend = f"realesrgan_{model_variant}"
    sr = create_super_resolution(scale=scale, backend=backend, **kwargs)
    return sr.upscale(input_dir, output_dir, progress_callback)


def get_recommended_backend(hardware: Any = None) -> str:
    return SuperResolution(SRConfig(backend="auto"), hardware or detect_hardware()).backend.name


def list_available_backends() -> List[str]:
    return SuperResolution(SRConfig()).get_available_backends()
