"""Operator-API mirror: the `SRBackend` interface of
`/root/reference/src/framewright/processors/enhancement/super_resolution.py` (:237-311) and the
`RealESRGANBackend` that implements it (:441-601), backed by the B200 engine.

`B200RealESRGANBackend` has the reference backend's constructor arguments (config, hardware, model_variant),
properties (`name`, `supported_scales`) and methods (`is_available`, `estimate_vram_usage`, `upscale_frame`,
`upscale_frames`, `clear_cache`) with the same argument meaning and error behaviour: per-frame failures are
collected in `SRResult.warnings` as "Frame <name>: <err>" and never raised.  It can be registered in the
reference's `SuperResolution.BACKENDS` table.
"""
from __future__ import annotations

import time
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, List, Optional

import numpy as np

from .pytorch_realesrgan import (PyTorchESRGANConfig, clear_upsampler_cache, enhance_frame_pytorch, get_upsampler,
                                 is_pytorch_esrgan_available)


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


class B200RealESRGANBackend(SRBackend):
    """`RealESRGANBackend` (reference :441-601) on the B200 engine."""

    _VARIANTS = {
        "x2plus": "RealESRGAN_x2plus",
        "x4plus": "RealESRGAN_x4plus",
        "anime": "RealESRGAN_x4plus_anime_6B",
        "animevideo": "realesr-animevideov3",
        "general": "realesr-general-x4v3",
    }

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
        if "x2" in self.model_variant:
            return [2]
        return [4]

    def is_available(self) -> bool:
        return is_pytorch_esrgan_available()

    def _get_model_name(self) -> str:
        # unknown variants fall back to x4plus exactly like the reference (:483-492)
        return self._VARIANTS.get(self.model_variant, "RealESRGAN_x4plus")

    def _ensure_config(self) -> None:
        if self._esrgan_config is None:
            self._esrgan_config = PyTorchESRGANConfig(
                model_name=self._get_model_name(),
                scale_factor=self.config.scale,
                tile_size=self.config.tile_size if self.config.tile_size is not None else 0,
                half_precision=self.config.half_precision,
                gpu_id=self.config.gpu_id,
            )

    def estimate_vram_usage(self, width: int, height: int, scale: int) -> int:
        """Reference formula (:507-512): base 2000 MB (1500 for anime) + 12 B/px x (1 + scale^2)."""
        base_vram = 1500 if "anime" in self.model_variant else 2000
        frame_vram = (width * height * 3 * 4 * (1 + scale * scale)) // (1024 * 1024)
        return base_vram + frame_vram

    def upscale_frame(self, frame: np.ndarray, scale: int = 4) -> np.ndarray:
        self._ensure_config()
        upsampler = get_upsampler(self._esrgan_config)
        output, _ = upsampler.enhance(frame, outscale=scale)
        return output

    def upscale_array(self, frames: np.ndarray) -> np.ndarray:
        """Frame-array in / frame-array out: [N,H,W,3] uint8 BGR -> [N,sH,sW,3] in one launch sequence."""
        self._ensure_config()
        return get_upsampler(self._esrgan_config).enhance_batch(frames)

    def upscale_frames(self, input_dir: Path, output_dir: Path, scale: int = 4,
                       progress_callback: Optional[Callable[[float], None]] = None) -> SRResult:
        result = SRResult(backend_used=self.name, scale_factor=scale)
        start_time = time.time()
        self._ensure_config()
        input_dir = Path(input_dir)
        output_dir = Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        result.output_dir = output_dir
        frames = sorted(input_dir.glob("*.png"))
        if not frames:
            frames = sorted(input_dir.glob("*.jpg"))
        if not frames:
            result.warnings.append("No frames found")
            return result
        total_frames = len(frames)
        peak_vram = 0
        for i, frame_path in enumerate(frames):
            try:
                output_path = output_dir / frame_path.name
                success, error = enhance_frame_pytorch(frame_path, output_path, self._esrgan_config)
                if success:
                    result.frames_processed += 1
                else:
                    result.frames_failed += 1
                    result.warnings.append(f"Frame {frame_path.name}: {error}")
                peak_vram = max(peak_vram, _device_bytes_in_use(self.config.gpu_id))
            except Exception as e:  # pragma: no cover - enhance_frame_pytorch never raises
                result.frames_failed += 1
                result.warnings.append(f"Frame {frame_path.name}: {str(e)}")
            if progress_callback:
                progress_callback((i + 1) / total_frames)
        result.processing_time_seconds = time.time() - start_time
        result.peak_vram_mb = peak_vram // (1024 * 1024)
        if result.processing_time_seconds > 0 and result.frames_processed > 0:
            result.avg_fps = result.frames_processed / result.processing_time_seconds
        return result

    def clear_cache(self) -> None:
        clear_upsampler_cache()


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


def upscale_frames(input_dir: Path, output_dir: Path, scale: int = 4, model_variant: Optional[str] = None,
                   progress_callback: Optional[Callable[[float], None]] = None) -> SRResult:
    """Convenience factory mirroring the reference's module-level `upscale_frames` (:1565-1587)."""
    variant = model_variant or ("x2plus" if scale == 2 else "x4plus")
    backend = B200RealESRGANBackend(SRConfig(scale=scale), None, variant)
    return backend.upscale_frames(Path(input_dir), Path(output_dir), scale=scale, progress_callback=progress_callback)
