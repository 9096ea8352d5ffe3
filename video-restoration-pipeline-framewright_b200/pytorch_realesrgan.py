"""Drop-in for `framewright.processors.pytorch_realesrgan` on a B200.

Mirrors the reference module's public surface one to one
(`/root/reference/src/framewright/processors/pytorch_realesrgan.py`):

    PyTorchESRGANConfig (+ validate)          :35-61
    is_pytorch_esrgan_available()             :64-82
    get_upsampler(config)                     :85-173
    enhance_frame_pytorch(in, out, config)    :176-247   -> (ok, error_message), never raises
    clear_upsampler_cache()                   :250-260
    NCNN_TO_PYTORCH_MODEL, convert_ncnn_model_name()  :263-275

Same names, argument meaning and error behaviour; the arithmetic runs in libb200sr.so instead of
PyPI `realesrgan`/`basicsr`.  Deliberate differences, all documented in DESIGN.md:
  * the upsampler cache is keyed by (model, tile, tile_pad, pre_pad, gpu) instead of being one global that
    ignores the config (reference :153-156 returns whatever was created first);
  * the model table is not rebuilt (5 random-initialised networks) on every call (reference :103-129);
  * `realesr-general-x4v3` / `realesr-animevideov3` are SRVGGNetCompact, as upstream ships them;
  * no `gc.collect()` + `empty_cache()` around every frame (reference :221, gpu_memory_optimizer.py:297-323).
"""
from __future__ import annotations

import logging
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

_PYTORCH_ESRGAN_AVAILABLE: Optional[bool] = None
_UPSAMPLER = None                      # most recently created upsampler (reference keeps one global)
_UPSAMPLERS: Dict[tuple, object] = {}
_CACHE_LOCK = threading.Lock()

VALID_MODELS = [
    "RealESRGAN_x4plus",
    "RealESRGAN_x4plus_anime_6B",
    "RealESRGAN_x2plus",
    "realesr-animevideov3",
    "realesr-general-x4v3",
]

# weight URLs the reference hands to RealESRGANer (pytorch_realesrgan.py:106,111,116,121,126); the file name
# selects the architecture, a local copy next to the process (or $B200SR_WEIGHTS_DIR) is loaded if present.
MODEL_URLS = {
    "RealESRGAN_x4plus": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.1.0/RealESRGAN_x4plus.pth",
    "RealESRGAN_x4plus_anime_6B": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.2.4/RealESRGAN_x4plus_anime_6B.pth",
    "RealESRGAN_x2plus": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.1/RealESRGAN_x2plus.pth",
    "realesr-animevideov3": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.5.0/realesr-animevideov3.pth",
    "realesr-general-x4v3": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.5.0/realesr-general-x4v3.pth",
}
MODEL_SCALES = {
    "RealESRGAN_x4plus": 4,
    "RealESRGAN_x4plus_anime_6B": 4,
    "RealESRGAN_x2plus": 2,
    "realesr-animevideov3": 4,
    "realesr-general-x4v3": 4,
}


@dataclass
class PyTorchESRGANConfig:
    """Configuration for Real-ESRGAN processing (field-for-field the reference dataclass)."""
    model_name: str = "RealESRGAN_x4plus"
    scale_factor: int = 4
    tile_size: int = 0  # 0 = auto
    tile_pad: int = 10
    pre_pad: int = 0
    half_precision: bool = True  # kept for API parity; the engine always uses 16-bit storage, fp32 accumulate
    gpu_id: int = 0

    def validate(self) -> None:
        if self.model_name not in VALID_MODELS:
            raise ValueError(
                f"Invalid model: {self.model_name}. "
                f"Supported models: {', '.join(VALID_MODELS)}"
            )
        if self.scale_factor not in [2, 4]:
            raise ValueError(f"Scale factor must be 2 or 4, got {self.scale_factor}")


def is_pytorch_esrgan_available() -> bool:
    """True when the B200 engine can run: the CUDA library loads and an sm_100 device is visible.  Memoised."""
    global _PYTORCH_ESRGAN_AVAILABLE
    if _PYTORCH_ESRGAN_AVAILABLE is not None:
        return _PYTORCH_ESRGAN_AVAILABLE
    try:
        import torch

        from . import _native

        _native.load()
        ok = torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
        if not ok:
            logger.warning("B200 Real-ESRGAN engine not available: no sm_100 CUDA device")
        _PYTORCH_ESRGAN_AVAILABLE = bool(ok)
    except Exception as e:  # ImportError / NativeLibraryError
        logger.warning(f"B200 Real-ESRGAN engine not available: {e}")
        _PYTORCH_ESRGAN_AVAILABLE = False
    return _PYTORCH_ESRGAN_AVAILABLE


def _local_weights(model_name: str) -> Optional[str]:
    import os

    from .upsampler import weights_dir

    for d in (weights_dir(), "weights", "."):
        if d:
            p = os.path.join(d, model_name + ".pth")
            if os.path.isfile(p):
                return p
    return None


def _available_vram_mb(gpu_id: int) -> Optional[float]:
    """Free device memory in MB (what the reference reads from GPUMemoryOptimizer.get_memory_stats)."""
    try:
        import torch

        if not torch.cuda.is_available():
            return None
        free, _total = torch.cuda.mem_get_info(gpu_id)
        return free / (1024 ** 2)
    except Exception:
        return None


def _auto_tile(gpu_id: int) -> int:
    """Reference rule (:136-151): >= 24 GB -> no tiling, >= 12 -> 400, >= 8 -> 256, else 128."""
    import torch

    gpu_mem_gb = torch.cuda.get_device_properties(gpu_id).total_memory / (1024 ** 3)
    if gpu_mem_gb >= 24:
        tile = 0
    elif gpu_mem_gb >= 12:
        tile = 400
    elif gpu_mem_gb >= 8:
        tile = 256
    else:
        tile = 128
    logger.info(f"Auto-selected tile size {tile} for {gpu_mem_gb:.1f}GB GPU")
    return tile


def get_upsampler(config: PyTorchESRGANConfig):
    """Get or create the upsampler for this config (object with `.enhance(img, outscale)`)."""
    global _UPSAMPLER
    if not is_pytorch_esrgan_available():
        raise RuntimeError(
            "B200 Real-ESRGAN engine not available: build libb200sr.so "
            "(python -c 'import __graft_entry__ as g; g.build()') and run on an sm_100 GPU"
        )
    if config.model_name not in MODEL_SCALES:
        raise ValueError(f"Unknown model: {config.model_name}")
    from .upsampler import RealESRGANer

    tile = config.tile_size
    if tile == 0:
        tile = _auto_tile(config.gpu_id)
    key = (config.model_name, int(tile or 0), int(config.tile_pad), int(config.pre_pad), int(config.gpu_id))
    with _CACHE_LOCK:
        up = _UPSAMPLERS.get(key)
        if up is None:
            logger.info(f"Creating B200 Real-ESRGAN upsampler with model {config.model_name}")
            local = _local_weights(config.model_name)
            up = RealESRGANer(
                scale=MODEL_SCALES[config.model_name],
                model_path=local or MODEL_URLS[config.model_name],
                dni_weight=None,
                model=None,
                model_name=config.model_name,
                tile=tile,
                tile_pad=config.tile_pad,
                pre_pad=config.pre_pad,
                half=config.half_precision,
                gpu_id=config.gpu_id,
            )
            _UPSAMPLERS[key] = up
        _UPSAMPLER = up
        return up


def enhance_frame_pytorch(
    input_path: Path,
    output_path: Path,
    config: PyTorchESRGANConfig,
) -> Tuple[bool, Optional[str]]:
    """Enhance one frame file -> file.  Returns (success, error_message); never raises.

    Frame-array form: the reference's `EnsembleSR` member (`processors/ensemble_sr.py:206-211`) calls this function as
    `enhance_frame_pytorch(frame, upsampler, config)` with a BGR ndarray and the object `get_upsampler` returned, and
    uses the return value as the upscaled frame.  (Against the reference's own function that call can only fail --
    it `imread`s `str(frame)`.)  Here it does what the caller means: ndarray in -> upscaled ndarray out through that
    upsampler, exceptions propagating to the caller's `try`, which turns them into `None`."""
    if isinstance(input_path, np.ndarray):
        config.validate()
        upsampler = output_path if hasattr(output_path, "enhance") else get_upsampler(config)
        return upsampler.enhance(input_path, outscale=config.scale_factor)[0]
    try:
        import cv2

        from .engine import EngineOutOfMemory

        config.validate()
        img = cv2.imread(str(input_path), cv2.IMREAD_UNCHANGED)
        if img is None:
            return False, f"Failed to read image: {input_path}"
        if config.tile_size == 0:  # auto mode mutates the config exactly like the reference (:208-218)
            available_mb = _available_vram_mb(config.gpu_id)
            if available_mb is not None:
                if available_mb > 8000:
                    config.tile_size = 0
                elif available_mb > 4000:
                    config.tile_size = 512
                elif available_mb > 2000:
                    config.tile_size = 384
                else:
                    config.tile_size = 256
                logger.debug(f"Auto-selected tile size {config.tile_size} based on {available_mb:.0f}MB VRAM")
        try:
            upsampler = get_upsampler(config)
            output, _ = upsampler.enhance(img, outscale=config.scale_factor)
        except EngineOutOfMemory as e:
            clear_upsampler_cache()
            return False, (
                f"GPU out of memory: {e}\n"
                f"Try: 1) Reduce tile_size, 2) Use smaller model, 3) Close other GPU applications"
            )
        cv2.imwrite(str(output_path), output)
        if not Path(output_path).exists():
            return False, "Output file was not created"
        return True, None
    except Exception as e:
        logger.error(f"B200 Real-ESRGAN failed: {e}")
        return False, str(e)


def enhance_frames_batch(frames: np.ndarray, config: PyTorchESRGANConfig) -> np.ndarray:
    """Frame-array in / frame-array out for a stack of same-size uint8 BGR frames [N,H,W,3] (one launch sequence)."""
    config.validate()
    return get_upsampler(config).enhance_batch(frames)


def clear_upsampler_cache():
    """Drop cached upsamplers and free their GPU memory."""
    global _UPSAMPLER
    with _CACHE_LOCK:
        ups = list(_UPSAMPLERS.values())
        _UPSAMPLERS.clear()
        had = _UPSAMPLER is not None or bool(ups)
        _UPSAMPLER = None
    for up in ups:
        try:
            up.close()
        except Exception:
            pass
    if had:
        try:
            import torch

            from .engine import PINNED_POOL

            PINNED_POOL.trim()
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        except ImportError:
            pass


# Map ncnn model names to pytorch model names (reference :263-270)
NCNN_TO_PYTORCH_MODEL = {
    "realesrgan-x4plus": "RealESRGAN_x4plus",
    "realesrgan-x4plus-anime": "RealESRGAN_x4plus_anime_6B",
    "realesr-animevideov3": "realesr-animevideov3",
    "realesrnet-x4plus": "realesr-general-x4v3",
    "realesrgan-x2plus": "RealESRGAN_x2plus",
}


def convert_ncnn_model_name(ncnn_name: str) -> str:
    """Convert ncnn model name to PyTorch model name (unknown -> RealESRGAN_x4plus)."""
    return NCNN_TO_PYTORCH_MODEL.get(ncnn_name, "RealESRGAN_x4plus")
