"""Tile-size helpers of the reference (`/root/reference/src/framewright/utils/gpu.py:386-512`:
`calculate_optimal_tile_size`, `get_adaptive_tile_sequence`) for THIS engine.

The reference sizes tiles from "MB per output megapixel" coefficients fitted to PyTorch + cuDNN (450 / 400 / 250 / 350).
They are wrong here in both directions: the B200 engine keeps every activation tensor of a frame resident (chunk-planar
dense blocks, residual pair, fp32 skip, the 2x / 4x tail tensors), about 6.9 KB per LR pixel for RRDBNet x4 -- 6.3 GB for
one 1280x720 frame -- but on a 180 GB part that still means "no tiling" for anything up to 8K input.
`engine_workspace_bytes` is the exact figure (`b200sr_workspace_bytes` computes the same on the device side;
tests/test_gpu_boundary.py checks they agree); the two reference functions keep their signatures and return values'
meaning (0 = no tiling needed; multiples of 32; never below 128)."""
from __future__ import annotations

import logging
import math
from typing import List, Optional, Tuple

from .archs import MODEL_ARCHS

logger = logging.getLogger(__name__)

# reference model names (ncnn style) -> this package's names
_NAMES = {
    "realesrgan-x4plus": "RealESRGAN_x4plus", "realesrgan-x4plus-anime": "RealESRGAN_x4plus_anime_6B",
    "realesrgan-x2plus": "RealESRGAN_x2plus", "realesr-animevideov3": "realesr-animevideov3",
    "realesrnet-x4plus": "realesr-general-x4v3",
}


def _align(b: int, a: int = 1024) -> int:
    return (b + a - 1) // a * a


def _region_bytes(kind: str, n: int, h: int, w: int) -> int:
    px = n * h * w
    pxt = n * h * ((w + 127) // 128) * 128
    if kind != "rrdb":
        return 2 * _align(px * 128) + _align(px * 16)
    tg = max(1, min(n, (1280 * 720) // max(h * w, 1)))      # frames per HR-tail group (csrc/b200sr.cu::tail_group)
    tpx = tg * h * w
    return (3 * _align(px * 384) + 2 * _align(pxt * 64) + _align(pxt * 256) + _align(px * 128)
            + 2 * _align(tpx * 512) + 2 * _align(tpx * 2048))


def engine_workspace_bytes(model_name: str, width: int, height: int, n: int = 1, tile: int = 0, tile_pad: int = 10,
                           pre_pad: int = 0) -> int:
    """Device bytes one engine lane holds for `n` frames of width x height (tile > 0: for the largest padded tile)."""
    name = _NAMES.get(model_name, model_name)
    arch = MODEL_ARCHS.get(name, MODEL_ARCHS["RealESRGAN_x4plus"])
    s = 2 if (arch.kind == "rrdb" and arch.scale == 2) else 1
    hp, wp = height + pre_pad, width + pre_pad
    hp, wp = (hp + s - 1) // s * s, (wp + s - 1) // s * s
    if tile > 0:
        hp, wp = min(hp, tile + 2 * tile_pad), min(wp, tile + 2 * tile_pad)
    return _region_bytes(arch.kind, n, (hp + s - 1) // s, (wp + s - 1) // s)


def engine_workspace_mb(model_name: str, width: int, height: int, n: int = 1, tile: int = 0, tile_pad: int = 10,
                        pre_pad: int = 0) -> int:
    return int(math.ceil(engine_workspace_bytes(model_name, width, height, n, tile, tile_pad, pre_pad) / 2 ** 20))


def _free_vram_mb() -> Optional[int]:
    try:
        import torch

        if torch.cuda.is_available():
            free, _ = torch.cuda.mem_get_info()
            return int(free >> 20)
    except Exception:
        pass
    return None


def calculate_optimal_tile_size(frame_resolution: Tuple[int, int], scale_factor: int,
                                available_vram_mb: Optional[int] = None, model_name: str = "realesrgan-x4plus",
                                safety_factor: float = 0.7) -> int:
    """Largest tile (multiple of 32, >= 128, <= the frame) whose engine workspace fits `safety_factor` of the free
    device memory; 0 when the whole frame fits (reference signature and return convention, :386-463)."""
    width, height = frame_resolution
    if available_vram_mb is None:
        available_vram_mb = _free_vram_mb() or 2048          # the reference's conservative default
    usable = int(available_vram_mb * safety_factor)
    if engine_workspace_mb(model_name, width, height) <= usable:
        logger.debug("No tiling needed")
        return 0
    tile = (min(width, height) // 32) * 32
    while tile > 128 and engine_workspace_mb(model_name, width, height, tile=tile) > usable:
        tile -= 32
    tile = max(128, tile)
    tile = min(tile, min(width, height))
    logger.info(f"Calculated tile size: {tile} (frame: {width}x{height}, VRAM: {usable}MB available)")
    return tile


def get_adaptive_tile_sequence(frame_resolution: Tuple[int, int], scale_factor: int,
                               starting_tile_size: Optional[int] = None, min_tile_size: int = 128) -> List[int]:
    """Decreasing tile sizes for the caller's out-of-memory retry ladder (reference :465-512: start, x0.75 ..., each
    rounded down to 32, ending with `min_tile_size`)."""
    if starting_tile_size is None:
        starting_tile_size = calculate_optimal_tile_size(frame_resolution, scale_factor)
    if starting_tile_size == 0:
        starting_tile_size = min(frame_resolution)
    sizes: List[int] = []
    current = starting_tile_size
    while current >= min_tile_size:
        rounded = (current // 32) * 32
        if rounded >= min_tile_size and rounded not in sizes:
            sizes.append(rounded)
        current = int(current * 0.75)
    if min_tile_size not in sizes:
        sizes.append(min_tile_size)
    return sizes
