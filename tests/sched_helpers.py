"""Picklable stand-ins for the B200 upsampler, used by the CPU tests of the multi-GPU scheduler (worker processes
import this module by name).  `scale` 2 nearest-neighbour "upscaling" + a marker so results are checkable."""
import time

import numpy as np


class FakeUpsampler:
    scale = 2

    def __init__(self, cfg):
        self.cfg = dict(cfg)
        self.gpu_id = int(cfg.get("gpu_id", 0))
        self.delay = float(cfg.get("tile_pad", 0)) * 1e-3 if self.gpu_id == 0 else 0.0   # tile_pad doubles as "GPU 0 is slow"
        self.closed = False

    def enhance_batch(self, frames, out=None):
        if self.delay:
            time.sleep(self.delay * len(frames))
        res = np.repeat(np.repeat(frames, 2, axis=1), 2, axis=2)
        if out is not None:
            out[...] = res
            return out
        return res

    def enhance(self, img, outscale=None):
        if img.ndim == 2:
            return np.repeat(np.repeat(img, 2, axis=0), 2, axis=1), "L"
        return np.repeat(np.repeat(img, 2, axis=0), 2, axis=1), "RGB"

    def close(self):
        self.closed = True


def fake_engine(cfg):
    return FakeUpsampler(cfg)


class CountingSource:
    """Synthetic frames generated in the worker from the index (no disk): frame i is filled with i % 251."""

    def __init__(self, n, h=6, w=8):
        self.n, self.h, self.w = n, h, w

    def __len__(self):
        return self.n

    def name(self, i):
        return f"frame_{i + 1:08d}.png"

    def open(self):
        pass

    def close(self):
        pass

    def load(self, i):
        return np.full((self.h, self.w, 3), i % 251, np.uint8)


class HangingUpsampler(FakeUpsampler):
    """GPU 1 never returns (a hung device): the parent must give up on it."""

    def enhance_batch(self, frames, out=None):
        if self.gpu_id == 1:
            time.sleep(3600)
        return super().enhance_batch(frames, out)


def hanging_engine(cfg):
    return HangingUpsampler(cfg)
