#!/usr/bin/env python
"""Host-plumbing ceiling of the multi-GPU scheduler, WITHOUT a GPU: the real `SchedulerPool` (spawned worker
processes, claim table, shared-memory frame arrays, ordered ring) around a stand-in engine that does no work, at the
headline frame size (1280x720 in, 5120x2880 out).  What it prints is the frame rate the Python / shared-memory side can
sustain -- the rate the GPUs would have to exceed for the plumbing to become the bound.

    python tools/plumbing_bench.py [--workers 4] [--frames 200]

Measured in the build container (8 host cores, 4 workers): ordered ring (`stream`, every output frame handed to the
consumer in order) 730-1120 frames/s; frame-array job (`run`, ArraySource -> ArraySink) 2600 frames/s -- against 27.7
frames/s per GPU and 220 at 8 GPUs on the device.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class NullUpsampler:
    """Touches one row in 64 of every output frame (the real engine's D2H copy fills the rest by DMA)."""
    scale = 4

    def __init__(self, cfg):
        self.cfg = cfg

    def enhance_batch(self, frames, out=None):
        n, h, w, c = frames.shape
        if out is None:
            out = np.empty((n, h * 4, w * 4, c), np.uint8)
        out[:, ::64] = 7
        return out

    def enhance(self, img, outscale=None):
        return self.enhance_batch(img[None])[0], "RGB"

    def close(self):
        pass


def null_engine(cfg):
    return NullUpsampler(cfg)


def main():
    import framewright_b200  # noqa: F401
    from framewright_b200.scheduler import ArraySink, ArraySource, SchedulerPool, SharedArray

    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=4, help="worker processes (stand-ins for GPUs)")
    ap.add_argument("--frames", type=int, default=200)
    args = ap.parse_args()
    cfg = {"model_name": "RealESRGAN_x4plus", "scale_factor": 4, "tile_size": 0, "tile_pad": 10, "pre_pad": 0}
    h, w, n = 720, 1280, args.frames
    pool = SchedulerPool(list(range(args.workers)), workers_per_gpu=2, start_timeout=120)
    frame = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
    sink = open(os.devnull, "wb")
    na = min(n, 64)
    sin, sout = SharedArray((na, h, w, 3)), SharedArray((na, 4 * h, 4 * w, 3))
    try:
        for rep in range(2):
            t0 = time.time()
            res = pool.stream((frame for _ in range(n)), cfg, lambda i, out: sink.write(memoryview(out).cast("B")),
                              num_frames=n, frame_shape=(h, w), scale=4, batch=2, engine_factory=null_engine)
            dt = time.time() - t0
            print(f"ordered ring   workers={args.workers} frames={n} ok={len(res.ok)}  {n / dt:8.1f} frames/s")
            sin.array[...] = rep
            t0 = time.time()
            res = pool.run(ArraySource(sin), ArraySink(sout), cfg, batch=2, engine_factory=null_engine)
            dt = time.time() - t0
            print(f"frame arrays   workers={args.workers} frames={na} ok={len(res.ok)}  {na / dt:8.1f} frames/s")
    finally:
        sin.release()
        sout.release()
        pool.close()


if __name__ == "__main__":
    main()
