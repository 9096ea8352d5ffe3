"""Multi-GPU frame scheduler for the upscaling path: one process and one engine per GPU, frames
sharded by index, no collective (frames are independent; RRDBNet / SRVGG have no cross-frame state).

API mirror of `/root/reference/src/framewright/utils/multi_gpu.py`:
    LoadBalanceStrategy (:30-37), GPUInfo (:39-93), DistributionResult (:95-137),
    MultiGPUDistributor(...).distribute_frames(frames, process_fn, output_dir, progress_callback) (:511-778),
    _assign_frames (:780-869), distribute_frames(...) convenience (:895-925).
The reference runs `workers_per_gpu x n_gpus` threads in one process and is never called by its own pipeline
(SURVEY.md finding 5).  Here `distribute_frames` keeps the signature and result type; when `process_fn` is
omitted the frames go through `upscale_shard` in one spawned worker process per GPU (CUDA_VISIBLE_DEVICES
pinned), each holding one B200 engine.

`shard_range` / `assign_contiguous` are the partition used by bench.py (rank r of world G owns frames
[r*N/G, (r+1)*N/G) -- the counts `_assign_frames` ROUND_ROBIN produces, in contiguous order).
"""
from __future__ import annotations

import logging
import multiprocessing as mp
import os
import time
from dataclasses import dataclass, field
from enum import Enum
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

logger = logging.getLogger(__name__)


class LoadBalanceStrategy(Enum):
    ROUND_ROBIN = "round_robin"
    LEAST_LOADED = "least_loaded"
    VRAM_AWARE = "vram_aware"
    WEIGHTED = "weighted"


@dataclass
class GPUInfo:
    id: int
    name: str
    total_vram_mb: int
    free_vram_mb: int
    utilization_pct: float
    temperature_c: Optional[float] = None
    pcie_bandwidth_gbps: Optional[float] = None
    compute_capability: Optional[str] = None

    @property
    def used_vram_mb(self) -> int:
        return self.total_vram_mb - self.free_vram_mb

    @property
    def vram_usage_pct(self) -> float:
        if self.total_vram_mb == 0:
            return 0.0
        return (self.used_vram_mb / self.total_vram_mb) * 100

    @property
    def is_healthy(self) -> bool:
        if self.temperature_c is not None and self.temperature_c > 90:
            return False
        return True

    @property
    def effective_capacity(self) -> float:
        vram_score = self.free_vram_mb / max(self.total_vram_mb, 1)
        util_score = 1.0 - (self.utilization_pct / 100.0)
        return (vram_score * 0.7) + (util_score * 0.3)


@dataclass
class DistributionResult:
    frames_per_gpu: Dict[int, List[Path]] = field(default_factory=dict)
    total_time: float = 0.0
    speedup_factor: float = 1.0
    gpu_utilization: Dict[int, float] = field(default_factory=dict)
    errors: Dict[str, str] = field(default_factory=dict)
    retried_frames: List[Path] = field(default_factory=list)

    @property
    def total_frames(self) -> int:
        return sum(len(frames) for frames in self.frames_per_gpu.values())

    @property
    def success_rate(self) -> float:
        total = self.total_frames + len(self.errors)
        if total == 0:
            return 100.0
        return (self.total_frames / total) * 100

    def summary(self) -> str:
        gpu_counts = ", ".join(f"GPU{gid}: {len(frames)}" for gid, frames in self.frames_per_gpu.items())
        return (
            f"Processed {self.total_frames} frames across {len(self.frames_per_gpu)} GPUs "
            f"({gpu_counts}) in {self.total_time:.1f}s "
            f"(speedup: {self.speedup_factor:.2f}x, success: {self.success_rate:.1f}%)"
        )


# ------------------------------------------------------------------------------------------ partition
def shard_range(num_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; sizes differ by at most one, earlier ranks get the extras."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(num_frames, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assign_contiguous(frames: Sequence, gpu_ids: Sequence[int]) -> Dict[int, list]:
    return {gid: list(frames[slice(*shard_range(len(frames), len(gpu_ids), i))]) for i, gid in enumerate(gpu_ids)}


def assign_frames(frames: Sequence, gpus: Sequence[GPUInfo], strategy: LoadBalanceStrategy) -> Dict[int, list]:
    """Reference `_assign_frames` semantics (:780-869) for every strategy."""
    assignments: Dict[int, list] = {g.id: [] for g in gpus}
    n = len(gpus)
    if n == 0:
        return assignments

    def round_robin(order):
        for i, frame in enumerate(frames):
            assignments[order[i % len(order)].id].append(frame)

    if strategy == LoadBalanceStrategy.ROUND_ROBIN:
        round_robin(list(gpus))
    elif strategy == LoadBalanceStrategy.LEAST_LOADED:
        round_robin(sorted(gpus, key=lambda g: g.utilization_pct))
    elif strategy == LoadBalanceStrategy.VRAM_AWARE:
        total_free = sum(g.free_vram_mb for g in gpus)
        if total_free == 0:
            round_robin(list(gpus))
        else:
            idx = 0
            for g in gpus:
                target = int(len(frames) * (g.free_vram_mb / total_free))
                for _ in range(target):
                    if idx < len(frames):
                        assignments[g.id].append(frames[idx])
                        idx += 1
            while idx < len(frames):
                assignments[gpus[idx % n].id].append(frames[idx])
                idx += 1
    elif strategy == LoadBalanceStrategy.WEIGHTED:
        caps = {g.id: g.effective_capacity for g in gpus}
        total = sum(caps.values())
        if total == 0:
            round_robin(list(gpus))
        else:
            idx = 0
            for gid, cap in caps.items():
                for _ in range(int(len(frames) * (cap / total))):
                    if idx < len(frames):
                        assignments[gid].append(frames[idx])
                        idx += 1
            best = max(caps.keys(), key=lambda k: caps[k])
            while idx < len(frames):
                assignments[best].append(frames[idx])
                idx += 1
    return assignments


# ------------------------------------------------------------------------------------------ discovery
def query_gpus() -> List[GPUInfo]:
    """Visible CUDA devices as GPUInfo (torch for names/memory; utilisation/temperature via pynvml if present)."""
    try:
        import torch
    except ImportError:
        return []
    if not torch.cuda.is_available():
        return []
    infos = []
    nvml = None
    try:
        import pynvml

        pynvml.nvmlInit()
        nvml = pynvml
    except Exception:
        nvml = None
    for i in range(torch.cuda.device_count()):
        p = torch.cuda.get_device_properties(i)
        try:
            free, total = torch.cuda.mem_get_info(i)
        except Exception:
            free, total = p.total_memory, p.total_memory
        util, temp = 0.0, None
        if nvml is not None:
            try:
                h = nvml.nvmlDeviceGetHandleByIndex(i)
                util = float(nvml.nvmlDeviceGetUtilizationRates(h).gpu)
                temp = float(nvml.nvmlDeviceGetTemperature(h, nvml.NVML_TEMPERATURE_GPU))
            except Exception:
                pass
        infos.append(GPUInfo(i, p.name, int(total // 2 ** 20), int(free // 2 ** 20), util, temp, None,
                             f"{p.major}.{p.minor}"))
    return infos


# ------------------------------------------------------------------------------------------ workers
def upscale_shard(frame_paths: Sequence[str], output_dir: str, gpu_id: int, model_name: str = "RealESRGAN_x4plus",
                  scale: int = 4, tile: int = 0, tile_pad: int = 10, pre_pad: int = 0,
                  batch: int = 4) -> List[Tuple[str, bool, Optional[str]]]:
    """Upscale one shard of frame files on one GPU with one engine; same-size frames are batched."""
    import cv2
    import numpy as np

    from .pytorch_realesrgan import PyTorchESRGANConfig, get_upsampler

    cfg = PyTorchESRGANConfig(model_name=model_name, scale_factor=scale, tile_size=tile, tile_pad=tile_pad,
                              pre_pad=pre_pad, gpu_id=gpu_id)
    cfg.validate()
    up = get_upsampler(cfg)
    out: List[Tuple[str, bool, Optional[str]]] = []
    os.makedirs(output_dir, exist_ok=True)
    pending: List[Tuple[str, np.ndarray]] = []

    def flush():
        if not pending:
            return
        try:
            res = up.enhance_batch(np.stack([im for _, im in pending]))
            for (p, _), o in zip(pending, res):
                dst = os.path.join(output_dir, os.path.basename(p))
                ok = cv2.imwrite(dst, o)
                out.append((dst, bool(ok), None if ok else "Output file was not created"))
        except Exception as e:  # one bad batch must not lose the shard
            for p, _ in pending:
                out.append((os.path.join(output_dir, os.path.basename(p)), False, str(e)))
        pending.clear()

    for p in frame_paths:
        img = cv2.imread(str(p), cv2.IMREAD_COLOR)
        if img is None:
            out.append((os.path.join(output_dir, os.path.basename(str(p))), False, f"Failed to read image: {p}"))
            continue
        if pending and pending[0][1].shape != img.shape:
            flush()
        pending.append((str(p), img))
        if len(pending) >= batch:
            flush()
    flush()
    return out


def _worker_entry(gpu_id: int, frame_paths, output_dir, kwargs, queue):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu_id)  # one process per GPU; the engine sees it as device 0
    try:
        res = upscale_shard(frame_paths, output_dir, 0, **kwargs)
        queue.put((gpu_id, res, None))
    except Exception as e:  # pragma: no cover
        queue.put((gpu_id, [], f"{type(e).__name__}: {e}"))


class MultiGPUDistributor:
    """`distribute_frames` with the reference's signature and result type."""

    def __init__(self, gpu_manager=None, strategy: LoadBalanceStrategy = LoadBalanceStrategy.ROUND_ROBIN,
                 workers_per_gpu: int = 1, max_retries: int = 2, enable_work_stealing: bool = True,
                 gpus: Optional[List[GPUInfo]] = None, **engine_kwargs):
        self.gpu_manager = gpu_manager
        self.strategy = strategy
        self.workers_per_gpu = workers_per_gpu
        self.max_retries = max_retries
        self.enable_work_stealing = enable_work_stealing
        self._gpus = gpus
        self._engine_kwargs = engine_kwargs
        self._result: Optional[DistributionResult] = None

    def _healthy_gpus(self) -> List[GPUInfo]:
        if self._gpus is not None:
            gpus = self._gpus
        elif self.gpu_manager is not None:
            return list(self.gpu_manager.get_healthy_gpus())
        else:
            gpus = query_gpus()
        return [g for g in gpus if g.is_healthy]

    def _assign_frames(self, frames: List[Path], gpus: List[GPUInfo]) -> Dict[int, List[Path]]:
        return assign_frames(frames, gpus, self.strategy)

    def distribute_frames(self, frames: List[Path],
                          process_fn: Optional[Callable[[Path, Path, int], Tuple[Path, bool, Optional[str]]]] = None,
                          output_dir: Path = Path("."),
                          progress_callback: Optional[Callable[[float, str], None]] = None) -> DistributionResult:
        if not frames:
            return DistributionResult()
        start = time.time()
        gpus = self._healthy_gpus()
        if not gpus:
            logger.error("No healthy GPUs available")
            return DistributionResult(errors={str(f): "No GPUs available" for f in frames})
        output_dir = Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        result = DistributionResult()
        result.frames_per_gpu = {g.id: [] for g in gpus}
        assignments = self._assign_frames(list(frames), gpus)
        done = 0

        if process_fn is not None:
            # caller-supplied per-frame function: run each GPU's list in its own thread (reference behaviour),
            # failed frames are retried once on the next GPU.
            import threading

            lock = threading.Lock()

            def run(gid: int, todo: List[Path]):
                nonlocal done
                for f in todo:
                    try:
                        outp, ok, err = process_fn(f, output_dir, gid)
                    except Exception as e:
                        outp, ok, err = None, False, str(e)
                    if not ok and len(gpus) > 1 and self.max_retries > 0:
                        alt = gpus[(list(result.frames_per_gpu).index(gid) + 1) % len(gpus)].id
                        try:
                            outp, ok, err = process_fn(f, output_dir, alt)
                        except Exception as e:
                            outp, ok, err = None, False, str(e)
                        with lock:
                            result.retried_frames.append(f)
                        if ok:
                            gid_done = alt
                        else:
                            gid_done = gid
                    else:
                        gid_done = gid
                    with lock:
                        if ok:
                            result.frames_per_gpu[gid_done].append(Path(outp) if outp is not None else f)
                        else:
                            result.errors[str(f)] = err or "Unknown error"
                        done += 1
                        if progress_callback:
                            progress_callback(done / len(frames), f"Processed {done}/{len(frames)} frames")

            threads = [threading.Thread(target=run, args=(gid, todo), daemon=True) for gid, todo in assignments.items()]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        else:
            ctx = mp.get_context("spawn")
            queue = ctx.Queue()
            procs = []
            for gid, todo in assignments.items():
                if not todo:
                    continue
                p = ctx.Process(target=_worker_entry,
                                args=(gid, [str(f) for f in todo], str(output_dir), self._engine_kwargs, queue))
                p.start()
                procs.append(p)
            for _ in procs:
                gid, res, fatal = queue.get()
                if fatal:
                    for f in assignments[gid]:
                        result.errors[str(f)] = fatal
                for src, (dst, ok, err) in zip(assignments[gid], res):
                    if ok:
                        result.frames_per_gpu[gid].append(Path(dst))
                    else:
                        result.errors[str(src)] = err or "Unknown error"
                done += len(assignments[gid])
                if progress_callback:
                    progress_callback(done / len(frames), f"Processed {done}/{len(frames)} frames")
            for p in procs:
                p.join()

        result.total_time = time.time() - start
        n = len(gpus)
        if n > 1 and result.total_frames > 0:
            mx = max(len(v) for v in result.frames_per_gpu.values())
            if mx > 0:
                result.speedup_factor = n * ((result.total_frames / n) / mx)
        for g in gpus:
            result.gpu_utilization[g.id] = g.utilization_pct
        logger.info(result.summary())
        self._result = result
        return result

    def get_result(self) -> Optional[DistributionResult]:
        return self._result


def distribute_frames(frames: List[Path], process_fn=None, output_dir: Path = Path("."),
                      strategy: LoadBalanceStrategy = LoadBalanceStrategy.ROUND_ROBIN, workers_per_gpu: int = 1,
                      progress_callback=None, **engine_kwargs) -> DistributionResult:
    """Convenience wrapper (reference :895-925)."""
    dist = MultiGPUDistributor(strategy=strategy, workers_per_gpu=workers_per_gpu, **engine_kwargs)
    return dist.distribute_frames(frames, process_fn, output_dir, progress_callback)
