"""Multi-GPU API of the upscaling path: the reference's `utils/multi_gpu.py` surface over the B200 scheduler.

API mirror of `/root/reference/src/framewright/utils/multi_gpu.py` (same names, arguments, result types):
    LoadBalanceStrategy (:30-37), GPUInfo (:39-93), DistributionResult (:95-137), WorkItem (:140-164),
    GPUManager (:166-427), WorkStealingQueue (:429-509),
    MultiGPUDistributor(...).distribute_frames(frames, process_fn, output_dir, progress_callback) (:511-778),
    _assign_frames (:780-869), detect_gpus / get_optimal_gpu / distribute_frames (:872-925),
    GPUSelector (:945-1043), MultiGPUManager (:1045-1281), list_gpus / select_gpu (:1283-1307).

Two execution paths behind `distribute_frames`:
  * `process_fn` given (the reference's contract: one call per frame, `(path, output_dir, gpu_id) -> (out, ok, err)`):
    `workers_per_gpu x n_gpus` threads over a `WorkStealingQueue` of `WorkItem`s, a failed item retried on a GPU
    not in its `failed_gpus` while `can_retry` -- the reference's semantics, for callers that bring their own function;
  * `process_fn` omitted: the product path.  Frames go through `scheduler.SchedulerPool` -- one persistent worker
    process and one engine per GPU, contiguous shards + tail stealing through a shared cursor table, frames batched
    through the engine's pipelined host path, per-frame completion callbacks, retry on another GPU, dead-worker
    detection.  `frames` may be a list of paths or any `scheduler.FrameSource` (shared-memory arrays, generators).

`shard_range` / `assign_contiguous` are the contiguous partition (rank r of G owns [r*N/G, (r+1)*N/G) -- the counts
`_assign_frames` ROUND_ROBIN produces, in contiguous order).
"""
from __future__ import annotations

import logging
import queue
import shutil
import subprocess
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from enum import Enum
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

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


@dataclass
class WorkItem:
    """One frame of a `process_fn` distribution (reference :140-164)."""
    frame_path: Path
    output_dir: Path
    priority: int = 0
    assigned_gpu: Optional[int] = None
    attempts: int = 0
    failed_gpus: List[int] = field(default_factory=list)

    @property
    def can_retry(self) -> bool:
        return self.attempts < 3


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
_SMI_QUERY = ("index,name,memory.total,memory.free,utilization.gpu,temperature.gpu,pcie.link.gen.current,"
              "pcie.link.width.current")


def _parse_smi(stdout: str, with_compute: bool = False) -> List[GPUInfo]:
    gpus = []
    for line in stdout.strip().split("\n"):
        if not line.strip():
            continue
        parts = [p.strip() for p in line.split(",")]
        try:
            info = GPUInfo(
                id=int(parts[0]), name=parts[1], total_vram_mb=int(parts[2]), free_vram_mb=int(parts[3]),
                utilization_pct=float(parts[4]) if parts[4] != "[N/A]" else 0.0,
                temperature_c=float(parts[5]) if parts[5] != "[N/A]" else None)
            if with_compute:
                info.compute_capability = parts[7] if len(parts) > 7 and parts[7] != "[N/A]" else None
            gpus.append(info)
        except (ValueError, IndexError) as e:
            logger.warning(f"Failed to parse GPU info: {e}")
    return gpus


def query_gpus() -> List[GPUInfo]:
    """Visible CUDA devices through torch (no nvidia-smi needed): ids are torch device indices."""
    try:
        import torch
    except ImportError:
        return []
    if not torch.cuda.is_available():
        return []
    infos = []
    for i in range(torch.cuda.device_count()):
        p = torch.cuda.get_device_properties(i)
        try:
            free, total = torch.cuda.mem_get_info(i)
        except Exception:
            free, total = p.total_memory, p.total_memory
        infos.append(GPUInfo(i, p.name, int(total // 2 ** 20), int(free // 2 ** 20), 0.0, None, None,
                             f"{p.major}.{p.minor}"))
    return infos


class GPUManager:
    """GPU detection, VRAM monitoring and health tracking (reference :166-427; nvidia-smi based, torch fallback)."""

    def __init__(self, gpu_ids: Optional[List[int]] = None, refresh_interval: float = 5.0):
        self._gpu_ids = gpu_ids
        self._refresh_interval = refresh_interval
        self._gpu_cache: Dict[int, GPUInfo] = {}
        self._last_refresh: float = 0
        self._lock = threading.Lock()
        self._monitoring = False
        self._monitor_thread: Optional[threading.Thread] = None
        self._refresh_gpu_info()

    @property
    def gpu_ids(self) -> List[int]:
        if self._gpu_ids is not None:
            return self._gpu_ids
        return list(self._gpu_cache.keys())

    @property
    def gpu_count(self) -> int:
        return len(self.gpu_ids)

    @property
    def is_multi_gpu(self) -> bool:
        return self.gpu_count > 1

    def _is_nvidia_smi_available(self) -> bool:
        return shutil.which("nvidia-smi") is not None

    def _smi(self, query: str, with_compute: bool = False) -> List[GPUInfo]:
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={query}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10)
            if r.returncode != 0:
                logger.error(f"nvidia-smi failed: {r.stderr}")
                return []
            gpus = _parse_smi(r.stdout, with_compute)
        except subprocess.TimeoutExpired:
            logger.warning("nvidia-smi timed out")
            return []
        except Exception as e:
            logger.warning(f"Failed to detect GPUs: {e}")
            return []
        return [g for g in gpus if self._gpu_ids is None or g.id in self._gpu_ids]

    def detect_gpus(self) -> List[GPUInfo]:
        if not self._is_nvidia_smi_available():
            gpus = [g for g in query_gpus() if self._gpu_ids is None or g.id in self._gpu_ids]
            if not gpus:
                logger.warning("nvidia-smi not available, no GPUs detected")
            return gpus
        return self._smi(_SMI_QUERY)

    def _refresh_gpu_info(self) -> None:
        gpus = self.detect_gpus()
        with self._lock:
            self._gpu_cache = {g.id: g for g in gpus}
            self._last_refresh = time.time()

    def _cache_stale(self) -> bool:
        return time.time() - self._last_refresh > self._refresh_interval

    def get_gpu_info(self, gpu_id: int, refresh: bool = False) -> Optional[GPUInfo]:
        if refresh or self._cache_stale():
            self._refresh_gpu_info()
        with self._lock:
            return self._gpu_cache.get(gpu_id)

    def get_all_gpu_info(self, refresh: bool = False) -> List[GPUInfo]:
        if refresh or self._cache_stale():
            self._refresh_gpu_info()
        with self._lock:
            return [self._gpu_cache[gid] for gid in self.gpu_ids if gid in self._gpu_cache]

    def get_optimal_gpu(self, strategy: LoadBalanceStrategy = LoadBalanceStrategy.VRAM_AWARE) -> int:
        gpus = self.get_all_gpu_info(refresh=True)
        if not gpus:
            return 0
        if strategy == LoadBalanceStrategy.LEAST_LOADED:
            gpus.sort(key=lambda g: g.utilization_pct)
        elif strategy == LoadBalanceStrategy.VRAM_AWARE:
            gpus.sort(key=lambda g: g.free_vram_mb, reverse=True)
        elif strategy == LoadBalanceStrategy.WEIGHTED:
            gpus.sort(key=lambda g: g.effective_capacity, reverse=True)
        return gpus[0].id

    def get_healthy_gpus(self) -> List[GPUInfo]:
        return [gpu for gpu in self.get_all_gpu_info(refresh=True) if gpu.is_healthy]

    def wait_for_vram(self, required_mb: int, gpu_id: Optional[int] = None, timeout: float = 60.0) -> bool:
        start_time = time.time()
        while time.time() - start_time < timeout:
            gpus = self.get_all_gpu_info(refresh=True)
            if gpu_id is not None:
                gpu = next((g for g in gpus if g.id == gpu_id), None)
                if gpu and gpu.free_vram_mb >= required_mb:
                    return True
            elif any(g.free_vram_mb >= required_mb for g in gpus):
                return True
            time.sleep(1.0)
        return False

    def start_monitoring(self) -> None:
        if self._monitoring:
            return
        self._monitoring = True
        self._monitor_thread = threading.Thread(target=self._monitor_loop, daemon=True, name="GPUMonitor")
        self._monitor_thread.start()

    def stop_monitoring(self) -> None:
        self._monitoring = False
        if self._monitor_thread:
            self._monitor_thread.join(timeout=5.0)
            self._monitor_thread = None

    def _monitor_loop(self) -> None:
        while self._monitoring:
            try:
                self._refresh_gpu_info()
                time.sleep(self._refresh_interval)
            except Exception as e:  # pragma: no cover
                logger.warning(f"GPU monitoring error: {e}")
                time.sleep(1.0)


class WorkStealingQueue:
    """Per-worker queues; an idle worker takes from another queue that still holds more than one item
    (reference :429-509)."""

    def __init__(self, num_workers: int):
        self._queues: Dict[int, queue.Queue] = {i: queue.Queue() for i in range(num_workers)}
        self._lock = threading.Lock()
        self._completed = 0
        self._total = 0

    def add_work(self, item: WorkItem, worker_id: int) -> None:
        with self._lock:
            self._queues[worker_id].put(item)
            self._total += 1

    def get_work(self, worker_id: int, timeout: float = 0.1) -> Optional[WorkItem]:
        try:
            return self._queues[worker_id].get(timeout=timeout)
        except queue.Empty:
            pass
        with self._lock:
            for other_id, q in self._queues.items():
                if other_id != worker_id and q.qsize() > 1:
                    try:
                        return q.get_nowait()
                    except queue.Empty:
                        continue
        return None

    def mark_complete(self) -> None:
        with self._lock:
            self._completed += 1

    @property
    def progress(self) -> float:
        with self._lock:
            return 0.0 if self._total == 0 else self._completed / self._total

    @property
    def is_complete(self) -> bool:
        with self._lock:
            # a re-queued (retried) item was counted twice in _total: complete means nothing queued and every
            # distinct item finished
            return all(q.empty() for q in self._queues.values()) and self._completed >= self._distinct

    _distinct = 0

    def set_distinct(self, n: int) -> None:
        self._distinct = n


# ------------------------------------------------------------------------------------------ distributor
class MultiGPUDistributor:
    """`distribute_frames` with the reference's signature and result type (module docstring: two paths)."""

    def __init__(self, gpu_manager: Optional[GPUManager] = None,
                 strategy: LoadBalanceStrategy = LoadBalanceStrategy.VRAM_AWARE, workers_per_gpu: int = 2,
                 max_retries: int = 2, enable_work_stealing: bool = True, gpus: Optional[List[GPUInfo]] = None,
                 batch: int = 2, **engine_kwargs):
        self.gpu_manager = gpu_manager
        self.strategy = strategy
        self.workers_per_gpu = workers_per_gpu
        self.max_retries = max_retries
        self.enable_work_stealing = enable_work_stealing
        self.batch = batch
        self._gpus = gpus
        self._engine_kwargs = engine_kwargs          # model_name, scale, tile, tile_pad, pre_pad (product path)
        self._result: Optional[DistributionResult] = None
        self._lock = threading.Lock()
        self._stop_event = threading.Event()
        self._pool = None
        self._pool_key = None

    # ---- GPUs and assignment
    def _healthy_gpus(self) -> List[GPUInfo]:
        if self._gpus is not None:
            return [g for g in self._gpus if g.is_healthy]
        if self.gpu_manager is None:
            self.gpu_manager = GPUManager()
        return list(self.gpu_manager.get_healthy_gpus())

    def _assign_frames(self, frames: List[Path], gpus: List[GPUInfo]) -> Dict[int, List[Path]]:
        return assign_frames(frames, gpus, self.strategy)

    def _shard_ranges(self, n: int, gpus: List[GPUInfo]) -> List[Tuple[int, int]]:
        """Contiguous shards whose SIZES follow the strategy (`_assign_frames` counts), in GPU order."""
        counts = [len(v) for v in assign_frames(list(range(n)), gpus, self.strategy).values()]
        ranges, lo = [], 0
        for c in counts:
            ranges.append((lo, lo + c))
            lo += c
        return ranges

    # ---- product path
    def _esr_config(self) -> Dict[str, Any]:
        kw = self._engine_kwargs
        return {"model_name": kw.get("model_name", "RealESRGAN_x4plus"),
                "scale_factor": int(kw.get("scale", kw.get("scale_factor", 4))),
                "tile_size": int(kw.get("tile", kw.get("tile_size", 0))), "tile_pad": int(kw.get("tile_pad", 10)),
                "pre_pad": int(kw.get("pre_pad", 0))}

    def _get_pool(self, gpu_ids: List[int]):
        from .scheduler import SchedulerPool

        key = (tuple(gpu_ids), self.workers_per_gpu)
        if self._pool is not None and (self._pool_key != key or set(self._pool.alive_gpus()) != set(gpu_ids)):
            self._pool.close()
            self._pool = None
        if self._pool is None:
            self._pool = SchedulerPool(gpu_ids, workers_per_gpu=max(2, self.workers_per_gpu + 1))
            self._pool_key = key
        return self._pool

    def close(self) -> None:
        """Stops the persistent worker processes (they otherwise live until this object is collected)."""
        if self._pool is not None:
            self._pool.close()
            self._pool = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def distribute_frames(self, frames, process_fn: Optional[Callable[[Path, Path, int], Tuple[Path, bool, Optional[str]]]] = None,
                          output_dir: Optional[Path] = None,
                          progress_callback: Optional[Callable[[float, str], None]] = None,
                          frame_callback: Optional[Callable[[int, str, bool, Optional[str], int], None]] = None,
                          sink=None, **hooks) -> DistributionResult:
        """frames: list of frame paths (reference) or a `scheduler.FrameSource`.  Without `process_fn` the frames run
        through the persistent per-GPU workers; results are written to `output_dir/<same name>` (or handed to `sink`).
        `frame_callback(index, name, ok, err, gpu_id)` fires as each frame completes."""
        from .scheduler import FrameSource, PathSource, PngSink

        n = len(frames)
        if n == 0:
            return DistributionResult()
        start = time.time()
        self._stop_event.clear()
        gpus = self._healthy_gpus()
        is_source = isinstance(frames, FrameSource) or (hasattr(frames, "load") and hasattr(frames, "name"))
        names = [frames.name(i) for i in range(n)] if is_source else [str(f) for f in frames]
        if not gpus:
            logger.error("No healthy GPUs available")
            return DistributionResult(errors={nm: "No GPUs available" for nm in names})
        gpu_ids = [g.id for g in gpus]
        if output_dir is not None:
            output_dir = Path(output_dir)
            output_dir.mkdir(parents=True, exist_ok=True)
        logger.info(f"Distributing {n} frames across {len(gpus)} GPUs (strategy={self.strategy.value})")
        if process_fn is not None:
            result = self._run_process_fn(list(frames), process_fn, output_dir or Path("."), gpus, progress_callback)
        else:
            source = frames if is_source else PathSource(frames)
            if sink is None:
                if output_dir is None:
                    raise ValueError("distribute_frames needs output_dir (or sink=) when process_fn is omitted")
                sink = PngSink(output_dir)
            pool = self._get_pool(gpu_ids)
            rr = pool.run(source, sink, self._esr_config(), batch=self.batch, progress_callback=progress_callback,
                          frame_callback=frame_callback, shard_ranges=self._shard_ranges(n, gpus),
                          steal=self.enable_work_stealing, **hooks)
            result = DistributionResult()
            result.frames_per_gpu = {gid: [] for gid in gpu_ids}
            for gid, idxs in rr.frames_per_gpu.items():
                for i in sorted(idxs):
                    info = rr.ok[i][1]
                    result.frames_per_gpu.setdefault(gid, []).append(
                        Path(info) if isinstance(info, str) else Path(names[i]))
            result.errors = {names[i]: e for i, e in rr.errors.items()}
            result.retried_frames = [Path(names[i]) for i in rr.retried]
            self.last_run = rr
        result.total_time = time.time() - start
        ngpu = len(gpus)
        if ngpu > 1 and result.total_frames > 0:
            mx = max(len(v) for v in result.frames_per_gpu.values())
            if mx > 0:
                result.speedup_factor = ngpu * ((result.total_frames / ngpu) / mx)
        for g in gpus:
            result.gpu_utilization[g.id] = g.utilization_pct
        logger.info(result.summary())
        self._result = result
        return result

    # ---- reference path: caller-supplied per-frame function, threads in this process
    def _run_process_fn(self, frames: List[Path], process_fn, output_dir: Path, gpus: List[GPUInfo],
                        progress_callback) -> DistributionResult:
        gpu_ids = [g.id for g in gpus]
        num_gpus = len(gpus)
        total_workers = num_gpus * max(1, self.workers_per_gpu)
        result = DistributionResult()
        result.frames_per_gpu = {gid: [] for gid in gpu_ids}
        wq = WorkStealingQueue(total_workers)
        wq.set_distinct(len(frames))
        steal = self.enable_work_stealing
        wid = 0
        for gpu_id, assigned in self._assign_frames(frames, gpus).items():
            for f in assigned:
                wq.add_work(WorkItem(frame_path=f, output_dir=output_dir, assigned_gpu=gpu_id), wid % total_workers)
                wid += 1
        state = {"completed": 0}
        errors: Dict[str, str] = {}
        retried: List[Path] = []
        lock = threading.Lock()

        def finish(msg: str) -> None:   # caller holds `lock`
            state["completed"] += 1
            wq.mark_complete()
            if progress_callback:
                progress_callback(state["completed"] / len(frames), msg)

        def worker_fn(w: int, gpu_id: int) -> None:
            while not self._stop_event.is_set():
                item = wq.get_work(w, timeout=0.05) if steal else None
                if not steal:
                    try:
                        item = wq._queues[w].get(timeout=0.05)
                    except queue.Empty:
                        item = None
                if item is None:
                    if wq.is_complete:
                        break
                    continue
                gpu = item.assigned_gpu if item.assigned_gpu is not None else gpu_id
                try:
                    out, ok, err = process_fn(item.frame_path, item.output_dir, gpu)
                except Exception as e:
                    with lock:
                        errors[str(item.frame_path)] = str(e)
                        finish(f"Error: {Path(item.frame_path).name}")
                    logger.error(f"Worker error processing {item.frame_path}: {e}")
                    continue
                with lock:
                    if ok:
                        result.frames_per_gpu[gpu].append(out)
                        finish(f"Processed {state['completed'] + 1}/{len(frames)} frames")
                        continue
                    item.attempts += 1
                    item.failed_gpus.append(gpu)
                    others = [g for g in gpu_ids if g not in item.failed_gpus]
                    if item.can_retry and len(item.failed_gpus) < num_gpus and others:
                        item.assigned_gpu = others[0]
                        retried.append(item.frame_path)
                        wq.add_work(item, w)
                        logger.warning(f"Retrying {Path(item.frame_path).name} on GPU {item.assigned_gpu}")
                        continue
                    errors[str(item.frame_path)] = err or "Unknown error"
                    finish(f"Error: {Path(item.frame_path).name}")

        with ThreadPoolExecutor(max_workers=total_workers) as ex:
            futs = [ex.submit(worker_fn, i, gpu_ids[i % num_gpus]) for i in range(total_workers)]
            for f in futs:
                try:
                    f.result(timeout=3600)
                except Exception as e:  # pragma: no cover
                    logger.error(f"Worker failed: {e}")
        result.errors = errors
        result.retried_frames = retried
        return result

    def stop(self) -> None:
        self._stop_event.set()

    def get_result(self) -> Optional[DistributionResult]:
        return self._result


def detect_gpus() -> List[GPUInfo]:
    return GPUManager().detect_gpus()


def get_optimal_gpu(strategy: LoadBalanceStrategy = LoadBalanceStrategy.VRAM_AWARE) -> int:
    return GPUManager().get_optimal_gpu(strategy)


def distribute_frames(frames, process_fn=None, output_dir: Optional[Path] = None,
                      strategy: LoadBalanceStrategy = LoadBalanceStrategy.VRAM_AWARE, workers_per_gpu: int = 2,
                      progress_callback=None, **engine_kwargs) -> DistributionResult:
    """Convenience wrapper (reference :895-925)."""
    dist = MultiGPUDistributor(strategy=strategy, workers_per_gpu=workers_per_gpu, **engine_kwargs)
    try:
        return dist.distribute_frames(frames, process_fn, output_dir, progress_callback)
    finally:
        dist.close()


def add_multi_gpu_config_fields() -> Dict[str, Any]:
    """Default multi-GPU configuration fields (reference :929-942)."""
    return {"enable_multi_gpu": False, "gpu_ids": None, "gpu_load_balance_strategy": "vram_aware",
            "workers_per_gpu": 2, "enable_work_stealing": True}


# ------------------------------------------------------------------------------------------ selection helpers
class GPUSelector:
    """Pick a GPU by index or by strategy (reference :945-1043)."""

    def __init__(self, gpu_manager: Optional[GPUManager] = None):
        self.gpu_manager = gpu_manager or GPUManager()

    def select_by_index(self, gpu_id: int) -> Optional[GPUInfo]:
        if gpu_id < 0:
            raise ValueError(f"Invalid GPU index: {gpu_id}. Must be non-negative.")
        gpu = self.gpu_manager.get_gpu_info(gpu_id, refresh=True)
        if gpu is None:
            raise ValueError(f"GPU {gpu_id} not found. Available GPUs: {self.gpu_manager.gpu_ids}")
        return gpu

    def select_best(self, strategy: LoadBalanceStrategy = LoadBalanceStrategy.VRAM_AWARE) -> int:
        return self.gpu_manager.get_optimal_gpu(strategy)

    def validate_gpu(self, gpu_id: int) -> bool:
        gpu = self.gpu_manager.get_gpu_info(gpu_id, refresh=True)
        return gpu is not None and gpu.is_healthy

    def get_available_gpus(self) -> List[GPUInfo]:
        return self.gpu_manager.get_all_gpu_info(refresh=True)

    def get_gpu_for_task(self, gpu_id: Optional[int] = None, multi_gpu: bool = False) -> Tuple[int, bool]:
        if gpu_id is not None:
            if not self.validate_gpu(gpu_id):
                logger.warning(f"Specified GPU {gpu_id} not available, falling back to auto-select")
                return self.select_best(), multi_gpu
            return gpu_id, multi_gpu
        return self.select_best(), bool(multi_gpu)


class MultiGPUManager:
    """GPUManager + compute capability, measured per-GPU speeds and the speed-proportional split (reference
    :1045-1281).  `get_dynamic_distribution` feeds `MultiGPUDistributor` shard sizes when GPUs differ."""

    def __init__(self, gpu_ids: Optional[List[int]] = None, refresh_interval: float = 5.0):
        self._base_manager = GPUManager(gpu_ids=gpu_ids, refresh_interval=refresh_interval)
        self._processing_speeds: Dict[int, float] = {}
        self._lock = threading.Lock()

    @property
    def gpu_ids(self) -> List[int]:
        return self._base_manager.gpu_ids

    @property
    def gpu_count(self) -> int:
        return self._base_manager.gpu_count

    @property
    def is_multi_gpu(self) -> bool:
        return self._base_manager.is_multi_gpu

    def detect_gpus_with_compute(self) -> List[GPUInfo]:
        if not self._base_manager._is_nvidia_smi_available():
            gpus = [g for g in query_gpus() if self._base_manager._gpu_ids is None or g.id in self._base_manager._gpu_ids]
            if not gpus:
                logger.warning("nvidia-smi not available, no GPUs detected")
            return gpus
        return self._base_manager._smi("index,name,memory.total,memory.free,utilization.gpu,temperature.gpu,"
                                       "pcie.link.gen.current,compute_cap", with_compute=True)

    def get_gpu_info(self, gpu_id: int) -> Optional[GPUInfo]:
        return self._base_manager.get_gpu_info(gpu_id, refresh=True)

    def get_all_gpu_info(self) -> List[GPUInfo]:
        return self._base_manager.get_all_gpu_info(refresh=True)

    def get_healthy_gpus(self) -> List[GPUInfo]:
        return self._base_manager.get_healthy_gpus()

    def update_processing_speed(self, gpu_id: int, frames_per_second: float) -> None:
        with self._lock:
            self._processing_speeds[gpu_id] = frames_per_second

    def get_dynamic_distribution(self, total_frames: int) -> Dict[int, int]:
        gpus = self.get_healthy_gpus()
        if not gpus:
            return {}
        with self._lock:
            speeds = self._processing_speeds.copy()
        ids = [g.id for g in gpus]
        weights: Dict[int, float]
        if speeds and all(i in speeds for i in ids) and sum(speeds.values()) > 0:
            weights = {i: speeds[i] for i in ids}
        elif sum(g.free_vram_mb for g in gpus) > 0:
            weights = {g.id: float(g.free_vram_mb) for g in gpus}
        else:
            weights = {i: 1.0 for i in ids}
        total = sum(weights.values())
        dist, assigned = {}, 0
        for i in ids[:-1]:
            dist[i] = int(total_frames * (weights[i] / total))
            assigned += dist[i]
        dist[ids[-1]] = total_frames - assigned
        return dist

    def format_gpu_table(self) -> str:
        gpus = self.detect_gpus_with_compute()
        if not gpus:
            return "No GPUs detected. Ensure NVIDIA drivers are installed."
        bar = "+------+---------------------------+---------+------------+--------+"
        lines = ["Available GPUs:", bar, "| ID   | Name                      | Memory  | Compute    | Status |", bar]
        for gpu in gpus:
            status = "Hot" if not gpu.is_healthy else ("Busy" if gpu.utilization_pct > 90 else "Ready")
            lines.append(f"| {gpu.id:<4} | {gpu.name:<25} | {f'{gpu.total_vram_mb / 1024:.0f} GB':<7} | "
                         f"{(gpu.compute_capability or 'N/A'):<10} | {status:<6} |")
        lines.append(bar)
        return "\n".join(lines)


def list_gpus() -> str:
    return MultiGPUManager().format_gpu_table()


def select_gpu(gpu_id: Optional[int] = None, multi_gpu: bool = False) -> Tuple[int, bool]:
    return GPUSelector().get_gpu_for_task(gpu_id, multi_gpu)
