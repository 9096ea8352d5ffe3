"""Frame-sharded multi-GPU scheduler: one persistent worker PROCESS per GPU, each holding one B200 engine.

This is the B200-first replacement for what the reference's `MultiGPUDistributor.distribute_frames`
(`/root/reference/src/framewright/utils/multi_gpu.py:549-770`) does with `workers_per_gpu x n_gpus` Python threads in
one process and one `process_fn` call (= one PNG read, one forward, one PNG write) per frame.  Frames are independent
(RRDBNet / SRVGG have no cross-frame state), so the path shards with no exchange step and no collective:

  * workers are spawned once (`SchedulerPool`), import torch and build their engine once, and serve any number of jobs;
    they take their GPU by index (nothing rewrites CUDA_VISIBLE_DEVICES);
  * a job's frame indices are cut into contiguous shards, one per GPU (`multi_gpu.shard_range`); every worker takes
    batches from the front of its own shard and, when that is empty, STEALS from the back of the shard with the most
    work left -- all through one small shared cursor table (the reference's analogue: `WorkStealingQueue`, :429-509);
  * inside a worker `workers_per_gpu` runner threads each loop claim -> load -> `enhance_batch` -> store; the engine's
    lanes (csrc/b200sr.cu) let two of them overlap copies and launch tails, the others decode / encode meanwhile;
  * frames reach the workers by index: files (`PathSource`), a shared-memory array (`SharedArray`, frame-array in /
    frame-array out with no PNG anywhere), a bounded shared-memory ring fed in order by the parent (`stream`: raw
    video pipes, `StreamingPipeline`), or a picklable generator (`bench.py --workload clip2000`);
  * every frame's completion is reported to the parent as it happens (`frame_callback(index, name, ok, err, gpu)` --
    what `CheckpointManager.update_frame` needs -- and `progress_callback(fraction, message)` per frame); a failed
    frame is retried on a GPU that has not failed it, at most 3 attempts in all (`WorkItem.can_retry`, :140-164); a
    worker that dies (native crash, OOM killer, GPU fault) is noticed by liveness polling, its claimed frames are
    retried elsewhere and its shard is stolen by the others -- `run` never hangs on a dead worker.
"""
from __future__ import annotations

import logging
import multiprocessing as mp
import os
import queue as queue_mod
import threading
import time
import traceback
from dataclasses import dataclass, field
from multiprocessing import shared_memory
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)

MAX_ATTEMPTS = 3          # WorkItem.can_retry: attempts < 3


# ------------------------------------------------------------------------------------------ sources and sinks
class FrameSource:
    """Frames by index.  Must be picklable (it travels to the worker processes)."""

    def __len__(self) -> int:  # pragma: no cover - interface
        raise NotImplementedError

    def name(self, i: int) -> str:
        return f"frame_{i + 1:08d}.png"

    def load(self, i: int) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def open(self) -> None:
        """Called once per worker process before the first load."""

    def close(self) -> None:
        pass


class PathSource(FrameSource):
    """Frame files, decoded in the worker (cv2.imread IMREAD_UNCHANGED, as pytorch_realesrgan.py:198)."""

    def __init__(self, paths: Sequence):
        self.paths = [str(p) for p in paths]

    def __len__(self) -> int:
        return len(self.paths)

    def name(self, i: int) -> str:
        return os.path.basename(self.paths[i])

    def load(self, i: int) -> np.ndarray:
        import cv2

        img = cv2.imread(self.paths[i], cv2.IMREAD_UNCHANGED)
        if img is None:
            raise IOError(f"Failed to read image: {self.paths[i]}")
        return img


class SharedArray:
    """A numpy array in POSIX shared memory, addressable from every worker; page-locked in processes that call
    `pin()` so the engine copies to / from it without staging (C ABI b200sr_host_register)."""

    def __init__(self, shape: Tuple[int, ...], dtype=np.uint8, name: Optional[str] = None, create: bool = True):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype).str
        nbytes = max(1, int(np.prod(self.shape)) * np.dtype(dtype).itemsize)
        self._shm = shared_memory.SharedMemory(create=create, size=nbytes, name=name) if create else \
            shared_memory.SharedMemory(name=name)
        self.name_ = self._shm.name
        self._owner = create
        self._pinned = False

    @property
    def array(self) -> np.ndarray:
        return np.ndarray(self.shape, dtype=np.dtype(self.dtype), buffer=self._shm.buf)

    def __getstate__(self):
        return {"shape": self.shape, "dtype": self.dtype, "name_": self.name_}

    def __setstate__(self, st):
        self.shape, self.dtype, self.name_ = st["shape"], st["dtype"], st["name_"]
        self._shm = shared_memory.SharedMemory(name=self.name_)
        self._owner = False        # (spawned workers share the parent's resource tracker: attaching does not unlink)
        self._pinned = False

    def pin(self) -> bool:
        if self._pinned:
            return True
        try:
            import ctypes

            from . import _native

            a = self.array
            rc = _native.load().b200sr_host_register(ctypes.c_void_p(a.ctypes.data), a.nbytes)
            self._pinned = rc == 0
        except Exception:
            self._pinned = False
        return self._pinned

    def unpin(self) -> None:
        if self._pinned:
            import ctypes

            from . import _native

            _native.load().b200sr_host_unregister(ctypes.c_void_p(self.array.ctypes.data))
            self._pinned = False

    def release(self) -> None:
        """Un-pin, unmap, and (creator only) unlink.  Idempotent."""
        if self._shm is None:
            return
        try:
            self.unpin()
        except Exception:
            pass
        try:
            self._shm.close()
            if self._owner:
                self._shm.unlink()
        except Exception:
            pass
        self._shm = None

    def __del__(self):  # pragma: no cover - safety net: never leave a registration behind an unmapped segment
        try:
            self.release()
        except Exception:
            pass


class ArraySource(FrameSource):
    """Frames [N,H,W,3] uint8 in a `SharedArray`.  `avail` (optional shared counter) bounds what may be claimed:
    the ring mode of `SchedulerPool.stream` writes slot i % S before raising it past i."""

    def __init__(self, shared: SharedArray, num_frames: Optional[int] = None, names: Optional[List[str]] = None):
        self.shared = shared
        self.n = int(num_frames if num_frames is not None else shared.shape[0])
        self.names = names

    def __len__(self) -> int:
        return self.n

    def name(self, i: int) -> str:
        return self.names[i] if self.names else super().name(i)

    def open(self) -> None:
        self.shared.pin()

    def close(self) -> None:
        # in the worker: un-pin BEFORE the mapping goes away -- a page-locked registration that outlives its mapping
        # would make a later segment mapped at the same address look pinned and DMA into freed pages
        if not self.shared._owner:
            self.shared.release()

    def slot(self, i: int) -> int:
        return i % self.shared.shape[0]

    def load(self, i: int) -> np.ndarray:
        return self.shared.array[self.slot(i)]


class FrameSink:
    """Where results go.  `store` runs in the worker; what it returns travels back with the completion message."""

    def open(self) -> None:
        pass

    def out_block(self, i0: int, k: int, shape: Tuple[int, ...]) -> Optional[np.ndarray]:
        """A (page-locked) [k, *shape] destination the engine may write the results of frames i0 .. i0+k-1 into
        directly, or None."""
        return None

    def store(self, i: int, name: str, out: np.ndarray) -> Any:  # pragma: no cover - interface
        raise NotImplementedError

    def close(self) -> None:
        pass


class PngSink(FrameSink):
    """`output_dir/<same name>` through cv2.imwrite (pytorch_realesrgan.py:230)."""

    def __init__(self, output_dir):
        self.output_dir = str(output_dir)

    def open(self) -> None:
        os.makedirs(self.output_dir, exist_ok=True)

    def store(self, i: int, name: str, out: np.ndarray) -> Any:
        import cv2

        dst = os.path.join(self.output_dir, name)
        if not cv2.imwrite(dst, out) or not os.path.exists(dst):
            raise IOError("Output file was not created")
        return dst


class ArraySink(FrameSink):
    """Results into a `SharedArray` [S,sH,sW,3] (slot = index % S): the engine's D2H copy lands there directly."""

    def __init__(self, shared: SharedArray):
        self.shared = shared

    def open(self) -> None:
        self.shared.pin()

    def close(self) -> None:
        if not self.shared._owner:      # worker side: un-pin, then unmap (see ArraySource.close)
            self.shared.release()

    def out_block(self, i0: int, k: int, shape: Tuple[int, ...]) -> Optional[np.ndarray]:
        a = self.shared.array
        s0 = i0 % a.shape[0]
        if tuple(a.shape[1:]) != tuple(shape) or s0 + k > a.shape[0]:
            return None
        return a[s0:s0 + k]

    def store(self, i: int, name: str, out: np.ndarray) -> Any:
        a = self.shared.array
        dst = a[i % a.shape[0]]
        if out.ctypes.data != dst.ctypes.data:
            dst[...] = out
        return None


class ChecksumSink(FrameSink):
    """Discards the pixels, returns a strided checksum (benchmarks: the result still crosses PCIe into host memory)."""

    def store(self, i: int, name: str, out: np.ndarray) -> Any:
        return int(out[::61, ::67].astype(np.int64).sum())


# ------------------------------------------------------------------------------------------ shared claim table
class ClaimTable:
    """[lo_g, hi_g) per shard in shared memory + `avail`, the exclusive upper bound of claimable indices."""

    CAP = 64   # indices one worker may hold at a time (runner threads x batch)

    def __init__(self, ctx, num_shards: int):
        self.num_shards = num_shards
        self.lock = ctx.Lock()
        self.cur = ctx.Array("q", 2 * num_shards, lock=False)
        self.avail = ctx.Value("q", 0, lock=False)
        # what every worker currently holds, in shared memory: if a worker dies, the parent reads its row here --
        # nothing depends on a message the dead process may never have flushed
        self.held = ctx.Array("q", num_shards * self.CAP, lock=False)

    def reset(self, ranges: Sequence[Tuple[int, int]], avail: int) -> None:
        with self.lock:
            for g in range(self.num_shards):
                lo, hi = ranges[g] if g < len(ranges) else (0, 0)
                self.cur[2 * g], self.cur[2 * g + 1] = lo, hi
            self.avail.value = avail
            for k in range(self.num_shards * self.CAP):
                self.held[k] = -1

    def _hold(self, shard: int, idxs: List[int]) -> None:   # caller holds the lock
        base, j = shard * self.CAP, 0
        for i in idxs:
            while j < self.CAP and self.held[base + j] >= 0:
                j += 1
            if j < self.CAP:
                self.held[base + j] = i

    def hold(self, shard: int, idxs: List[int]) -> None:
        with self.lock:
            self._hold(shard, idxs)

    def drop(self, shard: int, i: int) -> None:
        with self.lock:
            base = shard * self.CAP
            for j in range(self.CAP):
                if self.held[base + j] == i:
                    self.held[base + j] = -1
                    return

    def held_by(self, shard: int) -> List[int]:
        with self.lock:
            base = shard * self.CAP
            return [int(self.held[base + j]) for j in range(self.CAP) if self.held[base + j] >= 0]

    def set_avail(self, avail: int) -> None:
        with self.lock:
            self.avail.value = avail

    def remaining(self) -> int:
        with self.lock:
            return sum(max(0, self.cur[2 * g + 1] - self.cur[2 * g]) for g in range(self.num_shards))

    def claim(self, shard: int, batch: int, steal: bool = True, front: Optional[int] = None) -> List[int]:
        """Up to `batch` consecutive indices: from the front of the own shard, else from the back of the fullest one.
        `front` names the shard whose FRONT this worker draws from instead of its own (ring mode: every GPU takes the
        next chunk of the one ordered stream); what it takes is still recorded in its own row of the held-table."""
        with self.lock:
            av = self.avail.value
            src = shard if front is None else front
            lo, hi = self.cur[2 * src], min(self.cur[2 * src + 1], av)
            if lo < hi:
                k = min(batch, hi - lo)
                self.cur[2 * src] = lo + k
                idxs = list(range(lo, lo + k))
                self._hold(shard, idxs)
                return idxs
            if not steal:
                return []
            best, best_left = -1, 0
            for g in range(self.num_shards):
                left = min(self.cur[2 * g + 1], av) - self.cur[2 * g]
                if g != shard and left > best_left and self.cur[2 * g + 1] <= av:
                    best, best_left = g, left
            if best < 0:
                return []
            k = min(batch, best_left)
            hi_v = self.cur[2 * best + 1]
            self.cur[2 * best + 1] = hi_v - k
            idxs = list(range(hi_v - k, hi_v))
            self._hold(shard, idxs)
            return idxs


# ------------------------------------------------------------------------------------------ worker process
@dataclass
class JobSpec:
    job_id: int
    source: FrameSource
    sink: FrameSink
    config: Dict[str, Any]            # PyTorchESRGANConfig fields
    batch: int = 2
    shard_of_gpu: Dict[int, int] = field(default_factory=dict)
    steal: bool = True
    front_shard: Optional[int] = None      # ring mode: every worker claims from the front of this shard
    fail_on_gpus: Tuple[int, ...] = ()     # test hook: these GPUs report every frame as failed
    crash_on_gpus: Tuple[int, ...] = ()    # test hook: these workers die (os._exit) after their first claim
    engine_factory: Optional[Callable[[Dict[str, Any]], Any]] = None   # test hook: object with enhance_batch / enhance


def _default_engine(config: Dict[str, Any]):
    from .pytorch_realesrgan import PyTorchESRGANConfig, get_upsampler

    cfg = PyTorchESRGANConfig(**config)
    cfg.validate()
    return get_upsampler(cfg)


class _Reporter:
    """results.put + release of the index in the shared held-table (so a later death does not re-run it)."""

    def __init__(self, results, claims: "ClaimTable", shard: int):
        self.results, self.claims, self.shard = results, claims, shard

    def put(self, msg) -> None:
        self.results.put(msg)
        if msg[0] == "done":
            self.claims.drop(self.shard, msg[2])


def _run_batch(up, spec: JobSpec, idxs: List[int], gpu_id: int, results) -> None:
    """load -> enhance -> store for one claimed batch; every index gets exactly one ('done', ...) message."""
    src, sink = spec.source, spec.sink
    frames: List[Tuple[int, np.ndarray]] = []
    for i in idxs:
        try:
            if gpu_id in spec.fail_on_gpus:
                raise RuntimeError(f"injected failure on GPU {gpu_id}")
            frames.append((i, src.load(i)))
        except Exception as e:
            results.put(("done", spec.job_id, i, False, str(e), gpu_id, None))
    # same-size 3-channel uint8 frames run as one batch (one launch sequence); anything else (gray, alpha, 16-bit,
    # odd sizes) goes through `enhance` frame by frame
    groups: Dict[Tuple, List[Tuple[int, np.ndarray]]] = {}
    for i, f in frames:
        key = (f.shape, f.dtype.str) if (f.ndim == 3 and f.shape[2] == 3 and f.dtype == np.uint8) else ("single", i)
        groups.setdefault(key, []).append((i, f))
    for key, members in groups.items():
        try:
            if key[0] == "single":
                outs = [up.enhance(members[0][1], outscale=spec.config.get("scale_factor", 4))[0]]
            else:
                stack = np.stack([f for _, f in members]) if len(members) > 1 else members[0][1][None]
                scale = int(getattr(up, "scale", spec.config.get("scale_factor", 4)))
                h, w = stack.shape[1] * scale, stack.shape[2] * scale
                ids = [i for i, _ in members]
                out_arr = sink.out_block(ids[0], len(ids), (h, w, 3)) if ids == list(range(ids[0], ids[0] + len(ids))) \
                    else None
                outs = up.enhance_batch(stack, out=out_arr) if out_arr is not None else up.enhance_batch(stack)
        except Exception as e:
            msg = str(e)
            for i, _ in members:
                results.put(("done", spec.job_id, i, False, msg, gpu_id, None))
            continue
        for (i, _), o in zip(members, outs):
            try:
                info = sink.store(i, src.name(i), o)
                results.put(("done", spec.job_id, i, True, None, gpu_id, info))
            except Exception as e:
                results.put(("done", spec.job_id, i, False, str(e), gpu_id, None))
        del outs


def _worker_main(gpu_id: int, cmd_q, results, retry_q, claims: ClaimTable, job_done, threads_per_worker: int) -> None:
    """Worker process: owns GPU `gpu_id` (by index) for its whole life."""
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.set_device(gpu_id)
    except Exception:
        pass
    results.put(("ready", gpu_id, os.getpid()))
    engines: Dict[str, Any] = {}
    while True:
        msg = cmd_q.get()
        if msg is None or msg[0] == "stop":
            break
        spec: JobSpec = msg[1]
        try:
            key = repr(sorted(spec.config.items())) + repr(spec.engine_factory)
            if key not in engines:
                cfg = dict(spec.config, gpu_id=gpu_id)
                engines[key] = (spec.engine_factory or _default_engine)(cfg)
            up = engines[key]
            spec.source.open()
            spec.sink.open()
        except Exception as e:
            results.put(("worker_error", spec.job_id, gpu_id, f"{type(e).__name__}: {e}"))
            job_done.wait()
            continue
        shard = spec.shard_of_gpu.get(gpu_id, 0)
        crashed = threading.Event()
        reporter = _Reporter(results, claims, shard)

        def runner():
            while not job_done.is_set():
                idxs: List[int] = []
                try:   # frames another GPU failed come first (they are the oldest work)
                    i, failed = retry_q.get_nowait()
                    if gpu_id in failed:
                        retry_q.put((i, failed))
                        time.sleep(0.002)
                    else:
                        idxs = [i]
                        claims.hold(shard, idxs)
                except queue_mod.Empty:
                    pass
                if not idxs:
                    idxs = claims.claim(shard, spec.batch, steal=spec.steal, front=spec.front_shard)
                if not idxs:
                    time.sleep(0.002)
                    continue
                results.put(("claim", spec.job_id, gpu_id, idxs))
                if gpu_id in spec.crash_on_gpus:
                    crashed.set()
                    os._exit(17)           # (no flush: the parent must cope with a lost claim message)
                try:
                    _run_batch(up, spec, idxs, gpu_id, reporter)
                except Exception as e:  # pragma: no cover - _run_batch reports per frame
                    for i in idxs:
                        reporter.put(("done", spec.job_id, i, False, f"{type(e).__name__}: {e}", gpu_id, None))
                    logger.error("runner failed: %s", traceback.format_exc())

        ts = [threading.Thread(target=runner, daemon=True) for _ in range(max(1, threads_per_worker))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        try:
            spec.source.close()
            spec.sink.close()
        except Exception:
            pass
        results.put(("job_exit", spec.job_id, gpu_id))
    for up in engines.values():
        try:
            up.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------ parent side
@dataclass
class RunResult:
    """Per-frame outcome of one job (the distributor turns it into a `DistributionResult`)."""
    ok: Dict[int, Tuple[int, Any]] = field(default_factory=dict)        # index -> (gpu_id, sink info)
    errors: Dict[int, str] = field(default_factory=dict)                 # index -> last error
    retried: List[int] = field(default_factory=list)
    frames_per_gpu: Dict[int, List[int]] = field(default_factory=dict)
    stolen: int = 0
    total_time: float = 0.0
    dead_gpus: List[int] = field(default_factory=list)


class SchedulerPool:
    """One persistent worker process per GPU.  `run` pushes one job (a frame source + sink) through them."""

    def __init__(self, gpu_ids: Sequence[int], workers_per_gpu: int = 3, start_timeout: float = 300.0,
                 stall_timeout: float = 900.0):
        if not gpu_ids:
            raise ValueError("SchedulerPool needs at least one GPU id")
        self.gpu_ids = [int(g) for g in gpu_ids]
        self.workers_per_gpu = max(1, int(workers_per_gpu))
        # a worker that is alive but silent for this long while it holds frames (a hung GPU) is terminated and treated
        # like a dead one (the reference bounds every future with a timeout, utils/multi_gpu.py:741)
        self.stall_timeout = float(stall_timeout)
        self._ctx = mp.get_context("spawn")
        self._results = self._ctx.Queue()
        self._retry = self._ctx.Queue()
        self._claims = ClaimTable(self._ctx, len(self.gpu_ids))
        self._job_done = self._ctx.Event()
        self._cmd: Dict[int, Any] = {}
        self._procs: Dict[int, Any] = {}
        self._job_counter = 0
        self._lock = threading.Lock()
        self._closed = False
        for g in self.gpu_ids:
            q = self._ctx.Queue()
            p = self._ctx.Process(target=_worker_main, name=f"b200sr-gpu{g}", daemon=True,
                                  args=(g, q, self._results, self._retry, self._claims, self._job_done,
                                        self.workers_per_gpu))
            p.start()
            self._cmd[g], self._procs[g] = q, p
        ready, deadline = set(), time.time() + start_timeout
        while len(ready) < len(self.gpu_ids):
            try:
                m = self._results.get(timeout=0.5)
                if m[0] == "ready":
                    ready.add(m[1])
            except queue_mod.Empty:
                dead = [g for g, p in self._procs.items() if not p.is_alive() and g not in ready]
                if dead or time.time() > deadline:
                    self.close()
                    raise RuntimeError(f"scheduler workers failed to start (dead: {dead})")

    # ---- liveness
    def alive_gpus(self) -> List[int]:
        return [g for g, p in self._procs.items() if p.is_alive()]

    def run(self, source: FrameSource, sink: FrameSink, config: Dict[str, Any], batch: int = 2,
            progress_callback: Optional[Callable[[float, str], None]] = None,
            frame_callback: Optional[Callable[[int, str, bool, Optional[str], int], None]] = None,
            shard_ranges: Optional[Sequence[Tuple[int, int]]] = None, steal: bool = True,
            avail: Optional[int] = None, feeder: Optional[Callable[["SchedulerPool", Dict[str, Any]], None]] = None,
            front_shard: Optional[int] = None, **hooks) -> RunResult:
        """Runs one job to completion.  `shard_ranges[k]` is the contiguous index range GPU k starts from (default:
        `multi_gpu.shard_range`); `avail` + `feeder` + `front_shard` implement the ordered ring mode of `stream`."""
        from .multi_gpu import shard_range

        with self._lock:
            if self._closed:
                raise RuntimeError("scheduler pool is closed")
            n = len(source)
            res = RunResult(frames_per_gpu={g: [] for g in self.gpu_ids})
            if n == 0:
                return res
            t0 = time.time()
            alive = self.alive_gpus()
            if not alive:
                res.errors = {i: "No GPUs available" for i in range(n)}
                return res
            self._job_counter += 1
            job_id = self._job_counter
            if shard_ranges is None:
                shard_ranges = [shard_range(n, len(alive), k) for k in range(len(alive))]
            shard_of_gpu = {g: k for k, g in enumerate(alive)}
            owner_of = {}
            for k, (lo, hi) in enumerate(shard_ranges if front_shard is None else []):   # (ring mode: no owners)
                for i in range(lo, hi):
                    owner_of[i] = alive[k] if k < len(alive) else alive[0]
            self._claims.reset(list(shard_ranges), n if avail is None else avail)
            self._job_done.clear()
            while True:   # drop stale retry entries
                try:
                    self._retry.get_nowait()
                except queue_mod.Empty:
                    break
            spec = JobSpec(job_id, source, sink, dict(config), int(batch), shard_of_gpu, steal, front_shard, **hooks)
            for g in alive:
                self._cmd[g].put(("job", spec))
            in_flight: Dict[int, set] = {g: set() for g in self.gpu_ids}
            attempts: Dict[int, int] = {}
            failed_on: Dict[int, set] = {}
            finished, done_count, dead_seen = set(), 0, set()
            state = {"done": 0, "n": n, "results": res, "finished": finished}
            feeder_thread = None
            if feeder is not None:
                feeder_thread = threading.Thread(target=feeder, args=(self, state), daemon=True)
                feeder_thread.start()

            def final(i: int, ok: bool, err: Optional[str], gpu: int, info: Any) -> None:
                nonlocal done_count
                if i in finished:
                    return
                finished.add(i)
                done_count += 1
                state["done"] = done_count
                if ok:
                    res.ok[i] = (gpu, info)
                    res.frames_per_gpu.setdefault(gpu, []).append(i)
                    if owner_of.get(i, gpu) != gpu:
                        res.stolen += 1
                else:
                    res.errors[i] = err or "Unknown error"
                if frame_callback:
                    frame_callback(i, source.name(i), ok, None if ok else res.errors[i], gpu)
                if progress_callback:
                    progress_callback(done_count / n, f"Processed {done_count}/{n} frames" if ok
                                      else f"Error: {source.name(i)}")

            def failed(i: int, err: str, gpu: int) -> None:
                attempts[i] = attempts.get(i, 0) + 1
                failed_on.setdefault(i, set()).add(gpu)
                candidates = [g for g in self.alive_gpus() if g not in failed_on[i]]
                if attempts[i] < MAX_ATTEMPTS and candidates:
                    res.retried.append(i)
                    logger.warning("Retrying %s on GPU %s", source.name(i), candidates[0])
                    self._retry.put((i, tuple(sorted(failed_on[i]))))
                else:
                    final(i, False, err, gpu, None)

            last_heard = {g: time.time() for g in self.gpu_ids}
            try:
                self._run_loop(n, job_id, res, state, in_flight, finished, dead_seen, last_heard, shard_of_gpu, final,
                               failed, lambda: done_count)
            finally:
                # whatever happened (also an exception from a caller's callback): end the job in the workers, so that
                # the pool stays usable and no runner thread keeps claiming frames of a job nobody waits for
                self._job_done.set()
                if feeder_thread is not None:
                    feeder_thread.join(timeout=5)
                exited, deadline = set(), time.time() + 30
                want = set(g for g in alive if self._procs[g].is_alive() and g not in res.dead_gpus) | \
                    set(g for g in res.dead_gpus if self._procs[g].is_alive())
                while exited < want and time.time() < deadline:
                    try:
                        m = self._results.get(timeout=0.25)
                        if m[0] == "job_exit" and m[1] == job_id:
                            exited.add(m[2])
                    except queue_mod.Empty:
                        want = set(g for g in want if self._procs[g].is_alive())
                res.total_time = time.time() - t0
            return res

    def _run_loop(self, n, job_id, res, state, in_flight, finished, dead_seen, last_heard, shard_of_gpu, final, failed,
                  done) -> None:
        """The parent's side of one job: completion messages, hung / dead workers, an input that ended early."""
        while done() < n:
            try:
                m = self._results.get(timeout=0.25)
            except queue_mod.Empty:
                m = None
            if m is not None and len(m) > 2 and m[1] == job_id and m[0] in ("claim", "worker_error"):
                last_heard[m[2]] = time.time()
            elif m is not None and m[0] == "done" and m[1] == job_id:
                last_heard[m[5]] = time.time()
            for g, p in self._procs.items():   # hung worker: alive, holding frames, silent for too long
                if (g not in dead_seen and p.is_alive() and time.time() - last_heard[g] > self.stall_timeout
                        and (in_flight[g] or (g in shard_of_gpu and self._claims.held_by(shard_of_gpu[g])))):
                    logger.error("worker for GPU %s is silent for %.0f s: terminating it", g, self.stall_timeout)
                    p.terminate()
                    p.join(timeout=10)
            if m is not None and len(m) > 1 and m[1] == job_id:
                if m[0] == "claim":
                    in_flight[m[2]].update(i for i in m[3] if i not in finished)
                elif m[0] == "done":
                    _, _, i, ok, err, gpu, info = m
                    in_flight[gpu].discard(i)
                    if ok:
                        final(i, True, None, gpu, info)
                    elif i not in finished:
                        failed(i, err or "Unknown error", gpu)
                elif m[0] == "worker_error":
                    _, _, gpu, err = m
                    logger.error("GPU %s cannot run the job: %s", gpu, err)
                    dead_seen.add(gpu)
                    res.dead_gpus.append(gpu)
                    state["last_error"] = err
            # liveness: a dead worker never reports -- fail what it had claimed (retried elsewhere)
            for g, p in self._procs.items():
                if g not in dead_seen and not p.is_alive():
                    dead_seen.add(g)
                    res.dead_gpus.append(g)
                    logger.error("worker for GPU %s died (exit code %s)", g, p.exitcode)
                    lost = set(in_flight[g])
                    if g in shard_of_gpu:
                        lost.update(self._claims.held_by(shard_of_gpu[g]))
                    for i in sorted(lost):
                        if i not in finished:
                            failed(i, f"worker for GPU {g} died (exit code {p.exitcode})", g)
                    in_flight[g].clear()
            usable = [g for g in self.alive_gpus() if g not in dead_seen]
            if not usable:
                why = state.get("last_error") or "No GPUs available"
                for i in range(n):
                    if i not in finished:
                        final(i, False, why, -1, None)
            # ring mode: the input ended (or failed) before `num_frames` frames -- the rest can never arrive
            cut = state.get("truncate")
            if cut is not None and not state.get("truncated"):
                state["truncated"] = True
                logger.warning(cut[1])
                for i in range(cut[0], n):
                    if i not in finished:
                        final(i, False, cut[1], -1, None)

    # ---- frame-array in / frame-array out, ordered, bounded memory
    def stream(self, frames, config: Dict[str, Any], emit: Callable[[int, np.ndarray], None], num_frames: int,
               frame_shape: Tuple[int, int], scale: int, batch: int = 2, window: Optional[int] = None,
               progress_callback=None, frame_callback=None, **hooks) -> RunResult:
        """Ordered streaming: `frames` is an iterator of [H,W,3] uint8 arrays (a raw-video pipe, a decoder), results
        are handed to `emit(index, out)` IN ORDER.  The frames travel through two shared-memory rings of `window`
        slots (page-locked in the workers), so memory is bounded and nothing is pickled or written to disk."""
        h, w = frame_shape
        S = int(window or max(4 * batch * len(self.gpu_ids), 8))
        ring_in = SharedArray((S, h, w, 3))
        ring_out = SharedArray((S, h * scale, w * scale, 3))
        src, sink = ArraySource(ring_in, num_frames), ArraySink(ring_out)
        it = iter(frames)
        emitted = {"next": 0}
        cond = threading.Condition()

        def feeder(pool: "SchedulerPool", state: Dict[str, Any]) -> None:
            a = ring_in.array
            for i in range(num_frames):
                with cond:
                    while i - emitted["next"] >= S and not pool._job_done.is_set():
                        cond.wait(0.05)
                if pool._job_done.is_set():
                    return
                try:
                    f = next(it)
                    a[i % S][...] = f
                except StopIteration:
                    state["truncate"] = (i, f"input stream ended after {i} of {num_frames} frames")
                    return
                except Exception as e:   # a frame of the wrong shape, a broken pipe: the job must still end
                    state["truncate"] = (i, f"input stream failed at frame {i}: {type(e).__name__}: {e}")
                    return
                pool._claims.set_avail(i + 1)

        pending: Dict[int, bool] = {}

        def on_frame(i: int, name: str, ok: bool, err: Optional[str], gpu: int) -> None:
            pending[i] = ok
            with cond:
                while emitted["next"] in pending:
                    j = emitted["next"]
                    if pending.pop(j):
                        emit(j, ring_out.array[j % S])
                    emitted["next"] = j + 1
                cond.notify_all()
            if frame_callback:
                frame_callback(i, name, ok, err, gpu)

        try:
            # one shared shard [0, N): every GPU takes the next chunk from its FRONT (self-scheduling, in order; at most
            # `window` frames are ever available, so the GPUs work on neighbouring chunks of the stream)
            ranges = [(0, num_frames)] + [(0, 0)] * (len(self.gpu_ids) - 1)
            res = self.run(_SingleShard(src), sink, config, batch=batch, progress_callback=progress_callback,
                           frame_callback=on_frame, shard_ranges=ranges, steal=False, avail=0, feeder=feeder,
                           front_shard=0, **hooks)
        finally:
            ring_in.release()
            ring_out.release()
        return res

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        self._job_done.set()
        for g, q in self._cmd.items():
            try:
                q.put(("stop",))
            except Exception:
                pass
        for p in self._procs.values():
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _SingleShard(FrameSource):
    """Wraps a source for ring mode (all GPUs draw from the front of shard 0: `front_shard`)."""

    def __init__(self, inner: FrameSource):
        self.inner = inner

    def __len__(self):
        return len(self.inner)

    def name(self, i):
        return self.inner.name(i)

    def load(self, i):
        return self.inner.load(i)

    def open(self):
        self.inner.open()

    def close(self):
        self.inner.close()
