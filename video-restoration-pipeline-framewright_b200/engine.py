"""B200 engine: owns one `b200sr_engine` handle (C ABI, include/b200sr.h) on one GPU.

Replaces what `RealESRGANer(...)` holds after construction in the reference
(`/root/reference/src/framewright/processors/pytorch_realesrgan.py:160-170`): the network on the
device with its weights, plus the pre/tile/post-processing around the forward pass.  PyTorch is
used only for device memory and streams; all arithmetic runs in libb200sr.so.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _native
from .archs import MODEL_ARCHS, ArchDesc, conv_layers, prelu_layers


class EngineError(RuntimeError):
    pass


class EngineOutOfMemory(EngineError):
    """Device allocation failed; message starts with 'GPU out of memory' (callers grep for 'memory',
    /root/reference/src/framewright/restorer.py:1746)."""


class PinnedPool:
    """Recycling pool of page-locked host buffers (C ABI b200sr_host_alloc) behind numpy arrays.

    `enhance` returns a fresh ndarray per frame (44 MB at 720p x4).  Allocating it here lets the D2H copy land in
    the returned array itself -- no pinned -> pageable memcpy (~5 ms per 720p frame) -- and cudaHostAlloc (slow) is
    paid once per buffer: when the last view of an array dies, its buffer goes back to the pool.  The pool holds at
    most $B200SR_PINNED_POOL_MB (default 4096) of idle buffers; beyond that, or if pinning fails, buffers are freed /
    the array is ordinary pageable memory (the engine then stages through its own pinned buffers)."""

    def __init__(self):
        self._free: Dict[int, List[int]] = {}
        self._idle_bytes = 0
        self._lock = threading.Lock()
        self.cap_bytes = int(os.environ.get("B200SR_PINNED_POOL_MB", "4096")) << 20
        self.enabled = os.environ.get("B200SR_PINNED_POOL", "1") != "0"

    def _release(self, ptr: int, nbytes: int) -> None:
        with self._lock:
            if self._idle_bytes + nbytes <= self.cap_bytes:
                self._free.setdefault(nbytes, []).append(ptr)
                self._idle_bytes += nbytes
                return
        try:
            _native.load().b200sr_host_free(ctypes.c_void_p(ptr))
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def empty(self, shape: Tuple[int, ...], dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        if not self.enabled or nbytes < (1 << 16):
            return np.empty(shape, dtype=dtype)
        ptr = None
        with self._lock:
            lst = self._free.get(nbytes)
            if lst:
                ptr = lst.pop()
                self._idle_bytes -= nbytes
        if ptr is None:
            ptr = _native.load().b200sr_host_alloc(nbytes)
            if not ptr:
                return np.empty(shape, dtype=dtype)
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, int(ptr), nbytes)
        return np.frombuffer(buf, dtype=np.uint8, count=nbytes).view(dtype).reshape(shape)

    def trim(self) -> None:
        """Free every idle buffer (clear_upsampler_cache)."""
        with self._lock:
            ptrs = [p for lst in self._free.values() for p in lst]
            self._free.clear()
            self._idle_bytes = 0
        lib = _native.load()
        for p in ptrs:
            lib.b200sr_host_free(ctypes.c_void_p(p))


PINNED_POOL = PinnedPool()


def _fptr(t: torch.Tensor):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


class B200Engine:
    """One network on one B200.  `upscale_host` is thread-safe (the C ABI gives every call its own lanes);
    `upscale_device` calls must be stream-ordered with each other; weight loading / options / close need it idle."""

    def __init__(self, model_name_or_arch: Union[str, ArchDesc], state_dict: Dict[str, torch.Tensor], gpu_id: int = 0):
        arch = MODEL_ARCHS[model_name_or_arch] if isinstance(model_name_or_arch, str) else model_name_or_arch
        self.arch = arch
        self.gpu_id = int(gpu_id)
        self.scale = arch.scale
        self._lib = _native.load()  # raises NativeLibraryError if the CUDA library is missing
        if not torch.cuda.is_available():
            raise EngineError("no CUDA device: the B200 upscaling path has no CPU fallback")
        desc = _native.ModelDesc(
            _native.ARCH_RRDB if arch.kind == "rrdb" else _native.ARCH_SRVGG,
            arch.scale, arch.num_feat, arch.num_block, arch.num_grow_ch)
        handle = ctypes.c_void_p()
        rc = self._lib.b200sr_create(ctypes.byref(desc), self.gpu_id, ctypes.byref(handle))
        if rc != _native.OK:
            raise EngineError(f"b200sr_create failed (status {rc}): needs an sm_100 (B200) device {self.gpu_id}")
        self._h = handle
        self._load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    def _check(self, rc: int, what: str) -> None:
        if rc == _native.OK:
            return
        msg = self._lib.b200sr_last_error(self._h).decode("utf-8", "replace")
        if rc == _native.ERR_OOM:
            raise EngineOutOfMemory(f"GPU out of memory: {msg}")
        raise EngineError(f"{what} failed (status {rc}): {msg}")

    def _load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Strict load (upstream `load_state_dict(strict=True)`): every expected key, no extras."""
        layers = conv_layers(self.arch)
        prelus = prelu_layers(self.arch)
        expected = {n + ".weight" for n, _, _ in layers} | {n + ".bias" for n, _, _ in layers} | set(prelus)
        missing = sorted(expected - set(sd.keys()))
        extra = sorted(set(sd.keys()) - expected)
        if missing or extra:
            raise EngineError(f"state_dict mismatch: missing={missing[:4]} unexpected={extra[:4]}")
        assert self._lib.b200sr_num_convs(self._h) == len(layers)
        for i, (name, cin, cout) in enumerate(layers):
            w = sd[name + ".weight"].detach().to("cpu", torch.float32).contiguous()
            b = sd[name + ".bias"].detach().to("cpu", torch.float32).contiguous()
            if tuple(w.shape) != (cout, cin, 3, 3):
                raise EngineError(f"{name}.weight has shape {tuple(w.shape)}, expected {(cout, cin, 3, 3)}")
            self._check(self._lib.b200sr_set_conv(self._h, i, _fptr(w), _fptr(b), cout, cin), f"set_conv({name})")
        for i, key in enumerate(prelus):
            a = sd[key].detach().to("cpu", torch.float32).contiguous()
            self._check(self._lib.b200sr_set_prelu(self._h, i, _fptr(a), a.numel()), f"set_prelu({key})")
        self._check(self._lib.b200sr_finalize(self._h), "finalize")

    # ------------------------------------------------------------------ running
    def set_option(self, key: str, value: int) -> None:
        self._check(self._lib.b200sr_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def workspace_bytes(self, n: int, h: int, w: int, tile: int = 0, tile_pad: int = 10, pre_pad: int = 0) -> int:
        out = ctypes.c_size_t()
        self._check(self._lib.b200sr_workspace_bytes(self._h, n, h, w, tile, tile_pad, pre_pad, ctypes.byref(out)),
                    "workspace_bytes")
        return int(out.value)

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.b200sr_last_launch_count(self._h))

    def upscale_device(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None, tile: int = 0,
                       tile_pad: int = 10, pre_pad: int = 0) -> torch.Tensor:
        """frames: CUDA uint8 (or uint16: upstream's 16-bit image branch) [N,H,W,3] BGR (contiguous) -> CUDA tensor of
        the same dtype [N,H*s,W*s,3]; async on the current stream."""
        if frames.dtype not in (torch.uint8, torch.uint16) or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_cuda:
            raise EngineError("frames must be a CUDA uint8 / uint16 tensor [N,H,W,3]")
        if frames.device.index != self.gpu_id:
            raise EngineError(f"frames live on cuda:{frames.device.index}, engine on cuda:{self.gpu_id}")
        frames = frames.contiguous()
        n, h, w, _ = frames.shape
        s = self.scale
        if out is None:
            out = torch.empty((n, h * s, w * s, 3), dtype=frames.dtype, device=frames.device)
        elif tuple(out.shape) != (n, h * s, w * s, 3) or out.dtype != frames.dtype or not out.is_contiguous():
            raise EngineError("bad output tensor")
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        fn = self._lib.b200sr_enqueue_u16 if frames.dtype == torch.uint16 else self._lib.b200sr_enqueue_u8
        rc = fn(self._h, frames.data_ptr(), out.data_ptr(), n, h, w, int(tile), int(tile_pad), int(pre_pad),
                ctypes.c_void_p(stream))
        self._check(rc, "enqueue_u8")
        return out

    def upscale_host(self, frames: np.ndarray, out: Optional[np.ndarray] = None, tile: int = 0, tile_pad: int = 10,
                     pre_pad: int = 0) -> np.ndarray:
        """frames: host uint8 (or uint16) [N,H,W,3] or [H,W,3] BGR -> host array of the same dtype, through the C ABI's
        host-buffer call (pipelined H2D + forward + D2H inside; the result array is page-locked, `PinnedPool`)."""
        single = frames.ndim == 3
        f = np.ascontiguousarray(frames[None] if single else frames)
        if f.dtype not in (np.uint8, np.uint16) or f.ndim != 4 or f.shape[-1] != 3:
            raise EngineError("frames must be uint8 / uint16 [N,H,W,3] or [H,W,3]")
        n, h, w, _ = f.shape
        s = self.scale
        if out is None:
            out = PINNED_POOL.empty((n, h * s, w * s, 3), f.dtype)
        elif out.dtype != f.dtype or tuple(out.shape) != (n, h * s, w * s, 3) or not out.flags["C_CONTIGUOUS"]:
            raise EngineError("bad output array")
        fn = self._lib.b200sr_upscale_host_u16 if f.dtype == np.uint16 else self._lib.b200sr_upscale_host_u8
        rc = fn(self._h, f.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), n, h, w, int(tile),
                int(tile_pad), int(pre_pad))
        self._check(rc, "upscale_host_u8")
        return out[0] if single else out

    PROFILE_CLASSES = ["conv32_act", "conv64_act", "conv64_prelu", "conv64_rdb5", "conv64_rdb5_rrdb", "conv64_add",
                       "conv16_last_u8", "conv48_srvgg_last", "first_conv", "upsample2x", "rdb_fused", "hr_last_fused"]

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        """Per-kernel-class {ms, flops, launches} collected since the last call (needs set_option('profile', 1))."""
        n = len(self.PROFILE_CLASSES)
        ms = (ctypes.c_double * n)()
        fl = (ctypes.c_double * n)()
        cnt = (ctypes.c_int * n)()
        self._check(self._lib.b200sr_get_profile(self._h, n, ms, fl, cnt), "get_profile")
        return {k: {"ms": ms[i], "flops": fl[i], "launches": cnt[i]} for i, k in enumerate(self.PROFILE_CLASSES)
                if cnt[i] > 0}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.b200sr_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def debug_conv3x3(x_nhwc_bf16: torch.Tensor, cin: int, weight: torch.Tensor, bias: torch.Tensor, out: torch.Tensor,
                  out_choff: int = 0, slope: float = 1.0, prelu: Optional[torch.Tensor] = None, force_th: int = 0,
                  max_ctas: int = 0) -> None:
    """Test hook: run ONE tensor-core conv layer (C ABI b200sr_debug_conv3x3) on CUDA bf16 NHWC tensors."""
    lib = _native.load()
    n, h, w, in_pitch = x_nhwc_bf16.shape
    wt = weight.detach().to("cpu", torch.float32).contiguous()
    bs = bias.detach().to("cpu", torch.float32).contiguous()
    pr = prelu.detach().to("cpu", torch.float32).contiguous() if prelu is not None else None
    err = ctypes.create_string_buffer(512)
    stream = torch.cuda.current_stream(x_nhwc_bf16.device).cuda_stream
    rc = lib.b200sr_debug_conv3x3(
        x_nhwc_bf16.device.index, x_nhwc_bf16.data_ptr(), n, h, w, in_pitch, cin, _fptr(wt), _fptr(bs), wt.shape[0],
        1 if pr is not None else 0, float(slope), _fptr(pr) if pr is not None else None, out.data_ptr(),
        out.shape[-1], out_choff, 1 if x_nhwc_bf16.dtype == torch.float16 else 0, force_th, max_ctas,
        ctypes.c_void_p(stream), err, 512)
    if rc != _native.OK:
        raise EngineError(f"debug_conv3x3 failed ({rc}): {err.value.decode()}")
