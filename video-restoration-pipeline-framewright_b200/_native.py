"""ctypes binding of libb200sr.so (C ABI: include/b200sr.h) and its in-tree build.

There is no CPU fallback: if the library is missing or no sm_100 GPU is present, engine creation
raises (`NativeLibraryError` / `RuntimeError`) instead of silently computing elsewhere.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
# $B200SR_LIB selects another build of the library in the package directory (e.g. libb200sr_debug.so: compiled with
# -DB200SR_DEBUG, bounds traps in the fused kernel; `build_variant`)
LIB_PATH = os.path.join(_PKG_DIR, os.environ.get("B200SR_LIB", "libb200sr.so"))

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]

# Every symbol include/b200sr.h declares (tests check the built library exports all of them).
EXPORTED_SYMBOLS = [
    "b200sr_create", "b200sr_destroy", "b200sr_num_convs", "b200sr_num_prelus", "b200sr_set_conv",
    "b200sr_set_prelu", "b200sr_finalize", "b200sr_output_dims", "b200sr_workspace_bytes",
    "b200sr_enqueue_u8", "b200sr_upscale_host_u8", "b200sr_enqueue_u16", "b200sr_upscale_host_u16", "b200sr_last_launch_count", "b200sr_set_option",
    "b200sr_host_alloc", "b200sr_host_free", "b200sr_host_register", "b200sr_host_unregister",
    "b200sr_last_error", "b200sr_version", "b200sr_debug_conv3x3", "b200sr_get_profile",
    "b200sr_debug_plan_regions", "b200sr_debug_pack_weights", "b200sr_debug_choose_th", "b200sr_debug_rdb_items", "b200sr_debug_rdb_flag_rows", "b200sr_debug_rdb_stats", "b200sr_debug_rdb_trace",
]

OK, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_STATE = 0, 1, 2, 3, 4
ARCH_RRDB, ARCH_SRVGG = 0, 1


class NativeLibraryError(RuntimeError):
    """libb200sr.so is missing or unloadable (the product path has no fallback)."""


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("arch", ctypes.c_int),
        ("scale", ctypes.c_int),
        ("num_feat", ctypes.c_int),
        ("num_block", ctypes.c_int),
        ("num_grow_ch", ctypes.c_int),
    ]


def _sources():
    return [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR)) if f.endswith((".cu", ".cuh", ".h"))] + [
        os.path.join(os.path.dirname(_PKG_DIR), "include", "b200sr.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _sources())


def variant_flags(name: str):
    """Compile flags encoded in a variant's file name: *debug* -> bounds traps, *rdb2* -> fused RDB with 2 CTAs per SM,
    *rdb1* -> 1 CTA per SM."""
    flags = []
    if "debug" in name:
        flags.append("-DB200SR_DEBUG")
    if "rdb2" in name:
        flags.append("-DB200SR_RDB_CTAS=2")
    if "rdb1" in name:
        flags.append("-DB200SR_RDB_CTAS=1")
    return flags


def build_variant(name: str, flags, force: bool = False) -> str:
    """Another build of the library next to the default one, e.g. build_variant("libb200sr_debug.so", ["-DB200SR_DEBUG"])."""
    path = os.path.join(_PKG_DIR, name)
    if not force and os.path.exists(path) and all(
            not os.path.exists(s) or os.path.getmtime(s) <= os.path.getmtime(path) for s in _sources()):
        return path
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = path + ".tmp.%d" % os.getpid()
    proc = subprocess.run([nvcc] + NVCC_FLAGS + list(flags) + ["-o", tmp, os.path.join(CSRC_DIR, "b200sr.cu")],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise NativeLibraryError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, path)
    return path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/b200sr.cu for sm_100a into libb200sr.so next to this file (nvcc cross-compiles without a GPU)."""
    if os.path.basename(LIB_PATH) != "libb200sr.so":      # a variant selected through $B200SR_LIB
        return build_variant(os.path.basename(LIB_PATH), variant_flags(os.path.basename(LIB_PATH)), force)
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise NativeLibraryError("nvcc not found; cannot build libb200sr.so")
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    extra = os.environ.get("B200SR_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DB200SR_RDB_STATS for tools/rdb_stats.py
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", tmp, os.path.join(CSRC_DIR, "b200sr.cu")]
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise NativeLibraryError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """dlopen the library and declare the prototypes (idempotent)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for this path)")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as exc:  # pragma: no cover
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
        c_int, c_void_p, c_char_p, c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float
        fp = ctypes.POINTER(ctypes.c_float)
        lib.b200sr_create.argtypes = [ctypes.POINTER(ModelDesc), c_int, ctypes.POINTER(c_void_p)]
        lib.b200sr_create.restype = c_int
        lib.b200sr_destroy.argtypes = [c_void_p]
        lib.b200sr_destroy.restype = None
        lib.b200sr_num_convs.argtypes = [c_void_p]
        lib.b200sr_num_convs.restype = c_int
        lib.b200sr_num_prelus.argtypes = [c_void_p]
        lib.b200sr_num_prelus.restype = c_int
        lib.b200sr_set_conv.argtypes = [c_void_p, c_int, fp, fp, c_int, c_int]
        lib.b200sr_set_conv.restype = c_int
        lib.b200sr_set_prelu.argtypes = [c_void_p, c_int, fp, c_int]
        lib.b200sr_set_prelu.restype = c_int
        lib.b200sr_finalize.argtypes = [c_void_p]
        lib.b200sr_finalize.restype = c_int
        lib.b200sr_output_dims.argtypes = [c_void_p, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]
        lib.b200sr_output_dims.restype = c_int
        lib.b200sr_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                               ctypes.POINTER(ctypes.c_size_t)]
        lib.b200sr_workspace_bytes.restype = c_int
        lib.b200sr_enqueue_u8.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]
        lib.b200sr_enqueue_u8.restype = c_int
        lib.b200sr_upscale_host_u8.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]
        lib.b200sr_upscale_host_u8.restype = c_int
        lib.b200sr_enqueue_u16.argtypes = lib.b200sr_enqueue_u8.argtypes
        lib.b200sr_enqueue_u16.restype = c_int
        lib.b200sr_upscale_host_u16.argtypes = lib.b200sr_upscale_host_u8.argtypes
        lib.b200sr_upscale_host_u16.restype = c_int
        lib.b200sr_host_alloc.argtypes = [ctypes.c_size_t]
        lib.b200sr_host_alloc.restype = c_void_p
        lib.b200sr_host_free.argtypes = [c_void_p]
        lib.b200sr_host_free.restype = None
        lib.b200sr_host_register.argtypes = [c_void_p, ctypes.c_size_t]
        lib.b200sr_host_register.restype = c_int
        lib.b200sr_host_unregister.argtypes = [c_void_p]
        lib.b200sr_host_unregister.restype = c_int
        lib.b200sr_last_launch_count.argtypes = [c_void_p]
        lib.b200sr_last_launch_count.restype = c_int
        lib.b200sr_set_option.argtypes = [c_void_p, c_char_p, c_int]
        lib.b200sr_set_option.restype = c_int
        lib.b200sr_last_error.argtypes = [c_void_p]
        lib.b200sr_last_error.restype = c_char_p
        lib.b200sr_version.argtypes = []
        lib.b200sr_version.restype = c_char_p
        lib.b200sr_debug_conv3x3.argtypes = [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, fp, fp, c_int, c_int,
                                             c_float, fp, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                             c_char_p, c_int]
        lib.b200sr_debug_conv3x3.restype = c_int
        dp = ctypes.POINTER(ctypes.c_double)
        lib.b200sr_get_profile.argtypes = [c_void_p, c_int, dp, dp, ctypes.POINTER(c_int)]
        lib.b200sr_get_profile.restype = c_int
        lib.b200sr_debug_plan_regions.argtypes = [c_int] * 7 + [ctypes.POINTER(c_int), c_int]
        lib.b200sr_debug_plan_regions.restype = c_int
        lib.b200sr_debug_pack_weights.argtypes = [fp, c_int, c_int, c_int, c_void_p, ctypes.c_longlong]
        lib.b200sr_debug_pack_weights.restype = ctypes.c_longlong
        lib.b200sr_debug_choose_th.argtypes = [c_int] * 5
        lib.b200sr_debug_choose_th.restype = c_int
        lib.b200sr_debug_rdb_items.argtypes = [c_int, c_int, c_int, ctypes.POINTER(c_int), c_int]
        lib.b200sr_debug_rdb_items.restype = c_int
        lib.b200sr_debug_rdb_flag_rows.argtypes = []
        lib.b200sr_debug_rdb_flag_rows.restype = c_int
        lib.b200sr_debug_rdb_stats.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_longlong), c_int]
        lib.b200sr_debug_rdb_stats.restype = c_int
        lib.b200sr_debug_rdb_trace.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_longlong), c_int]
        lib.b200sr_debug_rdb_trace.restype = c_int
        _lib = lib
        return lib


def is_loaded() -> bool:
    return _lib is not None
