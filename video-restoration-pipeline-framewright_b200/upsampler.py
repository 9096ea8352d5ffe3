"""`RealESRGANer` duck type backed by the B200 engine.

The reference builds `RealESRGANer(scale, model_path, dni_weight, model, tile, tile_pad, pre_pad, half, gpu_id)`
(`/root/reference/src/framewright/processors/pytorch_realesrgan.py:160-170`, `cli.py:742-750`,
`face_restore.py:388-399`) and calls `.enhance(img, outscale=...) -> (output, img_mode)`
(`pytorch_realesrgan.py:223`, `enhancement/super_resolution.py:524`, `cli.py:769`).  This class keeps that
constructor and that call; the pre-process / tile loop / forward / post-process / uint8 conversion all
run on the GPU inside libb200sr.so (`b200sr_upscale_host_u8`).

Unlike the upstream object this one is re-entrant: `enhance` keeps no per-call state on `self`, and the engine's
host-buffer call is thread-safe (each call takes one of the engine's lanes: own stream, workspace and pinned
staging, csrc/b200sr.cu), so the `parallel_frames` threads of the reference (`restorer.py:1894`) overlap their
copies with each other's kernels.  `close()` waits for calls in flight; a closed upsampler raises `EngineError`.

Weights: a local checkpoint file is loaded as upstream does (params_ema | params, strict).  A URL is downloaded
into the weights directory like upstream's `load_file_from_url`; when that is impossible the constructor RAISES
(`FileNotFoundError`) -- it never falls back to random weights silently.  Synthetic weights (benchmarks, tests)
must be asked for explicitly: `synthetic_seed=<int>`, `state_dict=...` or `B200SR_SYNTHETIC_WEIGHTS=<seed>`.
"""
from __future__ import annotations

import logging
import os
import threading
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .archs import MODEL_ARCHS, ArchDesc, make_synthetic_state_dict
from .engine import B200Engine, EngineError

logger = logging.getLogger(__name__)


class ArchSpec:
    """What the `RRDBNet(...)` / `SRVGGNetCompact(...)` shim constructors return: an architecture
    descriptor plus an (optional) state dict -- not a torch module."""

    def __init__(self, arch: ArchDesc, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.arch = arch
        self._state_dict = state_dict

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        self._state_dict = dict(sd)
        return self

    def state_dict(self):
        return self._state_dict

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def half(self):
        return self


def _load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """Upstream loader: torch.load, prefer 'params_ema', else 'params', else the dict itself."""
    loadnet = torch.load(path, map_location="cpu")
    if isinstance(loadnet, dict):
        for key in ("params_ema", "params"):
            if key in loadnet:
                return loadnet[key]
    return loadnet


def weights_dir() -> str:
    """Where downloaded / user-provided checkpoints live: $B200SR_WEIGHTS_DIR or <package>/weights."""
    return os.environ.get("B200SR_WEIGHTS_DIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


def _resolve_checkpoint(model_path: str) -> str:
    """Local file -> itself; URL -> cached copy in weights_dir(), downloaded if missing (upstream
    `load_file_from_url`).  Raises FileNotFoundError when neither is possible (offline hosts)."""
    mp = str(model_path)
    if os.path.isfile(mp):
        return mp
    if mp.startswith(("http://", "https://")):
        fname = os.path.basename(mp.split("?")[0])
        for d in (weights_dir(), "weights", "."):
            cand = os.path.join(d, fname)
            if os.path.isfile(cand):
                return cand
        dst = os.path.join(weights_dir(), fname)
        try:
            import urllib.request

            os.makedirs(weights_dir(), exist_ok=True)
            tmp = dst + ".part"
            with urllib.request.urlopen(mp, timeout=float(os.environ.get("B200SR_DOWNLOAD_TIMEOUT", "20"))) as r, \
                    open(tmp, "wb") as f:
                while True:
                    chunk = r.read(1 << 20)
                    if not chunk:
                        break
                    f.write(chunk)
            os.replace(tmp, dst)
            return dst
        except Exception as exc:
            raise FileNotFoundError(
                f"model weights {fname} are not available: download of {mp} failed ({exc}); "
                f"place the file in {weights_dir()} (or set B200SR_WEIGHTS_DIR)") from exc
    # a bare / relative file name (face_restore.py:391 passes 'RealESRGAN_x4plus.pth'): look in the weights directories
    for d in (weights_dir(), "weights"):
        cand = os.path.join(d, os.path.basename(mp))
        if os.path.isfile(cand):
            return cand
    raise FileNotFoundError(f"model weights not found: {mp} (also looked in {weights_dir()})")


def _arch_for_model_path(model_path: Optional[str]) -> Optional[str]:
    if not model_path:
        return None
    base = os.path.basename(str(model_path))
    stem = base[:-4] if base.endswith(".pth") else base
    return stem if stem in MODEL_ARCHS else None


class RealESRGANer:
    """Drop-in for `realesrgan.RealESRGANer` on a B200."""

    def __init__(self, scale, model_path=None, dni_weight=None, model=None, tile=0, tile_pad=10, pre_pad=10,
                 half=False, device=None, gpu_id=None, state_dict=None, model_name=None, synthetic_seed=None):
        self.scale = int(scale)
        self.tile_size = int(tile or 0)
        self.tile_pad = int(tile_pad)
        self.pre_pad = int(pre_pad)
        self.mod_scale = None
        self.half = bool(half)  # accepted for API parity; the engine always computes bf16-in / fp32-accumulate
        # upstream: `dni_weight` blends two checkpoints of one architecture (realesr-general-x4v3 + its -wdn twin for
        # the denoise strength): model_path is then a list of two paths.  Also accepted: state_dict=[sd_a, sd_b].
        dni_pair = None
        if dni_weight is not None:
            if len(dni_weight) != 2:
                raise EngineError("dni_weight must hold two weights")
            if isinstance(state_dict, (list, tuple)) and len(state_dict) == 2:
                dni_pair = (state_dict[0], state_dict[1])
                state_dict = None
            elif isinstance(model_path, (list, tuple)) and len(model_path) == 2:
                dni_pair = tuple(_load_checkpoint(str(p)) for p in model_path)
            else:
                raise EngineError("dni_weight needs two checkpoints: model_path=[a, b] (local files) or state_dict=[a, b]")
            model_path = model_path[0] if isinstance(model_path, (list, tuple)) else model_path
        if gpu_id is None:
            gpu_id = device.index if isinstance(device, torch.device) and device.index is not None else 0
        self.gpu_id = int(gpu_id)
        self.device = torch.device(f"cuda:{self.gpu_id}")

        name = model_name or _arch_for_model_path(model_path)
        if isinstance(model, ArchSpec):
            arch = model.arch
            if state_dict is None:
                state_dict = model.state_dict()
        elif name is not None:
            arch = MODEL_ARCHS[name]
        else:
            raise EngineError("cannot determine the architecture: pass model=<RRDBNet(...) shim> or model_name=")
        # upstream ships realesr-general-x4v3 / realesr-animevideov3 as SRVGGNetCompact checkpoints even though
        # the reference constructs RRDBNet for those names (SURVEY.md finding 3): the name wins.
        if name in MODEL_ARCHS and MODEL_ARCHS[name] != arch:
            logger.info("model name %s selects %s over the passed architecture", name, MODEL_ARCHS[name].kind)
            arch = MODEL_ARCHS[name]
        if arch.scale != self.scale:
            raise EngineError(f"scale {self.scale} does not match the {arch.kind} network scale {arch.scale}")

        if dni_pair is not None:
            state_dict = self.dni(dni_pair[0], dni_pair[1], dni_weight)
        if state_dict is None:
            if synthetic_seed is None and os.environ.get("B200SR_SYNTHETIC_WEIGHTS", "") != "":
                synthetic_seed = int(os.environ["B200SR_SYNTHETIC_WEIGHTS"])
            if synthetic_seed is not None:
                # explicit opt-in only (benchmarks / tests): deterministic random-init weights of this architecture
                synth_name = name if name in MODEL_ARCHS else next(k for k, v in MODEL_ARCHS.items() if v == arch)
                logger.warning("using SYNTHETIC weights (seed %d) for %s: output is not a trained model's",
                               int(synthetic_seed), synth_name)
                state_dict = make_synthetic_state_dict(synth_name, int(synthetic_seed))
            elif model_path:
                state_dict = _load_checkpoint(_resolve_checkpoint(str(model_path)))   # raises FileNotFoundError
            else:
                raise FileNotFoundError("no weights: pass model_path=<file or URL>, state_dict=, or synthetic_seed=")
        self.arch = arch
        self._engine = B200Engine(arch, state_dict, gpu_id=self.gpu_id)
        self._cv = threading.Condition()
        self._inflight = 0
        self._closed = False

    @staticmethod
    def dni(net_a: Dict[str, torch.Tensor], net_b: Dict[str, torch.Tensor], dni_weight, key: str = "params",
            loc: str = "cpu") -> Dict[str, torch.Tensor]:
        """Deep network interpolation (upstream RealESRGANer.dni): w = dni_weight[0] * a + dni_weight[1] * b per
        tensor.  Takes state dicts (optionally wrapped as {"params": ...} / {"params_ema": ...} like the files)."""
        def unwrap(sd):
            for k in ("params_ema", key):
                if isinstance(sd, dict) and k in sd and isinstance(sd[k], dict):
                    return sd[k]
            return sd
        a, b = unwrap(net_a), unwrap(net_b)
        if set(a.keys()) != set(b.keys()):
            raise EngineError("dni: the two checkpoints have different tensors")
        return {k: float(dni_weight[0]) * a[k].float() + float(dni_weight[1]) * b[k].float() for k in a}

    # --------------------------------------------------------------------------------------------
    def _enter(self) -> None:
        with self._cv:
            if self._closed:
                raise EngineError("upsampler is closed")
            self._inflight += 1

    def _exit(self) -> None:
        with self._cv:
            self._inflight -= 1
            if self._inflight == 0:
                self._cv.notify_all()

    def _run_u8(self, img_bgr_u8: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:   # uint8 / uint16
        self._enter()
        try:
            return self._engine.upscale_host(img_bgr_u8, out=out, tile=self.tile_size, tile_pad=self.tile_pad,
                                             pre_pad=self.pre_pad)
        finally:
            self._exit()

    def enhance(self, img: np.ndarray, outscale: Optional[float] = None,
                alpha_upsampler: str = "realesrgan") -> Tuple[np.ndarray, str]:
        """BGR / gray / BGRA ndarray -> (ndarray scaled by `outscale` or the net scale, img_mode).  uint8 in ->
        uint8 out; an image whose maximum exceeds 256 is a 16-bit image as upstream (`max_range = 65535`): it runs
        through the engine's uint16 path and comes back as uint16."""
        import cv2

        if not isinstance(img, np.ndarray) or img.ndim not in (2, 3):
            raise EngineError("img must be an HxW or HxWxC ndarray")
        h_input, w_input = img.shape[0:2]
        if img.dtype != np.uint8:
            # upstream: img.astype(float32); max_range = 65535 if np.max(img) > 256 else 255
            img = img.astype(np.uint16) if np.max(img) > 256 else img.astype(np.uint8)
        if img.ndim == 2:
            img_mode = "L"
            out = self._run_u8(cv2.cvtColor(img, cv2.COLOR_GRAY2BGR))
            output = cv2.cvtColor(out, cv2.COLOR_BGR2GRAY)
        elif img.shape[2] == 4:
            img_mode = "RGBA"
            out = self._run_u8(np.ascontiguousarray(img[:, :, 0:3]))
            alpha = img[:, :, 3]
            if alpha_upsampler == "realesrgan":
                a = self._run_u8(cv2.cvtColor(alpha, cv2.COLOR_GRAY2BGR))
                out_alpha = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY)
            else:
                h, w = alpha.shape[0:2]
                out_alpha = cv2.resize(alpha, (w * self.scale, h * self.scale), interpolation=cv2.INTER_LINEAR)
            output = cv2.cvtColor(out, cv2.COLOR_BGR2BGRA)
            output[:, :, 3] = out_alpha
        elif img.shape[2] == 3:
            img_mode = "RGB"
            output = self._run_u8(np.ascontiguousarray(img))
        else:
            raise EngineError(f"unsupported channel count {img.shape[2]}")
        if outscale is not None and float(outscale) != float(self.scale):
            output = cv2.resize(output, (int(w_input * outscale), int(h_input * outscale)),
                                interpolation=cv2.INTER_LANCZOS4)
        return output, img_mode

    def enhance_batch(self, frames: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """[N,H,W,3] uint8 BGR -> [N,H*s,W*s,3]; frames are independent.  The engine cuts large batches into lane
        jobs and pipelines them (copies of one job under the kernels of the other).  `out`: optional destination
        (page-locked memory is written directly by the D2H copy)."""
        return self._run_u8(frames, out)

    @property
    def engine(self) -> B200Engine:
        return self._engine

    def close(self) -> None:
        """Waits for calls in flight (other threads inside `enhance`), then destroys the engine.  Idempotent."""
        with self._cv:
            if self._closed:
                return
            self._closed = True
            while self._inflight > 0:
                self._cv.wait()
        self._engine.close()
