"""The upscaling path behind the reference's optional operator interfaces (`SURVEY.md` section 8 (b) item 3):

  * `plugins.base.ProcessorPlugin` with `PluginCapability.UPSCALE`
    (`/root/reference/src/framewright/plugins/base.py:15-48, 186-250`; loaded, registered and instantiated by
    `plugins/manager.py:36-132, 172-205, 291-321`): `make_processor_plugin()` -> a class the reference's
    `PluginRegistry.register` accepts; `plugins_contrib/b200_realesrgan.py` is the one-file plugin its `PluginLoader`
    picks up from a plugin directory;
  * `engine.pipeline.FrameProcessor` / `VideoProcessor` (`engine/pipeline.py:110-149`: `process_frame(frame, **kwargs)
    -> frame`; `process_video(input_path, output_path, progress_callback, **kwargs) -> bool` -- the `processor` of a
    `PipelineStage`): `B200FrameProcessor`, `B200VideoProcessor`;
  * `infrastructure.gpu.backends.base.Backend` + `register_backend(BackendType, cls)` (`backends/base.py:65-214,
    808-816`): `make_compute_backend()` -> a `Backend` whose `load_model` / `run_inference` are the engine (the
    reference's own `CUDABackend.run_inference` returns its input, `:341-350`).

The reference's base classes are taken from the reference at call time (`framewright.*` must be importable where the
adapter is registered -- it is, inside a FrameWright process); this package itself never needs them.  The work is the
same everywhere: a `PyTorchESRGANConfig` from the settings, `get_upsampler`, frames through `enhance` /
`enhance_batch`.  No CPU path: without the CUDA library / an sm_100 device `initialize` fails loudly.
"""
from __future__ import annotations

import importlib
import logging
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .archs import MODEL_ARCHS
from .pytorch_realesrgan import (PyTorchESRGANConfig, clear_upsampler_cache, convert_ncnn_model_name, get_upsampler,
                                 is_pytorch_esrgan_available)
from .super_resolution import upscale_frame_list

logger = logging.getLogger(__name__)

PLUGIN_NAME = "b200-realesrgan"
PLUGIN_VERSION = "0.2"

# settings the adapters understand (the plugin's `settings_schema`); everything else in a settings dict is ignored
SETTINGS_SCHEMA: Dict[str, Any] = {
    "model_name": {"type": "string", "default": "RealESRGAN_x4plus", "enum": sorted(MODEL_ARCHS)},
    "scale": {"type": "number", "default": None, "description": "output scale (default: the network's; another value "
                                                                "resizes the result, as RealESRGANer.enhance(outscale=)"},
    "tile_size": {"type": "integer", "default": 0, "description": "0 = whole frame"},
    "tile_pad": {"type": "integer", "default": 10},
    "pre_pad": {"type": "integer", "default": 0},
    "max_batch": {"type": "integer", "default": 16, "description": "same-size frames per engine call in process_batch"},
}


def gpu_id_of(device: Any, default: int = 0) -> int:
    """'cuda' / 'cuda:3' / 3 / torch.device -> GPU index.  'cpu' (the reference's default device string,
    `plugins/base.py:105`) is refused: this path has no CPU implementation."""
    if isinstance(device, int):
        return device
    text = str(device)
    if text.startswith("cuda"):
        return int(text.split(":", 1)[1]) if ":" in text else default
    raise RuntimeError(f"the B200 upscaling path runs on a CUDA device only (got device {text!r}); "
                       "there is no CPU fallback")


class UpscaleSession:
    """Settings -> config -> upsampler -> frames.  Shared by the three adapters; free of reference imports.
    With `model_path` (a checkpoint file the caller names, `Backend.load_model(name, model_path)`) the session owns
    its upsampler; without, it uses the process-wide cache of `get_upsampler` like every other caller."""

    def __init__(self, device: Any = "cuda:0", settings: Optional[Dict[str, Any]] = None, model_path: Any = None):
        self.gpu_id = gpu_id_of(device)
        self.model_path = None if model_path is None else str(model_path)
        self.settings: Dict[str, Any] = {k: v["default"] for k, v in SETTINGS_SCHEMA.items()}
        self._config: Optional[PyTorchESRGANConfig] = None
        self._own = None
        self._lock = threading.Lock()
        self.update(settings or {})

    def update(self, settings: Dict[str, Any]) -> None:
        before = dict(self.settings)
        for k, v in settings.items():
            if k == "model" or k == "model_name":
                name = str(v)
                self.settings["model_name"] = name if name in MODEL_ARCHS else convert_ncnn_model_name(name)
            elif k in ("scale", "scale_factor", "outscale"):
                self.settings["scale"] = v
            elif k in ("tile", "tile_size"):
                self.settings["tile_size"] = int(v or 0)
            elif k in SETTINGS_SCHEMA:
                self.settings[k] = v
        if self.settings != before:
            self._config = None
            self._drop_own()

    @property
    def net_scale(self) -> int:
        return MODEL_ARCHS[self.settings["model_name"]].scale

    @property
    def scale(self) -> float:
        s = self.settings["scale"]
        return float(self.net_scale if s is None else s)

    def config(self) -> PyTorchESRGANConfig:
        if self._config is None:
            cfg = PyTorchESRGANConfig(model_name=self.settings["model_name"], scale_factor=self.net_scale,
                                      tile_size=int(self.settings["tile_size"]), tile_pad=int(self.settings["tile_pad"]),
                                      pre_pad=int(self.settings["pre_pad"]), gpu_id=self.gpu_id)
            cfg.validate()
            self._config = cfg
        return self._config

    def upsampler(self):
        if self.model_path is None:
            return get_upsampler(self.config())
        with self._lock:
            if self._own is None:
                from .upsampler import RealESRGANer

                cfg = self.config()
                self._own = RealESRGANer(scale=self.net_scale, model_path=self.model_path, model_name=cfg.model_name,
                                         tile=cfg.tile_size, tile_pad=cfg.tile_pad, pre_pad=cfg.pre_pad,
                                         gpu_id=self.gpu_id)
            return self._own

    def _drop_own(self) -> None:
        with self._lock:
            own, self._own = self._own, None
        if own is not None:
            own.close()

    def open(self):
        """Builds (or finds in the cache) the engine now, so that a missing library / device / checkpoint fails at
        initialisation and not on the first frame."""
        if not is_pytorch_esrgan_available():
            raise RuntimeError("B200 upscaling path unavailable: libb200sr.so could not be loaded or there is no "
                               "sm_100 device (no CPU fallback)")
        return self.upsampler()

    def frame(self, frame: np.ndarray) -> np.ndarray:
        return self.upsampler().enhance(frame, outscale=self.scale)[0]

    def frames(self, frames: Sequence[np.ndarray]) -> List[np.ndarray]:
        return upscale_frame_list(self.upsampler(), list(frames), self.scale, max_batch=int(self.settings["max_batch"]))

    def output_size(self, input_size: Tuple[int, int]) -> Tuple[int, int]:
        h, w = input_size[:2]
        return (int(h * self.scale), int(w * self.scale))     # as `RealESRGANer.enhance` sizes its result

    def close(self) -> None:
        if self.model_path is None:
            clear_upsampler_cache()
        self._drop_own()


# ---------------------------------------------------------------------------------------------------------------------
# plugins.base.ProcessorPlugin
# ---------------------------------------------------------------------------------------------------------------------
def make_processor_plugin(base_module: Any = None):
    """Returns `B200RealESRGANPlugin`, a subclass of the reference's `ProcessorPlugin` (`base_module` =
    `framewright.plugins.base`, imported if not given; the class is created once per base module)."""
    base = base_module or importlib.import_module("framewright.plugins.base")
    cached = _PLUGIN_CLASSES.get(base)
    if cached is not None:
        return cached

    class B200RealESRGANPlugin(base.ProcessorPlugin):
        """Real-ESRGAN upscaling (RRDBNet x4 / x2, SRVGGNetCompact) on the B200 engine."""

        @classmethod
        def get_metadata(cls):
            return base.PluginMetadata(
                name=PLUGIN_NAME, version=PLUGIN_VERSION,
                description="Real-ESRGAN upscaling (RRDBNet x4/x2, SRVGGNetCompact) as hand-written sm_100a kernels",
                capabilities={base.PluginCapability.UPSCALE},
                python_packages=["torch", "numpy"],
                min_vram_mb=8000, recommended_vram_mb=24000,
                supports_cpu=False, supports_cuda=True, supports_mps=False,
                settings_schema=dict(SETTINGS_SCHEMA))

        def _on_initialize(self) -> None:
            # (`initialize` has stored device + settings; an exception here leaves `is_initialized` False and makes
            #  `PluginManager.get_plugin` log it and return None, manager.py:318-321)
            self._session = UpscaleSession(self._device, self._settings)
            self._session.open()

        def _on_settings_changed(self, settings: Dict[str, Any]) -> None:
            if getattr(self, "_session", None) is not None:
                self._session.update(settings)

        def _on_cleanup(self) -> None:
            if getattr(self, "_session", None) is not None:
                self._session.close()
                self._session = None

        def _need(self) -> UpscaleSession:
            if not self._initialized or getattr(self, "_session", None) is None:
                raise RuntimeError(f"plugin {PLUGIN_NAME} is not initialized (call initialize(device='cuda:N') first)")
            return self._session

        def process_frame(self, frame: np.ndarray, frame_number: int, context: Optional[Dict[str, Any]] = None) -> np.ndarray:
            return self._need().frame(frame)

        def process_batch(self, frames: List[np.ndarray], start_frame: int,
                          context: Optional[Dict[str, Any]] = None) -> List[np.ndarray]:
            return self._need().frames(frames)

        def supports_batch(self) -> bool:
            return True

        def get_temporal_radius(self) -> int:
            return 0                      # frames are independent (no temporal state in RRDBNet / SRVGG)

        def estimate_output_size(self, input_size: tuple) -> tuple:
            s = getattr(self, "_session", None) or UpscaleSession("cuda:0", self._settings)
            return s.output_size(input_size)

        def get_progress_weight(self) -> float:
            return 10.0                   # the heavy stage of a restoration (17.9 M MAC per input pixel)

    _PLUGIN_CLASSES[base] = B200RealESRGANPlugin
    return B200RealESRGANPlugin


_PLUGIN_CLASSES: Dict[Any, Any] = {}      # base module -> class (one class per set of base classes)


def register_plugin(manager_or_registry: Any, base_module: Any = None):
    """`register_plugin(PluginManager())` / `register_plugin(registry)`: registers the plugin class and returns it."""
    cls = make_processor_plugin(base_module)
    registry = getattr(manager_or_registry, "registry", manager_or_registry)
    registry.register(cls)
    return cls


# ---------------------------------------------------------------------------------------------------------------------
# engine.pipeline.FrameProcessor
# ---------------------------------------------------------------------------------------------------------------------
class B200FrameProcessor:
    """`PipelineStage(name="upscale", processor=B200FrameProcessor(model_name=..., device="cuda:0"))`: satisfies the
    reference's `FrameProcessor` protocol.  Keyword arguments are the stage's `StageConfig.params` (the same for every
    frame of a stage): known ones update the settings and stay in force.  One caller thread at a time per object."""

    def __init__(self, device: Any = "cuda:0", **settings: Any):
        self._session = UpscaleSession(device, settings)

    def process_frame(self, frame: np.ndarray, **kwargs: Any) -> np.ndarray:
        if kwargs:
            known = {k: v for k, v in kwargs.items() if k in _SETTING_ALIASES}
            if known:
                self._session.update(known)
        return self._session.frame(frame)

    def process_frames(self, frames: Sequence[np.ndarray], **kwargs: Any) -> List[np.ndarray]:
        if kwargs:
            self._session.update({k: v for k, v in kwargs.items() if k in _SETTING_ALIASES})
        return self._session.frames(frames)

    __call__ = process_frame

    def close(self) -> None:
        self._session.close()


_SETTING_ALIASES = set(SETTINGS_SCHEMA) | {"model", "scale_factor", "outscale", "tile"}


class B200VideoProcessor:
    """The reference's `VideoProcessor` protocol (`engine/pipeline.py:126-149`): `process_video(input_path, output_path,
    progress_callback, **kwargs) -> bool`, the form `Pipeline._run_processor` prefers (`:1156-1171`).  (Its frame-by-frame
    fallback for a `FrameProcessor`, `:1192-1269`, writes into a `VideoWriter` opened at the INPUT size, which cannot hold an
    upscaled frame -- an upscaling stage has to be a video processor there.)

    Decode (cv2.VideoCapture), the engine and encode (cv2.VideoWriter at the scaled size) run in three threads with
    bounded queues between them, so the codec work of batch k+1 / k-1 overlaps the forward pass of batch k.  Failures
    raise (the pipeline's stage logic retries / fails on exceptions, `:1083-1138`); success returns True."""

    def __init__(self, device: Any = "cuda:0", fourcc: str = "mp4v", queue_batches: int = 3, **settings: Any):
        self._session = UpscaleSession(device, settings)
        self.fourcc = fourcc
        self.queue_batches = max(1, int(queue_batches))
        self.frames_processed = 0

    def process_video(self, input_path: Any, output_path: Any, progress_callback: Any = None, **kwargs: Any) -> bool:
        import queue as queue_mod

        import cv2

        if kwargs:
            self._session.update({k: v for k, v in kwargs.items() if k in _SETTING_ALIASES})
        session = self._session
        cap = cv2.VideoCapture(str(input_path))
        if not cap.isOpened():
            raise ValueError(f"Cannot open video: {input_path}")
        fps = cap.get(cv2.CAP_PROP_FPS) or 25.0
        width, height = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        out_h, out_w = session.output_size((height, width))
        writer = cv2.VideoWriter(str(output_path), cv2.VideoWriter_fourcc(*self.fourcc), fps, (out_w, out_h))
        if not writer.isOpened():
            cap.release()
            raise ValueError(f"Cannot create video: {output_path}")
        batch = max(1, int(session.settings["max_batch"]))
        decoded: "queue_mod.Queue" = queue_mod.Queue(self.queue_batches)
        encoded: "queue_mod.Queue" = queue_mod.Queue(self.queue_batches)
        failure: List[BaseException] = []
        stop = threading.Event()

        def put(q, item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue_mod.Full:
                    continue
            return False

        def decode() -> None:
            try:
                frames: List[np.ndarray] = []
                while not stop.is_set():
                    ok, frame = cap.read()
                    if not ok:
                        break
                    frames.append(frame)
                    if len(frames) == batch:
                        if not put(decoded, frames):
                            return
                        frames = []
                if frames:
                    put(decoded, frames)
            except BaseException as e:
                failure.append(e)
            finally:
                put(decoded, None)

        def encode() -> None:
            try:
                while True:
                    frames = encoded.get()
                    if frames is None:
                        return
                    for f in frames:
                        writer.write(np.ascontiguousarray(f))
            except BaseException as e:
                failure.append(e)
                stop.set()

        threads = [threading.Thread(target=decode, name="b200sr-decode", daemon=True),
                   threading.Thread(target=encode, name="b200sr-encode", daemon=True)]
        for t in threads:
            t.start()
        done = 0
        try:
            session.open()
            while not failure:
                try:
                    frames = decoded.get(timeout=0.1)
                except queue_mod.Empty:
                    continue
                if frames is None:
                    break
                if not put(encoded, session.frames(frames)):
                    break
                done += len(frames)
                if progress_callback is not None and total > 0:
                    progress_callback(min(1.0, done / total))
        except BaseException as e:
            failure.append(e)
        finally:
            if failure:
                stop.set()
            try:
                encoded.put(None, timeout=5)
            except queue_mod.Full:
                stop.set()
            for t in threads:
                t.join(timeout=30)
            cap.release()
            writer.release()
        self.frames_processed = done
        if failure:
            raise failure[0]
        if progress_callback is not None:
            progress_callback(1.0)
        return True

    def close(self) -> None:
        self._session.close()


# ---------------------------------------------------------------------------------------------------------------------
# infrastructure.gpu.backends.base.Backend
# ---------------------------------------------------------------------------------------------------------------------
def make_compute_backend(base_module: Any = None, detector_module: Any = None):
    """Returns `B200Backend`, a subclass of the reference's compute `Backend` for `register_backend(BackendType.CUDA,
    B200Backend)`.  `load_model(name)` builds the engine for one of the five Real-ESRGAN models, `run_inference(name,
    frames)` runs it on a BGR uint8 frame, a frame stack [N,H,W,3] or a list of frames."""
    base = base_module or importlib.import_module("framewright.infrastructure.gpu.backends.base")
    det = detector_module or importlib.import_module("framewright.infrastructure.gpu.detector")
    cached = _BACKEND_CLASSES.get(base)
    if cached is not None:
        return cached

    class B200Backend(base.Backend):
        def __init__(self, device_id: int = 0):
            super().__init__(device_id)
            self._sessions: Dict[str, UpscaleSession] = {}

        @property
        def backend_type(self):
            return det.BackendType.CUDA

        @property
        def name(self) -> str:
            return "B200 (sm_100a, tcgen05 / TMEM / TMA)"

        def initialize(self) -> bool:
            if self._initialized:
                return True
            self._initialized = bool(is_pytorch_esrgan_available())
            if not self._initialized:
                logger.warning("B200 backend unavailable: libb200sr.so not loadable or no sm_100 device")
            return self._initialized

        def cleanup(self) -> None:
            sessions, self._sessions = self._sessions, {}
            for s in sessions.values():
                s.close()
            self._initialized = False

        def get_memory_info(self) -> Dict[str, float]:
            try:
                import torch

                free, total = torch.cuda.mem_get_info(self.device_id)
                mb = 1024.0 * 1024.0
                return {"total_mb": total / mb, "used_mb": (total - free) / mb, "free_mb": free / mb}
            except Exception:
                return {"total_mb": 0.0, "used_mb": 0.0, "free_mb": 0.0}

        def get_capabilities(self):
            if self._capabilities is None:
                total = int(self.get_memory_info()["total_mb"])
                self._capabilities = base.BackendCapabilities(
                    name=self.name, backend_type=self.backend_type, vendor=det.GPUVendor.NVIDIA,
                    supports_fp16=True, supports_fp32=False, supports_int8=False, supports_dynamic_shapes=True,
                    supports_batching=True, max_memory_mb=total, recommended_memory_mb=int(total * 0.8),
                    max_batch_size=64, max_tile_size=0, supported_models=sorted(MODEL_ARCHS))
            return self._capabilities

        def allocate_memory(self, size_mb: float) -> bool:
            # the engine sizes its own workspace per frame shape (b200sr_workspace_bytes); this answers "would it fit"
            return self.get_memory_info()["free_mb"] >= float(size_mb)

        def free_memory(self) -> None:
            try:
                import torch

                torch.cuda.empty_cache()
            except Exception:
                pass

        def load_model(self, model_name: str, model_path=None, **kwargs) -> bool:
            # which of the five networks: an explicit model= / model_name= option, the name itself (PyTorch or ncnn
            # spelling, pytorch_realesrgan.py:263-275), else the checkpoint's file name; anything else is not this path's
            from .pytorch_realesrgan import NCNN_TO_PYTORCH_MODEL
            from .upsampler import _arch_for_model_path

            wanted = kwargs.pop("model", None) or kwargs.pop("model_name", None) or model_name
            name = wanted if wanted in MODEL_ARCHS else NCNN_TO_PYTORCH_MODEL.get(str(wanted)) \
                or _arch_for_model_path(None if model_path is None else str(model_path))
            if name is None:
                logger.error("B200 backend: %s is not a Real-ESRGAN model of this path", model_name)
                return False
            try:
                session = UpscaleSession(self.device_id, dict(kwargs, model_name=name), model_path=model_path)
                session.open()
            except Exception as e:
                logger.error("B200 backend: loading %s failed: %s", model_name, e)
                return False
            old = self._sessions.pop(model_name, None)
            if old is not None and old.model_path is not None:
                old.close()
            self._sessions[model_name] = session
            return True

        def unload_model(self, model_name: str) -> None:
            session = self._sessions.pop(model_name, None)
            if session is None:
                return
            if session.model_path is not None:
                session.close()                       # its own engine
            elif not any(s.model_path is None for s in self._sessions.values()):
                session.close()                       # the last user of the shared cache clears it
            self.free_memory()

        def run_inference(self, model_name: str, inputs: Any, **kwargs) -> Any:
            session = self._sessions.get(model_name)
            if session is None:
                raise ValueError(f"Model {model_name} not loaded")       # the reference's message (:346-347)
            if isinstance(inputs, np.ndarray) and inputs.ndim == 4:
                return np.stack(session.frames(list(inputs)))
            if isinstance(inputs, (list, tuple)):
                return session.frames(inputs)
            return session.frame(inputs)

    _BACKEND_CLASSES[base] = B200Backend
    return B200Backend


_BACKEND_CLASSES: Dict[Any, Any] = {}


def register_compute_backend(base_module: Any = None, detector_module: Any = None):
    """`register_backend(BackendType.CUDA, B200Backend)` on the reference's registry; returns the class."""
    base = base_module or importlib.import_module("framewright.infrastructure.gpu.backends.base")
    det = detector_module or importlib.import_module("framewright.infrastructure.gpu.detector")
    cls = make_compute_backend(base, det)
    base.register_backend(det.BackendType.CUDA, cls)
    return cls
