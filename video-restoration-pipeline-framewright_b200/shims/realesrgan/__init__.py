"""`realesrgan` shim: RealESRGANer backed by libb200sr.so (see framewright_b200/upsampler.py)."""
from framewright_b200.upsampler import RealESRGANer  # noqa: F401

__version__ = "0.3.0+b200sr"   # read by framewright.utils.dependencies.check_realesrgan (:318-342)
__b200sr_shim__ = True
__all__ = ["RealESRGANer"]
