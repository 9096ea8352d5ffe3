"""Shim packages that let the UNMODIFIED reference files drive the B200 engine.

The reference imports its (absent, un-pinned) third-party dependencies by name:
    from realesrgan import RealESRGANer                      (pytorch_realesrgan.py:73,99; cli.py:707)
    from basicsr.archs.rrdbnet_arch import RRDBNet           (pytorch_realesrgan.py:74,100; cli.py:708)
`install()` puts this directory's `realesrgan` / `basicsr` packages at the front of sys.path (or registers
them in sys.modules), which is the same seam the reference's own tests use (tests/test_cli.py:287-292).
"""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def install() -> None:
    """Make `import realesrgan` / `import basicsr` resolve to the B200 shims."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    for name in ("realesrgan", "basicsr", "basicsr.archs", "basicsr.archs.rrdbnet_arch", "realesrgan.archs",
                 "realesrgan.archs.srvgg_arch"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__b200sr_shim__", False):
            del sys.modules[name]
