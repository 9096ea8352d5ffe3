"""pytest plugin (tests only): makes `framewright.utils.multi_gpu` / `framewright.processors.pytorch_realesrgan`
resolve to THIS repo's mirrors, so that the reference's own, unmodified test files can be run against them
(`python -m pytest -p ref_alias_plugin --noconftest /root/reference/tests/test_multi_gpu.py`)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import framewright_b200  # noqa: E402,F401
from framewright_b200 import multi_gpu, pytorch_realesrgan  # noqa: E402

for name in ("framewright", "framewright.utils", "framewright.processors"):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules.setdefault(name, m)
sys.modules["framewright.utils.multi_gpu"] = multi_gpu
sys.modules["framewright.utils"].multi_gpu = multi_gpu
sys.modules["framewright.processors.pytorch_realesrgan"] = pytorch_realesrgan
sys.modules["framewright.processors"].pytorch_realesrgan = pytorch_realesrgan
