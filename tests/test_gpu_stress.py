"""Race / bounds hardening of the fused residual-dense-block kernel's cross-CTA protocol, in place of
compute-sanitizer (closed on this pool): tools/stress_fused.py -- 60 randomized iterations (ragged shapes around every
strip / flag-block / column-tile boundary, random batch, `max_ctas` from 148 down to ONE resident CTA, fused vs
per-conv bytes) -- executed with the DEBUG build of the library, which traps on any out-of-range completion-counter
index, claimed-item field or TMA coordinate."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("lib", ["libb200sr_debug.so", "libb200sr.so"])
def test_fused_schedule_stress(lib):
    env = dict(os.environ, B200SR_LIB=lib)
    iters = "60" if "debug" in lib else "30"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_fused.py"), "--iters", iters, "--seed",
                        "11" if "debug" in lib else "12"], env=env, capture_output=True, text=True, timeout=900)
    tail = p.stdout[-2500:] + p.stderr[-1500:]
    assert p.returncode == 0, tail
    assert "STRESS ok" in p.stdout and "MISMATCH" not in p.stdout and "B200SR_DEBUG trap" not in p.stdout, tail
    if "debug" in lib:
        assert "DEBUG: bounds traps on" in p.stdout, tail
