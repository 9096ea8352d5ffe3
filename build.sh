#!/bin/bash
# Build libb200sr.so (sm_100a) in-tree; equivalent to `python -c "import __graft_entry__ as g; g.build()"`.
set -e
cd "$(dirname "$0")"
python - <<'PY'
import framewright_b200
from framewright_b200 import _native
print(_native.build(force=True, verbose=True))
PY
