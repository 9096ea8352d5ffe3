#!/usr/bin/env python
"""Writes profiles/<name>: mnemonic counts of the built library's SASS (no GPU needed): python tools/sass_summary.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "video-restoration-pipeline-framewright_b200", "libb200sr.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02b_sass_summary.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
mn = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG.4D", "UTMALDG.5D", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA.16816", "HGMMA",
      "SYNCS.ARRIVE", "MEMBAR", "CCTL.IVALL", "REDG", "ATOMG", "FFMA2", "LDCU", "STS.128"]
lines = [f"# cuobjdump -sass video-restoration-pipeline-framewright_b200/libb200sr.so | grep -c <mnemonic>   (tools/sass_summary.py)"]
for m in mn:
    lines.append(f"{m:14s} {len(re.findall(r'\b' + re.escape(m) + r'\b', sass))}")
lines.append("")
lines.append("# per kernel: UTCHMMA / LDTM / UTMALDG / UBLKCP counts")
cur, per = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur:
        for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"):
            if re.search(r"\b" + k, ln):
                per[cur][k] += 1
for k, c in per.items():
    lines.append(f"{k:110s} UTCHMMA={c['UTCHMMA']} LDTM={c['LDTM']} UTMALDG={c['UTMALDG']} UBLKCP={c['UBLKCP']}")
open(out, "w").write("\n".join(lines) + "\n")
print(out)
