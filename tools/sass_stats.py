#!/usr/bin/env python3
"""Instruction histogram per kernel from `cuobjdump -sass` (dev-container check before GPU time).

    python tools/sass_stats.py libflagstats_b200/libflagstats_cuda.so [substring-of-mangled-name]
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None
funcs = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in funcs.items():
    if pat not in name:
        continue
    h = collections.Counter()
    for _, t in ins:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        h[t.split()[0].split(".")[0]] += 1
    print(name, "total", len(ins))
    print("  " + "  ".join(f"{k}:{v}" for k, v in h.most_common(14)))
