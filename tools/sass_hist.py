#!/usr/bin/env python
"""Static opcode histogram of one kernel of a built library: python tools/sass_hist.py <lib.so> <kernel-name-substring>
(the plan-400 / plan-512 kernels are one straight-line loop body per pass plus setup, so static counts track the dynamic ones)."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist, tot = None, collections.Counter(), 0
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(2).split(".")[0]
            full = m.group(2)
            key = full if op in ("LDS", "STS", "LDL", "STL", "SHFL") else op
            hist[key] += 1
            tot += 1
print(pat, "total", tot)
for k, v in hist.most_common(45):
    print(f"  {k:14s} {v}")
