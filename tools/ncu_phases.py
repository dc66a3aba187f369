#!/usr/bin/env python
"""Share of stall samples (~ time at this occupancy), executed warp instructions and shared-memory wavefronts per PHASE of a
straight-line tile loop, from the SASS source page of an `ncu --set full` capture.  Phases are ranges of SASS instruction indices
(address order = program order); `--markers` prints the memory / sync instructions with their index and the cumulative sample
share, which is how the boundaries of a build are found.
usage: python tools/ncu_phases.py <rep> --markers
       python tools/ncu_phases.py <rep> "name:first-last" ..."""
import csv, re, subprocess, sys
rep, args = sys.argv[1], sys.argv[2:]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r)
hdr = rows[hi]
si, sa, ie, iw = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
body = [r for r in rows[hi + 1:] if len(r) > iw]
def f(x):
    try: return float(x)
    except ValueError: return 0.0
tot = [sum(f(r[k]) for r in body) for k in (sa, ie, iw)]
print(f"{len(body)} SASS instructions, {tot[0]:.0f} samples, {tot[1]:.0f} warp instructions, {tot[2]:.0f} shared-memory wavefronts")
if args and args[0] == "--markers":
    cum = 0.0
    for i, r in enumerate(body):
        cum += f(r[sa])
        s = r[si].strip()
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        if op in ("LDS", "STS", "UBLKCP", "SYNCS", "MUFU", "SHFL", "BAR", "FENCE", "MEMBAR", "ELECT", "CREDUX", "BRA", "LDG", "STG"):
            print(f"{i:5d} {100 * cum / tot[0]:6.1f}%  {f(r[ie]):10.0f}  {s[:80]}")
else:
    print("| phase | SASS range | stall samples | warp instructions | shared-memory wavefronts |\n|---|---|---:|---:|---:|")
    for a in args:
        name, rng = a.rsplit(":", 1)
        lo, hi_ = (int(x) for x in rng.split("-"))
        s = [sum(f(r[k]) for r in body[lo:hi_]) for k in (sa, ie, iw)]
        print(f"| {name} | {lo}-{hi_} | {100 * s[0] / tot[0]:.1f} % | {100 * s[1] / tot[1]:.1f} % | {100 * s[2] / max(tot[2], 1):.1f} % |")
