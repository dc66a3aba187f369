#!/usr/bin/env python
"""Dynamic SASS opcode histogram (warp instructions per pass) from an .ncu-rep source page.
usage: python tools/ncu_ophist.py rep.ncu-rep passes"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; passes = float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
h = collections.Counter(); samples = collections.Counter()
for r in rows[2:]:
    src = r[ix["Source"]].strip()
    t = src.split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.rstrip(";")
    h[op] += int(r[ix["Instructions Executed"]]); samples[op] += int(r[ix["# Samples"]])
tot = sum(h.values())
print(f"total {tot/passes:.1f} warp-instructions per pass")
for op, n in h.most_common(45):
    print(f"  {op:24s} {n/passes:8.2f}  samples {samples[op]}")
