#!/usr/bin/env python
"""List every shared-memory SASS instruction of an .ncu-rep source page in program order with its per-pass wavefronts.
usage: python tools/ncu_smem.py rep.ncu-rep [passes]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
passes = float(sys.argv[2]) if len(sys.argv) > 2 else max(int(r[ix["Instructions Executed"]]) for r in data if "LDS" in r[ix["Source"]] or "STS" in r[ix["Source"]])
tot = 0; totid = 0
cls = {}
for n, r in enumerate(data):
    wf = int(r[ix["L1 Wavefronts Shared"]] or 0)
    if wf == 0: continue
    ideal = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    ex = int(r[ix["Instructions Executed"]])
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    print(f"{n:5d} exec/pass {ex/passes:5.2f} wf/pass {wf/passes:6.2f} ideal {ideal/passes:6.2f}  {src[:70]}")
    tot += wf; totid += ideal
    c = cls.setdefault(op, [0, 0, 0]); c[0] += ex; c[1] += wf; c[2] += ideal
print(f"total wf/pass {tot/passes:.1f} ideal {totid/passes:.1f}")
for op, c in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print(f"  {op:12s} exec/pass {c[0]/passes:7.2f} wf/pass {c[1]/passes:7.2f} ideal {c[2]/passes:7.2f}")
