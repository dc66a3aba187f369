#!/usr/bin/env python
"""Static issue-stall accounting: decode the stall field (bits 105..108) of every SASS instruction of one kernel in the
built library and weight it with the per-instruction execution counts of an .ncu-rep source page.
usage: python tools/sass_stalls.py lib.so mangled_kernel_substring rep.ncu-rep passes"""
import csv, io, re, subprocess, sys, collections
lib, key, rep, passes = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if "Function :" in l and key in l)
end = next((i for i in range(start + 1, len(sass)) if "Function :" in sass[i]), len(sass))
ins = []
i = start
pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
while i < end:
    m = pat.match(sass[i])
    if m:
        hi = re.search(r"/\* 0x([0-9a-f]{16}) \*/", sass[i + 1])
        hiw = int(hi.group(1), 16)
        stall = (hiw >> 41) & 0xF
        yld = (hiw >> 45) & 1
        ins.append((int(m.group(1), 16), m.group(2).strip(), stall, yld))
        i += 2
    else:
        i += 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: k for k, h in enumerate(hdr)}
data = rows[2:]
assert len(data) == len(ins), (len(data), len(ins))
tot_stall = 0; tot_ins = 0
byop = collections.Counter(); byop_n = collections.Counter()
hist = collections.Counter()
for (addr, txt, stall, yld), r in zip(ins, data):
    n = int(r[ix["Instructions Executed"]])
    tot_stall += n * max(stall, 1); tot_ins += n
    t = txt.split(); op = t[1] if t[0].startswith("@") else t[0]
    byop[op] += n * max(stall, 1); byop_n[op] += n
    hist[stall] += n
print(f"instructions/pass {tot_ins/passes:.1f}   sum of stall fields/pass {tot_stall/passes:.1f}")
print("stall histogram (cycles: instr/pass):", {k: round(v / passes, 1) for k, v in sorted(hist.items())})
for op, v in byop.most_common(14):
    print(f"  {op:20s} n/pass {byop_n[op]/passes:7.1f}  stall cycles/pass {v/passes:8.1f}  avg {v/max(byop_n[op],1):.2f}")
