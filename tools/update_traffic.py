#!/usr/bin/env python
"""profiles/traffic.json from the launch lists of tools/measure_round.sh: mean DRAM bytes per launch of the hot kernel of each workload.
usage: python tools/update_traffic.py gpurun_out/m r2   (copies the CSVs to profiles/<tag>_launches_<workload>.csv)"""
import csv, json, os, shutil, sys
src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for wl in ("cfg2", "cfg3", "cfg4shard"):
    path = os.path.join(src, f"launches_{wl}.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    by = {}
    for r in rows:
        by.setdefault((r[0], r[4]), {})[r[12]] = float(r[14])
    hot = [(k, v) for k, v in by.items() if "melspec400" in k[1] or "melspec512" in k[1]]
    n = len(hot)
    rd = sum(v["dram__bytes_read.sum"] for _, v in hot) / n
    wr = sum(v["dram__bytes_write.sum"] for _, v in hot) / n
    t = sum(v["gpu__time_duration.sum"] for _, v in hot) / n
    others = sorted({k[1] for k in by if "melspec400" not in k[1] and "melspec512" not in k[1]})
    dst = os.path.join(ROOT, "profiles", f"{tag}_launches_{wl}.csv")
    shutil.copy(path, dst)
    out[wl] = {"kernel": hot[0][0][1].replace("void ", "").replace("(KParams)", ""), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
               "traffic": int(rd + wr), "ncu_us_per_launch": t / 1e3, "launches_averaged": n, "other_library_kernels_in_step": others,
               "source": f"profiles/{tag}_launches_{wl}.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum --clock-control none -k regex:melspec, mean over the hot kernel's launches)"}
    print(wl, out[wl]["kernel"], f"read {rd/1e6:.1f} MB write {wr/1e6:.1f} MB  {t/1e3:.1f} us  n={n}", others)
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
