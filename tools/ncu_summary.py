#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into the handful of numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 30]"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_tma_st.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d:
                print(f"  {k:90s} {d[k]:>16s} {units[hdr.index(k)]}")
        st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")}
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]:
            print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.3f} warps/issue")
    src = page(rep, "source")
    if len(src) > 2:
        hdr = src[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = src[2:]
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((h, sum(int(r[ix[h]]) for r in data)) for h in stalls), key=lambda kv: -kv[1])
        print(f"== source page: {tot} samples over {len(data)} SASS instructions")
        print("  " + ", ".join(f"{h[6:]}={v}" for h, v in agg if v))
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
            s = sorted(((h[6:], int(r[ix[h]])) for h in stalls if int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:2]
            print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:64]:64s} {s}")
        print("  -- shared-memory wavefronts by instruction (excess = bank conflicts)")
        for r in sorted(data, key=lambda r: -int(r[ix["L1 Wavefronts Shared Excessive"]] or 0))[:12]:
            print(f"  {r[ix['L1 Wavefronts Shared Excessive']]:>9s} {r[ix['L1 Wavefronts Shared']]:>9s} {r[ix['L1 Wavefronts Shared Ideal']]:>9s}  {r[ix['Source']].strip()[:64]}")


if __name__ == "__main__":
    main()
