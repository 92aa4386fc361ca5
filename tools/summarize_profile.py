#!/usr/bin/env python3
"""Turn ncu outputs into the markdown summaries kept under profiles/.
  summarize_profile.py launches <launches.csv>        per-kernel time shares of one bench step
  summarize_profile.py full <report.ncu-rep> [regex]  key metrics + stall breakdown of one kernel
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
        agg.setdefault(name, []).append(v)
    ours = {k: v for k, v in agg.items() if k.startswith("nm_")}
    tot = sum(sum(v) for v in ours.values())
    print("| kernel | launches | avg us | share of nm_* time |")
    print("|---|---|---|---|")
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    other = {k: v for k, v in agg.items() if not k.startswith("nm_")}
    print("\nother launches in the capture (workload generation by torch, not part of a step): %d" % sum(len(v) for v in other.values()))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("### %s  (grid %s x block %s)\n" % (d.get("Kernel Name", "?").split("(")[0], d.get("Grid Size"), d.get("Block Size")))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d and d[k] != "":
                print("| %s | %s | %s |" % (k, d[k], u[k]))
        st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v}
        print("\nwarp-stall reasons (warps stalled per issue-active cycle):\n")
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]:
            print("* %s: %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2])
