#!/usr/bin/env python3
"""Summaries for profiles/: (1) launch list CSV (ncu --metrics gpu__time_duration.sum) -> per-kernel
share table; (2) full .ncu-rep -> key counters of the EM kernel.
    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv > profiles/launches_r01.txt
    python tools/summarize_ncu.py full gpurun_out/em_r01_c.ncu-rep > profiles/em_r01_c.txt"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    t = collections.defaultdict(float)
    n = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        val = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = r[ik].split("(")[0]
        t[name] += val * scale
        n[name] += 1
    tot = sum(t.values())
    print("%-70s %8s %12s %8s" % ("kernel", "launches", "total ms", "share"))
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        print("%-70s %8d %12.3f %7.2f%%" % (k[:70], n[k], v, 100 * v / tot))
    print("%-70s %8d %12.3f" % ("TOTAL", sum(n.values()), tot))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("# kernel:", r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-90s %-12s %s" % (w, units[i], r[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
    h = collections.Counter()
    for r in data:
        toks = r[ia].strip().split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        h[op] += int(r[ie])
    tot = sum(h.values())
    print("# executed warp-instructions by opcode (SASS), total %d" % tot)
    for op, c in h.most_common(16):
        print("%-10s %14d %6.2f%%" % (op, c, 100.0 * c / tot))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
