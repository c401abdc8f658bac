#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the design notes quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [regex ...]"""
import csv
import io
import re
import subprocess
import sys

DEFAULT = [r"gpu__time_duration\.sum", r"dram__bytes_(read|write)\.sum$", r"dram__throughput.avg.pct", r"lts__t_bytes\.sum$",
           r"sm__throughput.avg.pct", r"sm__warps_active.avg.pct", r"launch__registers_per_thread", r"launch__grid_size",
           r"launch__occupancy_limit", r"sm__inst_executed_pipe_(fma|fmaheavy|lsu|alu)\b.*pct", r"sm__pipe_fma.*cycles_active.*pct",
           r"sm__pipe_tensor.*cycles_active.*pct",
           r"smsp__issue_active.avg.pct", r"smsp__inst_executed.sum$", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$",
           r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$", r"smsp__average_warps_issue_stalled.*_per_issue_active",
           r"smsp__average_warp_latency_issue_stalled", r"sm__cycles_elapsed.max", r"l1tex__t_sector_hit_rate",
           r"lts__t_sector_hit_rate", r"smsp__inst_executed_op_shared", r"sm__sass_inst_executed_op_global_red",
           r"launch__shared_mem_per_block", r"smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep = sys.argv[1]
    pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(p.search(h) for p in pats):
                print(f"  {h.split('.', 2)[-1] if h.count('.') > 2 else h:95s} {units[i]:14s} {r[i]}")


if __name__ == "__main__":
    main()
