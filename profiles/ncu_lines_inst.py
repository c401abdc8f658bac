#!/usr/bin/env python
"""Aggregate an .ncu-rep's executed warp instructions by CUDA source line (needs -lineinfo + --import-source).
usage: python profiles/ncu_lines_inst.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur, hdr = None, None
    samples, insts, text = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            si = hdr.index("Warp Stall Sampling (All Samples)")
            ii = hdr.index("Instructions Executed")
        elif hdr and r[0].isdigit():
            key = (cur, int(r[0]))
            samples[key] += int(r[si]) if r[si].isdigit() else 0
            try:
                insts[key] += float(r[ii])
            except ValueError:
                pass
            text[key] = r[1].strip()[:100]
    tot = sum(samples.values())
    tot_i = sum(insts.values())
    print(f"total stall samples {tot}, warp instructions {tot_i:.0f}")
    for k, v in insts.most_common(top):
        print(f"{100 * v / tot_i:5.1f}% inst {100 * samples[k] / tot:5.1f}% samples  {k[0]}:{k[1]:<4d} {text[k]}")


if __name__ == "__main__":
    main()
