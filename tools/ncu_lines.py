#!/usr/bin/env python
"""Per CUDA-C line totals (samples, warp instructions) of one kernel from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py report.ncu-rep <kernel-regex> [top_n]"""
import csv, io, subprocess, sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{rx}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, agg, src = None, None, {}, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; hdr = None; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {k: i for i, k in enumerate(r)}; continue
        if hdr is None or not r[0].isdigit():
            continue
        # a line row has Source filled at index 1; SASS rows follow with the same Line No
        try:
            smp = float(r[hdr["# Samples"]] or 0); ins = float(r[hdr["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        key = (cur_file, int(r[0]))
        if r[2] == "-":     # the CUDA line itself carries the totals of its SASS
            agg[key] = (smp, ins); src[key] = r[1].strip()
    tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[1] for v in agg.values())
    print("lines", len(agg), "samples", tot_s, "warp-insts", tot_i)
    for title, idx in (("by samples", 0), ("by instructions", 1)):
        print("---", title)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
            print("%6.2f%% smp %6.2f%% inst  %s:%d  %s" % (100 * v[0] / max(tot_s, 1), 100 * v[1] / max(tot_i, 1), k[0], k[1], src[k][:100]))


if __name__ == "__main__":
    main()
