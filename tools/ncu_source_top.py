#!/usr/bin/env python
"""Top SASS lines by stall samples / executed instructions from an `ncu --page source --csv` export.
usage: ncu_source_top.py file.csv <kernel-substring> [top_n] [occurrence]"""
import csv
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(path)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    sel = [s for s in secs if pat in s["name"]]
    if not sel:
        print("no kernel matching", pat, "in", [s["name"][:60] for s in secs]); return
    s = sel[occ]
    h = {k: i for i, k in enumerate(s["hdr"])}
    def f(r, k):
        try: return float(r[h[k]])
        except Exception: return 0.0
    tot_s = sum(f(r, "# Samples") for r in s["rows"]); tot_i = sum(f(r, "Instructions Executed") for r in s["rows"])
    print(s["name"][:100]); print("SASS lines", len(s["rows"]), "samples", tot_s, "warp-insts", tot_i)
    print("--- by samples")
    for r in sorted(s["rows"], key=lambda r: -f(r, "# Samples"))[:top]:
        print("%6.2f%% smp %6.2f%% inst thr=%4.1f  %s" % (100 * f(r, "# Samples") / max(tot_s, 1), 100 * f(r, "Instructions Executed") / max(tot_i, 1),
                                                   f(r, "Avg. Threads Executed"), r[h["Source"]].strip()[:90]))
    print("--- by instructions executed")
    for r in sorted(s["rows"], key=lambda r: -f(r, "Instructions Executed"))[:top]:
        print("%6.2f%% smp %6.2f%% inst thr=%4.1f  %s" % (100 * f(r, "# Samples") / max(tot_s, 1), 100 * f(r, "Instructions Executed") / max(tot_i, 1),
                                                   f(r, "Avg. Threads Executed"), r[h["Source"]].strip()[:90]))


def by_inst():
    pass


if __name__ == "__main__":
    main()
