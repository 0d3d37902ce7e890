#!/usr/bin/env python
"""Copies the artefacts of a GPU visit from gpurun_out/ (scratch) into profiles/ (tracked) under round names and derives
the summaries the docs quote: launch shares of the last captured step, the ncu full-set table, DRAM bytes per launch.
usage: python tools/collect_profiles.py <round> <tag-of-gpu_round.sh> [key=gpurun_out-file ...]
e.g.   python tools/collect_profiles.py r02 r02m scale_n8_peer=r02q_n8_peer.json"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_line_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main():
    rnd, tag = sys.argv[1], sys.argv[2]
    extra = dict(a.split("=", 1) for a in sys.argv[3:])
    def cp(src, dst):
        s = os.path.join(G, src)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(P, dst)); print("copied", src, "->", dst)
    cp(f"{tag}_bench.json", f"{rnd}_bench_final.json")
    cp(f"{tag}_bench_ref.json", f"{rnd}_bench_reference_arm.json")
    cp(f"{tag}_launches.csv", f"{rnd}_launches_final.csv")
    cp(f"{tag}_smoke.log", f"{rnd}_smoke.log")
    for k, v in extra.items():
        cp(v, f"{rnd}_{k}" + os.path.splitext(v)[1])
    # launch shares of the last step in the launch list
    lp = os.path.join(G, f"{tag}_launches.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
        hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        seq = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:] if len(r) > vi and r[vi].replace(",", "").replace(".", "").isdigit()]
        # last step = launches after the last points_in_boxes_kernel
        starts = [i for i, (n, _) in enumerate(seq) if n.startswith("points_in_boxes_kernel") or "points_in_boxes_kernel" in n]
        step = seq[starts[-1]:] if starts else seq
        agg = collections.OrderedDict()
        for n, us in step:
            short = n.split("(")[0].split("::")[-1]
            c, t = agg.get(short, (0, 0.0)); agg[short] = (c + 1, t + us)
        tot = sum(t for _, t in agg.values())
        with open(os.path.join(P, f"{rnd}_launches_final_shares.txt"), "w") as f:
            f.write("ncu launch list, last captured step of tools/prof_step.py (gpu__time_duration.sum, --clock-control none; cold-cache, serialised):\n\n")
            for n, (c, t) in agg.items():
                f.write("%-60s x%3d %9.1f us %5.1f%%\n" % (n[:60], c, t, 100 * t / tot))
            f.write("\ntotal %.1f us in %d launches\n" % (tot, sum(c for c, _ in agg.values())))
        print("wrote launch shares")
    raw = os.path.join(G, f"{tag}_full.raw.csv")
    if os.path.exists(raw):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{rnd}_ncu_full_final_summary.txt"), "w").write(
            "ncu --set full --clock-control none --import-source on, tools/prof_step.py (one row per captured launch; tools/ncu_summary.py)\n\n" + out)
        rows = list(csv.reader(open(raw)))
        h = rows[0]
        def col(name):
            return h.index(name) if name in h else -1
        kn, gs = col("Kernel Name"), col("Grid Size")
        rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
        traffic = collections.OrderedDict()
        units = rows[1]
        def to_bytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        for r in rows[2:]:
            if len(r) <= max(rd, wr) or rd < 0:
                continue
            name = r[kn].split("(")[0].split("::")[-1]
            try:
                b = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
            except ValueError:
                continue
            e = traffic.setdefault(name, {"dram_bytes_per_launch": 0.0, "grid": r[gs], "launches_captured": 0})
            e["dram_bytes_per_launch"] += b; e["launches_captured"] += 1
        for e in traffic.values():
            e["dram_bytes_per_launch"] /= max(e["launches_captured"], 1)
        tpath = os.path.join(P, "ncu_traffic.json")
        if os.path.exists(tpath):      # kernels of other captures (e.g. the C4 FPS kernel) keep their entries
            for k, v in json.load(open(tpath)).items():
                traffic.setdefault(k, v)
        json.dump(traffic, open(tpath, "w"), indent=1)
        print("wrote ncu summary + ncu_traffic.json", list(traffic)[:6])


if __name__ == "__main__":
    main()
