#!/bin/bash
# One bench line per BASELINE configuration (short runs).  usage (under gpurun): bash tools/gpu_configs.sh <tag> [extra bench args]
set -u
tag=${1:-cfg}; shift || true
mkdir -p gpurun_out
for c in C3 C4 C5; do
  timeout 600 python bench.py --config $c "$@" > gpurun_out/${tag}_bench_${c}.json 2> gpurun_out/${tag}_bench_${c}.err; echo "$c rc=$?"
  tail -3 gpurun_out/${tag}_bench_${c}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${c}.json").read().strip().splitlines()[-1])
    print("$c value",round(d["value"],1),d["unit"],"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],3),"roof",d["roofline"]["frac"])
    for s in d["stages"]: print("  %-28s %8.4f ms  share %.3f  %s" % (s["group"],s["ms_per_step"],s["share_of_step"],("%.1f %s frac %.3f"%(s["achieved"],s["unit"],s["frac"])) if "achieved" in s else ""))
    print("  cpu", d["cpu_baseline"])
except Exception as e: print("no bench line", e)
PY
done
