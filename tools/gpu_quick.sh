#!/bin/bash
# Short GPU-box visit: parity tests + bench line (no profiler).  usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
set -u
tag=${1:-quick}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/${tag}_pytest.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
fi
echo "pytest rc=$?"; tail -25 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"roof",d["roofline"]["frac"])
    for s in d["stages"]: print("  %-28s %8.4f ms  share %.3f  %s" % (s["group"],s["ms_per_step"],s["share_of_step"],("%.1f %s frac %.3f"%(s["achieved"],s["unit"],s["frac"])) if "achieved" in s else ""))
    print("cpu", d["cpu_baseline"])
except Exception as e: print("no bench line", e)
PY
