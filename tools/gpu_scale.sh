#!/bin/bash
# Multi-GPU bench lines.  usage (under gpurun --gpus N): bash tools/gpu_scale.sh <tag> <N> [variants...]
# variants: peer nccl none (how the end-of-path collection runs)
set -u
tag=$1; n=$2; shift 2
vars=${*:-peer nccl none}
mkdir -p gpurun_out
for v in $vars; do
  case $v in
    none) extra="--no-gather" ;;
    *) extra="--gather $v" ;;
  esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 100 --warmup 5 $extra > gpurun_out/${tag}_n${n}_${v}.json 2> gpurun_out/${tag}_n${n}_${v}.err
  echo "$v rc=$?"
  tail -3 gpurun_out/${tag}_n${n}_${v}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n${n}_${v}.json").read().strip().splitlines()[-1])
    print("$v n=$n value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"gather",d["config"]["gather"])
    pr=d.get("per_rank_ms")
    if pr:
        print("  cols",pr["columns"])
        for r in pr["rows"]: print("  ",[round(x,3) for x in r])
except Exception as e: print("no bench line", e)
PY
done
