#!/bin/bash
# compute-sanitizer pass over a subset of the GPU parity tests (memcheck + racecheck per kernel).
# usage (under gpurun): bash tools/gpu_sanitize.sh <tag> "<pytest -k expr>" [tools...]
# Summaries land in gpurun_out/<tag>_<tool>.log; copy the tails worth keeping into profiles/.
set -u
tag=$1; kexpr=$2; shift 2
tools=${*:-memcheck racecheck}
mkdir -p gpurun_out
for tool in $tools; do
  timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 \
      python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$kexpr" > "gpurun_out/${tag}_${tool}.log" 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" "gpurun_out/${tag}_${tool}.log" | tail -8
done
