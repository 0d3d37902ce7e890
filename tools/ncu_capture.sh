#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count>   (run on the GPU box under gpurun)
# Full-set capture of the selected kernels; exports the raw + source pages as CSV next to the report and
# drops the report when it is too large to travel back (gpurun_out/ is capped at 64 MiB).
set -u
tag=$1; rx=$2; skip=$3; cnt=$4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" -f -o "gpurun_out/$tag" \
    python tools/prof_step.py --steps 5 > "gpurun_out/$tag.log" 2>&1
echo "ncu rc=$?"
ncu -i "gpurun_out/$tag.ncu-rep" --page raw --csv > "gpurun_out/$tag.raw.csv" 2>/dev/null
ncu -i "gpurun_out/$tag.ncu-rep" --page source --csv > "gpurun_out/$tag.source.csv" 2>/dev/null
sz=$(stat -c %s "gpurun_out/$tag.ncu-rep" 2>/dev/null || echo 0)
if [ "$sz" -gt 12000000 ]; then rm -f "gpurun_out/$tag.ncu-rep"; echo "report dropped ($sz B), CSV pages kept"; fi
ls -la gpurun_out | grep "$tag"
