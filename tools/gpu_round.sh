#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, full-set capture of the top kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/prof_step.py --steps 2 > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?"
timeout 400 bash tools/ncu_capture.sh ${tag}_full "vcn_chain_kernel|points_in_boxes_kernel|dynvox_|knn_scan_kernel|knn_prepare_kernel|largest_cluster_kernel|splice_mask_kernel|vcn_linear_tc|crop_" 60 40
