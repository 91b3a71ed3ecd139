#!/bin/bash
# ncu evidence for the fused kernel (run under gpurun, 1 GPU):  tools/profile.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list with device time per launch
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# (2) full capture of the fused kernel (timed-region launches: skip the warm-ups)
ncu --set full --clock-control none --import-source on -k regex:pndf_fused -s 3 -c 2 -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
