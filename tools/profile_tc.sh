#!/bin/bash
# ncu evidence for the tensor-core path (run under gpurun, 1 GPU):  tools/profile_tc.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
# (1) launch list with device time per launch: bench.py's default workload takes the tensor-core path
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_tc_${TAG}.log 2>&1
# (2) full capture of every kernel of the second step (the first is the warm-up)
ncu --set full --clock-control none --import-source on -k regex:'tc_' -s 16 -c 15 -o gpurun_out/prof_tc_${TAG} -f \
    python tools/tc_step.py lrelu 65536 2 > gpurun_out/tc_step_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
