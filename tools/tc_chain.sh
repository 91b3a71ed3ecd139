#!/bin/bash
# measured 3xTF32 tcgen05 chain experiment: run under gpurun, results -> gpurun_out/tc_chain_<act>.txt
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tc_chain_probe tools/tc_chain_probe.cu 2> gpurun_out/tc_chain_build.log || exit 1
for act in lrelu softplus; do
  python tools/tc_chain_export.py $act /tmp/chain_$act.bin > /dev/null
  timeout 300 gpurun_out/tc_chain_probe /tmp/chain_$act.bin $act > gpurun_out/tc_chain_$act.txt 2>&1
  cat gpurun_out/tc_chain_$act.txt
done
