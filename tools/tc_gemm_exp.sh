#!/bin/bash
# where does the tensor-core GEMM wait?  the standalone harness built three ways (numbers of the two experiment builds are WRONG
# on purpose -- only their times mean anything)
mkdir -p gpurun_out
for v in ${VARIANTS:-"" "-DPNDF_TC_EXP_HALF_FEED" "-DPNDF_TC_EXP_NO_DRAIN" "-DPNDF_TC_EXP_HALF_FEED+-DPNDF_TC_EXP_NO_DRAIN"}; do
  v=${v//+/ }
  echo "=== build flags: [$v]"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 $v -o gpurun_out/tc_gemm_test tools/tc_gemm_test.cu 2> gpurun_out/tc_gemm_build.log || { cat gpurun_out/tc_gemm_build.log; exit 1; }
  timeout 300 gpurun_out/tc_gemm_test 2>&1 | grep -v "^ *$"
done > gpurun_out/tc_gemm_exp.txt 2>&1
cat gpurun_out/tc_gemm_exp.txt
