#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tc_gemm_test tools/tc_gemm_test.cu 2> gpurun_out/tc_gemm_build.log || { cat gpurun_out/tc_gemm_build.log; exit 1; }
timeout 300 gpurun_out/tc_gemm_test > gpurun_out/tc_gemm_test.txt 2>&1; cat gpurun_out/tc_gemm_test.txt
