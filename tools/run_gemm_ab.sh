#!/bin/bash
# A/B of GEMM variants: trailing-update shaped products and the N = 16384 Cholesky / cluster single point.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/gemm_ab.log
for lib in "" "$@"; do
  if [ -n "$lib" ]; then export MCHRG_B200_LIB=$PWD/$lib; else unset MCHRG_B200_LIB; fi
  echo "=== ${lib:-default}" >> gpurun_out/gemm_ab.log
  timeout 300 python tools/prof_gemm.py 15360 512 2>&1 | grep dgemm >> gpurun_out/gemm_ab.log
  timeout 300 python tools/prof_gemm.py 8192 8192 2>&1 | grep dgemm >> gpurun_out/gemm_ab.log
  timeout 300 python tools/time_cluster.py 16384 2>&1 | tail -4 >> gpurun_out/gemm_ab.log
done
cat gpurun_out/gemm_ab.log
