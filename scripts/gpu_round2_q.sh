#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 192x8 128x8; do
  B200_GEMM_CLUSTER=$v timeout 120 python scripts/gemm_bench.py 2>&1 | tail -4
done > gpurun_out/gemm_cluster8.log 2>&1
cat gpurun_out/gemm_cluster8.log
