#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
B200_GEMM_EPI=direct timeout 120 python scripts/gemm_bench.py 2>&1 | tail -4
B200_GEMM_PAIR=1 B200_GEMM_EPI=direct timeout 120 python scripts/gemm_bench.py 2>&1 | tail -4
B200_GEMM_PAIR=1 B200_GEMM_EPI=direct timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k gemm 2>&1 | tail -3
} > gpurun_out/gemm_pair2.log 2>&1
cat gpurun_out/gemm_pair2.log
