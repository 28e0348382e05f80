#!/bin/bash
# round 2, call 16: multicast-cluster GEMM variants + int8 path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in 0 128x2 128x4 192x2 192x4; do
  B200_GEMM_CLUSTER=$v timeout 120 python scripts/gemm_bench.py 2>&1 | tail -5
done
} > gpurun_out/gemm_cluster.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "int8 or quantize_i8 or gemm" > gpurun_out/t_i8.log 2>&1
timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "quantized" -s > gpurun_out/t_q.log 2>&1
tail -3 gpurun_out/t_i8.log gpurun_out/t_q.log
cat gpurun_out/gemm_cluster.log
