#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -s -k "old_style or gemm_matches or forward_scores" 2>&1 | grep -v "^$" | tail -n 12
