#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/step_timeline.py > gpurun_out/step_timeline.log 2>&1
TL_SLOTS=1 timeout 300 python scripts/step_timeline.py 2>&1 | head -40 > gpurun_out/step_timeline1.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
cat gpurun_out/step_timeline.log
