#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for a in 0 2 4 8 6 10; do GEMM_ABLATE=$a timeout 100 python scripts/gemm_profile.py; done > gpurun_out/gemm_profile.log 2>&1
cat gpurun_out/gemm_profile.log
