#!/bin/bash
# Run on the GPU box through gpurun: each group in its own process + timeout so one hung kernel
# cannot take the rest of the call with it.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1 to=$2; shift 2
  timeout -s KILL $to python -m pytest "$@" -q -m gpu -p no:cacheprovider -rA --timeout 240 > gpurun_out/$name.log 2>&1
  echo "== $name exit $?"; tail -n 25 gpurun_out/$name.log
}
run gemm_mma 300 tests/test_gpu_kernels.py -k "gemm and mma"
run gemm_tc 300 tests/test_gpu_kernels.py -k "gemm and tcgen05"
run stem 200 tests/test_gpu_kernels.py -k "conv_stem or error"
run lstm 300 tests/test_gpu_kernels.py -k "lstm"
run decode 300 tests/test_gpu_kernels.py -k "crf_decode"
B200_GEMM_IMPL=mma run pipeline_mma 600 tests/test_gpu_pipeline.py -s -k "forward_scores or decode_of_own"
run pipeline 900 tests/test_gpu_pipeline.py -s
