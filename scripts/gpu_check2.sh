#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to python -m pytest "$@" -q -m gpu -p no:cacheprovider -rA --timeout 200 > gpurun_out/$name.log 2>&1
  echo "== $name exit $?"; grep -E "PASSED|FAILED|ERROR|passed|failed|observed|expected|Error|error:" gpurun_out/$name.log | head -40; }
run probe 120 tests/test_gpu_kernels.py -s -k "tmem_conventions"
run lstm_tc 300 tests/test_gpu_kernels.py -s -k "lstm_384 and tcgen05"
run lstm_mma 300 tests/test_gpu_kernels.py -k "lstm_384 and mma"
run pipeline 900 tests/test_gpu_pipeline.py -s
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -4 gpurun_out/bench.err
