#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run t_kernels 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py -q -m gpu -p no:cacheprovider --timeout 300 -x
run t_pipe 900 python -m pytest tests/test_gpu_pipeline.py tests/test_cli.py -q -m gpu -p no:cacheprovider --timeout 600 -x
TAILN=2 run bench 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=3 run sup 400 python scripts/bench_sup.py
