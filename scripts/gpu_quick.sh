#!/bin/bash
# quick confidence run: every GPU test, isolated GEMM shapes, hac bench line, sup timing
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run t_all 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600
TAILN=4 run gemm 200 python scripts/gemm_bench.py
TAILN=1 run bench 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=1 run sup 400 python scripts/bench_sup.py
