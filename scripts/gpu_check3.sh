#!/bin/bash
# round-1 late checks: fused SwiGLU GEMM, split-warp-set LSTM A/B, same-box torch GPU reference
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_tf 600 python -m pytest tests/test_gpu_transformer.py -q -m gpu -p no:cacheprovider --timeout 300
run tl_base 200 python scripts/lstm_timeline.py
B200_LSTM_SPLIT=1 run tl_split 200 python scripts/lstm_timeline.py
B200_LSTM_SPLIT=1 run t_lstm_split 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k lstm -p no:cacheprovider --timeout 300
TAILN=2 run bench_base 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
B200_LSTM_SPLIT=1 TAILN=2 run bench_split 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=3 run sup 400 python scripts/bench_sup.py
TAILN=3 run ref_hac 400 python scripts/gpu_reference.py --model hac
TAILN=3 run ref_sup 600 python scripts/gpu_reference.py --model sup
