#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=8 run t_sup3 600 python -m pytest tests/test_gpu_transformer.py -q -p no:cacheprovider -x
TAILN=12 run attn_vs_fa3 300 python scripts/attention_vs_flashattn.py
TAILN=16 run attn_tl4 200 python scripts/attention_timeline.py
