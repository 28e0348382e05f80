#!/bin/bash
# schedule experiments: GEMM CTA cap x stagger x LSTM stream priority
mkdir -p gpurun_out
for cfg in "0 0 0" "16 0 0" "24 0 0" "40 0 0" "0 0 1" "16 0 1" "24 0 1" "40 0 1" "16 1 1" "24 1 1" "24 1 0"; do
  set -- $cfg
  out=$(B200_TILE_GEMM_CTAS=$1 B200_TILE_STAGGER=$2 B200_LSTM_PRIO=$3 timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))")
  echo "cap=$1 stagger=$2 prio=$3 -> $out"
done
