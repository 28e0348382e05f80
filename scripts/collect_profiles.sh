#!/bin/bash
# copies the round-2 evidence that scripts/gpu_final_a.sh / gpu_final_b.sh left in gpurun_out/ into profiles/ (tracked)
cd "$(dirname "$0")/.."
for f in r02_pytest_gpu.txt r02_bench.json r02_bench.err r02_bench_reference.json r02_lstm_timeline.txt r02_attention_timeline.txt \
         r02_attention_vs_flashattn.json r02_gemm_bench.txt r02_launches_hac.csv r02_launches_sup.csv r02_bench_n2.json r02_bench_reference_n2.json; do
  [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/$f
done
python scripts/ncu_summary.py
ls profiles/
