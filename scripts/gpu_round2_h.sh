#!/bin/bash
# GPU call: beam search tests; ncu captures (attention_tc, lstm_rec_tc6, gemm_ws, crf_decode) + launch lists
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_beam 600 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "beam" -s
grep -h "beam kernel ==\|beam-32 vs exact" gpurun_out/t_beam.log
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/prof_attention_tc -f python scripts/profile_step.py sup 1 64 > gpurun_out/ncu_attn.log 2>&1; tail -n 2 gpurun_out/ncu_attn.log
ncu --set full --clock-control none --import-source on -k regex:lstm_rec_tc6 -s 2 -c 1 -o gpurun_out/prof_lstm_rec_tc6 -f python scripts/profile_step.py hac 1 > gpurun_out/ncu_rec.log 2>&1; tail -n 2 gpurun_out/ncu_rec.log
ncu --set full --clock-control none --import-source on -k regex:gemm_ws_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_ws -f python scripts/profile_step.py hac 1 > gpurun_out/ncu_gemm.log 2>&1; tail -n 2 gpurun_out/ncu_gemm.log
ncu --set full --clock-control none --import-source on -k regex:crf_decode -c 1 -o gpurun_out/prof_crf_decode -f python scripts/profile_step.py hac 1 > gpurun_out/ncu_dec.log 2>&1; tail -n 2 gpurun_out/ncu_dec.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hac.csv python scripts/profile_step.py hac 1 > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sup.csv python scripts/profile_step.py sup 1 > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
