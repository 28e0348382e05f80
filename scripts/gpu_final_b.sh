#!/bin/bash
# Round-2 evidence, part B: ncu launch lists of one full step (hac, sup) and --set full captures of the dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_hac.csv python scripts/profile_step.py hac 1 > gpurun_out/ncu_l1.log 2>&1; tail -n 1 gpurun_out/ncu_l1.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_sup.csv python scripts/profile_step.py sup 1 > gpurun_out/ncu_l2.log 2>&1; tail -n 1 gpurun_out/ncu_l2.log
cap() { local name=$1 regex=$2 which=$3 skip=$4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/prof_$name -f python scripts/profile_step.py $which 1 > gpurun_out/ncu_$name.log 2>&1; tail -n 1 gpurun_out/ncu_$name.log; }
cap lstm_rec_tc6 lstm_rec_tc6 hac 2
cap gemm_pair gemm_pair hac 2
cap crf_decode crf_decode hac 0
cap conv_stem conv_stem hac 0
cap attention_tc2 attention_tc2 sup 3
cap gemm_pair_streaming gemm_pair sup 5
ls -la gpurun_out/*.ncu-rep
