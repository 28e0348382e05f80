#!/bin/bash
# Round-end evidence: bench line, ncu launch list of the same workload, full ncu captures of the dominant kernels.
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -2 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python scripts/profile_step.py 1 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_rec_tc -s 3 -c 1 -o gpurun_out/prof_lstm_rec_tc_final -f python scripts/profile_step.py 1 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:crf_decode -c 1 -o gpurun_out/prof_crf_decode_final -f python scripts/profile_step.py 1 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 40 -c 1 -o gpurun_out/prof_gemm_ws_final -f python scripts/profile_step.py 1 > gpurun_out/ncu_c.log 2>&1
for f in a b c; do tail -n 1 gpurun_out/ncu_$f.log; done
