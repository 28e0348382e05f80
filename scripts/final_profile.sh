#!/bin/bash
# Round-end evidence: full GPU test suite, smoke, bench line (+ reference arm), ncu launch list of the same workload,
# full ncu captures of the dominant kernels, sup timing and same-box GPU references.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -n 2 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python scripts/bench_sup.py > gpurun_out/sup.log 2>&1; tail -n 1 gpurun_out/sup.log
timeout 300 python scripts/gpu_reference.py --model hac > gpurun_out/ref_hac.log 2>&1; tail -n 1 gpurun_out/ref_hac.log
timeout 300 python scripts/gpu_reference.py --model sup > gpurun_out/ref_sup.log 2>&1; tail -n 1 gpurun_out/ref_sup.log
timeout 120 python scripts/lstm_timeline.py > gpurun_out/tl_final.log 2>&1; tail -n 11 gpurun_out/tl_final.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python scripts/profile_step.py 1 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_rec_tc -s 3 -c 1 -o gpurun_out/prof_lstm_rec_tc_final -f python scripts/profile_step.py 1 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:crf_decode -c 1 -o gpurun_out/prof_crf_decode_final -f python scripts/profile_step.py 1 > gpurun_out/ncu_b.log 2>&1
for f in a b; do tail -n 1 gpurun_out/ncu_$f.log; done
