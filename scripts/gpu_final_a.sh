#!/bin/bash
# Round-2 evidence, part A: full GPU test suite, smoke, the default bench line and the reference arm (what the driver runs),
# the recurrent-kernel timeline.  Outputs under gpurun_out/r02_*; copied to profiles/ by hand afterwards.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest -m gpu exit $?"; tail -n 4 gpurun_out/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -n 16 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference.json
timeout 200 python scripts/lstm_tile_timeline.py > gpurun_out/r02_lstm_timeline.txt 2>&1; tail -n 12 gpurun_out/r02_lstm_timeline.txt
timeout 200 python scripts/attention_timeline.py > gpurun_out/r02_attention_timeline.txt 2>&1
timeout 300 python scripts/attention_vs_flashattn.py > gpurun_out/r02_attention_vs_flashattn.json 2>&1
timeout 200 python scripts/gemm_bench.py > gpurun_out/r02_gemm_bench.txt 2>&1; tail -n 4 gpurun_out/r02_gemm_bench.txt
