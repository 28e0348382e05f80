#!/bin/bash
# last check of the final tree: build() from scratch is done on the CPU side; here the GPU suite, smoke and the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; grep "resident\|e2e\|quantize" gpurun_out/r02_bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ws_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_ws -f python scripts/profile_step.py hac 1 > gpurun_out/ncu_gemm_ws.log 2>&1; tail -n 1 gpurun_out/ncu_gemm_ws.log
