#!/bin/bash
# host staging one batch ahead of the device slots: full GPU tests, three short hac runs, the default bench line + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/r02_pytest_gpu.txt
for i in 1 2 3; do
timeout 400 python bench.py --workload hac --no-cpu-baseline > gpurun_out/hac_rep$i.json 2> gpurun_out/hac_rep$i.err; grep "resident\|e2e" gpurun_out/hac_rep$i.err
done
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; grep "resident\|e2e\|quantize\|cpu baseline\|config 1" gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
