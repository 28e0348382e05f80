#!/bin/bash
# after switching the pair kernels off by default: GEMM tests, three short hac runs (stability of e2e), the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py -q -p no:cacheprovider -k "gemm or swiglu" 2>&1 | tail -n 2
for i in 1 2 3; do
timeout 400 python bench.py --workload hac --no-cpu-baseline > gpurun_out/hac_rep$i.json 2> gpurun_out/hac_rep$i.err; grep "resident\|e2e" gpurun_out/hac_rep$i.err
done
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; grep "resident\|e2e\|quantize\|cpu baseline\|config 1" gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$?"
