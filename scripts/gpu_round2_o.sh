#!/bin/bash
# round 2, call 17: 12 epilogue warps in the GEMMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in 0 128x2; do
  B200_GEMM_CLUSTER=$v timeout 120 python scripts/gemm_bench.py 2>&1 | tail -5
done
} > gpurun_out/gemm_epi12.log 2>&1
cat gpurun_out/gemm_epi12.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/t_epi12.log 2>&1
tail -n 3 gpurun_out/t_epi12.log
timeout 600 python bench.py --workload hac > gpurun_out/bench_epi12.json 2> gpurun_out/bench_epi12.err
tail -n 4 gpurun_out/bench_epi12.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_epi12.json").read().strip().splitlines()[-1])
print("hac ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print(d["config"].get("stage_launch_ms_summed_per_step"))
PY
