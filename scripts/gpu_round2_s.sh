#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
B200_GEMM_PAIR=1 timeout 120 python scripts/gemm_bench.py 2>&1 | tail -4
} > gpurun_out/gemm_pair3.log 2>&1
cat gpurun_out/gemm_pair3.log
for p in 1 0 1; do
B200_GEMM_PAIR=$p timeout 600 python bench.py --workload hac --no-cpu-baseline > gpurun_out/bench_pair$p.json 2> gpurun_out/bench_pair$p.err
grep "resident\|e2e\|quantize" gpurun_out/bench_pair$p.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_pair$p.json").read().strip().splitlines()[-1])
print("pair=$p hac ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print({k: round(v, 3) for k, v in d["config"].get("stage_launch_ms_summed_per_step", {}).items()} if isinstance(d["config"].get("stage_launch_ms_summed_per_step"), dict) else list(d["config"].keys()))
PY
done
