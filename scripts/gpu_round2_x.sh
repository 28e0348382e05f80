#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py -q -p no:cacheprovider -x -k "crf_decode or beam or sup" 2>&1 | tail -n 3
timeout -s KILL 400 python bench.py --steps 6 --warmup 3 --workload sup --no-cpu-baseline 2>gpurun_out/sup_dec2.err > gpurun_out/sup_dec2.json
python - <<PY
import json
d = json.load(open("gpurun_out/sup_dec2.json"))
print("sup ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], d["stage_ms_per_step"])
PY
