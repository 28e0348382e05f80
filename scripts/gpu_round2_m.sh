#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=6 run t_sup 900 python -m pytest tests/test_gpu_transformer.py -q -p no:cacheprovider -x
grep -h "sup 18\|sup width" gpurun_out/t_sup.log
TAILN=14 run attn_vs_fa 300 python scripts/attention_vs_flashattn.py
TAILN=20 run attn_tl 200 python scripts/attention_timeline.py
timeout -s KILL 400 python bench.py --steps 6 --warmup 3 --workload sup --no-cpu-baseline 2>gpurun_out/sup.err > gpurun_out/sup.json
python - <<PY
import json
d = json.load(open("gpurun_out/sup.json"))
print("sup ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], d["stage_ms_per_step"], d["roofline"])
PY
