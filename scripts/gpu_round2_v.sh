#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=4 run t_pairs 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py -q -p no:cacheprovider -x -k "gemm or sup or attention"
TAILN=16 run attn_tl3 200 python scripts/attention_timeline.py
for p in 1 0; do
B200_GEMM_PAIR=$p timeout -s KILL 400 python bench.py --steps 6 --warmup 3 --workload sup --no-cpu-baseline 2>gpurun_out/sup_pair$p.err > gpurun_out/sup_pair$p.json
python - <<PY
import json
d = json.load(open("gpurun_out/sup_pair$p.json"))
print("pair=$p sup ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], d["stage_ms_per_step"])
PY
done
