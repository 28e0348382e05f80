#!/bin/bash
# GPU call: attention with 8 softmax warps, decode with cp.async prefetch rings -- tests, sup + hac bench
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=8 run t_all 1800 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -x
for mode in "B200_BENCH_SLOTS=1" "B200_BENCH_SLOTS=2"; do
  echo "--- sup $mode"
  tag=$(echo $mode | tr ' =' '__')
  env $mode timeout -s KILL 400 python bench.py --steps 6 --warmup 3 --workload sup --no-cpu-baseline 2>gpurun_out/sup_$tag.err > gpurun_out/sup_$tag.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sup_$tag.json"))
    print("$mode", "ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], d["stage_ms_per_step"], d["stage_tflops"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/sup_$tag.err").read()[-1500:])
PY
done
for mode in "B200_BENCH_SLOTS=1" "B200_BENCH_SLOTS=2"; do
  echo "--- hac $mode"
  tag=$(echo $mode | tr ' =' '__')
  env $mode timeout -s KILL 300 python bench.py --steps 12 --warmup 4 --workload hac --no-cpu-baseline 2>gpurun_out/bench_$tag.err > gpurun_out/bench_$tag.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print("$mode", "ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], "single %.2f" % d["e2e"]["single_call_ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          "launch_ms %.3f" % d["roofline"]["launch_ms"], d["stage_launch_ms_summed_per_step"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/bench_$tag.err").read()[-1500:])
PY
done
