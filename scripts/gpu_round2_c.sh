#!/bin/bash
# GPU call: tile recurrent kernel with the L2-staged multicast exchange -- unit tests, timeline, suite, bench
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=15 run t_tile_mc 300 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "tile or column_blocks" -x
TAILN=12 run tl_tile_mc 120 python scripts/lstm_tile_timeline.py
TAILN=25 run t_all 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900
grep -h "headline shape\|sequences:\|sup 18\|sup width\|reverse-compl" gpurun_out/t_all.log
for mode in "B200_TILE_STREAMS=1" "B200_TILE_STREAMS=0"; do
  echo "--- $mode"
  env $mode timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload hac --no-cpu-baseline 2>gpurun_out/bench_$mode.err > gpurun_out/bench_$mode.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$mode.json"))
    print("$mode", "ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          "launch_ms %.3f" % d["roofline"]["launch_ms"], d["stage_launch_ms_summed_per_step"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/bench_$mode.err").read()[-1500:])
PY
done
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_full.err > gpurun_out/bench_full.json; echo "full bench rc=$?"; tail -12 gpurun_out/bench_full.err
