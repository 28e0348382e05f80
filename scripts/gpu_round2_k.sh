#!/bin/bash
# GPU call: 2x32 tile shape of the recurrent kernel (tests, timeline, bench) + attention timeline
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=22 run attn_tl 200 python scripts/attention_timeline.py
export B200_LSTM_SHAPE=2x32
TAILN=5 run t_tile_2x32 400 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "tile" -x
TAILN=10 run tl_2x32 120 python scripts/lstm_tile_timeline.py
TAILN=6 run t_pipe_2x32 900 python -m pytest tests/test_gpu_pipeline.py -q -p no:cacheprovider -x
grep -h "headline shape\|sequences:" gpurun_out/t_pipe_2x32.log
for mode in "B200_BENCH_SLOTS=2" "B200_BENCH_SLOTS=3"; do
  echo "--- hac 2x32 $mode"
  tag=$(echo $mode | tr ' =' '__')
  env $mode timeout -s KILL 300 python bench.py --steps 12 --warmup 4 --workload hac --no-cpu-baseline 2>gpurun_out/bench232_$tag.err > gpurun_out/bench232_$tag.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench232_$tag.json"))
    print("$mode", "ms/step %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], "single %.2f" % d["e2e"]["single_call_ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          "launch_ms %.3f" % d["roofline"]["launch_ms"], d["stage_launch_ms_summed_per_step"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/bench232_$tag.err").read()[-1500:])
PY
done
