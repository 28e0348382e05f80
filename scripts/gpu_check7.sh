#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2
  timeout -s KILL $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run t_kernels 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "stem" -p no:cacheprovider --timeout 300
run t_pipe 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider --timeout 600 -x -k "forward_scores or basecall_pipeline or full_size"
for e in "X=1" "B200_STEM_IMPL=fma"; do
  echo "bench $e: $(env $e timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k: round(v,3) for k,v in d['stage_launch_ms_summed_per_step'].items() if k in ('conv_stem','crf_decode')})")"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_stem -c 1 -o gpurun_out/prof_conv_stem -f python scripts/profile_step.py 1 > gpurun_out/ncu_s.log 2>&1; tail -n 1 gpurun_out/ncu_s.log
