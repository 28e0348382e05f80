#!/bin/bash
# 2-GPU call: the driver's launch line for N=2 (both arms), then single-GPU attention check and the default bench line
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "N=2 rc=$?"
tail -4 gpurun_out/bench_n2.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n2.json"))
    print("N=2 value %.3e ms/step %.2f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "sup", d["configs"]["config3_sup"]["value"], [b["value"] for b in d["configs"]["config5_sup_sweep"]])
except Exception as e:
    print("N=2 parse failed", e)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/ref_n2.json 2> gpurun_out/ref_n2.err; echo "ref N=2 rc=$?"; cut -c1-300 gpurun_out/ref_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_transformer.py -q -p no:cacheprovider -x 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 200 python scripts/attention_timeline.py 2>&1 | head -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"; tail -14 gpurun_out/bench_default.err
