#!/bin/bash
# Round-2 evidence, part C (2 GPUs): the driver's launch line for N=2, both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.out 2> gpurun_out/r02_bench_n2.err; echo "N=2 rc=$?"
grep "^{" gpurun_out/r02_bench_n2.out | tail -n 1 > gpurun_out/r02_bench_n2.json
grep "resident\|e2e" gpurun_out/r02_bench_n2.err | tail -n 8
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2 value %.3e ms/step %.2f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "sup", d["configs"]["config3_sup"]["value"], [b["value"] for b in d["configs"]["config5_sup_sweep"]])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n2.out 2> gpurun_out/r02_bench_reference_n2.err; echo "ref N=2 rc=$?"
grep "^{" gpurun_out/r02_bench_reference_n2.out | tail -n 1 > gpurun_out/r02_bench_reference_n2.json; cut -c1-200 gpurun_out/r02_bench_reference_n2.json
