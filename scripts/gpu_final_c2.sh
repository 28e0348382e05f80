#!/bin/bash
# 2 GPUs, after the stream fix: the driver's launch line for N=2 (product arm only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.out 2> gpurun_out/r02_bench_n2.err; echo "N=2 rc=$?"
grep "^{" gpurun_out/r02_bench_n2.out | tail -n 1 > gpurun_out/r02_bench_n2.json
grep "hac resident\|hac e2e" gpurun_out/r02_bench_n2.err | tail -n 4
