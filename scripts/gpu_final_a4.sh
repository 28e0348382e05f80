#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -k "score_batches or basecall or cli" 2>&1 | tail -n 2
for i in 1 2 3; do
timeout 400 python bench.py --workload hac --no-cpu-baseline > gpurun_out/hac_rep$i.json 2> gpurun_out/hac_rep$i.err; grep "e2e" gpurun_out/hac_rep$i.err
done
timeout 400 python bench.py --workload sup --no-cpu-baseline --steps 6 > gpurun_out/sup_rep.json 2> gpurun_out/sup_rep.err; grep "L=9996" gpurun_out/sup_rep.err
