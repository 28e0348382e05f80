#!/bin/bash
# compute-sanitizer memcheck over tests that launch the product-path kernels (bounded: the tool serialises launches).
# The first-generation recurrent kernel (cross-check only) is left out: memcheck reports its DSMEM bulk copy to the CTA's own
# rank ("not located in remote CTA").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "many_row_blocks or column_blocks or int8 or crf_decode_matches or beam_search_matches or conv_stem" > gpurun_out/sanitizer_kernels.log 2>&1
echo "sanitizer kernels rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_kernels.log | tail -3
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 3 python -m pytest tests/test_gpu_pipeline.py -q -p no:cacheprovider -k "forward_scores_match_oracle or decode_of_own_scores or quantized" > gpurun_out/sanitizer_pipeline.log 2>&1
echo "sanitizer pipeline rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_pipeline.log | tail -3
grep -m 12 "Invalid\|at \|by thread" gpurun_out/sanitizer_pipeline.log gpurun_out/sanitizer_kernels.log | head -30
