#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "chunk_on_the_device" 2>&1 | tail -n 3
timeout 300 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -k "basecall or cli" 2>&1 | tail -n 2
