#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
timeout 300 python scripts/image_bench.py 2>&1 | grep -v Warning | tail -2 | tee gpurun_out/image_bench.json
