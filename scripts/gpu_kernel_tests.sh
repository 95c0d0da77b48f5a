#!/bin/bash
# Runs every GPU test module in its own process (a trapped kernel must not poison the others) and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/env.txt 2>&1
nproc >> gpurun_out/env.txt; free -g | head -2 >> gpurun_out/env.txt
rc=0
for f in ${@:-tests/test_gemm_gpu.py tests/test_rows_gpu.py tests/test_attention_gpu.py tests/test_decode_gpu.py}; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -p no:cacheprovider --timeout 600 -x --no-header -rA 2>&1 | tail -150 > gpurun_out/$name.log
  echo "== $name exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$name.log
done
