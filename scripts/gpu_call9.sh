#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 600 -k "single_kernel" -s 2>&1 | tail -30 > gpurun_out/mega_test.log; echo "== mega tests exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt; grep -E "mega vs|passed|failed|Error|timeout|trap" gpurun_out/mega_test.log | head
for mode in mega kernels; do
  VITA_B200_DECODE=$mode timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  echo "== bench $mode exit $?" | tee -a gpurun_out/summary.txt
  python -c "
import json; d=json.load(open('gpurun_out/bench_$mode.json')); print('$mode', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline']['frac'], d['gpu_launches'])"
  tail -3 gpurun_out/bench_$mode.err
done
