#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash scripts/gpu_kernel_tests.sh tests/test_decode_tc_gpu.py tests/test_attention_gpu.py tests/test_gemm_gpu.py
for impl in tc; do
  VITA_B200_GEMV=$impl timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err
  echo "== bench $impl exit $?" | tee -a gpurun_out/summary.txt
  python -c "
import json; d=json.load(open('gpurun_out/bench_$impl.json')); print('$impl', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline']['frac'])"
  tail -3 gpurun_out/bench_$impl.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?" | tee -a gpurun_out/summary.txt
