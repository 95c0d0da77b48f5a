#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash scripts/gpu_kernel_tests.sh tests/test_decode_tc_gpu.py tests/test_model_gpu.py
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "== bench tc exit $?" | tee -a gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('tc', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline']['frac'])"
tail -3 gpurun_out/bench_tc.err
VITA_B200_PDL=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_nopdl.json 2> gpurun_out/bench_tc_nopdl.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc_nopdl.json')); print('tc-nopdl', d['value'], d['phases_ms'], d['decode']['hbm_frac'])"
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1; cat gpurun_out/gemm_bench.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?" | tee -a gpurun_out/summary.txt
