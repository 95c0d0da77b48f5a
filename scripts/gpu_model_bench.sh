#!/bin/bash
# GPU call #2: remaining kernel tests, end-to-end parity, smoke, bench (+ reference arm), ncu launch list and captures.
mkdir -p gpurun_out
bash scripts/gpu_kernel_tests.sh tests/test_rows_gpu.py tests/test_model_gpu.py
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== bench ref exit $?" | tee -a gpurun_out/summary.txt
tail -c 1500 gpurun_out/bench_ref.json
# every launch of one short step with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?" | tee -a gpurun_out/summary.txt
# full captures of the two dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_gemv_kernel.*GateUp -s 8 -c 3 \
  -o gpurun_out/prof_gateup -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_gateup.log 2>&1
echo "== ncu gateup exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel.*Lb1 -s 4 -c 3 \
  -o gpurun_out/prof_moegemm -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_moegemm.log 2>&1
echo "== ncu moe gemm exit $?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out
