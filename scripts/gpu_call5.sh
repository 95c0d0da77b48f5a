#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash scripts/gpu_kernel_tests.sh tests/test_decode_tc_gpu.py tests/test_model_gpu.py
grep -E "layer [01]:|prefill logits:|decode: usable" gpurun_out/test_model_gpu.log
for impl in tc simt; do
  VITA_B200_GEMV=$impl timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$impl.json 2> gpurun_out/bench_$impl.err
  echo "== bench $impl exit $?" | tee -a gpurun_out/summary.txt
  python -c "
import json; d=json.load(open('gpurun_out/bench_$impl.json')); print('$impl', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline']['frac'])"
  tail -3 gpurun_out/bench_$impl.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:TcGateUpOp -s 8 -c 2 \
  -o gpurun_out/prof_tc_gateup -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_gateup.log 2>&1
echo "== ncu tc gateup exit $?" | tee -a gpurun_out/summary.txt
