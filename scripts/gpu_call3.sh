#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash scripts/gpu_kernel_tests.sh tests/test_decode_gpu.py tests/test_attention_gpu.py tests/test_model_gpu.py
grep -E "layer [01]:|prefill logits:|decode: usable" gpurun_out/test_model_gpu.log
timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:GateUpOp -s 8 -c 2 \
  -o gpurun_out/prof_gateup -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_gateup.log 2>&1
echo "== ncu gateup exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_bf16_tn_kernel<256, true|gemm_bf16_tn_kernel<256, 1" -s 4 -c 2 \
  -o gpurun_out/prof_moegemm -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_moegemm.log 2>&1
echo "== ncu moe gemm exit $?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out | grep ncu-rep
