#!/bin/bash
# Round-2 validation on one B200: whole GPU suite, smoke, the default bench, GEMM / FlashAttention shape timings, ncu launch
# list + --set full captures of the dominant decode kernel and of the tensor-bound kernels (-> gpurun_out/, then
# scripts/summarize_profiles.py r02 -> profiles/).
mkdir -p gpurun_out; rm -f gpurun_out/prof_*.ncu-rep
S=gpurun_out/summary.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest exit $?" | tee -a $S; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300 | tee -a $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $S
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
echo "== bench exit $?" | tee -a $S
python - <<'PY' | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["phases_ms"], "hbm_frac", round(d["decode"]["hbm_frac"], 4),
          "roof", round(d["roofline"]["frac"], 4), "long", d.get("prefill_long"), "prefill", d["prefill"], "cpu", d.get("cpu_baseline", {}).get("value"),
          d["clocks"], "launches", d["gpu_launches"])
    print("parity", d.get("parity"))
except Exception as e:
    print("bench unreadable", e)
PY
tail -3 gpurun_out/bench_r02.err | cut -c1-300
python scripts/gemm_bench.py > gpurun_out/gemm_shapes.txt 2>&1; cut -c1-900 gpurun_out/gemm_shapes.txt | tee -a $S
for a in "mix 4096" "mix 512" "vit 1025" "whale 248"; do timeout 200 python scripts/fa_check.py $a bench 2>&1 | grep fa_ | tee -a $S; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --layers 4 --new-tokens 4 --no-cpu-baseline --no-parity --no-long-prefill > gpurun_out/ncu_launches.log 2>&1
echo "== ncu launches exit $?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:TcGateUpOp -s 8 -c 1 \
  -o gpurun_out/prof_tc_gateup -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline --no-parity --no-long-prefill > gpurun_out/ncu_gateup.log 2>&1
echo "== ncu gateup exit $?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 4 -c 4 -f -o gpurun_out/prof_gemm_s4096 \
  python scripts/prof_target.py gemm > gpurun_out/ncu_gemm.log 2>&1
echo "== ncu gemm exit $?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_flash \
  python scripts/prof_target.py flash > gpurun_out/ncu_flash.log 2>&1
echo "== ncu flash exit $?" | tee -a $S
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv | tee -a $S
# configs[4] on one GPU: 16 concurrent audio queries, batched decode; and with staggered arrivals through the batcher
timeout 600 python scripts/bench_cfg5.py > gpurun_out/cfg5_bs16.json 2> gpurun_out/cfg5_bs16.err; tail -1 gpurun_out/cfg5_bs16.json | cut -c1-600 | tee -a $S
timeout 600 python scripts/bench_cfg5.py --arrivals 12 > gpurun_out/cfg5_arrivals.json 2> gpurun_out/cfg5_arrivals.err; tail -1 gpurun_out/cfg5_arrivals.json | cut -c1-900 | tee -a $S
