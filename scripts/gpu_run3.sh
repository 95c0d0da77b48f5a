#!/bin/bash
# third GPU call: the two fixed tests, PR1 on the minted fixture + GPU seed search (information), decode A/B of the
# completion counters, the bench (N = 1), GEMM shapes, ncu evidence for the tensor-bound kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_depth_gpu.py tests/test_decode_gpu.py tests/test_pr1_gpu.py tests/test_rows_gpu.py -q -s > gpurun_out/pytest_fix.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_fix.log; grep -E "passed|failed|first_row|oracle margins|top-2 logit|rc=" gpurun_out/pytest_fix.log | cut -c1-600
timeout 600 python -m oracle.make_golden_pr1 search --first 40 --max 140 --want 4 > gpurun_out/pr1_search.log 2>&1; tail -8 gpurun_out/pr1_search.log | cut -c1-300
python scripts/gemm_bench.py > gpurun_out/gemm_shapes.txt 2>&1; tail -3 gpurun_out/gemm_shapes.txt | cut -c1-600
VITA_B200_CHAIN_COUNTERS=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-long-prefill > gpurun_out/bench_nochain.json 2> gpurun_out/bench_nochain.err
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -5 gpurun_out/bench_r02.err
python - <<'PY'
import json
for f in ("bench_nochain", "bench_r02"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d["value"], "decode ms/tok", d["phases_ms"]["decode_per_token"], "hbm_frac", d["decode"]["hbm_frac"],
              "enc", d["phases_ms"]["encoders_and_splice"], "prefill", d["phases_ms"]["mixtral_prefill"], "e2e", d["e2e"]["value"],
              "long", (d.get("prefill_long") or {}).get("tensor_frac"), "roof", d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
        print("   parity", d.get("parity"))
    except Exception as e:
        print(f, "unreadable", e)
PY
# ncu: launch list of a short bench + full-set captures of the tensor-bound kernels (one GPU, a handful of launches each)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --layers 4 --new-tokens 4 --no-cpu-baseline --no-parity --no-long-prefill > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 4 -c 4 -f -o gpurun_out/prof_gemm_s4096 python scripts/prof_target.py gemm > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_flash python scripts/prof_target.py flash > gpurun_out/ncu_flash.log 2>&1
tail -2 gpurun_out/ncu_gemm.log gpurun_out/ncu_flash.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
