#!/bin/bash
# call 8: PR1 fixture diagnostics (logit values), GEMM shapes with MT off, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pr1_gpu.py::test_pr1_free_running_greedy_token_ids_exact -q -s -p no:cacheprovider > gpurun_out/pytest_c8.log 2>&1
echo "pr1 rc=$?"; grep -E "passed|failed|top-2 logit|cuda top-2|oracle margins" gpurun_out/pytest_c8.log | cut -c1-2500
python scripts/gemm_bench.py auto > gpurun_out/gemm_shapes_auto.txt 2>&1; cut -c1-1200 gpurun_out/gemm_shapes_auto.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_c8.json 2> gpurun_out/bench_c8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c8.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), d["phases_ms"], "hbm_frac", round(d["decode"]["hbm_frac"], 4), "long", d["prefill_long"]["ms"], round(d["prefill_long"]["tensor_frac"], 4), "prefill tf", round(d["prefill"]["tensor_frac"], 4))
PY
