#!/bin/bash
# call 3: GEMM tail split (tests + shapes A/B), model tests, routing-aligned full-depth check, PR1 seed search
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_rows_gpu.py -q -x -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1
echo "gemm/model tests rc=$?"; tail -4 gpurun_out/pytest_gemm.log | cut -c1-300
VITA_B200_GEMM_TAIL_SPLIT=0 python scripts/gemm_bench.py auto > gpurun_out/gemm_shapes_nosplit.txt 2>&1
python scripts/gemm_bench.py auto > gpurun_out/gemm_shapes_split.txt 2>&1
python - <<'PY'
import json
a = json.loads(open("gpurun_out/gemm_shapes_nosplit.txt").read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/gemm_shapes_split.txt").read().strip().splitlines()[-1])
for k in a:
    print(f"{k:20s} whole {a[k]['us']:8.1f} us {a[k]['tflops']:7.1f} TF | tail-split {b[k]['us']:8.1f} us {b[k]['tflops']:7.1f} TF")
PY
timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_full_depth.log 2>&1
echo "full_depth rc=$?"; grep -E "^\{|passed|failed" gpurun_out/pytest_full_depth.log | cut -c1-2500
timeout 900 python -m oracle.make_golden_pr1 search --scales 8,16 --first 0 --max 300 --want 100000 > gpurun_out/pr1_search.log 2>&1
echo "pr1 search rc=$?"; grep -E "scale [0-9]+:|Error" gpurun_out/pr1_search.log | cut -c1-400
