#!/bin/bash
# call 5: parked waits + early route A/B, new align kernel, PR1 (fixture + aligned), full-depth again with early route
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rows_gpu.py tests/test_decode_tc_gpu.py tests/test_pr1_gpu.py tests/test_model_gpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_c5.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^FAILED|oracle margins|top-2 logit|max_row_rel_err" gpurun_out/pytest_c5.log | cut -c1-900
timeout 600 python scripts/decode_ab.py --only default,early_route_off,early_route_lead4 --rounds 3 > gpurun_out/decode_ab3.txt 2>&1
echo "decode_ab rc=$?"; grep -E "min |Error" gpurun_out/decode_ab3.txt | cut -c1-300
timeout 300 python scripts/prefill_timeline.py 4096 4 > gpurun_out/prefill_timeline_4096.txt 2>&1; grep -E "prefill|moe_align|router" gpurun_out/prefill_timeline_4096.txt | cut -c1-200
