#!/bin/bash
# call 6: RoPE epilogue (tests), unit-router aligned PR1, robust PR1 seed search, parked-wait A/B, S=506/4096 timelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_attention_gpu.py tests/test_vllm_adapter_gpu.py tests/test_serving_engine_gpu.py "tests/test_pr1_gpu.py::test_pr1_any_seed_against_the_routing_aligned_oracle" -q -s -p no:cacheprovider > gpurun_out/pytest_c6.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^FAILED|max_row_rel_err" gpurun_out/pytest_c6.log | cut -c1-1500
timeout 600 python scripts/decode_ab.py --only default,park_off,early_route_on --rounds 3 > gpurun_out/decode_ab4.txt 2>&1
echo "decode_ab rc=$?"; grep -E "min |Error" gpurun_out/decode_ab4.txt | cut -c1-300
timeout 300 python scripts/prefill_timeline.py 4096 4 > gpurun_out/prefill_timeline_4096.txt 2>&1; head -8 gpurun_out/prefill_timeline_4096.txt | cut -c1-200
timeout 300 python scripts/prefill_timeline.py 506 8 > gpurun_out/prefill_timeline_506.txt 2>&1; head -8 gpurun_out/prefill_timeline_506.txt | cut -c1-200
timeout 900 python -m oracle.make_golden_pr1 search --scales 16,8 --first 0 --max 400 --want 3 > gpurun_out/pr1_search.log 2>&1
echo "pr1 search rc=$?"; grep -E "scale [0-9]+:|Error|rope table|robust\": true" gpurun_out/pr1_search.log | cut -c1-600
