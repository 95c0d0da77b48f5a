#!/bin/bash
# Diagnostic call: routing-aligned full-depth check, PR1 router-scale sweep, decode-chain A/B (completion counters)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_full_depth.log 2>&1
echo "full_depth rc=$?"; grep -E "^\{|passed|failed|Error|assert" gpurun_out/pytest_full_depth.log | cut -c1-3000
timeout 900 python -m oracle.make_golden_pr1 search --scales 1,8,16,128 --first 40 --max 64 --want 1000 > gpurun_out/pr1_search.log 2>&1
echo "pr1 search rc=$?"; grep -E "scale [0-9]+:|Error" gpurun_out/pr1_search.log | cut -c1-400
timeout 600 python scripts/decode_ab.py --only default,counters_off,counters_off_nopdl --rounds 3 > gpurun_out/decode_ab.txt 2>&1
echo "decode_ab rc=$?"; grep -E "min " gpurun_out/decode_ab.txt | cut -c1-300
