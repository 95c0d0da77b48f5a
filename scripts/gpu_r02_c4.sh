#!/bin/bash
# call 4: early routing hand-over (tests + A/B), prefill timelines, GEMM shapes with the cost-model heuristic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_tc_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py tests/test_demo_replay_gpu.py tests/test_gemm_gpu.py -q -x -p no:cacheprovider > gpurun_out/pytest_c4.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_c4.log | cut -c1-300
timeout 600 python scripts/decode_ab.py --only default,early_route_off,early_route_lead4 --rounds 3 > gpurun_out/decode_ab2.txt 2>&1
echo "decode_ab rc=$?"; grep -E "min |Error" gpurun_out/decode_ab2.txt | cut -c1-300
timeout 300 python scripts/prefill_timeline.py 506 8 > gpurun_out/prefill_timeline_506.txt 2>&1; cat gpurun_out/prefill_timeline_506.txt | cut -c1-200
timeout 300 python scripts/prefill_timeline.py 4096 4 > gpurun_out/prefill_timeline_4096.txt 2>&1; cat gpurun_out/prefill_timeline_4096.txt | cut -c1-200
python scripts/gemm_bench.py auto > gpurun_out/gemm_shapes_auto.txt 2>&1; cut -c1-1200 gpurun_out/gemm_shapes_auto.txt
