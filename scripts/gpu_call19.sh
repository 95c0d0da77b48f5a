#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_tc_gpu.py tests/test_attention_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -12 > gpurun_out/pytest_19.log; echo "== pytest exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_19.log
timeout 900 python scripts/decode_ab.py --rounds 3 --new-tokens 128 > gpurun_out/decode_ab3.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab3.log | tail -10 | cut -c1-150
