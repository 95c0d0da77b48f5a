#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -12 > gpurun_out/pytest_22.log; echo "== pytest exit ${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_22.log
for t in 1; do
VITA_B200_LIB=$PWD/vita_b200/lib/libvita_b200_trace.so timeout 600 python scripts/decode_trace.py --layers 6 --opt attn_tagged=$t > gpurun_out/decode_trace_t$t.log 2>&1; echo "== trace tagged=$t exit $?"
grep -v Warning gpurun_out/decode_trace_t$t.log | tail -12 | cut -c1-170
done
timeout 600 python scripts/decode_ab.py --rounds 3 --new-tokens 128 > gpurun_out/decode_ab6.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab6.log | tail -6 | cut -c1-150
