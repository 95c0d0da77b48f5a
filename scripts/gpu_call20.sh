#!/bin/bash
mkdir -p gpurun_out
TAG=$PWD/vita_b200/lib/libvita_b200_tag.so
VITA_B200_LIB=$TAG timeout 900 python -m pytest tests/test_decode_tc_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -12 > gpurun_out/pytest_20.log; echo "== pytest(tag lib) exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_20.log
timeout 900 python scripts/decode_ab.py --rounds 3 --new-tokens 128 --only default,fast+wide,nopdl > gpurun_out/decode_ab4.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab4.log | tail -4 | cut -c1-150
VITA_B200_LIB=$TAG timeout 900 python scripts/decode_ab.py --rounds 3 --new-tokens 128 --only default,wide_route,nopdl > gpurun_out/decode_ab4_tag.log 2>&1; echo "== ab tag exit $?"
grep -v Warning gpurun_out/decode_ab4_tag.log | tail -4 | cut -c1-150
