#!/bin/bash
mkdir -p gpurun_out
VITA_B200_LIB=$PWD/vita_b200/lib/libvita_b200_trace.so timeout 600 python scripts/decode_trace.py --layers 6 > gpurun_out/decode_trace_final.log 2>&1; echo "== trace exit $?"
grep -v Warning gpurun_out/decode_trace_final.log | tail -12 | cut -c1-170
timeout 600 python scripts/decode_ab.py --rounds 3 --new-tokens 128 --only default,nopdl > gpurun_out/decode_ab7.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab7.log | tail -3 | cut -c1-150
VITA_B200_LIB=$PWD/vita_b200/lib/libvita_b200_nomax.so timeout 600 python scripts/decode_ab.py --rounds 3 --new-tokens 128 --only default,nopdl > gpurun_out/decode_ab7_nomax.log 2>&1; echo "== ab nomax exit $?"
grep -v Warning gpurun_out/decode_ab7_nomax.log | tail -3 | cut -c1-150
timeout 600 python -m pytest tests/test_decode_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
