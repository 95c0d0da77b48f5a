#!/bin/bash
mkdir -p gpurun_out
VITA_B200_LIB=$PWD/vita_b200/lib/libvita_b200_trace.so timeout 600 python scripts/decode_trace.py --layers 6 > gpurun_out/decode_trace_final.log 2>&1; echo "== trace exit $?"
grep -v Warning gpurun_out/decode_trace_final.log | tail -12 | cut -c1-170
timeout 600 python scripts/decode_ab.py --rounds 3 --new-tokens 128 > gpurun_out/decode_ab8.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab8.log | tail -7 | cut -c1-150
