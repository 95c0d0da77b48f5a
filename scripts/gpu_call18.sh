#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/decode_ab.py --rounds 3 --new-tokens 128 > gpurun_out/decode_ab2.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab2.log | tail -13 | cut -c1-150
VITA_B200_LIB=$PWD/vita_b200/lib/libvita_b200_mb2.so timeout 900 python scripts/decode_ab.py --rounds 3 --new-tokens 128 --only default,pdl_only,nopdl > gpurun_out/decode_ab2_mb2.log 2>&1; echo "== ab mb2 exit $?"
grep -v Warning gpurun_out/decode_ab2_mb2.log | tail -5 | cut -c1-150
