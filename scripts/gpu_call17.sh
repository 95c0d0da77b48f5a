#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/decode_ab.py --rounds 3 --new-tokens 128 > gpurun_out/decode_ab.log 2>&1; echo "== ab exit $?"
grep -v Warning gpurun_out/decode_ab.log | tail -12
