#!/bin/bash
# one gpurun call: correctness matrix (each case in its own process, bounded), then timing
mkdir -p gpurun_out
L=gpurun_out/fa_bringup.log
: > $L
run() { echo "== $*" >> $L; timeout 120 env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run VITA_B200_FA_NQ=1 python scripts/fa_check.py mix 128
run VITA_B200_FA_NQ=1 python scripts/fa_check.py mix 256
run VITA_B200_FA_NQ=2 python scripts/fa_check.py mix 256
run VITA_B200_FA_NQ=1 python scripts/fa_check.py vit 128
run VITA_B200_FA_NQ=1 python scripts/fa_check.py vit 1025
run VITA_B200_FA_NQ=2 python scripts/fa_check.py vit 1025
run VITA_B200_FA_NQ=1 python scripts/fa_check.py whale 248
run VITA_B200_FA_NQ=2 python scripts/fa_check.py whale 248
run VITA_B200_FA_NQ=0 python scripts/fa_check.py mix 506
run VITA_B200_FA_NQ=0 python scripts/fa_check.py mix 4096 bench
run VITA_B200_FA_NQ=1 python scripts/fa_check.py mix 4096 bench
run VITA_B200_FA_NQ=0 python scripts/fa_check.py vit 1025 bench
# descriptor variants, only informative if the cases above are wrong
run VITA_B200_FA_NQ=1 VITA_B200_FA_V_LBO=1024 VITA_B200_FA_V_SBO=16384 python scripts/fa_check.py mix 128
run VITA_B200_FA_NQ=1 VITA_B200_FA_V_LBO=16384 VITA_B200_FA_V_SBO=2048 python scripts/fa_check.py mix 128
timeout 900 python -m pytest tests/test_attention_gpu.py -x -q >> $L 2>&1
echo "pytest attention rc=$?" >> $L
tail -60 $L
timeout 900 python -m pytest tests/test_demo_replay_gpu.py tests/test_decode_gpu.py -x -q > gpurun_out/demo_replay.log 2>&1
echo "pytest demo rc=$?" >> gpurun_out/demo_replay.log
tail -30 gpurun_out/demo_replay.log
