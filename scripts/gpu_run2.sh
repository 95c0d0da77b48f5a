#!/bin/bash
# second GPU call: FA two-tile mode, the whole GPU test-suite, PR1 prompt-seed search
mkdir -p gpurun_out
L=gpurun_out/fa_bringup2.log
: > $L
run() { echo "== $*" >> $L; timeout 150 env "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run VITA_B200_FA_NQ=2 python scripts/fa_check.py mix 256
run VITA_B200_FA_NQ=2 python scripts/fa_check.py vit 1025
run VITA_B200_FA_NQ=2 python scripts/fa_check.py whale 248
run VITA_B200_FA_NQ=0 python scripts/fa_check.py mix 4096 bench
run VITA_B200_FA_NQ=0 python scripts/fa_check.py mix 1300
run VITA_B200_FA_NQ=0 python scripts/fa_check.py vit 1025 bench
run VITA_B200_FA_NQ=1 python scripts/fa_check.py vit 1025 bench
run VITA_B200_FA_NQ=0 VITA_B200_FA_POLY=1 python scripts/fa_check.py mix 4096 bench
run VITA_B200_FA_NQ=0 VITA_B200_FA_POLY=2 python scripts/fa_check.py mix 4096 bench
run VITA_B200_FA_NQ=0 VITA_B200_FA_POLY=1 python scripts/fa_check.py vit 1025 bench
grep -v "mbarrier timeout" $L | tail -40
# the whole GPU suite with this round's unverified features switched ON (defaults stay conservative until this is green)
export VITA_B200_FA_NQ=0 VITA_B200_CHAIN_COUNTERS=1 VITA_B200_ENC_GRAPH=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --deselect tests/test_full_depth_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -x -s > gpurun_out/pytest_full_depth.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_full_depth.log
tail -30 gpurun_out/pytest_full_depth.log
timeout 900 python -m oracle.make_golden_pr1 search --max 300 --want 3 > gpurun_out/pr1_search.log 2>&1
tail -20 gpurun_out/pr1_search.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
