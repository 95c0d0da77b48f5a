#!/bin/bash
# multi-GPU call (gpurun --gpus N): expert-parallel parity + timing, then the bench with its ep record
N=${1:-2}
mkdir -p gpurun_out
export VITA_B200_FA_NQ=${VITA_B200_FA_NQ:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in seq p2p; do
  VITA_B200_EP=$mode timeout 600 $TR --master-port 29541 tests/ep_check.py --layers 2 --seq 300 > gpurun_out/ep${N}_${mode}_parity.log 2>&1
  echo "ep_check $mode rc=$?" >> gpurun_out/ep${N}_${mode}_parity.log; grep -E "EP x|mode|rc=|Error|error" gpurun_out/ep${N}_${mode}_parity.log | tail -5
done
L=$([ "$N" -ge 8 ] && echo 32 || echo 8)
for mode in seq p2p; do
  VITA_B200_EP=$mode timeout 900 $TR --master-port 29542 tests/ep_check.py --layers 2 --seq 1000 --time-seq 4096 --time-layers $L > gpurun_out/ep${N}_${mode}_time.log 2>&1
  echo "ep_time $mode rc=$?" >> gpurun_out/ep${N}_${mode}_time.log; grep -E "EP x|mode|rc=|Error|error" gpurun_out/ep${N}_${mode}_time.log | tail -5
done
VITA_B200_EP=seq timeout 1500 $TR --master-port 29543 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
