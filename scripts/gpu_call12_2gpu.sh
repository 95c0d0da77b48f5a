#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for mode in p2p nccl; do
  VITA_B200_EP=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$((RANDOM % 10)) \
    tests/ep_check.py --layers 2 --seq 300 --time-seq 4096 --time-layers 8 > gpurun_out/ep_check_$mode.log 2>&1
  echo "== ep_check $mode exit $?" | tee -a gpurun_out/summary.txt; grep -E "^EP|Error|error|Traceback|timeout" gpurun_out/ep_check_$mode.log | head -8
done
tail -20 gpurun_out/ep_check_p2p.log | cut -c1-300
