#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/env8.txt
VITA_B200_EP=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  tests/ep_check.py --layers 2 --seq 300 --time-seq 4096 --time-layers 32 > gpurun_out/ep8_p2p.log 2>&1
echo "== ep8 p2p exit $?" | tee -a gpurun_out/summary.txt; grep -E "^EP|Error|error|Traceback|timeout" gpurun_out/ep8_p2p.log | head -8
VITA_B200_EP=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
  tests/ep_check.py --layers 2 --seq 300 --time-seq 4096 --time-layers 32 > gpurun_out/ep8_nccl.log 2>&1
echo "== ep8 nccl exit $?" | tee -a gpurun_out/summary.txt; grep -E "^EP|Error|error|Traceback|timeout" gpurun_out/ep8_nccl.log | head -8
