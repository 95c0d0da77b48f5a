#!/bin/bash
# 2 GPUs: expert-parallel prefill parity + timing, and the torchrun path of bench.py (replicas)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/env2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  tests/ep_check.py --layers 2 --seq 300 --time-seq 4096 --time-layers 8 > gpurun_out/ep_check.log 2>&1
echo "== ep_check exit $?" | tee -a gpurun_out/summary.txt; grep -E "^EP|Error|error" gpurun_out/ep_check.log | head
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "== bench 2gpu exit $?" | tee -a gpurun_out/summary.txt
python -c "
import json
for l in open('gpurun_out/bench_2gpu.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('2gpu', d['value'], d['n_gpus'], d['phases_ms'], d['e2e'])"
tail -3 gpurun_out/bench_2gpu.err
