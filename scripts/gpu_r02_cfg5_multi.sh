#!/bin/bash
# configs[4] on N GPUs: 16 audio queries with staggered arrivals, request parallel (queries sharded over the ranks)
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29588 scripts/bench_cfg5.py --arrivals 12 > gpurun_out/cfg5_arrivals_${N}gpu.json 2> gpurun_out/cfg5_arrivals_${N}gpu.err
echo "rc=$?"; tail -1 gpurun_out/cfg5_arrivals_${N}gpu.json | cut -c1-1200; tail -3 gpurun_out/cfg5_arrivals_${N}gpu.err | cut -c1-300
