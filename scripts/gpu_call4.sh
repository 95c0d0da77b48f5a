#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash scripts/gpu_kernel_tests.sh tests/test_decode_gpu.py tests/test_attention_gpu.py tests/test_model_gpu.py
grep -E "layer [01]:|prefill logits:|decode: usable" gpurun_out/test_model_gpu.log
for pdl in 1 0; do
  VITA_B200_PDL=$pdl timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
  echo "== bench pdl=$pdl exit $?" | tee -a gpurun_out/summary.txt
  python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$pdl.json')); print('pdl=$pdl', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline']['frac'])"
  tail -3 gpurun_out/bench_pdl$pdl.err
done
