#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; echo "== pytest gpu exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_cfg5.py > gpurun_out/cfg5.json 2> gpurun_out/cfg5.err; echo "== cfg5 exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/cfg5.json; tail -3 gpurun_out/cfg5.err
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('bench', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline'], d['prefill'], d['prefill_long'], d.get('cpu_baseline'))"
tail -3 gpurun_out/bench.err
