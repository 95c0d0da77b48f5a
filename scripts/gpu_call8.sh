#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; echo "== pytest gpu exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('bench', d['value'], d['phases_ms'], d['decode']['hbm_frac'], d['roofline']['frac'], d['prefill'], d.get('cpu_baseline'))"
tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== bench ref exit $?" | tee -a gpurun_out/summary.txt
tail -c 1200 gpurun_out/bench_ref.json
