#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
VITA_B200_PDL=1 timeout 900 python -m pytest tests/test_decode_tc_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_pdl.log; echo "== pytest(PDL=1) exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/pytest_pdl.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "== bench $name exit $?" | tee -a gpurun_out/summary.txt
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', round(d['value'],2), d['phases_ms'], round(d['decode']['hbm_frac'],4), round(d['roofline']['frac'],4), d.get('prefill_long'))"
  tail -2 gpurun_out/bench_$name.err
}
EXTRA="--no-long-prefill" run pdl1 VITA_B200_PDL=1
