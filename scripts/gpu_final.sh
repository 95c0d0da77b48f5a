#!/bin/bash
# Round-end validation: full GPU test suite, smoke, default bench (with CPU baseline), ncu launch list + full captures.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/prof_*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_final.log; echo "== pytest exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('final', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['phases_ms'], round(d['decode']['hbm_frac'],4), d['roofline'], d.get('prefill_long'), d['cpu_baseline'], d['clocks'])"
tail -2 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline --no-long-prefill > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launches exit $?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:TcGateUpOp -s 8 -c 1 \
  -o gpurun_out/prof_tc_gateup -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline --no-long-prefill > gpurun_out/ncu_gateup.log 2>&1
echo "== ncu tc gateup exit $?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:decode_attn_kernel -s 8 -c 1 \
  -o gpurun_out/prof_decode_attn -f python bench.py --steps 1 --warmup 1 --layers 4 --new-tokens 4 --no-cpu-baseline --no-long-prefill > gpurun_out/ncu_attn.log 2>&1
echo "== ncu decode attn exit $?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out/*.ncu-rep
