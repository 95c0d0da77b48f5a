mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ep_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 1 --warmup 3 --no-parity --no-long-prefill > gpurun_out/bench_2gpu_default.json 2> gpurun_out/bench_2gpu_default.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_2gpu_default.json").read().strip().splitlines()[-1])
print("value", d["value"], "n", d["n_gpus"]); e = d.get("ep") or {}
print({k: e.get(k) for k in ("mode", "ms", "phases_ms", "prefill_tflops_aggregate", "speedup_prefill", "bit_identical_last_row_logits_vs_single_gpu", "error")})
PY
