#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -6
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-long-prefill 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],2), round(d['phases_ms']['decode_per_token'],4), round(d['decode']['hbm_frac'],4), d['gpu_launches'])"
