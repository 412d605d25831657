#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_final2
mkdir -p "$OUT"
timeout 300 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4 | tee "$OUT/pytest_gpu.log"
timeout 100 python bench.py --gpus 1 --steps 300 --warmup 30 2>/dev/null | tail -1 | tee "$OUT/bench_n1.jsonl" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp32 %.4f tf32 %.4f e2e %.4f' % (d['ms_per_step'], d['tf32_mode']['ms_per_step'], d['e2e']['ms_per_step']))"
