#!/usr/bin/env bash
# same-box A/B: batched RMW update epilogue vs TMA reduce-add + split kernel (two alternations to see the noise)
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_k2
mkdir -p "$OUT"
for env in "" "SSB_WGRAD_RMW=0" "" "SSB_WGRAD_RMW=0"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 --no-alt 2>/dev/null | tail -1 | tee -a "$OUT/bench_ab.jsonl" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp32 %.4f  e2e %.4f  nodes %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['graph_nodes']))"
done
timeout 200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_engine.py -q 2>&1 | tail -3
