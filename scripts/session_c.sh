#!/usr/bin/env bash
# Round-2 session C (1 GPU): DERIVE chain kernel - full GPU suite, headline bench with and without it, timeline.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_c
mkdir -p "$OUT"
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -E "^smoke|Error|error" | tee "$OUT/smoke.log"
echo "== bench: default (derive) vs SSB_CHAIN_NO_DERIVE=1"
for env in "" "SSB_CHAIN_NO_DERIVE=1" "SSB_WGRAD_GROUP=0"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_derive.jsonl"
done
echo "== driver-style short run"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 2>/dev/null | tail -1 | tee -a "$OUT/bench_short.jsonl"
echo "== chain timeline (fp32)"
timeout 120 python scripts/chain_timeline.py 2>&1 | tail -40 | tee "$OUT/chain_timeline.log"
