#!/usr/bin/env bash
# Round-2 session D (1 GPU): DERIVE with batched splitter + gated wgrad wave.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_d
mkdir -p "$OUT"
echo "== pytest -m gpu (no -x: one stale assertion must not hide the rest)"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 2>&1 | tail -40 | tee "$OUT/pytest_gpu.log"
echo "== bench variants"
for env in "" "SSB_CHAIN_NO_DERIVE=1" "SSB_WGRAD_GATE=1" "SSB_WGRAD_GATE=1 SSB_CHAIN_NO_DERIVE=1"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench.jsonl"
done
echo "== chain timeline (fp32, derive)"
timeout 120 python scripts/chain_timeline.py 2>&1 | tail -20 | tee "$OUT/chain_timeline.log"
