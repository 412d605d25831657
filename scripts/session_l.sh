#!/usr/bin/env bash
# Round-2 session L (1 GPU): final evidence on the final code - GPU suite, smoke, profiles (ncu, launch lists), sanitizer.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_l
mkdir -p "$OUT"
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 2>&1 | tail -8 | tee "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -E "^smoke|rror" | tee "$OUT/smoke.log"
echo "== profiles"
bash scripts/capture_profiles.sh 2>&1 | tail -12 | tee "$OUT/capture.log"
echo "== sanitizer"
bash scripts/sanitize.sh 2>&1 | tee "$OUT/sanitize.log"
echo "== reference arm on this box (N=1), for the record"
timeout 300 python bench.py --impl reference --steps 50 --warmup 5 2>/dev/null | tail -1 | tee "$OUT/bench_reference_n1.jsonl" | cut -c1-400
