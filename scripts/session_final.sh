#!/usr/bin/env bash
# final evidence on the final code (1 GPU): suite, smoke, ncu captures + launch lists
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_final
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4 | tee "$OUT/pytest_gpu.log"
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -E "^smoke|rror" | tee "$OUT/smoke.log"
bash scripts/capture_profiles.sh 2>&1 | tail -4 | tee "$OUT/capture.log"
