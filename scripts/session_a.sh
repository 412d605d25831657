#!/usr/bin/env bash
# Round-2 session A (1 GPU): the driver's own checks first (pytest -m gpu, smoke), then the validation of the
# opt-in code written blind in round 1.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_a
mkdir -p "$OUT"
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8 | tee "$OUT/smoke.log"
echo "== first_gpu_session"
bash scripts/first_gpu_session.sh 2>&1 | tee "$OUT/first_session.log"
