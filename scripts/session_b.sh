#!/usr/bin/env bash
# Round-2 session B (2 GPUs): the multi-GPU suite (never run by the driver: its box has one GPU), then the
# validation of the NVLS / peer-memory paths written blind in round 1 and the DP transport comparison.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_b
mkdir -p "$OUT"
echo "== multi-GPU tests (2 GPUs; larger worlds skip)"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --runxfail -x 2>&1 | tail -30 | tee "$OUT/pytest_multi.log"
echo "== first_multi_gpu_session"
N_GPUS=2 bash scripts/first_multi_gpu_session.sh 2>&1 | tee "$OUT/first_multi.log"
