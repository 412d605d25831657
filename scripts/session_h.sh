#!/usr/bin/env bash
# Round-2 session H (1 GPU): evidence - full GPU suite, ncu captures, launch lists, sanitizer passes, precision table,
# wide-GEMM microbenchmark.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_h
mkdir -p "$OUT"
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 2>&1 | tail -15 | tee "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -E "^smoke|rror" | tee "$OUT/smoke.log"
echo "== precision: 3xTF32 vs cuBLAS fp32"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -k 3xtf32 2>&1 | grep -E "3xTF32|passed|failed" | tee "$OUT/precision.log"
echo "== profiles"
bash scripts/capture_profiles.sh 2>&1 | tail -30 | tee "$OUT/capture.log"
echo "== kernel bench (stress shapes, split-K auto)"
timeout 300 python scripts/kernel_bench.py --shapes stress --iters 20 --k-splits -1 | tee gpurun_out/kernel_bench.jsonl | tail -20
echo "== sanitizer"
bash scripts/sanitize.sh 2>&1 | tee "$OUT/sanitize.log"
