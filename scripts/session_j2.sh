#!/usr/bin/env bash
# Round-2 session J2 (2 GPUs): re-validation of the pipeline + DP paths after the chain kernel was templated on FOLD.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_j2
mkdir -p "$OUT"
echo "== tests: pp2, peer transport, alternation, tight single step (dp2 LL / nvls / pp2), LL bitwise dp2"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "test_pp2 or pp_peer or alternation or single_step or dp_ll" --maxfail=5 2>&1 | tail -8 | tee "$OUT/pytest.log"
echo "== bench: pp2 gpipe folded, dp2"
timeout 300 python bench.py --gpus 2 --pp 2 --schedule gpipe --n-mubatches 8 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench.jsonl" | cut -c1-200
timeout 300 python bench.py --gpus 2 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench.jsonl" | cut -c1-200
