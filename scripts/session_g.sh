#!/usr/bin/env bash
# Round-2 session G (2 GPUs): pipeline transport on micro-batch streams + the tight single-step oracle tests.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_g
mkdir -p "$OUT"
echo "== tests: pp2 (nccl + peer), alternation, tight single-step"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "pp2 or pp_peer or alternation or single_step" --maxfail=5 2>&1 | tail -30 | tee "$OUT/pytest_pp.log"
echo "== bench pp2 gpipe 8 micro-batches: nccl vs peer"
for tr in nccl peer; do
    timeout 300 python bench.py --gpus 2 --pp 2 --schedule gpipe --n-mubatches 8 --pp-transport "$tr" --steps 200 --warmup 30 2>/dev/null \
        | tail -1 | tee -a "$OUT/bench_pp2.jsonl"
done
echo "== bench pp2 1f1b 8 micro-batches, peer"
timeout 300 python bench.py --gpus 2 --pp 2 --schedule pipedream --n-mubatches 8 --pp-transport peer --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp2.jsonl"
