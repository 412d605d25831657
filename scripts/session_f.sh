#!/usr/bin/env bash
# Round-2 session F (2 GPUs): LL kernel with batched phase B + split gates.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_f
mkdir -p "$OUT"
echo "== LL tests"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "dp_ll or test_dp2" --maxfail=3 2>&1 | tail -30 | tee "$OUT/pytest_ll.log"
echo "== bench dp2: LL gated / LL behind the chain / flag protocol"
for env in "" "SSB_DP_GATE=0" "SSB_DP_LL=0"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 2 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp2.jsonl"
done
echo "== driver-style short run"
timeout 300 python bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp2_short.jsonl"
