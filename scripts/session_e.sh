#!/usr/bin/env bash
# Round-2 session E (2 GPUs): LL two-shot fused DP kernel - correctness vs the flag protocol, then exposed-comm A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_e
mkdir -p "$OUT"
echo "== LL tests first (fail fast)"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "dp_ll or dp2" --maxfail=3 2>&1 | tail -30 | tee "$OUT/pytest_ll.log"
echo "== bench dp2: LL gated / LL behind the chain / flag protocol"
for env in "" "SSB_DP_GATE=0" "SSB_DP_LL=0"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 2 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp2.jsonl"
done
echo "== single GPU reference point on this box"
timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 --no-alt 2>/dev/null | tail -1 | tee -a "$OUT/bench_n1.jsonl"
echo "== remaining multi-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --runxfail -k "not dp_ll and not dp2" --maxfail=5 2>&1 | tail -15 | tee "$OUT/pytest_multi_rest.log"
