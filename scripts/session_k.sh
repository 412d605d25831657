#!/usr/bin/env bash
# Round-2 session K (1 GPU): same-box A/B of the single-GPU step: accumulators, RMW update epilogue, loss-head prefetch.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_k
mkdir -p "$OUT"
for env in "" "SSB_WGRAD_RMW=0" "SSB_ACC_SPLIT=0" "SSB_HEAD_PREFETCH=0" "SSB_WGRAD_RMW=0 SSB_ACC_SPLIT=0 SSB_HEAD_PREFETCH=0" "SSB_WGRAD_GROUP=0 SSB_WGRAD_RMW=0 SSB_ACC_SPLIT=0 SSB_HEAD_PREFETCH=0 SSB_LOSS_ZEROCOPY=0"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_ab.jsonl" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp32 %.4f  tf32 %.4f  e2e %.4f  nodes %s' % (d['ms_per_step'], d['tf32_mode']['ms_per_step'], d['e2e']['ms_per_step'], d['config']['graph_nodes']))"
done
echo "== chain timeline fp32 (default)"
timeout 120 python scripts/chain_timeline.py 2>&1 | grep -E "^gemm|producer" | cut -c1-200 | tee "$OUT/chain_timeline.log"
