#!/usr/bin/env bash
# Round-2 session J (2 GPUs): pipeline boundaries folded into the chain kernel - correctness, then A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_j
mkdir -p "$OUT"
echo "== tests: pp2 all schedules, peer transport, train/eval alternation, tight single step"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "pp2 or pp_peer or alternation or single_step" --maxfail=5 2>&1 | tail -20 | tee "$OUT/pytest_pp.log"
echo "== bench pp2 gpipe 8 micro-batches: folded / push+wait kernels / nccl"
for env in "" "SSB_PP_FOLD=0"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 2 --pp 2 --schedule gpipe --n-mubatches 8 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp2.jsonl"
done
timeout 300 python bench.py --gpus 2 --pp 2 --schedule pipedream --n-mubatches 8 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp2.jsonl"
timeout 300 python bench.py --gpus 2 --pp 2 --schedule naive --n-mubatches 4 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp2.jsonl"
echo "== train.py end to end on 2 GPUs (pp=2, eval + train alternate, native engine)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 train.py --pp 2 --schedule gpipe --synthetic --steps 600 2>&1 | grep -E "Epoch|Error|error" | tail -5 | tee "$OUT/train_pp2.log"
