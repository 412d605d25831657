#!/usr/bin/env bash
# Round-2 session I (2 GPUs): validation of the update kernels that refresh W_lo themselves, the batched fences of the
# wide-layer DP kernel, the rotating accumulators (precision table) - full single-GPU suite + 2-GPU suite + A/B benches.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_i
mkdir -p "$OUT"
echo "== pytest -m gpu (whole suite; multi-GPU cases that need > 2 GPUs skip)"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 2>&1 | tail -15 | tee "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | grep -E "^smoke|rror" | tee "$OUT/smoke.log"
echo "== precision: 3xTF32 vs cuBLAS fp32 (rotating accumulators)"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -k 3xtf32 2>&1 | grep -E "3xTF32|passed|failed" | tee "$OUT/precision.log"
echo "== bench N=1, N=2"
timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench.jsonl"
timeout 300 python bench.py --gpus 2 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench.jsonl"
echo "== wide model dp2 (two-shot flag protocol with batched fences) vs nccl"
for comm in fused nccl; do
    timeout 300 python bench.py --gpus 2 --comm "$comm" --hidden 4096 --n-layers 4 --seed-mode index --steps 20 --warmup 5 2>/dev/null | tail -1 | tee -a "$OUT/bench_wide.jsonl"
done
echo "== pp2 gpipe peer (deferred wgrad wave)"
timeout 300 python bench.py --gpus 2 --pp 2 --schedule gpipe --n-mubatches 8 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp2.jsonl"
echo "== kernel bench (stress shapes; cold vs stream)"
timeout 300 python scripts/kernel_bench.py --shapes stress --iters 20 --k-splits -1 | tee gpurun_out/kernel_bench.jsonl | cut -c1-330
