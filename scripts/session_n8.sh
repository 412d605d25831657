#!/usr/bin/env bash
# Round-2 session on 8 GPUs: dp=8 LL kernel, the 1->8 curve, BASELINE configs 2-5, timeline.  Kept short: 8x cost.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_n8
mkdir -p "$OUT"
T='tests/test_gpu_multi.py::'
echo "== tests: tight single step dp8 (LL kernel vs CPU oracle), dp2 x pp4"
timeout 600 python -m pytest -q --maxfail=2 "${T}test_single_step_matches_the_cpu_oracle_tightly[8-1-naive-fused]" "${T}test_dp2_pp4_gpipe" 2>&1 | tail -8 | tee "$OUT/pytest.log"
echo "== scaling curve, driver style (--steps 20 --warmup 3)"
for n in 1 2 4 8; do
    timeout 300 python bench.py --gpus $n --steps 20 --warmup 3 --no-alt 2>/dev/null | tail -1 | tee -a "$OUT/scale_short.jsonl"
done
echo "== dp8 long: LL (default) vs flag protocol"
timeout 300 python bench.py --gpus 8 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp8.jsonl"
SSB_DP_LL=0 timeout 300 python bench.py --gpus 8 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp8.jsonl"
echo "== LL timeline dp8"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29743 scripts/dp_ll_timeline.py 2>/dev/null | grep -v "^\*\*\*" | tee "$OUT/dp_ll_timeline_dp8.log"
echo "== BASELINE config 3: pp=4 GPipe 8 micro-batches (fp32, folded peer transport), on 4 of the 8 GPUs"
timeout 300 python bench.py --gpus 4 --pp 4 --schedule gpipe --n-mubatches 8 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp4.jsonl"
echo "== BASELINE config 4: dp=2 x pp=4 GPipe 8 micro-batches (fp32)"
timeout 300 python bench.py --gpus 8 --pp 4 --schedule gpipe --n-mubatches 8 --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp2pp4.jsonl"
echo "== BASELINE config 5: dp=4 x pp=2 PipeDream-flush, hidden 8192 x 15 Linears (fp32)"
timeout 600 python bench.py --gpus 8 --pp 2 --schedule pipedream --hidden 8192 --n-layers 15 --seed-mode index --steps 10 --warmup 3 --repeats 2 2>/dev/null | tail -1 | tee -a "$OUT/bench_stress.jsonl"
