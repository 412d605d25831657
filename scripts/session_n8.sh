#!/usr/bin/env bash
# Round-2 session on 8 GPUs: dp=8 LL kernel, the 1->8 curve, BASELINE configs 2, 4, 5, timelines.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_n8
mkdir -p "$OUT"
echo "== tests: dp8 LL bitwise vs flag protocol, tight single step dp8, dp2 x pp4"
timeout 900 python -m pytest -q --maxfail=3 'tests/test_gpu_multi.py::test_dp_ll_kernel_is_bitwise_identical_to_flag_protocol[8]' 'tests/test_gpu_multi.py::test_single_step_matches_the_cpu_oracle_tightly[8-1-naive-fused]' 'tests/test_gpu_multi.py::test_dp2_pp4_gpipe' 2>&1 | tail -15 | tee "$OUT/pytest.log"
echo "== scaling curve, driver style (--steps 20 --warmup 3) and long"
for n in 1 2 4 8; do
    timeout 300 python bench.py --gpus $n --steps 20 --warmup 3 2>/dev/null | tail -1 | tee -a "$OUT/scale_short.jsonl"
done
timeout 300 python bench.py --gpus 8 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp8.jsonl"
SSB_DP_LL=0 timeout 300 python bench.py --gpus 8 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp8.jsonl"
echo "== LL timeline dp8"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29743 scripts/dp_ll_timeline.py 2>/dev/null | grep -v "^\*\*\*" | tee "$OUT/dp_ll_timeline_dp8.log"
echo "== BASELINE config 4: dp=2 x pp=4 GPipe 8 micro-batches (fp32)"
for tr in nccl peer; do
    timeout 300 python bench.py --gpus 8 --pp 4 --schedule gpipe --n-mubatches 8 --pp-transport "$tr" --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp2pp4.jsonl"
done
echo "== BASELINE config 5: dp=4 x pp=2 PipeDream-flush, hidden 8192 x 15 Linears (fp32)"
timeout 600 python bench.py --gpus 8 --pp 2 --schedule pipedream --hidden 8192 --n-layers 15 --seed-mode index --steps 20 --warmup 3 --repeats 3 2>/dev/null | tail -1 | tee -a "$OUT/bench_stress.jsonl"
