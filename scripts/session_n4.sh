#!/usr/bin/env bash
# Round-2 session on 4 GPUs: dp=4 LL kernel (correctness + exposed comm), pp=4 pipeline (BASELINE config 3), timelines.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_n4
mkdir -p "$OUT"
echo "== tests: dp4 LL bitwise vs flag protocol, pp4, pp4 peer"
timeout 300 python -m pytest tests/test_gpu_engine.py -q -k "per_microbatch" --maxfail=3 2>&1 | tail -3
timeout 900 python -m pytest -q --maxfail=3 'tests/test_gpu_multi.py::test_dp_ll_kernel_is_bitwise_identical_to_flag_protocol[4]' 'tests/test_gpu_multi.py::test_pp4_1f1b' 'tests/test_gpu_multi.py::test_pp_peer_transport_matches_oracle[1-4-gpipe]' 'tests/test_gpu_multi.py::test_dp2_pp2_gpipe_fused' 'tests/test_gpu_multi.py::test_single_step_matches_the_cpu_oracle_tightly[2-2-gpipe-fused]' 'tests/test_gpu_multi.py::test_dp4_fused' 2>&1 | tail -15 | tee "$OUT/pytest.log"
echo "== bench dp4: LL gated / flag protocol"
for env in "" "SSB_DP_LL=0"; do
    env $env timeout 300 python bench.py --gpus 4 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp4.jsonl"
done
echo "== LL timeline dp4"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 scripts/dp_ll_timeline.py 2>/dev/null | grep -v "^\*\*\*" | tee "$OUT/dp_ll_timeline_dp4.log"
echo "== BASELINE config 3: pp=4 GPipe 8 micro-batches (fp32), nccl vs peer, + 1F1B"
for tr in nccl peer; do
    timeout 300 python bench.py --gpus 4 --pp 4 --schedule gpipe --n-mubatches 8 --pp-transport "$tr" --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp4.jsonl"
done
timeout 300 python bench.py --gpus 4 --pp 4 --schedule pipedream --n-mubatches 8 --pp-transport peer --steps 200 --warmup 30 2>/dev/null | tail -1 | tee -a "$OUT/bench_pp4.jsonl"
echo "== reference arm, same config (pp=4 GPipe 8 micro-batches)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29742 bench.py --impl reference --gpus 4 --pp 4 --schedule gpipe --n-mubatches 8 --steps 50 --warmup 5 2>/dev/null | grep '^{' | tee -a "$OUT/bench_pp4_reference.jsonl"
