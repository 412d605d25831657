#!/usr/bin/env bash
# First multi-GPU session of a new round (2 GPUs is enough): validate the NVLS path written without hardware
# and compare the three DP transports on the headline workload.
#   gpurun --gpus 2 --timeout 900 -- bash scripts/first_multi_gpu_session.sh
set -uo pipefail
cd "$(dirname "$0")/.."
N=${N_GPUS:-2}
OUT=gpurun_out/first_multi_session
mkdir -p "$OUT"

echo "== 1. NVLS tests (xfail markers ignored)"
timeout 600 python -m pytest tests/test_gpu_multi.py -k nvls -q --runxfail -x 2>&1 | tail -30 | tee "$OUT/nvls_tests.log"

echo "== 1b. peer-memory pipeline transport tests + pp=2 bench (nccl vs peer)"
timeout 600 python -m pytest tests/test_gpu_multi.py -k "pp_peer" -q --runxfail -x 2>&1 | tail -30 | tee "$OUT/pp_peer_tests.log"
for tr in nccl peer; do
    timeout 300 python bench.py --gpus "$N" --pp "$N" --schedule gpipe --n-mubatches 8 --pp-transport "$tr" --steps 200 --warmup 30 2>/dev/null \
        | tail -1 | tee -a "$OUT/bench_pp_transports.jsonl"
done

echo "== 2. switch all-reduce vs NCCL"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29731 \
    scripts/nvls_bench.py 2>&1 | grep '^{' | tee "$OUT/nvls_bench.jsonl"

echo "== 3. headline bench, one line per DP transport"
for comm in fused nccl nvls; do
    timeout 300 python bench.py --gpus "$N" --comm "$comm" --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp_transports.jsonl"
done

echo "-- nvls + grouped wgrad + the single-GPU opt-ins that passed (edit the list after first_gpu_session.sh)"
SSB_WGRAD_GROUP=1 timeout 300 python bench.py --gpus "$N" --comm nvls --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp_transports.jsonl"

echo "== 4. wide model (two-shot regime): hidden 4096 x 4"
for comm in fused nccl nvls; do
    timeout 300 python bench.py --gpus "$N" --comm "$comm" --hidden 4096 --n-layers 4 --seed-mode index --precision tf32 \
        --steps 20 --warmup 5 2>/dev/null | tail -1 | tee -a "$OUT/bench_dp_transports_wide.jsonl"
done
