#!/usr/bin/env bash
# First GPU session of a new round: validate the opt-in code written without hardware, then collect the
# probe data the next kernel designs depend on.  One GPU, ~6-8 minutes.
#   gpurun --timeout 900 -- bash scripts/first_gpu_session.sh
# Everything lands in gpurun_out/first_session/.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/first_session
mkdir -p "$OUT"

echo "== 1. experimental kernels + aux subsystems (xfail markers ignored: real outcome wanted)"
timeout 300 python -m pytest tests/test_gpu_zz_aux.py -q -k "not experimental" 2>&1 | tail -5 | tee "$OUT/aux_tests.log"
for group in splitk weight_lo chain_multicast loss_zero_copy wgrad_group_launch two_node_step; do
    echo "-- experimental group: $group (own process)"
    timeout 300 python -m pytest tests/experimental_cases.py -k "$group" -q -x -p no:cacheprovider 2>&1 | tail -15 | tee -a "$OUT/experimental_tests.log"
done

echo "== 2. headline bench: default vs opt-in variants (fp32)"
for env in "" "SSB_FUSE_WLO=1" "SSB_CHAIN_MC=1" "SSB_LOSS_ZEROCOPY=1" "SSB_WGRAD_GROUP=1" "SSB_WGRAD_GROUP=1 SSB_FUSE_WLO=1 SSB_LOSS_ZEROCOPY=1" "SSB_WGRAD_GROUP=1 SSB_FUSE_WLO=1 SSB_LOSS_ZEROCOPY=1 SSB_CHAIN_MC=1"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 1 --steps 300 --warmup 50 2>/dev/null | tail -1 | tee -a "$OUT/bench_variants.jsonl"
done

echo "== 3. wide layers: plain vs split-K (kernel level, then one stage of the stress model)"
timeout 300 python scripts/kernel_bench.py --shapes stress --iters 20 --k-splits -1 | tee "$OUT/kernel_bench_splitk.jsonl"
for env in "" "SSB_SPLITK=1"; do
    echo "-- env: ${env:-default}"
    env $env timeout 300 python bench.py --gpus 1 --hidden 8192 --n-layers 4 --seed-mode index --precision tf32 \
        --steps 20 --warmup 5 2>/dev/null | tail -1 | tee -a "$OUT/bench_wide_variants.jsonl"
done

echo "== 4. hardware probes (TMA ingest vs ring depth, multicast, DSMEM, graph edges)"
bash scripts/microbench/run.sh > "$OUT/microbench.log" 2>&1
cp gpurun_out/microbench.jsonl "$OUT/" 2>/dev/null || true
tail -5 "$OUT/microbench.log"
