#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_m2
mkdir -p "$OUT"
echo "== OLD (e506f76) chain timeline fp32"
(cd _old_tree && SSB_WGRAD_GROUP=1 SSB_LOSS_ZEROCOPY=1 timeout 120 python scripts/chain_timeline.py 2>&1 | grep -E "^gemm" | cut -c1-120) | tee "$OUT/timeline_old.log"
echo "== HEAD chain timeline fp32"
timeout 120 python scripts/chain_timeline.py 2>&1 | grep -E "^gemm" | cut -c1-120 | tee "$OUT/timeline_head.log"
