#!/usr/bin/env bash
# one-call, same-box comparison of build variants of the fp32 single-GPU step (side trees with their own in-tree builds)
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_m3
mkdir -p "$OUT"
run() {
    local dir=$1; shift; local label=$1; shift
    (cd "$dir" && env "$@" timeout 120 python bench.py --gpus 1 --steps 300 --warmup 30 --no-alt 2>/dev/null | tail -1) | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', 'fp32 %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'])" | tee -a "$OUT/variants.log"
}
run _old_tree "old(e506f76)" SSB_WGRAD_GROUP=1 SSB_LOSS_ZEROCOPY=1
run . "HEAD" A=1
run _v2 "v2(chain: no accumulator code)" A=1
run _v3 "v3(chain: no sync_debug code)" A=1
run _v4 "v4(chain: no ready-signal code)" A=1
run _v5 "v5(tc_gemm: no accumulator code)" A=1
run _old_tree "old(e506f76)" SSB_WGRAD_GROUP=1 SSB_LOSS_ZEROCOPY=1
run . "HEAD" A=1
