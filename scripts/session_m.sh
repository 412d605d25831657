#!/usr/bin/env bash
# same-box comparison of the single-GPU step: session-H code (side tree with its own in-tree build) vs HEAD
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out/session_m
mkdir -p "$OUT"
echo "== HEAD: single-GPU suite"
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4 | tee "$OUT/pytest_gpu.log"
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -s -k 3xtf32 2>&1 | grep -E "3xTF32 vs cuBLAS fp32 \[fwd|passed|failed" | tee "$OUT/precision.log"
run() {  # dir label env...
    local dir=$1; shift; local label=$1; shift
    (cd "$dir" && env "$@" timeout 200 python bench.py --gpus 1 --steps 300 --warmup 50 2>/dev/null | tail -1) | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', 'fp32 %.4f' % d['ms_per_step'], 'tf32 %.4f' % d.get('tf32_mode',{}).get('ms_per_step',0), 'e2e %.4f' % d['e2e']['ms_per_step'], 'nodes', d['config']['graph_nodes'])" | tee -a "$OUT/bisect.log"
}
for rep in 1 2; do
    [ -d _old_tree ] && run _old_tree "e506f76(sessionH)" SSB_WGRAD_GROUP=1 SSB_LOSS_ZEROCOPY=1
    run . "HEAD" A=1
    run . "HEAD,ACC_SPLIT=0" SSB_ACC_SPLIT=0
done
