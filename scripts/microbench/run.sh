#!/usr/bin/env bash
# Hardware probes for the next kernel designs (see probe.cu).  One GPU, a few seconds each.
#   gpurun --timeout 600 -- bash scripts/microbench/run.sh
# Results: gpurun_out/microbench.jsonl (one JSON object per line)
set -uo pipefail
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Icsrc -o /tmp/ssb_probe scripts/microbench/probe.cu || exit 1
: > gpurun_out/microbench.jsonl
for probe in edges dsmem ingest; do
    timeout 240 /tmp/ssb_probe "$probe" | tee -a gpurun_out/microbench.jsonl
done
# split-K vs plain wide GEMMs (experimental kernels, see NOTES.md)
timeout 300 python scripts/kernel_bench.py --shapes stress --iters 20 --k-splits -1 | tee -a gpurun_out/microbench.jsonl
