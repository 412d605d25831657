#!/bin/bash
# The five configurations BASELINE.json names, through the same harness (bench.py) so every number is
# device-timed, max over ranks, with the end-to-end and clock fields.  Needs an 8 x B200 node.
# Optional: COMM=fused|nccl|nvls (DP transport), PPT=nccl|peer (stage-boundary transport) - the opt-in transports
# must have passed scripts/first_multi_gpu_session.sh first.
set -u
P=${PRECISION:-fp32}
X="--comm ${COMM:-fused} --pp-transport ${PPT:-nccl}"
echo "# 1. 4-layer MLP, sequential dp=1 pp=1 on the CPU (plumbing; portable Python VM)"
python train.py --device cpu --layer-sizes 784 128 64 32 10 --steps 50 --no-eval --synthetic
echo "# 2. 8-layer-sizes MLP (reference default), dp=8 (fused in-kernel all-reduce)"
python bench.py --gpus 8 --precision $P $X
echo "# 3. pp=4 GPipe, 8 micro-batches on 4 GPUs"
python bench.py --gpus 4 --pp 4 --schedule gpipe --n-mubatches 8 --precision $P $X
echo "# 4. dp=2 x pp=4 GPipe on 8 GPUs (both comm paths active)"
python bench.py --gpus 8 --pp 4 --schedule gpipe --n-mubatches 8 --precision $P $X
echo "# 5. dp=4 x pp=2 PipeDream-flush, hidden=8192 (16 layer sizes -> 15 Linears; len(sizes) must divide by pp)"
python bench.py --gpus 8 --pp 2 --schedule pipedream --hidden 8192 --n-layers 15 --seed-mode index --steps 10 --warmup 3 --precision $P $X
