#!/bin/bash
# Run on a B200 via gpurun (1 GPU).  Everything lands in gpurun_out/; summaries are copied
# into profiles/ by scripts/summarize_profiles.py on the CPU side.
set -u
mkdir -p gpurun_out
echo "== kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q 2>&1 | tail -15
echo "== bench"; timeout 300 python bench.py --steps 500 --warmup 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
echo "== kernel bench"; timeout 600 python scripts/kernel_bench.py --shapes all --iters 20 > gpurun_out/kernel_bench.jsonl 2>&1; tail -3 gpurun_out/kernel_bench.jsonl
echo "== launch list (one step, eager so every node is a plain launch)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 63 -c 21 --csv --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --pool-batches 8 > gpurun_out/ncu_launch.log 2>&1
echo "== full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|loss_head" -s 63 -c 21 -o gpurun_out/prof_step \
    python bench.py --steps 2 --warmup 3 --no-graph --pool-batches 8 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
