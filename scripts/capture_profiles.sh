#!/bin/bash
# Run on a B200 via gpurun (1 GPU).  Everything lands in gpurun_out/; summaries are copied
# into profiles/ by scripts/summarize_profiles.py on the CPU side.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches_*.csv
echo "== bench"; for p in fp32 tf32; do timeout 200 python bench.py --steps 500 --warmup 50 --precision $p > gpurun_out/bench_n1_$p.json 2> gpurun_out/bench_n1_$p.err; tail -c 400 gpurun_out/bench_n1_$p.json; done
echo "== launch list (steps of the flagship config, eager so every node is a plain launch; fp32 precision)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 12 --csv --log-file gpurun_out/launches_step_fp32.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-alt --no-graph --pool-batches 8 > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 9 --csv --log-file gpurun_out/launches_step_tf32.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-alt --no-graph --pool-batches 8 --precision tf32 > gpurun_out/ncu_launch2.log 2>&1
echo "== full captures: chain kernel + grouped wgrad (one step each), fp32 and tf32"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|tc_wgrad_group" -s 6 -c 2 -o gpurun_out/prof_step_fp32 \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-alt --no-graph --pool-batches 8 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|tc_wgrad_group" -s 6 -c 2 -o gpurun_out/prof_step_tf32 \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-alt --no-graph --pool-batches 8 --precision tf32 > gpurun_out/ncu_full2.log 2>&1
echo "== full captures: the per-layer GEMM family on a wide layer (fwd / dgrad split-K, wgrad)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 12 -c 6 -o gpurun_out/prof_wide_gemms \
    python scripts/kernel_bench.py --shapes stress --iters 2 --k-splits -1 > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | head -40
