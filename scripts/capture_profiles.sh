#!/bin/bash
# Run on a B200 via gpurun (1 GPU).  Everything lands in gpurun_out/; summaries are copied
# into profiles/ by scripts/summarize_profiles.py on the CPU side.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches_*.csv
echo "== bench"; for p in fp32 tf32; do timeout 200 python bench.py --steps 500 --warmup 50 --precision $p > gpurun_out/bench_n1_$p.json 2> gpurun_out/bench_n1_$p.err; tail -c 400 gpurun_out/bench_n1_$p.json; done
echo "== launch list (one step of the flagship config, eager so every node is a plain launch; fp32 precision)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 10 --csv --log-file gpurun_out/launches_step_fp32.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --pool-batches 8 > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 8 --csv --log-file gpurun_out/launches_step_tf32.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --pool-batches 8 --precision tf32 > gpurun_out/ncu_launch2.log 2>&1
echo "== full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|tc_gemm" -s 9 -c 9 -o gpurun_out/prof_step_fp32 \
    python bench.py --steps 2 --warmup 3 --no-graph --pool-batches 8 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain" -s 3 -c 1 -o gpurun_out/prof_chain_tf32 \
    python bench.py --steps 2 --warmup 3 --no-graph --pool-batches 8 --precision tf32 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
