#!/bin/bash
# usage (under gpurun, one GPU): bash scripts/final_capture.sh  -> bench lines, launch list, ncu captures in gpurun_out/
set -x
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_final_l5.err | tail -1 > gpurun_out/bench_final_l5.json
python scripts/show_bench.py < gpurun_out/bench_final_l5.json
timeout 300 python bench.py --steps 3 --warmup 3 --levels 4 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_final_l4.json
python scripts/show_bench.py < gpurun_out/bench_final_l4.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches_all.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2a_relax_reg -s 30 -c 1 -o gpurun_out/r01_k2a_final_l5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2_feas -s 30 -c 1 -o gpurun_out/r01_k2_final_l5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k34_kernel -s 30 -c 1 -o gpurun_out/r01_k34_final_l5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
