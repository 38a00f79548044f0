#!/bin/bash
mkdir -p gpurun_out
for v in 0; do
PPGPU_WALK_LAST=$v PPGPU_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solves > gpurun_out/r02h_bench_wl$v.json 2> gpurun_out/r02h_bench_wl$v.err
tail -2 gpurun_out/r02h_bench_wl$v.err
python scripts/show_bench.py < gpurun_out/r02h_bench_wl$v.json
done
