#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inherit.py tests/test_gpu_walk.py -q -x --timeout 500 2>&1 | tail -3
timeout 200 python scripts/shard_probe2.py 8 0 5 2>&1 | tail -3 | cut -c1-400
PPGPU_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solves > gpurun_out/r02j_bench_l5.json 2> gpurun_out/r02j_bench.err
tail -1 gpurun_out/r02j_bench.err
python scripts/show_bench.py < gpurun_out/r02j_bench_l5.json | head -1
