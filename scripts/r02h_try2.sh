#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inherit.py -q -x --timeout 800 2>&1 | tail -3
PPGPU_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solves > gpurun_out/r02h_bench_s3.json 2> gpurun_out/r02h_bench_s3.err
tail -1 gpurun_out/r02h_bench_s3.err
python scripts/show_bench.py < gpurun_out/r02h_bench_s3.json | head -1
python -c "
import json
d=json.loads(open('gpurun_out/r02h_bench_s3.json').read().strip().splitlines()[-1])
print(d['roofline'].get('inherited')['certified'], d['roofline'].get('k2w'))"
