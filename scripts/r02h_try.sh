#!/bin/bash
# usage (under gpurun): bash scripts/r02h_try.sh  -> witness-inheritance tests + the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inherit.py tests/test_gpu_walk.py -q -x --timeout 800 2>&1 | tail -15 | tee gpurun_out/r02h_pytest.txt
PPGPU_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solves > gpurun_out/r02h_bench_l5.json 2> gpurun_out/r02h_bench.err
tail -3 gpurun_out/r02h_bench.err
python scripts/show_bench.py < gpurun_out/r02h_bench_l5.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_l5.json').read().strip().splitlines()[-1])
print(d['roofline'].get('inherited'), d['roofline'].get('k2w'))
PY
