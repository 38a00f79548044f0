#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_install.py tests/test_gpu_sampled.py -q -x --timeout 250 2>&1 | tail -3 | tee gpurun_out/r02k_pytest.txt
PPGPU_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_l5.json 2> gpurun_out/r02k_bench.err
tail -1 gpurun_out/r02k_bench.err
python scripts/show_bench.py < gpurun_out/r02k_bench_l5.json | head -1
