#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_walk.py tests/test_gpu_sampled.py -x -q --timeout 500 2>&1 | tail -3
timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | cut -c1-900
PPGPU_K2W_HYBRID=0 timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 1 2>&1 | tail -4 | head -2 | cut -c1-400
