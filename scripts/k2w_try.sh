#!/bin/bash
# quick loop for the K2w walk: tests, per-family times with the register / shared-memory dictionary and with the walk off
timeout 600 python -m pytest tests/test_gpu_walk.py -x -q --timeout 300 2>&1 | tail -5
timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | cut -c1-900
PPGPU_K2W_REG=0 timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 1 2>&1 | tail -4 | cut -c1-900
