#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampled.py tests/test_gpu_graph.py -x -q --timeout 600 2>&1 | tail -5
timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | cut -c1-600
PPGPU_K34_COMPACT=0 timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 1 2>&1 | tail -4 | cut -c1-300
