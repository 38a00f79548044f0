#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_walk.py -x -q --timeout 300 2>&1 | tail -3
timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | cut -c1-900
