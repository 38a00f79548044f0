#!/bin/bash
# usage (under gpurun): bash scripts/run_gpu_suite.sh  -> tests, smoke, bench
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout 800 2>&1 | tail -15
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | python scripts/show_bench.py
