#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_walk.py -x -q --timeout 300 2>&1 | tail -2
for cfg in "0 1024" "1 512" "1 1024" "1 2048"; do
  set -- $cfg
  echo "== groups $1 split $2"
  PPGPU_K2W_GROUPS=$1 PPGPU_K2W_SPLIT=$2 timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | head -3 | cut -c1-1000
done
