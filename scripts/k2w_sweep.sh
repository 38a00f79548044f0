#!/bin/bash
for cfg in "0 1024" "0 512" "0 2048" "1 1024"; do
  set -- $cfg
  echo "== groups $1 item $2"
  PPGPU_K2W_GROUPS=$1 PPGPU_K2W_ITEM=$2 timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | head -3 | cut -c1-1000
done
