#!/bin/bash
for item in 1024 2048 4096; do
  for chunk in 4194304 16777216; do
    echo "== item $item chunk $chunk"
    PPGPU_K2W_REG=0 PPGPU_K2W_ITEM=$item PPGPU_CHUNK=$chunk timeout 300 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | tail -4 | head -2 | cut -c1-330
  done
done
