#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2w_walk -s 0 -c 1 -o gpurun_out/k2w_l4 \
    python scripts/fam_times.py synthetic_30_6_40_s0 4 1 > gpurun_out/k2w_prof.log 2>&1
tail -3 gpurun_out/k2w_prof.log | cut -c1-400
