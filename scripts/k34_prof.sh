#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k34c_kernel -s 3 -c 1 -o gpurun_out/k34c_l5 \
    python scripts/fam_times.py synthetic_30_6_40_s0 5 1 > gpurun_out/k34c_prof.log 2>&1
tail -3 gpurun_out/k34c_prof.log | cut -c1-300
