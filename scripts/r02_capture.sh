#!/bin/bash
# usage (under gpurun, one GPU): bash scripts/r02_capture.sh  -> tests, smoke, bench lines, launch list, ncu captures in gpurun_out/
set -x
T=${TAG:-r02}
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/${T}_pytest.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/${T}_bench_l5.err | tail -1 > gpurun_out/${T}_bench_l5.json
python scripts/show_bench.py < gpurun_out/${T}_bench_l5.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/${T}_bench_ref.err | tail -1 > gpurun_out/${T}_bench_ref.json
cut -c1-600 gpurun_out/${T}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_all.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solves > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2w_walk -s 7 -c 1 -o gpurun_out/${T}_k2w_l5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solves > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k34c_kernel -s 19 -c 1 -o gpurun_out/${T}_k34c_l5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solves > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_prefilter -s 19 -c 1 -o gpurun_out/${T}_k3p_l5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solves > /dev/null 2>&1
ls -la gpurun_out/ | tail -14
