#!/bin/bash
# usage (under gpurun, one GPU): bash scripts/r02i_capture.sh -> ncu captures first (the bench line reads the traffic they
# give), then tests, smoke, bench lines, launch list; everything in gpurun_out/
set -x
T=${TAG:-r02i}
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solves"
# K2w launches per solve: levels 3, 4, 5 -> the level-5 launch of the timed step is the 12th; inherit: levels 4, 5 -> the 8th
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2w_walk -s 11 -c 1 -o gpurun_out/${T}_k2w_l5 $B > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:inherit_kernel -s 7 -c 1 -o gpurun_out/${T}_inherit_l5 $B > /dev/null 2>&1
python scripts/traffic_from_rep.py gpurun_out/${T}_k2w_l5.ncu-rep profiles/r02_k2w_walk_traffic.json 70595661 \
  "ncu --set full, k2w_walk_kernel<K2wSmemDict<3,37>>, the level-5 launch of the bench workload AFTER witness inheritance (70.6 M candidates, ~10 M still open; scripts/r02i_capture.sh); algorithmic bytes: 2 x 16 B mask reads + 1 B status read per candidate + 1 B status OR per certificate"
python scripts/traffic_from_rep.py gpurun_out/${T}_inherit_l5.ncu-rep profiles/r02_inherit_traffic.json 70595661 \
  "ncu --set full, inherit_kernel, the level-5 launch of the bench workload (scripts/r02i_capture.sh)"
cp profiles/r02_k2w_walk_traffic.json profiles/r02_inherit_traffic.json gpurun_out/
python scripts/ncu_keys.py gpurun_out/${T}_k2w_l5.ncu-rep > gpurun_out/${T}_k2w_ncu.txt 2>&1
python scripts/ncu_keys.py gpurun_out/${T}_inherit_l5.ncu-rep > gpurun_out/${T}_inherit_ncu.txt 2>&1
cat gpurun_out/${T}_k2w_ncu.txt gpurun_out/${T}_inherit_ncu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/${T}_pytest.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/${T}_bench_l5.err | tail -1 > gpurun_out/${T}_bench_l5.json
python scripts/show_bench.py < gpurun_out/${T}_bench_l5.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/${T}_bench_ref.err | tail -1 > gpurun_out/${T}_bench_ref.json
cut -c1-400 gpurun_out/${T}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_all.csv $B > /dev/null 2>&1
ls -la gpurun_out/ | grep ${T}
