#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "0 1" "1 1"; do
  set -- $cfg
  PPGPU_ASSEMBLY_DEVICE=$1 PPGPU_ASSEMBLY_PAUSE_GC=$2 PPGPU_BENCH_VERBOSE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-full-solves > gpurun_out/r02k_bench_$1$2.json 2> gpurun_out/r02k_bench_$1$2.err
  echo "device=$1 pause_gc=$2"; grep "e2e ms" gpurun_out/r02k_bench_$1$2.err; python scripts/show_bench.py < gpurun_out/r02k_bench_$1$2.json | head -1
done
