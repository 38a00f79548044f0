#!/bin/bash
# usage (under gpurun --gpus N): N=2 bash scripts/r02h_n2.sh -> sharded bench line (digest-checked against a 1-GPU run in the same process)
N=${N:-2}
mkdir -p gpurun_out
PPGPU_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 3 --warmup 3 --no-full-solves > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_n$N.err
grep "rank 0\|PARITY\|Error\|error" gpurun_out/r02h_n$N.err | tail -5
tail -1 gpurun_out/r02h_bench_n$N.json | python scripts/show_bench.py
tail -1 gpurun_out/r02h_bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['parity'], d['e2e']['value'])"
