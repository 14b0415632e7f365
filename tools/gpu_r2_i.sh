#!/bin/bash
set -x
mkdir -p gpurun_out
N=8
GNDT_BENCH_TARGET_POINTS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2i_bench8.json 2> gpurun_out/r2i_bench8.err; echo "bench8 rc=$?"
cat gpurun_out/r2i_bench8.json | cut -c1-3000; grep -E "rank0\]" gpurun_out/r2i_bench8.err | head -20 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_times.py 2>/dev/null | grep world
