#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_worker.py > gpurun_out/r2g_worker.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2g_worker.log
grep -E "ok|rc=|Error" gpurun_out/r2g_worker.log | head
GNDT_BENCH_TARGET_POINTS=20000000 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err; echo "bench2 rc=$?"
cat gpurun_out/r2g_bench2.json | cut -c1-4000; grep -E "rank0\]" gpurun_out/r2g_bench2.err | head -20 | cut -c1-300
