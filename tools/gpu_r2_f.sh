#!/bin/bash
# round 2, call F (2 GPUs): native strip exchange parity + 2-GPU bench
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_worker.py > gpurun_out/r2f_worker.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2f_worker.log
tail -25 gpurun_out/r2f_worker.log


