#!/bin/bash
mkdir -p gpurun_out
N=8
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_pipe.py 2>/dev/null | grep world; }
CUDA_DEVICE_MAX_CONNECTIONS=32 DEPTH=3 GNDT_XCHG_SPARE=0 run
CUDA_DEVICE_MAX_CONNECTIONS=32 DEPTH=3 GNDT_XCHG_SPARE=32 run
