#!/bin/bash
mkdir -p gpurun_out
N=8
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_pipe.py 2>/dev/null | grep world; }
GNDT_XCHG_SPARE=0 GNDT_XCHG_SKIP=4 run
GNDT_XCHG_SPARE=0 GNDT_XCHG_SKIP=0 run
GNDT_XCHG_SPARE=32 GNDT_XCHG_SKIP=0 run
GNDT_XCHG_SPARE=64 GNDT_XCHG_SKIP=0 run
GNDT_XCHG_SPARE=32 GNDT_XCHG_CTAS=96 GNDT_XCHG_SKIP=0 run
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/h2d_bw.py 2>/dev/null | grep world
