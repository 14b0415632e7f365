#!/bin/bash
set -x
mkdir -p gpurun_out
N=8
for spare in 0 32; do
  GNDT_XCHG_SPARE=$spare GNDT_BENCH_TARGET_POINTS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/r2o_bench${N}_s$spare.json 2> gpurun_out/r2o_bench${N}_s$spare.err; echo "bench rc=$?"
  python tools/print_bench.py gpurun_out/r2o_bench${N}_s$spare.json "spare $spare" || grep -E "rank0\]" gpurun_out/r2o_bench${N}_s$spare.err | head -12 | cut -c1-300
done
