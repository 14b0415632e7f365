#!/bin/bash
set -x
mkdir -p gpurun_out
N=${NGPU:-2}
for ctas in 16 32 48 96; do
  if [ $ctas = 0 ]; then unset GNDT_XCHG_CTAS; else export GNDT_XCHG_CTAS=$ctas; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_times.py 2>/dev/null | grep world
done
unset GNDT_XCHG_CTAS
GATHER=voxels,slopes,columns timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_times.py 2>/dev/null | grep world
