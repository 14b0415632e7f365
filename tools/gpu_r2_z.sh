#!/bin/bash
# N=4: copy-engine transport against the SM push (pipelined step time), then the strip worker on 4 ranks
N=${NGPU:-4}
run() { timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_pipe.py 2>&1 | grep "world\|rror" | head -3; }
TRANSPORT=ce run
TRANSPORT=sm run
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tests/multi_gpu_worker.py 2>&1 | grep "multi-gpu\|parity ok\|rror\|ssert" | head -12 | cut -c1-200
