#!/bin/bash
set -x
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench rc=$?"
python tools/print_bench.py gpurun_out/bench_r2_${N}gpu.json "N=$N" || grep -E "rank0\]" gpurun_out/bench_r2_${N}gpu.err | head -12 | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_${N}gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('parity'))[:600]); print(json.dumps(d.get('target_cfg3'))[:900])"
