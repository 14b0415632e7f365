#!/bin/bash
# N=2: copy-engine transport — parity (worker test), pipelined step time against the SM push, the bench line
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_world2_strips_nccl" -m gpu -x -q 2>&1 | grep -v "^$" | tail -15 | cut -c1-400
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/xchg_pipe.py 2>&1 | grep "world\|rror" | head -3; }
TRANSPORT=ce run
TRANSPORT=sm run
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench rc=$?"
python tools/print_bench.py gpurun_out/bench_r2_${N}gpu.json "N=$N" || tail -20 gpurun_out/bench_r2_${N}gpu.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_${N}gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d['config'].get('exchange'))[:400]); print(json.dumps(d.get('parity'))[:300]); print(json.dumps(d.get('target_cfg3'))[:600])"
