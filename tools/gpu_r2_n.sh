#!/bin/bash
set -x
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_worker.py > gpurun_out/r2n_worker.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2n_worker.log
grep -E "ok|rc=|Error" gpurun_out/r2n_worker.log | head
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2n_multi.log 2>&1; tail -3 gpurun_out/r2n_multi.log
for spare in 0 32; do
  GNDT_XCHG_SPARE=$spare GNDT_BENCH_TARGET_POINTS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/r2n_bench${N}_s$spare.json 2> gpurun_out/r2n_bench${N}_s$spare.err; echo "bench rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2n_bench${N}_s$spare.json').read().strip().splitlines()[-1])
print('spare $spare value %.3g ms/step %.3f serial %.3f e2e %.3g (%.2f ms) d2h %d' % (d['value'], d['ms_per_step'], d['serial_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step']))" || grep -E "rank0\]" gpurun_out/r2n_bench${N}_s$spare.err | head -12 | cut -c1-300
done
