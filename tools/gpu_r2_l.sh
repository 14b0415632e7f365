#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log | cut -c1-300
timeout 600 python tools/ab_variants.py --cfgs cfg2:10000000,cfg3:20000000,cfg1:1000000 default default@GNDT_STAGES=0 > gpurun_out/r2l_ab.log 2>&1; cat gpurun_out/r2l_ab.log | cut -c1-400
timeout 900 python tools/stream_times.py r2 60 > gpurun_out/r2l_stream.log 2>&1; tail -3 gpurun_out/r2l_stream.log | cut -c1-700
timeout 2400 python tools/full_size_parity.py cfg5 > gpurun_out/r2l_cfg5.log 2>&1; tail -c 300 gpurun_out/r2l_cfg5.log
