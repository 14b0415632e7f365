#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -4 gpurun_out/r2t_pytest.log | cut -c1-300
bash tools/gpu_r2_final.sh
