#!/bin/bash
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_$tool.log python tools/sanitize.py > gpurun_out/r2_sanitizer_$tool.out 2>&1; echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.out
  tail -3 gpurun_out/r2_sanitizer_$tool.log
done
timeout 900 python tools/full_size_parity.py cfg2 > gpurun_out/r2k_cfg2.log 2>&1; tail -c 600 gpurun_out/r2k_cfg2.log
timeout 1500 python tools/full_size_parity.py cfg3 > gpurun_out/r2k_cfg3.log 2>&1; tail -c 600 gpurun_out/r2k_cfg3.log
timeout 2400 python tools/full_size_parity.py cfg5 > gpurun_out/r2k_cfg5.log 2>&1; tail -c 600 gpurun_out/r2k_cfg5.log
