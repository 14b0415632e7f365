#!/bin/bash
set -x
mkdir -p gpurun_out
V=_variants
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
timeout 1500 python tools/ab_variants.py --cfgs cfg2:10000000,cfg3:20000000 default $V/libgndt_red5.so $V/libgndt_red3.so $V/libgndt_fin3.so $V/libgndt_fin5.so $V/libgndt_red5fin5.so > gpurun_out/r2e_ab.log 2>&1
cat gpurun_out/r2e_ab.log
