#!/bin/bash
# round 2, call A: parity of the restructured sort + shape A/B + launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 1200 python tools/ab_variants.py --cfgs cfg2:10000000,cfg3:20000000,cfg1:1000000 default _variants/libgndt_b384x8x3m9.so _variants/libgndt_c256x16x2.so _variants/libgndt_d512x16x1.so _variants/libgndt_e512x8x2m9.so _variants/libgndt_f256x12x3m9.so _variants/libgndt_g384x8x2m8.so > gpurun_out/r2a_ab.log 2>&1
cat gpurun_out/r2a_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python tools/profile_build.py > gpurun_out/r2a_prof.log 2>&1
tail -3 gpurun_out/r2a_prof.log
