#!/bin/bash
# round 2, call C: parity after grouped look-back / PDL / lean bounds; PDL and stage-event A/B; launch list
set -x
mkdir -p gpurun_out
V=_variants
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
timeout 1500 python tools/ab_variants.py --cfgs cfg2:10000000 default default@GNDT_STAGES=0 default@GNDT_STAGES=0@GNDT_NO_PDL=1 default@GNDT_NO_PDL=1 $V/libgndt_k384x8x3m8.so $V/libgndt_f256x12x3m9.so $V/libgndt_s512x8x2m9g8.so $V/libgndt_t512x6x3m9.so > gpurun_out/r2c_ab.log 2>&1
cat gpurun_out/r2c_ab.log
timeout 900 python tools/ab_variants.py --cfgs cfg3:20000000,cfg1:1000000 default default@GNDT_STAGES=0 > gpurun_out/r2c_ab2.log 2>&1
cat gpurun_out/r2c_ab2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv python tools/profile_build.py > gpurun_out/r2c_prof.log 2>&1
