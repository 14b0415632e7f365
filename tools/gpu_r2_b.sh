#!/bin/bash
# round 2, call B: shape A/B (interleaved look-back), per-pass launch lists, one full ncu capture of the passes
set -x
mkdir -p gpurun_out
V=_variants
timeout 1500 python tools/ab_variants.py --cfgs cfg2:10000000 default default@GNDT_SORT_WAVES=1000 $V/libgndt_e512x8x2m9.so $V/libgndt_e512x8x2m9.so@GNDT_SORT_WAVES=1000 $V/libgndt_f256x12x3m9.so $V/libgndt_h256x8x4m9.so $V/libgndt_h256x8x4m9.so@GNDT_SORT_WAVES=1000 $V/libgndt_k384x8x3m8.so $V/libgndt_k384x8x3m8.so@GNDT_SORT_WAVES=1000 $V/libgndt_l512x8x2m9g32.so $V/libgndt_o256x16x2m9.so $V/libgndt_p512x8x2m8.so $V/libgndt_q1024x4x1m9.so > gpurun_out/r2b_ab.log 2>&1
cat gpurun_out/r2b_ab.log
for v in e512x8x2m9 k384x8x3m8 h256x8x4m9; do
  GNDT_LIB=$PWD/$V/libgndt_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_$v.csv python tools/profile_build.py > /dev/null 2>&1
done
GNDT_LIB=$PWD/$V/libgndt_e512x8x2m9.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sort_pass -s 6 -c 3 -o gpurun_out/r2b_prof_e -f python tools/profile_build.py > gpurun_out/r2b_prof_e.log 2>&1
tail -2 gpurun_out/r2b_prof_e.log
