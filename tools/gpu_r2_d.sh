#!/bin/bash
# round 2, call D: full ncu capture of every kernel of the second build (default lib)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -s 15 -c 15 -o gpurun_out/r2d_prof -f python tools/profile_build.py > gpurun_out/r2d_prof.log 2>&1
tail -2 gpurun_out/r2d_prof.log
