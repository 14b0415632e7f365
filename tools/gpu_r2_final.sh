#!/bin/bash
# round 2 final single-GPU evidence: launch list + full ncu capture of one build, all configs at full size, the bench line
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/profile_build.py > gpurun_out/r2_prof_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 16 -c 14 -o gpurun_out/prof_r2 -f python tools/profile_build.py > gpurun_out/r2_prof_full.log 2>&1
tail -2 gpurun_out/r2_prof_full.log
timeout 1800 python tools/bench_configs.py --tag r2 --big > gpurun_out/r2_configs.log 2>&1; tail -16 gpurun_out/r2_configs.log | cut -c1-250
cp profiles/configs_r2.json gpurun_out/configs_r2.json
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err; echo "bench rc=$?"
cat gpurun_out/bench_r2_1gpu.json | cut -c1-2500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2>/dev/null; cat gpurun_out/bench_r2_ref.json | cut -c1-600
