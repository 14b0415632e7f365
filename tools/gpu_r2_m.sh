#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log | cut -c1-300
timeout 600 python tools/ab_variants.py --cfgs cfg2:10000000,cfg3:20000000 default _variants/libgndt_lbt0.so default _variants/libgndt_lbt0.so > gpurun_out/r2m_ab.log 2>&1; cat gpurun_out/r2m_ab.log | cut -c1-400
