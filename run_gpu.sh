#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_create_rays -s 1 -c 1 -f -o gpurun_out/prof_k1_r01final python bench.py --frame-scale 0.0625 --steps 1 --warmup 1 --skip-e2e --skip-splat --skip-cpu --skip-thinlens --skip-crypto > gpurun_out/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat -s 1 -c 1 -f -o gpurun_out/prof_k2_r01final python bench.py --frame-scale 0.25 --steps 1 --warmup 1 --skip-e2e --skip-cpu --skip-thinlens --skip-crypto > gpurun_out/ncu_k2.log 2>&1
tail -2 gpurun_out/ncu_k2.log | cut -c1-300
