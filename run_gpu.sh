#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err; cut -c1-900 gpurun_out/bench_full.json
