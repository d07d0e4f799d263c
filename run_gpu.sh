#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 > gpurun_out/smoke.log; cat gpurun_out/smoke.log
timeout 600 python -m pytest tests/test_golden.py -q -m gpu --tb=short 2>&1 | tail -15 > gpurun_out/golden.log; cat gpurun_out/golden.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
