#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/launches_bench.log 2>&1
tail -c 300 gpurun_out/launches_bench.log
