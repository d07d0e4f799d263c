#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python scripts/sweep_lenses.py > gpurun_out/c4_sweep.txt 2> gpurun_out/c4_sweep.err
tail -3 gpurun_out/c4_sweep.err; wc -l gpurun_out/c4_sweep.txt
