#!/bin/bash
# round 2, GPU call U (8 GPUs): peer combine with the receiver owning no slab: parity at N = 8, config C5
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12 ) > gpurun_out/u_check.txt
( timeout 600 $TR --master-port 29513 scripts/run_c5.py --combine both --tag n8_both_rootless 2>gpurun_out/u_c5.err | grep "^C5" ) > gpurun_out/u_c5.txt
tail -3 gpurun_out/u_check.txt; cut -c1-900 gpurun_out/u_c5.txt; tail -3 gpurun_out/u_c5.err
