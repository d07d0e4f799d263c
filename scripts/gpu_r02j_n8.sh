#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 --skip-thinlens --skip-crypto 2>gpurun_out/j_bench_n8.err | grep '^{' | tail -1 ) > gpurun_out/j_bench_n8.json
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 5 --warmup 3 --skip-thinlens --skip-crypto --skip-e2e 2>gpurun_out/j_bench_n4.err | grep '^{' | tail -1 ) > gpurun_out/j_bench_n4.json
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/run_c5.py --steps 2 --combine scatter 2>gpurun_out/j_c5_scatter.err | grep '^C5' | tail -1 ) > gpurun_out/j_c5_scatter.txt
python -c "
import json
for n in (8,4):
    d=json.load(open('gpurun_out/j_bench_n%d.json'%n)); print(d['summary']); r=d['roofline']; print('  splat ms',r['splat_ms'],'acc',r['splat_accumulate_ms'])"; cat gpurun_out/j_c5_scatter.txt
