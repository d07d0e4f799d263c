#!/bin/bash
# round 2, GPU call E (8 GPUs): bench at N=8 and config C5 with both combines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/e_topo.txt 2>&1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 --skip-thinlens --skip-crypto 2>gpurun_out/e_bench_n8.err | grep '^{' | tail -1 ) > gpurun_out/e_bench_n8.json
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/run_c5.py --steps 2 --combine scatter 2>gpurun_out/e_c5_scatter.err | grep '^C5' | tail -1 ) > gpurun_out/e_c5_scatter.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 scripts/run_c5.py --steps 2 --combine reduce 2>gpurun_out/e_c5_reduce.err | grep '^C5' | tail -1 ) > gpurun_out/e_c5_reduce.txt
python -c "
import json; d=json.load(open('gpurun_out/e_bench_n8.json')); print(d['summary']); print({k:v for k,v in d['e2e'].items() if 'link' in k and 'note' not in k})"; cat gpurun_out/e_c5_scatter.txt gpurun_out/e_c5_reduce.txt
