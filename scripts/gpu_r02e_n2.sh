#!/bin/bash
# round 2, GPU call E (2 GPUs): NCCL combine parity, adaptor tests, bench at N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_adaptor_gpu.py -q --tb=short 2>&1 | tail -40 ) > gpurun_out/e_pytest_n2.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --skip-thinlens --skip-crypto 2>gpurun_out/e_bench_n2.err | grep '^{' | tail -1 ) > gpurun_out/e_bench_n2.json
( timeout 300 compute-sanitizer --tool initcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -12 ) > gpurun_out/e_initcheck.txt
tail -15 gpurun_out/e_pytest_n2.txt; python -c "
import json; d=json.load(open('gpurun_out/e_bench_n2.json')); print(d['summary']); print({k:v for k,v in d['e2e'].items() if 'link' in k and 'note' not in k})"; tail -3 gpurun_out/e_initcheck.txt
