#!/bin/bash
# round 2, GPU call V (2 GPUs): collective failure flag of the peer mapping, bench fallback logic, thin-lens splat after the red_add revert
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -4 ) > gpurun_out/v_check.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip-thinlens --skip-crypto 2>gpurun_out/v_bench.err | grep '^{' | tail -1 ) > gpurun_out/v_bench_n2.json
python scripts/ab_kernels.py --tag thin_final --thin --skip-k1 2>&1 | grep "^AB" | cut -c1-330 > gpurun_out/v_ab.txt
tail -2 gpurun_out/v_check.txt; cat gpurun_out/v_ab.txt
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/v_bench_n2.json")); print(d["summary"]); print(d["config"]["splat_partition"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/v_bench.err").read()[-1500:])
PY
