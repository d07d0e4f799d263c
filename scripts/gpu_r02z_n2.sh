#!/bin/bash
# round 2, GPU call Z (2 GPUs): the driver's N = 2 launch of bench.py on the final tree + the multi-rank combine checks
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/z_bench.err | grep '^{' | tail -1 ) > gpurun_out/z_bench_n2.json
( timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/z_pytest_mgpu.txt
tail -2 gpurun_out/z_pytest_mgpu.txt
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/z_bench_n2.json")); print(d["summary"]); print(d["config"]["splat_partition"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/z_bench.err").read()[-1500:])
PY
