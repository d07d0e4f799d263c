#!/bin/bash
# round 2, GPU call Q (2 GPUs): peer-memory combine: parity (multi_gpu_check) and the N = 2 splat step, peer vs reduce-scatter
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -25 ) > gpurun_out/q_check.txt
for comb in peer scatter; do
( LB_BENCH_COMBINE=$comb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu --skip-thinlens --skip-crypto 2>gpurun_out/q_bench_$comb.err | grep '^{' | tail -1 ) > gpurun_out/q_bench_$comb.json
done
tail -6 gpurun_out/q_check.txt
python - <<'PY'
import json
for c in ("peer","scatter"):
    try:
        d=json.load(open(f"gpurun_out/q_bench_{c}.json"))
        print(c, d["summary"], d["roofline"].get("splat_ms"), d["roofline"].get("splat_accumulate_ms"))
    except Exception as e:
        print(c, "failed", e); print(open(f"gpurun_out/q_bench_{c}.err").read()[-1500:])
PY
