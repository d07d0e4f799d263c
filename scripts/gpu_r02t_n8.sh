#!/bin/bash
# round 2, GPU call T (8 GPUs): peer-memory combine at N = 8: parity, config C5 (both combines on the same partials), the bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12 ) > gpurun_out/t_check.txt
( timeout 600 $TR --master-port 29513 scripts/run_c5.py --combine both --tag n8_both 2>gpurun_out/t_c5.err | grep "^C5" ) > gpurun_out/t_c5.txt
( timeout 600 $TR --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/t_bench.err | grep '^{' | tail -1 ) > gpurun_out/t_bench_n8.json
tail -3 gpurun_out/t_check.txt; cut -c1-900 gpurun_out/t_c5.txt; tail -3 gpurun_out/t_c5.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/t_bench_n8.json")); print(d["summary"]); print({k:v for k,v in d["roofline"].items() if k in ("splat_ms","splat_accumulate_ms","splat_value")})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/t_bench.err").read()[-1500:])
PY
