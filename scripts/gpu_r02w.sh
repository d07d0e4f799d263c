#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --skip-thinlens --skip-crypto --skip-cpu > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/w_bench.json").read().strip().splitlines()[-1])
print(d["summary"]); print({k:v for k,v in d["e2e"].items() if k.startswith("splat")})
PY
( timeout 600 python -m pytest tests/test_filter_gpu.py tests/test_crypto_gpu.py -m gpu -q -x 2>&1 | tail -3 )
