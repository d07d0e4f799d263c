#!/bin/bash
# round 2, GPU call AC: cryptomatte pass-through adds merged inside the warp (classify): cryptomatte / filter / adaptor parity tests, bench cryptomatte leg
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_crypto_gpu.py tests/test_filter_gpu.py tests/test_adaptor_gpu.py -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/ac_pytest.txt
timeout 200 python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-splat --skip-thinlens > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err
tail -3 gpurun_out/ac_pytest.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ac_bench.json").read().strip().splitlines()[-1]); c=d["cryptomatte"]; print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in("ms_per_step","value","crypto_dropped","ms")}) for k,v in c.items() if k!="config"})
PY
