#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
python scripts/ab_kernels.py --tag po_redg2 --skip-k1
python scripts/ab_kernels.py --tag thin_redg2 --thin --skip-k1
} 2>&1 | grep "^AB\|Error\|error" > gpurun_out/s_ab.txt
cut -c1-330 gpurun_out/s_ab.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/s_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s_bench.json").read().strip().splitlines()[-1])
print(d["summary"]); print(json.dumps(d["cryptomatte"])[:600]); print({k:v for k,v in d["roofline"].items() if k.startswith("splat_accum") or k in ("traffic","frac","splat_frac","splat_ms")})
PY
