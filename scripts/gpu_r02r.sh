#!/bin/bash
# round 2, GPU call R: REDG reductions in the PO splat kernel, shared cryptomatte total plane: parity + timing + bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/r_pytest.txt
{
python scripts/ab_kernels.py --tag po_redg --skip-k1
python scripts/ab_kernels.py --tag thin_redg --thin --skip-k1
} 2>&1 | grep "^AB\|Error\|error" > gpurun_out/r_ab.txt
( timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/r_bench.err | tail -1 ) > gpurun_out/r_bench.json
tail -4 gpurun_out/r_pytest.txt; cut -c1-420 gpurun_out/r_ab.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r_bench.json"))
print(d["summary"]); print(json.dumps(d["cryptomatte"])[:600]); print({k:v for k,v in d["roofline"].items() if k.startswith("splat_accum") or k in ("traffic","frac","splat_frac","splat_ms")})
PY
tail -2 gpurun_out/r_bench.err
