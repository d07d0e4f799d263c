#!/bin/bash
# round 2, last GPU call AD: the final tree -- whole GPU suite, smoke, bench line (cryptomatte frames now run their own classify /
# thin-lens splat instantiations; every other kernel is byte-identical to the one measured before)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/ad_pytest.txt
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/ad_smoke.txt
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench.err
tail -4 gpurun_out/ad_pytest.txt; cat gpurun_out/ad_smoke.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ad_bench.json").read().strip().splitlines()[-1]); print(d["summary"]); c=d["cryptomatte"]
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in("ms_per_step","value","crypto_dropped","ms")}) for k,v in c.items() if k!="config"})
print("thinlens", d["thinlens"]["splat"]["ms_per_step"], d["thinlens"]["rays"]["value"], "traffic", d["roofline"]["traffic"])
PY
