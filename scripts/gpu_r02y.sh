#!/bin/bash
# round 2, GPU call Y: whole GPU suite (incl. the operator-driven frame through both plugins), smoke, bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/y_pytest.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/y_smoke.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
tail -4 gpurun_out/y_pytest.txt; cat gpurun_out/y_smoke.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/y_bench.json").read().strip().splitlines()[-1]); print(d["summary"])
PY
