#!/bin/bash
# round 2, final GPU call X: whole GPU suite, smoke, bench line, launch list of the bench command, sanitizer on the final kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/x_pytest.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/x_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/x_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/x_launches_bench.log 2>&1
( timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -6 ) > gpurun_out/x_memcheck.txt
( LB_SPLAT_TILE=1 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -6 ) > gpurun_out/x_racecheck_tile.txt
( LB_SPLAT_TILE=1 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -6 ) > gpurun_out/x_memcheck_tile.txt
tail -3 gpurun_out/x_pytest.txt; cat gpurun_out/x_smoke.txt; tail -2 gpurun_out/x_memcheck.txt gpurun_out/x_racecheck_tile.txt gpurun_out/x_memcheck_tile.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/x_bench.json").read().strip().splitlines()[-1]); print(d["summary"])
PY
wc -l gpurun_out/x_launches.csv
