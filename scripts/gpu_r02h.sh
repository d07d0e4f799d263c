#!/bin/bash
# round 2, GPU call H: chunked splat work units + fast K1 epilogue: parity, timing, bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/h_pytest.txt
{
python scripts/ab_kernels.py --tag final --skip-k2
python scripts/ab_kernels.py --tag final --skip-k1
python scripts/ab_kernels.py --tag final_lens43 --lens 43 --spp 4 --skip-k2
python scripts/ab_kernels.py --tag final_lens0 --lens 0 --spp 8 --skip-k2
} 2>&1 | grep -E "^AB|Error|error" > gpurun_out/h_ab.txt
( timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/h_bench.err | tail -1 ) > gpurun_out/h_bench.json
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -4 ) > gpurun_out/h_racecheck.txt
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -4 ) > gpurun_out/h_memcheck.txt
tail -8 gpurun_out/h_pytest.txt; cat gpurun_out/h_ab.txt | cut -c1-500; python -c "
import json; d=json.load(open('gpurun_out/h_bench.json')); print(d['summary'])"; tail -2 gpurun_out/h_racecheck.txt gpurun_out/h_memcheck.txt
