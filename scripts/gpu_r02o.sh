#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_thinlens_gpu.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/o_pytest.txt
( timeout 900 python scripts/sweep_lenses.py --json gpurun_out/c4_sweep.json 2>&1 ) > gpurun_out/o_sweep.txt
( timeout 900 python bench.py 2>gpurun_out/o_bench.err ) > gpurun_out/o_bench.json
tail -4 gpurun_out/o_pytest.txt; tail -3 gpurun_out/o_sweep.txt; cut -c1-1500 gpurun_out/o_bench.json; tail -3 gpurun_out/o_bench.err
