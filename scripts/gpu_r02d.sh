#!/bin/bash
# round 2, GPU call D: adaptor tests with details, whole GPU suite, sanitizers on the torch-free driver, bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_adaptor_gpu.py -q -x --tb=short 2>&1 | tail -60 ) > gpurun_out/d_adaptor.txt
( timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_adaptor_gpu.py 2>&1 | tail -60 ) > gpurun_out/d_pytest.txt
( timeout 300 scripts/_build/sanitize_driver 2>&1 | tail -8 ) > gpurun_out/d_driver.txt
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -25 ) > gpurun_out/d_memcheck.txt
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -25 ) > gpurun_out/d_racecheck.txt
( timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -25 ) > gpurun_out/d_initcheck.txt
( timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/d_bench.err | tail -1 ) > gpurun_out/d_bench.json
python scripts/ab_kernels.py --tag thin --thin --spp 16 2>&1 | grep "^AB" > gpurun_out/d_ab.txt
tail -25 gpurun_out/d_adaptor.txt; tail -12 gpurun_out/d_pytest.txt; cat gpurun_out/d_driver.txt; tail -5 gpurun_out/d_memcheck.txt; tail -5 gpurun_out/d_racecheck.txt; tail -3 gpurun_out/d_initcheck.txt; cat gpurun_out/d_ab.txt
