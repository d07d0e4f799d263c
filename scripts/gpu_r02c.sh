#!/bin/bash
# round 2, GPU call C: whole GPU suite (no -x), variants, compute-sanitizer on the smoke workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 ) > gpurun_out/c_pytest.txt
{
for v in k2w4 k2w6 k1w3 k1w5; do LB_LIBRARY=$PWD/variants/lib_$v.so python scripts/ab_kernels.py --tag $v; done
} 2>&1 | grep -E "^AB|Error|error" > gpurun_out/c_ab.txt
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke 2>&1 | tail -15 ) > gpurun_out/c_memcheck.txt
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke 2>&1 | tail -15 ) > gpurun_out/c_racecheck.txt
tail -30 gpurun_out/c_pytest.txt; cat gpurun_out/c_ab.txt; tail -4 gpurun_out/c_memcheck.txt; tail -4 gpurun_out/c_racecheck.txt
