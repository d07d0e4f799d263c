#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
python scripts/ab_kernels.py --tag base --skip-k1
LB_LIBRARY=$PWD/variants/lib_ring.so python scripts/ab_kernels.py --tag ring --skip-k1
LB_LIBRARY=$PWD/variants/lib_ring.so python scripts/ab_kernels.py --tag ring_k1 --skip-k2
} 2>&1 | grep "^AB" > gpurun_out/k_ab.txt
( LB_LIBRARY=$PWD/variants/lib_ring.so timeout 900 python -m pytest tests/test_filter_gpu.py tests/test_golden.py tests/test_branches_gpu.py -m gpu -q -k "not table" 2>&1 | tail -5 ) > gpurun_out/k_pytest.txt
cat gpurun_out/k_ab.txt | cut -c1-330; tail -3 gpurun_out/k_pytest.txt
