#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python scripts/ab_kernels.py --tag tickets --skip-k1 2>&1 | grep "^AB" > gpurun_out/i_ab.txt
( timeout 900 python -m pytest tests/test_filter_gpu.py tests/test_golden.py tests/test_crypto_gpu.py tests/test_branches_gpu.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/i_pytest.txt
cat gpurun_out/i_ab.txt | cut -c1-400; tail -3 gpurun_out/i_pytest.txt
