#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
python scripts/accum_bench.py
LB_NO_CDF_GUIDE=1 python scripts/ab_kernels.py --tag thin_noguide --thin
python scripts/ab_kernels.py --tag thin_guide --thin
LB_NO_CDF_GUIDE=1 python scripts/ab_kernels.py --tag po_noguide --skip-k1
python scripts/ab_kernels.py --tag po_guide --skip-k1
} 2>&1 | grep "^AB\|^ACCUM\|Error\|error" > gpurun_out/l_ab.txt
( timeout 900 python -m pytest tests/test_thinlens_gpu.py tests/test_camera_gpu.py tests/test_filter_gpu.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/l_pytest.txt
cat gpurun_out/l_ab.txt | cut -c1-600; tail -5 gpurun_out/l_pytest.txt
