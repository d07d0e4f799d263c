#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
LB_ADD_ZEROS=1 python scripts/run_c5.py --emulate-world 8 --order row --tag po_addzeros_row
python scripts/run_c5.py --emulate-world 8 --order row --tag po_skip_row
python scripts/run_c5.py --emulate-world 8 --order bucket --tag po_skip_bucket
LB_ADD_ZEROS=1 python scripts/run_c5.py --emulate-world 8 --order row --thin --tag thin_addzeros_row
python scripts/run_c5.py --emulate-world 8 --order row --thin --tag thin_skip_row
python scripts/run_c5.py --emulate-world 8 --order bucket --thin --tag thin_skip_bucket
LB_ADD_ZEROS=1 python scripts/run_c5.py --emulate-world 8 --order bucket --thin --tag thin_addzeros_bucket
} 2>&1 | grep "^C5\|Error\|error" > gpurun_out/m_c5.txt
( timeout 1200 python -m pytest tests/test_thinlens_gpu.py tests/test_camera_gpu.py tests/test_filter_gpu.py tests/test_golden.py tests/test_crypto_gpu.py tests/test_branches_gpu.py -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/m_pytest.txt
cat gpurun_out/m_c5.txt | cut -c1-700; tail -5 gpurun_out/m_pytest.txt
