#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_thinlens_gpu.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/n_pytest.txt
{
for mode in 0 1; do
LB_SPLAT_TILE=$mode python scripts/ab_kernels.py --tag thin_c3_tile$mode --thin --skip-k1
LB_SPLAT_TILE=$mode python scripts/ab_kernels.py --tag thin_f57_r0.4_tile$mode --thin --skip-k1 --focus 57 --disc-radius 0.4
LB_SPLAT_TILE=$mode python scripts/ab_kernels.py --tag thin_f45_r0.4_tile$mode --thin --skip-k1 --focus 45 --disc-radius 0.4
LB_SPLAT_TILE=$mode python scripts/ab_kernels.py --tag thin_8k_f57_tile$mode --thin --skip-k1 --focus 66 --disc-radius 0.2 --k2-size 7680x4320
done
python scripts/ab_kernels.py --tag thin_f57_r0.4_auto --thin --skip-k1 --focus 57 --disc-radius 0.4
} 2>&1 | grep "^AB\|Error\|error" > gpurun_out/n_ab.txt
cat gpurun_out/n_ab.txt | cut -c1-700; tail -8 gpurun_out/n_pytest.txt
