#!/bin/bash
# round 2, GPU call A: parity of the folded kernels + A/B timing of kernel variants + ncu of the new kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/a_pytest.txt
{
python scripts/ab_kernels.py --tag fold_default
LB_NO_FOLD=1 python scripts/ab_kernels.py --tag nofold
for v in imm k2f4 k2f6 k1f3 k1f5; do LB_LIBRARY=$PWD/variants/lib_$v.so python scripts/ab_kernels.py --tag $v; done
python scripts/ab_kernels.py --tag fold_lens43 --lens 43 --spp 4
LB_NO_FOLD=1 python scripts/ab_kernels.py --tag nofold_lens43 --lens 43 --spp 4
python scripts/ab_kernels.py --tag fold_lens0 --lens 0 --spp 8
LB_NO_FOLD=1 python scripts/ab_kernels.py --tag nofold_lens0 --lens 0 --spp 8
} 2>&1 | grep -E "^AB|Error|error" > gpurun_out/a_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat_f -c 1 -o gpurun_out/r02_k2f -f python scripts/ab_kernels.py --skip-k1 > gpurun_out/a_ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_create_rays_f -s 1 -c 1 -o gpurun_out/r02_k1f -f python scripts/ab_kernels.py --skip-k2 --spp 4 > gpurun_out/a_ncu_k1.log 2>&1
tail -3 gpurun_out/a_pytest.txt; cat gpurun_out/a_ab.txt
