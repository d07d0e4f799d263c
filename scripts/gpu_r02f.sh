#!/bin/bash
# round 2, GPU call F: Horner-form bodies vs the shared-monomial DAG bodies, occupancy variants, parity of the new default
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
python scripts/ab_kernels.py --tag horner
LB_KERNEL_GEN=1 python scripts/ab_kernels.py --tag horner_table
for v in dag hk2b6 hk2b4 hk1b5 hk1b3; do LB_LIBRARY=$PWD/variants/lib_$v.so python scripts/ab_kernels.py --tag $v; done
python scripts/ab_kernels.py --tag horner_lens43 --lens 43 --spp 4
python scripts/ab_kernels.py --tag horner_lens0 --lens 0 --spp 8
} 2>&1 | grep -E "^AB|Error|error" > gpurun_out/f_ab.txt
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/f_pytest.txt
( timeout 300 compute-sanitizer --tool initcheck --error-exitcode 9 scripts/_build/sanitize_driver 2>&1 | tail -6 ) > gpurun_out/f_initcheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_create_rays -s 1 -c 1 -o gpurun_out/r02h_k1 -f python scripts/ab_kernels.py --skip-k2 --spp 4 > gpurun_out/f_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat -c 1 -o gpurun_out/r02h_k2 -f python scripts/ab_kernels.py --skip-k1 > gpurun_out/f_ncu_k2.log 2>&1
cat gpurun_out/f_ab.txt; tail -8 gpurun_out/f_pytest.txt; tail -3 gpurun_out/f_initcheck.txt
