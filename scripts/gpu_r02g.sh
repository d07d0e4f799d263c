#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
python scripts/ab_kernels.py --tag horner --skip-k1
for v in hk2b6 hk2b4 hk1b5 dag; do LB_LIBRARY=$PWD/variants/lib_$v.so python scripts/ab_kernels.py --tag $v --skip-k1; done
python scripts/ab_kernels.py --tag horner_again --skip-k1
} 2>&1 | grep -E "^AB|Error|error" > gpurun_out/g_ab.txt
cat gpurun_out/g_ab.txt
