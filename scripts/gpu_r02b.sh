#!/bin/bash
# round 2, GPU call B: whole GPU suite, bench line, kernel generations A/B, ncu captures (K1, K2, thin-lens splat)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/b_pytest.txt
( timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/b_smoke.txt
{
python scripts/ab_kernels.py --tag gen2_w550
LB_KERNEL_GEN=1 python scripts/ab_kernels.py --tag gen1_table
LB_KERNEL_GEN=0 python scripts/ab_kernels.py --tag gen0_v1 --skip-k1
python scripts/ab_kernels.py --tag gen2_lens43 --lens 43 --spp 4
python scripts/ab_kernels.py --tag gen2_lens0 --lens 0 --spp 8
python scripts/ab_kernels.py --tag thin --thin --spp 16
} 2>&1 | grep -E "^AB|Error|error" > gpurun_out/b_ab.txt
( timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/b_bench.err | tail -1 ) > gpurun_out/b_bench.json
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/b_bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_create_rays -s 1 -c 1 -o gpurun_out/r02_k1 -f python scripts/ab_kernels.py --skip-k2 --spp 4 > gpurun_out/b_ncu_k1.log 2>&1
python scripts/ncusum.py gpurun_out/r02_k1.ncu-rep --traffic k_create_rays --units 33177600 --out gpurun_out/kernel_traffic.json > gpurun_out/b_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat -c 1 -o gpurun_out/r02_k2 -f python scripts/ab_kernels.py --skip-k1 > gpurun_out/b_ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat_thinlens -c 1 -o gpurun_out/r02_thin_splat -f python scripts/ab_kernels.py --skip-k1 --thin > gpurun_out/b_ncu_thin.log 2>&1
tail -5 gpurun_out/b_pytest.txt; cat gpurun_out/b_smoke.txt; cat gpurun_out/b_ab.txt; cut -c1-600 gpurun_out/b_bench.json
