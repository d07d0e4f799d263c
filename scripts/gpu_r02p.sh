#!/bin/bash
# round 2, GPU call P: whole GPU suite; ncu captures of the final K1 / K2 / thin-lens splat kernels; launch list of the bench command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/p_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_create_rays -s 1 -c 1 -o gpurun_out/r02p_k1 -f python scripts/ab_kernels.py --skip-k2 --spp 4 > gpurun_out/p_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat -c 1 -o gpurun_out/r02p_k2 -f python scripts/ab_kernels.py --skip-k1 > gpurun_out/p_ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat_thinlens -c 1 -o gpurun_out/r02p_thin_splat -f python scripts/ab_kernels.py --skip-k1 --thin > gpurun_out/p_ncu_thin.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/p_launches_bench.log 2>&1
tail -5 gpurun_out/p_pytest.txt; ls -la gpurun_out/r02p_*; wc -l gpurun_out/p_launches.csv
