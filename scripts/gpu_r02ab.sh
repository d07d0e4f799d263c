#!/bin/bash
# round 2, GPU call AB: ncu --set full of the thin-lens splat kernel on the cryptomatte frame (RGBA + 3 ranked cryptomatte AOVs,
# 4 depth sub-samples): the L2 reduction / atomic / load counters behind "the table updates are L2-operation-bound"
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# launches of k_filter_splat_thinlens with --skip-thinlens --skip-splat --steps 1: rgba_only warm-up + step, then crypto warm-up + step
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_filter_splat_thinlens -s 3 -c 1 -o gpurun_out/r02ab_crypto_splat -f \
  python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-splat --skip-thinlens > gpurun_out/ab_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_filter_splat_thinlens -s 1 -c 1 -o gpurun_out/r02ab_rgba_splat -f \
  python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-splat --skip-thinlens > gpurun_out/ab_ncu2.log 2>&1
tail -3 gpurun_out/ab_ncu.log; ls -la gpurun_out/*.ncu-rep
