mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --frame-scale 0.25 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_create_rays -s 1 -c 1 -f -o gpurun_out/prof_k1 python bench.py --frame-scale 0.0625 --steps 1 --warmup 1 --skip-e2e --skip-splat --skip-cpu > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter_splat -s 1 -c 1 -f -o gpurun_out/prof_k2 python bench.py --frame-scale 0.0625 --steps 1 --warmup 1 --skip-e2e --skip-cpu > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out/
