mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "filter or golden" > gpurun_out/test.log 2>&1; tail -4 gpurun_out/test.log
python bench.py --skip-e2e --skip-cpu --steps 3 --warmup 1 > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k2.json')); print('rays/s %.3e'%d['value']); print(json.dumps(d['splat']))"; tail -3 gpurun_out/bench_k2.err
ncu --set full --clock-control none --import-source on -k regex:k_filter_splat -s 1 -c 1 -f -o gpurun_out/prof_k2c python bench.py --frame-scale 0.125 --steps 1 --warmup 1 --skip-e2e --skip-cpu > gpurun_out/ncu_k2c.log 2>&1
