mkdir -p gpurun_out
python bench.py --skip-e2e --skip-cpu --steps 3 --warmup 1 > gpurun_out/bench_tl.json 2> gpurun_out/bench_tl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tl.json')); print(json.dumps(d['thinlens']))"; tail -3 gpurun_out/bench_tl.err
