mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/test.log 2>&1; tail -5 gpurun_out/test.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print('rays/s %.3e e2e %.3e clocks %s cpu %s'%(d['value'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'])); print(json.dumps(d['splat']['roofline'])); print(d['splat']['value'])"; tail -3 gpurun_out/bench_full.err
