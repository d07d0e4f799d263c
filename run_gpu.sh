mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/test.log 2>&1; tail -4 gpurun_out/test.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
