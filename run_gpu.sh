mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "every_lens" > gpurun_out/test.log 2>&1; tail -4 gpurun_out/test.log
python scripts/sweep_lenses.py > gpurun_out/sweep.txt 2>&1; tail -48 gpurun_out/sweep.txt
python scripts/run_c5.py --spp 4 --steps 1 2>&1 | tail -3
