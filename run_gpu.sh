#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_thinlens_gpu.py tests/test_crypto_gpu.py -q -m gpu --tb=short 2>&1 | tail -8 > gpurun_out/tl.log
timeout 600 python bench.py --skip-e2e --skip-cpu --skip-splat --skip-crypto --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('thin rays/s %.4g (hbm frac %.3f)  thin splats/s %.4g' % (d['thinlens']['rays']['value'], d['thinlens']['rays']['roofline']['frac'], d['thinlens']['splat']['value']))
" >> gpurun_out/tl.log 2>&1
cat gpurun_out/tl.log
