#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/k2pack.log
LB_LIBRARY=/root/repo/pota_b200/liblentil_b200_k3.so timeout 900 python -m pytest tests/test_filter_gpu.py tests/test_crypto_gpu.py -q -m gpu --tb=short -k "not every_lens and not all_lens" 2>&1 | tail -25 >> gpurun_out/k2pack.log
for v in _k2 _k3 _k4; do
  echo "== variant '$v'" >> gpurun_out/k2pack.log
  LB_LIBRARY=/root/repo/pota_b200/liblentil_b200$v.so timeout 300 python bench.py --skip-e2e --skip-cpu --skip-thinlens --skip-crypto --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('K1 rays/s %.4g  K2 splats/s %.4g (%.1f ms) splats %d attempts %d' % (d['value'], d['splat']['value'], d['splat']['ms_per_step'], d['splat']['splats_per_step'], d['splat']['attempts_per_step']))
" >> gpurun_out/k2pack.log 2>&1
done
cat gpurun_out/k2pack.log
