#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for v in "" _a _b _c _d; do
  lib=/root/repo/pota_b200/liblentil_b200$v.so
  echo "== variant '$v'" >> gpurun_out/tune.log
  LB_LIBRARY=$lib timeout 300 python bench.py --skip-e2e --skip-cpu --skip-thinlens --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('K1 rays/s %.4g  K2 splats/s %.4g  (%.1f ms/step; splat %s)' % (d['value'], d['splat']['value'], d['ms_per_step'], d['splat'].get('ms_per_step')))
" >> gpurun_out/tune.log 2>&1
done
cat gpurun_out/tune.log
