#!/bin/bash
# round 2, GPU call AA: A/B of resident-CTA bounds with the final bodies -- 550 nm splat kernel at 5 (default) / 6 / 7 CTAs per SM
# (96 / 80 / 72 registers, no spill up to 6), K1 at 4 (default) / 5, service batch 8 / 12
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/aa_ab.txt
for v in default v6 v7 v6s default; do
  if [ "$v" = default ]; then unset LB_LIBRARY; else export LB_LIBRARY=$PWD/variants/liblentil_$v.so; fi
  [ "$v" = default ] || [ -f "$LB_LIBRARY" ] || continue
  skip=""; [ "$v" = v7 ] && skip="--skip-k1"; [ "$v" = v6s ] && skip="--skip-k1"; [ -s gpurun_out/aa_ab.txt ] && [ "$v" = default ] && skip="--skip-k1"
  timeout 120 python scripts/ab_kernels.py --tag $v $skip 2>&1 | grep "^AB" >> gpurun_out/aa_ab.txt
done
python - <<'PY'
import json
for ln in open("gpurun_out/aa_ab.txt"):
    d = json.loads(ln[3:])
    print(d["tag"], "K1 %.4g rays/s" % d.get("k1_rays_per_s", 0), "K2 %.2f ms" % d.get("k2_ms", 0), d.get("k2_reps_ms"), d.get("k2_splats"), d.get("k2_attempts"), d.get("k1_checksum"), d.get("k2_energy"))
PY
