"""Config C4 (BASELINE.json configs[3]): every lens of the pack x f-stop x focus-distance grid, camera rays/s
per lens against its polynomial size.  Run on a GPU box; writes a table to stdout (kept in profiles/).

    python scripts/sweep_lenses.py [--rays 8294400] [--json gpurun_out/c4_sweep.json]

The JSON file holds every cell (rays/s, mean tries per ray, dead fraction): a vignetted main ray is traced again with a new lens
sample up to vignetting_retries times (lentil.h:296-316), so a cell's rays/s falls with its mean tries -- the f/1.4 cells.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pota_b200 import abi, workloads  # noqa: E402
from pota_b200.camera import RAY_OUT_FIELDS, Camera, lens_names  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=3840 * 2160)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    record = {"config": "C4: every lens of the pack x fstop {1.4, 2.8, 5.6, 11} x focus {50, 150, 500, 1e9} cm", "rays_per_cell": a.rays, "lenses": []}
    n = a.rays
    dev = torch.device("cuda", 0)
    ins = workloads.camera_samples(3840, 2160, 1, dev, 0, n, "pixel")
    out = {k: torch.empty((3, n), dtype=torch.float32, device=dev) for k in RAY_OUT_FIELDS}
    out["tries"] = torch.empty(n, dtype=torch.int32, device=dev)
    names = lens_names()
    print(f"# C4 sweep: {n} rays per cell (3840x2160x1spp), unrolled kernels, device-resident; rays/s = median over the 16 cells")
    print("# lens | terms(eval,ap,ap_jac) | F_eval F_ap | setup ms | Grays/s min / median / max over fstop{1.4,2.8,5.6,11} x focus{50,150,500,1e9}cm | dead%")
    for k, name in enumerate(names):
        focal = float(name.split("__")[-1].replace("mm", ""))
        sensor = min(36.0, 0.7 * focal)
        rates, dead, setup, cells = [], [], [], []
        for fstop in (1.4, 2.8, 5.6, 11.0):
            for focus in (50.0, 150.0, 500.0, 1.0e9):
                p = abi.CameraParams.defaults(camera_type=1, lens_model=k, fstop=fstop, focus_dist=focus, sensor_width=sensor)
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                cam = Camera(p, device=0)
                e1.record()
                args = [ins[key] for key in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]
                cam.create_rays(*args, out=out)
                e1.record()
                cam.create_rays(*args, out=out)
                e2.record()
                torch.cuda.synchronize()
                rates.append(n / (e1.elapsed_time(e2) * 1e-3) / 1e9)
                setup.append(e0.elapsed_time(e1))
                dead.append(float((out["weight"][0] == 0).float().mean()))
                cells.append({"fstop": fstop, "focus_cm": focus, "grays_per_s": round(rates[-1], 3), "mean_tries": round(float(out["tries"].float().mean()) + 1.0, 3),
                              "dead_frac": round(dead[-1], 4), "setup_ms": round(setup[-1], 1)})
                w = cam.lens_work
                cam.close()
        record["lenses"].append({"lens": k, "name": name, "terms_eval": w.terms_eval, "terms_ap": w.terms_ap, "terms_ap_jac": w.terms_ap_jac,
                                 "F_eval": w.F_eval, "F_ap": w.F_ap, "grays_per_s_median": round(sorted(rates)[8], 3), "cells": cells})
        rates.sort()
        print(f"{k:2d} {name:46s} | {w.terms_eval:3d} {w.terms_ap:3d} {w.terms_ap_jac:3d} | {w.F_eval:5.0f} {w.F_ap:5.0f} | {sorted(setup)[8]:6.1f} | "
              f"{rates[0]:5.2f} / {rates[8]:5.2f} / {rates[-1]:5.2f} | {100 * sum(dead) / len(dead):4.1f}", flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(record, f)


if __name__ == "__main__":
    main()
