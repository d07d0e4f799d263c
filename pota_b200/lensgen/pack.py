"""Build the lens pack: pota_b200/lenses/<lens_id>.json for the 44 ids of pota_h_lenses.h.

    python -m pota_b200.lensgen.pack [--only ID ...] [--jobs N]

The JSON files are committed; nothing at build/run time needs numpy or a refit.  Each file holds
the lens constants the reference reads from `lens_constants.h` (fields of
/root/reference/src/lentil.h:106-120) and the nine fitted polynomials.  Coefficients are rounded
to float32 so that the FP64 oracle and the FP32 kernels evaluate *the same* polynomial.
"""
from __future__ import annotations

import argparse
import json
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from .fit import fit_lens
from .prescriptions import LENS_DB_DIRS, LENS_IDS, lens_spec
from .raytrace import Lens

PACK_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lenses")
POLY_NAMES = ["out_x", "out_y", "out_dx", "out_dy", "out_t", "ap_x", "ap_y", "ap_dx", "ap_dy"]


def _entrance_pupil_radius(lens: Lens) -> float:
    """Largest exit height of rays leaving the on-axis infinity-focus point that clear the lens."""
    s = np.linspace(1e-4, 0.6, 4000)
    z = np.zeros_like(s)
    o = lens.trace_from(z, z, z, s, np.full_like(s, 0.55), lens.z_sensor)
    ok = o["ok"]
    bad = np.where(~ok)[0]
    last = (bad[0] - 1) if bad.size else s.size - 1
    return float(abs(o["pos"][1, max(last, 0)]))


def lens_constants(lens: Lens, spec) -> dict:
    ep = _entrance_pupil_radius(lens)
    return {
        "lens_name": spec.lens_id,
        "lens_outer_pupil_radius": float(lens.h[0]),
        "lens_inner_pupil_radius": float(lens.h[-1]),
        "lens_length": float(lens.length),
        "lens_back_focal_length": float(lens.bfl),
        "lens_effective_focal_length": float(lens.efl),
        "lens_aperture_pos": float(lens.zv[lens.stop] - lens.z_sensor),
        "lens_aperture_housing_radius": float(lens.h[lens.stop]),
        "lens_inner_pupil_curvature_radius": float(lens.R[-1]),
        "lens_outer_pupil_curvature_radius": float(lens.R[0]),
        "lens_inner_pupil_geometry": "spherical",
        "lens_outer_pupil_geometry": "spherical",
        "lens_field_of_view": float(2.0 * np.arctan(18.0 / lens.efl)),
        "lens_fstop": float(lens.efl / (2.0 * ep)),
        "lens_aperture_radius_at_fstop": float(lens.h[lens.stop]),
    }


def _f32(c: float) -> float:
    return float(np.float32(c))


def build_one(lens_id: str) -> str:
    spec = lens_spec(lens_id)
    lens = Lens(spec.surfaces, spec.focal_mm)
    polys, rms = fit_lens(lens, spec.max_degree, spec.max_terms)
    consts = lens_constants(lens, spec)
    # constants are consumed in double by the reference; keep them float32-exact too so that
    # host FP64 setup and device FP32 kernels see identical values
    for k, v in consts.items():
        if isinstance(v, float):
            consts[k] = _f32(v)
    doc = {
        "lens_id": lens_id,
        "index": spec.index,
        "db_dir": LENS_DB_DIRS[spec.index],
        "focal_mm": spec.focal_mm,
        "max_degree": spec.max_degree,
        "max_terms": spec.max_terms,
        "prescription_mm": [list(r) for r in lens.rows],
        "constants": consts,
        "fit_rms": rms,
        "polys": {n: [[_f32(c), list(e)] for c, e in polys[n] if _f32(c) != 0.0] for n in POLY_NAMES},
    }
    os.makedirs(PACK_DIR, exist_ok=True)
    path = os.path.join(PACK_DIR, lens_id + ".json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return f"{lens_id}: efl {lens.efl:.2f} bfl {lens.bfl:.2f} f/{consts['lens_fstop']:.2f} rms out_x {rms['out_x']:.3g} ap_x {rms['ap_x']:.3g}"


def load_pack(pack_dir: str = PACK_DIR) -> list[dict]:
    """All lenses in LensModel enum order."""
    out = []
    for lens_id in LENS_IDS:
        with open(os.path.join(pack_dir, lens_id + ".json")) as f:
            out.append(json.load(f))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    ids = a.only or LENS_IDS
    with ProcessPoolExecutor(a.jobs) as ex:
        for line in ex.map(build_one, ids):
            print(line, flush=True)


if __name__ == "__main__":
    main()
