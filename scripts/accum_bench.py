"""Where should the splat accumulate live?  The accumulate alone (RGBA + filter weight per splat at pseudo-random pixels) as
global vector reductions into planes of several sizes (L2-resident ... HBM-resident) and as a shared-memory tile with float /
integer atomics (include/lentil_b200.h: lb_bench_splat_accum).  One JSON line."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pota_b200.camera import lib  # noqa: E402

L = lib()
v = ctypes.c_double()
out = {}
for mb in (41, 100, 400, 1600, 5400):
    assert L.lb_bench_splat_accum(0, 0, mb, ctypes.byref(v)) == 0
    out[f"global_red_v4_plus_f32_{mb}MB_gsplats"] = round(v.value, 2)
    assert L.lb_bench_red_peak(0, mb, ctypes.byref(v)) == 0
    out[f"global_red_v4_{mb}MB_gbs"] = round(v.value, 1)
for mode, name in ((1, "smem_tile_float_cas"), (2, "smem_tile_u32_native")):
    assert L.lb_bench_splat_accum(0, mode, 41, ctypes.byref(v)) == 0
    out[f"{name}_gsplats"] = round(v.value, 2)
print("ACCUM " + json.dumps(out), flush=True)
