"""The C-ABI library loads on a CPU-only box, exports every symbol include/lentil_b200.h declares, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import subprocess
import re

import pytest

from pota_b200 import abi, camera
from pota_b200.lensgen.prescriptions import LENS_IDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lentil_b200.h")).read()
    return sorted(set(re.findall(r"LB_API\s+[\w\s\*]+?\b(lb_\w+)\s*\(", src)))


def test_header_symbols_are_exported(product_lib):
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(product_lib, name), f"{name} declared in lentil_b200.h but not exported"
    assert sorted(camera.EXPORTS) == declared, "pota_b200.camera.EXPORTS is out of sync with the header"


def test_lens_names_match_reference_enum(product_lib):
    # LensModel enum order of /root/reference/include/auto_generated_lens_includes/pota_h_lenses.h:4-47
    assert camera.lens_names() == LENS_IDS
    assert product_lib.lb_lens_count() == 44
    ref_hdr = "/root/reference/include/auto_generated_lens_includes/pota_h_lenses.h"
    if os.path.exists(ref_hdr):
        ids = [l.strip().rstrip(",") for l in open(ref_hdr) if l.strip() and not l.strip().startswith("//")]
        assert ids == LENS_IDS


def test_params_default_matches_reference_defaults(product_lib):
    p = abi.CameraParams()
    product_lib.lb_camera_params_default(C.byref(p))
    d = abi.CameraParams.defaults()
    assert bytes(p) == bytes(d)
    # lentil_camera.cpp:19-52
    assert (p.camera_type, p.bidir_sample_mult, p.units, p.sensor_width, p.enable_dof, p.fstop, p.focus_dist) == (0, 5, 1, 36.0, 1, 0.0, 150.0)
    assert (p.vignetting_retries, p.wavelength, p.focal_length_lentil, p.lens_model) == (15, 550.0, 35.0, LENS_IDS.index("cooke__speed_panchro__1920__40mm"))


def test_struct_sizes(tmp_path):
    """ctypes mirrors vs the C compiler's view of include/lentil_b200.h (sizes and the offset of each last field)."""
    assert C.sizeof(abi.CameraParams) == 28 * 4
    assert C.sizeof(abi.RayIn) == 6 * 8 and C.sizeof(abi.RayOut) == 8 * 8
    pairs = [("lb_camera_params", abi.CameraParams), ("lb_bokeh_image", abi.BokehImage), ("lb_camera_state", abi.CameraState),
             ("lb_ray_in", abi.RayIn), ("lb_ray_out", abi.RayOut), ("lb_aov_desc", abi.AovDesc), ("lb_frame_desc", abi.FrameDesc),
             ("lb_samples", abi.Samples), ("lb_filter_stats", abi.FilterStats), ("lb_lens_work", abi.LensWork)]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "lentil_b200.h"\nint main(void){\n'
    for cname, ct in pairs:
        last = ct._fields_[-1][0]
        src += f'  printf("%zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));\n'
    src += "  return 0;\n}\n"
    (tmp_path / "sz.c").write_text(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(tmp_path / "sz.c"), "-o", str(tmp_path / "sz")], check=True)
    out = subprocess.run([str(tmp_path / "sz")], check=True, capture_output=True, text=True).stdout.split("\n")
    for (cname, ct), line in zip(pairs, out):
        size, off = (int(v) for v in line.split())
        assert C.sizeof(ct) == size, cname
        assert getattr(ct, ct._fields_[-1][0]).offset == off, cname


def test_no_cpu_fallback(product_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(camera.LentilError, match="no CUDA device"):
        camera.Camera(abi.CameraParams.defaults(camera_type=1))
    peak = C.c_double()
    assert product_lib.lb_bench_fp32_peak(0, C.byref(peak)) == abi.LB_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pota_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text and "libref" not in text, os.path.join(dirpath, f)
