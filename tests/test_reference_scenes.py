"""The parameter sets of the reference's own test scenes: every `lentil_camera` node that /root/reference/tests/*/*.ass carry under
today's node name (fixture tests/golden/reference_scene_cameras.json, extracted by tests/golden/make_scene_cameras.py: the
parameters the scene sets on top of the defaults of lentil_camera.cpp:19-52, its camera-to-world matrix, resolution and AA) is run
through the compiled reference (oracle/_ref/libref.so) and the oracle -- setup, camera_create_ray, and filter_pixel +
driver_process_bucket of a highlight frame placed in WORLD space under the scene's camera matrix -- as exported (the default
camera_type, ThinLens) and with the same numbers under PolynomialOptics (po_bidir_debug is a polynomial-optics scene by name).
Both sides are double-precision CPU code: bit-identical, as in test_oracle_vs_ref.py.  No GPU needed."""
import json
import os

import numpy as np
import pytest

from oracle import orc, ref
from pota_b200 import abi, workloads

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so not built")

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_scene_cameras.json")) as _f:
    SCENES = json.load(_f)
INT_PARAMS = {"camera_type", "bidir_sample_mult", "units", "enable_dof", "aperture_blades_lentil", "lens_model", "abb_chromatic_type",
              "bokeh_enable_image", "vignetting_retries", "enable_bidir_transmission", "enable_skydome"}


def scene_params(name, camera_type):
    kw = {}
    for k, v in SCENES[name]["params"].items():
        if k == "bokeh_image_path":
            continue  # only read when bokeh_enable_image is on (lentil.h:222-228); none of the scenes switches it on
        kw[k] = int(v) if k in INT_PARAMS else float(v)
    kw["camera_type"] = camera_type
    if camera_type == abi.LB_CAMERA_POLYNOMIAL_OPTICS:
        kw.setdefault("lens_model", 5)
    return abi.CameraParams.defaults(**kw)


def test_fixture_is_the_reference_scenes():
    assert len(SCENES) == 4 and all(len(s["camera_to_world"]) == 4 for s in SCENES.values())
    assert SCENES["po_bidir_debug/po_bidir_debug.ass"]["params"]["fstop"] == 0.5
    assert {s["options"]["AA_samples"] for s in SCENES.values()} == {3}


@pytest.mark.parametrize("camera_type", [abi.LB_CAMERA_THINLENS, abi.LB_CAMERA_POLYNOMIAL_OPTICS])
@pytest.mark.parametrize("name", sorted(SCENES))
def test_scene_camera_reference_vs_oracle(name, camera_type):
    sc = SCENES[name]
    p = scene_params(name, camera_type)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    so, sr = o.state, r.state
    assert so.aperture_radius == sr.aperture_radius and so.tan_fov == sr.tan_fov
    if camera_type == abi.LB_CAMERA_POLYNOMIAL_OPTICS:
        assert so.sensor_shift == sr.sensor_shift
    # camera_create_ray over the scene's aspect ratio
    W = 160
    H = W * sc["options"]["yres"] // sc["options"]["xres"]
    n = W * H
    ins = workloads.camera_samples(W, H, 1, "cpu", 0, n, "linear")
    arrs = [ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]
    a, b = o.create_rays(*arrs), r.create_rays(*arrs)
    first = a["tries"] == 0  # retries draw from the reference's process-global xor128: the one stated deviation
    assert first.mean() > 0.2
    for k in orc.RAY_OUT_FIELDS:
        np.testing.assert_array_equal(a[k][:, first], b[k][:, first], err_msg=k)
    # filter_pixel + driver_process_bucket at the scene's AA, the frame given in world space under the scene's camera matrix
    aa = sc["options"]["AA_samples"]
    spp = aa * aa
    fr = workloads.highlight_frame(W, H, spp, so.tan_fov, "cpu", z_plane=float(sc["params"]["focus_dist"]) * 2.0,
                                   pitch=float(sc["params"]["focus_dist"]) * 0.16, radius=float(sc["params"]["focus_dist"]) * 0.004)
    c2w = np.asarray(sc["camera_to_world"], np.float64)  # Arnold: row vectors, translation in the last row
    w2c = np.linalg.inv(c2w).astype(np.float32)
    pos = fr["pos_cs"].numpy().copy()
    hit = pos[:, 3] < 1e29
    pw = np.concatenate([pos[:, :3].astype(np.float64), np.ones((pos.shape[0], 1))], axis=1) @ c2w
    pos[hit, :3] = pw[hit, :3].astype(np.float32)
    aovs = [("RGBA", 0, 1), ("N", 1, 0)]
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), pos, 1.0 / spp)
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(*args, world_to_camera=w2c)
    r.filter_begin(W, H, aovs, spp=spp)
    r.filter_accumulate(*args, world_to_camera=w2c)
    st = o.filter_stats()
    assert st["redistributed"] > 20 and st["splats"] > 200, st
    for k in range(len(aovs)):
        bo, wo = o.buffers(k)
        br, wr = r.buffers(k)
        np.testing.assert_array_equal(bo, br, err_msg=f"buffer of {aovs[k][0]}")
        np.testing.assert_array_equal(wo, wr)
        np.testing.assert_array_equal(o.resolve(k), r.resolve(k))
