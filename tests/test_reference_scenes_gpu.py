"""GPU side of tests/test_reference_scenes.py: the `lentil_camera` parameter sets of the reference's own test scenes (fixture
tests/golden/reference_scene_cameras.json), as exported (ThinLens) and under PolynomialOptics, through the C ABI against the CPU
oracle -- setup solvers exact, camera rays within the tolerance of tests/test_camera_gpu.py, and a highlight frame placed in
WORLD space under the scene's camera matrix within the image bounds of tests/test_filter_gpu.py / test_thinlens_gpu.py."""
import numpy as np
import pytest

from pota_b200 import abi, workloads
from tests.test_reference_scenes import SCENES, scene_params

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("camera_type", [abi.LB_CAMERA_THINLENS, abi.LB_CAMERA_POLYNOMIAL_OPTICS])
@pytest.mark.parametrize("name", ["po_bidir_debug/po_bidir_debug.ass", "tl_redistribution_bug/redistribution_bug.ass",
                                  "tl_redistribution_bug/redistribution_bug_merged_aov_02.ass"])
def test_scene_camera_gpu_vs_oracle(name, camera_type):
    from oracle import orc
    from pota_b200.camera import Camera
    from tests.test_camera_gpu import _check_parity
    from tests.test_filter_gpu import psnr, rel_l1

    sc = SCENES[name]
    p = scene_params(name, camera_type)
    o, g = orc.OracleCamera(p), Camera(p, None, device=0)
    so, sg = o.state, g.state
    assert sg.aperture_radius == so.aperture_radius and sg.tan_fov == so.tan_fov and sg.sensor_shift == so.sensor_shift
    po = camera_type == abi.LB_CAMERA_POLYNOMIAL_OPTICS
    W = 160
    H = W * sc["options"]["yres"] // sc["options"]["xres"]
    if po:  # camera_create_ray
        ins = workloads.camera_samples(W, H, 4, "cpu", 0, W * H * 4, "linear")
        keys = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")
        ref = o.create_rays(*[ins[k].numpy() for k in keys], nthreads=8)
        got = {k: v.cpu().numpy() for k, v in g.create_rays(*[ins[k].cuda() for k in keys]).items()}
        _check_parity(ref, got, min_ok_frac=0.999, min_live=0.05)
    aa = sc["options"]["AA_samples"]
    spp = aa * aa
    fd = float(sc["params"]["focus_dist"])
    fr = workloads.highlight_frame(W, H, spp, so.tan_fov, "cpu", z_plane=fd * 2.0, pitch=fd * 0.16, radius=fd * 0.004)
    c2w = np.asarray(sc["camera_to_world"], np.float64)
    w2c = np.linalg.inv(c2w).astype(np.float32)
    pos = fr["pos_cs"].numpy().copy()
    hit = pos[:, 3] < 1e29
    pw = np.concatenate([pos[:, :3].astype(np.float64), np.ones((pos.shape[0], 1))], axis=1) @ c2w
    pos[hit, :3] = pw[hit, :3].astype(np.float32)
    aovs = [("RGBA", 0, 1)]
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), pos, 1.0 / spp, world_to_camera=w2c, nthreads=8)
    g.filter_begin(W, H, aovs)
    g.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), torch.from_numpy(pos).cuda(), 1.0 / spp, world_to_camera=w2c)
    torch.cuda.synchronize()
    st, sgt = o.filter_stats(), g.filter_stats()
    for k in ("samples", "redistributed", "passthrough"):
        assert st[k] == sgt[k], (k, st, sgt)
    assert st["redistributed"] > 20
    assert abs(st["splats"] - sgt["splats"]) <= 2e-3 * st["splats"] + 2, (st, sgt)
    bo, wo = o.buffers(0)
    bg, wg = g.buffers(0)
    assert rel_l1(bg, bo) <= 3e-3 and rel_l1(wg, wo) <= 3e-3, (rel_l1(bg, bo), rel_l1(wg, wo))
    assert psnr(g.resolve(0).cpu().numpy(), o.resolve(0)) >= 50.0
