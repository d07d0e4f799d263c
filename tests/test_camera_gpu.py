"""GPU parity of camera_create_ray (K1) against the CPU oracle, through the C ABI.

Tolerance (BASELINE.json north_star): per-ray outputs within 1e-4 relative, FP32 kernels vs the FP64
oracle on identical inputs and identical (float32-exact) coefficients.
"""
import numpy as np
import pytest

from pota_b200 import abi, workloads
from tests.util import po_params, rel_err_vec

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-4


def _run_both(params, n=200_000, bokeh=None):
    from oracle import orc
    from pota_b200.camera import Camera

    width = max(1, int(round((n * 16 / 9) ** 0.5)))  # a 16:9 frame of ~n pixels so the batch covers the whole sensor
    height = -(-n // width)
    ins = workloads.camera_samples(width, height, 1, "cpu", 0, n, seed_mode="linear")
    ocam = orc.OracleCamera(params, bokeh)
    gcam = Camera(params, bokeh, device=0)
    so, sg = ocam.state, gcam.state
    # the setup solvers run in FP64 on the device and must reproduce the host-double reference exactly
    assert sg.aperture_radius == so.aperture_radius
    assert sg.sensor_shift == so.sensor_shift
    assert sg.tan_fov == so.tan_fov
    assert sg.focus_check_ok == so.focus_check_ok
    np_in = {k: v.numpy() for k, v in ins.items()}
    ref = ocam.create_rays(np_in["sx"], np_in["sy"], np_in["dsx"], np_in["dsy"], np_in["lensx"], np_in["lensy"], nthreads=8)
    dev = {k: v.cuda() for k, v in ins.items()}
    out = gcam.create_rays(dev["sx"], dev["sy"], dev["dsx"], dev["dsy"], dev["lensx"], dev["lensy"])
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    return ref, got, gcam, dev


def _check_parity(ref, got, min_ok_frac=0.9999, min_live=0.5):
    n = ref["tries"].shape[0]
    assert np.array_equal(ref["weight"] == 0, got["weight"] == 0), "weight=0 (failed ray) pattern differs"
    same_tries = ref["tries"] == got["tries"]
    assert same_tries.mean() >= min_ok_frac, f"tries differ on {(~same_tries).sum()} of {n} rays"
    live = (ref["weight"][0] != 0) & same_tries
    assert live.sum() >= min_live * n, f"only {live.sum()} of {n} rays are live"
    np.testing.assert_array_equal(ref["weight"][:, live], got["weight"][:, live])
    for k in ("origin", "dir"):
        e = rel_err_vec(got[k][:, live], ref[k][:, live])
        frac = (e <= TOL).mean()
        assert frac >= min_ok_frac, f"{k}: {100 * (1 - frac):.4f}% of rays above {TOL} (max {e.max():.3e})"
        assert np.median(e) <= 1e-6
    # The differentials are finite differences of float32 vectors over a 1e-3 step (lentil_camera.cpp:84,
    # 115-118): the reference's own output is quantised to ulp(vector)/1e-3, a few % of its value.  They
    # are compared in that unit; beyond a few units are rays whose main and offset traces stopped their
    # Newton loop at different iterations (in either implementation).
    for k, base in (("dOdx", "origin"), ("dOdy", "origin"), ("dDdx", "dir"), ("dDdy", "dir")):
        unit = np.spacing(np.abs(ref[base][:, live]).max(axis=0).astype(np.float32)) / 1e-3
        d = np.abs(got[k][:, live] - ref[k][:, live]).max(axis=0) / unit
        assert np.median(d) <= 1.0, (k, np.median(d))
        assert np.quantile(d, 0.99) <= 4.0, (k, np.quantile(d, 0.99))
        assert (d > 8.0).mean() <= 5e-3, (k, (d > 8.0).mean())


@pytest.mark.parametrize("lens_model,fstop,focus", [(5, 2.8, 150.0), (0, 1.4, 50.0), (40, 5.6, 500.0), (16, 0.0, 150.0), (37, 11.0, 1.0e7)])
def test_create_rays_parity(lens_model, fstop, focus, kernel_kind):
    ref, got, gcam, _ = _run_both(po_params(lens_model=lens_model, fstop=fstop, focus_dist=focus))
    assert gcam.kernel_kind == kernel_kind
    _check_parity(ref, got)


def test_create_rays_no_dof_and_units():
    ref, got, *_ = _run_both(po_params(enable_dof=0, units=abi.LB_UNITS_M), n=50_000)
    _check_parity(ref, got, min_live=0.1)  # axis-parallel rays only clear the rear element near the sensor centre


def test_create_rays_short_lens_small_sensor():
    # 16 mm stand-in: image circle smaller than a 36 mm sensor, so a 10 mm sensor is used (many vignetting retries)
    ref, got, *_ = _run_both(po_params(lens_model=28, sensor_width=10.0), n=100_000)
    _check_parity(ref, got, min_live=0.9)
    assert (ref["tries"] > 0).sum() > 50  # the retry path (counter RNG re-draw) is exercised


def test_create_rays_blades():
    ref, got, *_ = _run_both(po_params(aperture_blades_lentil=6), n=50_000)
    _check_parity(ref, got)


def test_create_rays_bokeh_image():
    img = workloads.disc_bokeh_image(64)
    ref, got, *_ = _run_both(po_params(bokeh_enable_image=1), n=50_000, bokeh=img)
    _check_parity(ref, got)


def test_create_rays_host_path_matches_device_path():
    from pota_b200.camera import RAY_OUT_FIELDS

    params = po_params()
    ref, got, gcam, dev = _run_both(params, n=100_003)
    n = 100_003
    host_in = {k: v.cpu().pin_memory() for k, v in dev.items()}
    out = {k: torch.empty((3, n), dtype=torch.float32).pin_memory() for k in RAY_OUT_FIELDS}
    out["tries"] = torch.empty(n, dtype=torch.int32).pin_memory()
    gcam.create_rays_host(host_in["sx"], host_in["sy"], host_in["dsx"], host_in["dsy"], host_in["lensx"], host_in["lensy"], out)
    for k in list(RAY_OUT_FIELDS) + ["tries"]:
        np.testing.assert_array_equal(out[k].numpy(), got[k], err_msg=k)


def test_empty_and_ragged_batches():
    from pota_b200.camera import Camera

    cam = Camera(po_params(), device=0)
    z = torch.empty(0, device="cuda")
    out = cam.create_rays(z, z, z, z, z, z)
    assert out["origin"].shape == (3, 0)
    for n in (1, 31, 33, 129):
        ins = {k: v.cuda() for k, v in workloads.camera_samples(64, 64, 1, "cpu", 0, n, "linear").items()}
        out = cam.create_rays(ins["sx"], ins["sy"], ins["dsx"], ins["dsy"], ins["lensx"], ins["lensy"])
        torch.cuda.synchronize()
        assert torch.isfinite(out["dir"]).all()


def test_ray_results_do_not_depend_on_the_batch_split(kernel_kind):
    """The unrolled kernel traces rays j and j + ceil(n/2) of a call in one thread (packed FP32 lanes): a ray's result
    must not depend on which ray it is paired with, i.e. on how the caller cuts its samples into calls."""
    from pota_b200.camera import RAY_OUT_FIELDS, Camera

    cam = Camera(po_params(fstop=1.4, focus_dist=50.0, lens_model=0), device=0)  # f/1.4: a good share of rays retries
    assert cam.kernel_kind == kernel_kind
    n = 20_001
    ins = {k: v.cuda() for k, v in workloads.camera_samples(200, 101, 1, "cpu", 0, n, "linear").items()}
    keys = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")
    whole = cam.create_rays(*[ins[k] for k in keys])
    torch.cuda.synchronize()
    assert (whole["tries"] > 0).float().mean() > 0.05
    for cut in (1, 7_000, 10_001, 20_000):
        a = cam.create_rays(*[ins[k][:cut].contiguous() for k in keys], ray_id_base=0)
        b = cam.create_rays(*[ins[k][cut:].contiguous() for k in keys], ray_id_base=cut)
        torch.cuda.synchronize()
        for f in list(RAY_OUT_FIELDS) + ["tries"]:
            got = torch.cat([a[f], b[f]], dim=-1)
            assert torch.equal(got.view(torch.int32), whole[f].view(torch.int32)), (cut, f)


def test_reverse_rays():
    from pota_b200.camera import Camera

    cam = Camera(po_params(), device=0)
    Po = torch.tensor([[1.0, 2.0, -10.0, 0.0], [0.5, -0.25, -1e-6, 0.0]], device="cuda")
    Ps = cam.reverse_rays(Po).cpu().numpy()
    t = cam.state.tan_fov
    exp = np.array([[1.0 / (10.0 * t), 2.0 / (10.0 * t)], [0.5 / 1e-3, -0.25 / 1e-3]], np.float32)
    np.testing.assert_allclose(Ps, exp, rtol=1e-6)


@pytest.mark.parametrize("lens_model", range(44))
def test_every_lens_of_the_pack(lens_model):
    """Config C4: all 44 LensModel ids (polynomial degree 5..11, 12..64 terms), unrolled kernels, rays + a small splat frame."""
    from oracle import orc
    from pota_b200.camera import Camera

    name = Camera.__module__ and __import__("pota_b200.camera", fromlist=["lens_names"]).lens_names()[lens_model]
    focal = float(name.split("__")[-1].replace("mm", ""))
    sensor = min(36.0, 0.7 * focal)  # the short stand-ins do not cover a 36 mm sensor
    p = po_params(lens_model=lens_model, fstop=2.0, focus_dist=100.0, sensor_width=sensor, bidir_sample_mult=8)
    ref, got, gcam, _ = _run_both(p, n=20_000)
    assert gcam.kernel_kind == "unrolled"
    _check_parity(ref, got, min_ok_frac=0.999, min_live=0.8)
    # reverse path on the same lens
    ocam = orc.OracleCamera(p)
    W, H, spp = 96, 54, 4
    fr = workloads.highlight_frame(W, H, spp, ocam.state.tan_fov, "cpu", z_plane=40.0, pitch=4.0, radius=0.2)
    aovs = [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)]
    ocam.filter_begin(W, H, aovs)
    ocam.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp, nthreads=8)
    gcam.filter_begin(W, H, aovs)
    gcam.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), fr["pos_cs"].cuda(), 1.0 / spp)
    so, sg = ocam.filter_stats(), gcam.filter_stats()
    assert so["redistributed"] == sg["redistributed"] and so["redistributed"] > 0, (so, sg)
    assert abs(so["splats"] - sg["splats"]) <= 5e-3 * so["splats"] + 3, (so, sg)
    bo, _ = ocam.buffers(0)
    bg, _ = gcam.buffers(0)
    assert np.abs(bg - bo).sum() / np.abs(bo).sum() <= 1e-2


@pytest.mark.parametrize("outer", [1, 2])
def test_cylindrical_outer_pupil_gpu(outer, kernel_kind):
    from oracle import orc
    from pota_b200.camera import Camera

    p = po_params(fstop=2.0, focus_dist=80.0)
    ocam, gcam = orc.OracleCamera(p), Camera(p, device=0)
    ocam.set_pupil_geometry(outer)
    gcam.set_pupil_geometry(outer)
    n = 50_000
    ins = workloads.camera_samples(300, 170, 1, "cpu", 0, n, "linear")
    ref = ocam.create_rays(*[ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")], nthreads=8)
    out = gcam.create_rays(*[ins[k].cuda() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")])
    torch.cuda.synchronize()
    _check_parity(ref, {k: v.cpu().numpy() for k, v in out.items()}, min_ok_frac=0.999)
