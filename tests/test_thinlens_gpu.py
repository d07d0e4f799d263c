"""GPU parity of the thin-lens camera model against the CPU oracle, through the C ABI.

Reference: Camera::trace_ray_fw_thinlens (lentil.h:431-569) and the ThinLens case of filter_pixel
(lentil_filter.cpp:303-447).  The model is a few float operations per ray; the kernels are compiled with
-fmad=false and reproduce the host arithmetic operation by operation, so results are expected to be
bit-identical except where device libm differs from glibc (powf of the vignetting/bias terms, sin/cos of the
coma rotation): tolerances are 1e-6 relative on rays and a 1e-3 relative-L1 bound on framebuffers.
"""
import numpy as np
import pytest

from pota_b200 import abi, workloads

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TL_CASES = [
    dict(),
    dict(fstop=1.4, focus_dist=80.0, focal_length_lentil=50.0),
    dict(abb_coma=0.6, abb_distortion=0.3, optical_vignetting=2.0, bokeh_circle_to_square=0.5, bokeh_anamorphic=0.3),
    dict(abb_spherical=0.3, aperture_blades_lentil=6),
    dict(bokeh_enable_image=1, units=abi.LB_UNITS_M, abb_coma=0.2),
    dict(abb_chromatic=0.5, abb_chromatic_type=1),
]
IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")


def _params(**kw):
    base = dict(camera_type=abi.LB_CAMERA_THINLENS, fstop=2.8, focus_dist=150.0)
    base.update(kw)
    return abi.CameraParams.defaults(**base)


@pytest.mark.parametrize("kw", TL_CASES)
def test_thinlens_rays(kw):
    from oracle import orc
    from pota_b200.camera import RAY_OUT_FIELDS, Camera

    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = _params(**kw)
    o, g = orc.OracleCamera(p, img), Camera(p, img, device=0)
    assert o.state.aperture_radius == g.state.aperture_radius and o.state.tan_fov == g.state.tan_fov
    n = 100_000
    w = int(round((n * 16 / 9) ** 0.5))
    ins = workloads.camera_samples(w, -(-n // w), 1, "cpu", 0, n, "linear")
    ref = o.create_rays(*[ins[k].numpy() for k in IN_KEYS], nthreads=8)
    out = g.create_rays(*[ins[k].cuda() for k in IN_KEYS])
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    same = ref["tries"] == got["tries"]
    assert same.mean() >= 0.9995, (~same).sum()
    np.testing.assert_array_equal(ref["weight"][:, same], got["weight"][:, same])
    for k in ("origin", "dir"):
        np.testing.assert_allclose(got[k][:, same], ref[k][:, same], rtol=1e-6, atol=1e-7, err_msg=k)
        assert (got[k][:, same] == ref[k][:, same]).mean() > 0.99, k  # bit-identical for almost every ray
    for k in ("dOdx", "dOdy", "dDdx", "dDdy"):
        assert (got[k][:, same] == ref[k][:, same]).mean() > 0.98, k
    # host path == device path
    h_in = {k: ins[k].pin_memory() for k in IN_KEYS}
    h_out = {k: torch.empty((3, n), dtype=torch.float32).pin_memory() for k in RAY_OUT_FIELDS}
    g.create_rays_host(*[h_in[k] for k in IN_KEYS], out=h_out)
    np.testing.assert_array_equal(h_out["dir"].numpy(), got["dir"])


@pytest.mark.parametrize("kw", TL_CASES)
def test_thinlens_redistribution(kw):
    from oracle import orc
    from pota_b200.camera import Camera

    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    kw = dict(kw)
    kw.setdefault("fstop", 1.4)
    kw.setdefault("focus_dist", 35.0)
    p = _params(bidir_sample_mult=8, **kw)
    o, g = orc.OracleCamera(p, img), Camera(p, img, device=0)
    W, H, spp = 240, 135, 4
    z = 0.75 if kw.get("units") == abi.LB_UNITS_M else 75.0
    fr = workloads.highlight_frame(W, H, spp, o.state.tan_fov, "cpu", z_plane=z, pitch=z * 0.072, radius=z * 0.0018, n_extra_aov=1)
    aovs = [("RGBA", 0, 1), ("light0", 0, 0), ("N", 1, 0)]
    vn = [None, fr["aov_values"][0].numpy(), fr["aov_values"][0].numpy()]
    vg = [None, fr["aov_values"][0].cuda(), fr["aov_values"][0].cuda()]
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp, aov_values=vn)
    g.filter_begin(W, H, aovs)
    g.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), fr["pos_cs"].cuda(), 1.0 / spp, aov_values=vg)
    so, sg = o.filter_stats(), g.filter_stats()
    for k in ("samples", "redistributed", "passthrough"):
        assert so[k] == sg[k], (k, so, sg)
    assert so["redistributed"] > 100
    for k in ("splats", "attempts"):
        assert abs(so[k] - sg[k]) <= 1e-3 * so[k] + 2, (k, so, sg)
    for a in (0, 1):
        bo, wo = o.buffers(a)
        bg, wg = g.buffers(a)
        assert np.abs(bg - bo).sum() / np.abs(bo).sum() <= 1e-3
        np.testing.assert_allclose(wg.sum(dtype=np.float64), wo.sum(dtype=np.float64), rtol=1e-4)
    ro, rg = o.resolve(2), g.resolve(2).cpu().numpy()
    assert (np.abs(ro - rg).max(axis=2) > 1e-6).mean() <= 2e-3


@pytest.mark.parametrize("size", [64, 250, 301])
def test_bokeh_guided_search_equals_full_search(size, monkeypatch):
    """bokehSample's two upper_bound searches (imagebokeh.h:341-412) through the guide tables return the index the full
    binary search returns: rays are compared bit for bit between a camera built with and one built without the tables
    (LB_NO_CDF_GUIDE=1), lens samples inside [0, 1), on its edges, outside it and NaN."""
    from pota_b200.camera import Camera

    rng = np.random.default_rng(size)
    img = workloads.disc_bokeh_image(size) * rng.uniform(0.0, 1.0, (size, size, 1)).astype(np.float32)
    img[rng.uniform(size=(size, size)) < 0.3] = 0.0  # empty pixels and whole empty rows: runs of equal CDF entries
    img[5] = 0.0
    p = _params(bokeh_enable_image=1, fstop=1.4)
    n = 200_000
    w = int(round((n * 16 / 9) ** 0.5))
    ins = workloads.camera_samples(w, -(-n // w), 1, "cpu", 0, n, "linear")
    special = torch.tensor([0.0, 1.0, -0.25, 1.5, float("nan"), 1.0 - 2.0**-24, 2.0**-30, 0.5], dtype=torch.float32)
    ins["lensx"][: special.numel()] = special
    ins["lensy"][100 : 100 + special.numel()] = special
    ins["lensx"][200:456] = torch.arange(256, dtype=torch.float32) / 256.0  # bucket edges
    ins["lensy"][200:456] = torch.arange(256, dtype=torch.float32).flip(0) / 256.0
    outs = []
    for no_guide in ("0", "1"):
        monkeypatch.setenv("LB_NO_CDF_GUIDE", no_guide)
        g = Camera(p, img, device=0)
        out = g.create_rays(*[ins[k].cuda() for k in IN_KEYS])
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in out.items()})
    for k in outs[0]:
        np.testing.assert_array_equal(outs[0][k].view(np.uint32), outs[1][k].view(np.uint32), err_msg=k)
    assert np.unique(outs[0]["origin"][0]).size > size // 2  # the lens samples did spread over the kernel (one value per image column)


@pytest.mark.parametrize("kw", [dict(bokeh_enable_image=1), dict(abb_chromatic=0.5, abb_chromatic_type=1), dict(optical_vignetting=2.0, abb_coma=0.3)])
def test_thinlens_tile_accumulate(kw, monkeypatch):
    """Shared-memory tile accumulation (Camera::add_to_buffer lentil.h:823-851 behind a per-CTA window, thinlens_kernels.cu):
    the tile kernel (LB_SPLAT_TILE=1; opt-in, the L2 reductions of the direct kernel measured faster on B200) and the direct
    kernel produce the same splats -- identical counts, framebuffers equal up to float summation order -- both within the
    parity bound of the oracle; with LB_SPLAT_TILE=auto the device picks the tile kernel when the bokeh discs fit the window
    and the direct kernel when they do not."""
    from oracle import orc
    from pota_b200.camera import Camera

    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = _params(bidir_sample_mult=8, fstop=1.4, focus_dist=35.0, **kw)
    o = orc.OracleCamera(p, img)
    W, H, spp = 240, 135, 4
    # one light a few pixels wide: the work items of a batch are neighbours (with several lights on a row, a row-major batch
    # holds pixels of all of them and only the first light's discs fall into the window)
    fr = workloads.highlight_frame(W, H, spp, o.state.tan_fov, "cpu", z_plane=75.0, pitch=75.0 * 0.3, radius=75.0 * 0.008, grid=(1, 1), n_extra_aov=1)
    aovs = [("RGBA", 0, 1), ("light0", 0, 0), ("N", 1, 0)]
    vn = [None, fr["aov_values"][0].numpy(), fr["aov_values"][0].numpy()]
    vg = [None, fr["aov_values"][0].cuda(), fr["aov_values"][0].cuda()]
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp, aov_values=vn)
    so = o.filter_stats()
    res = {}
    for mode in ("0", "1", "auto"):
        monkeypatch.setenv("LB_SPLAT_TILE", mode)
        g = Camera(p, img, device=0)
        g.filter_begin(W, H, aovs)
        g.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), fr["pos_cs"].cuda(), 1.0 / spp, aov_values=vg)
        res[mode] = (g.filter_stats(), [g.buffers(a) for a in (0, 1)], g.resolve(2).cpu().numpy())
    s0, s1, sa = res["0"][0], res["1"][0], res["auto"][0]
    assert s0["tile_splats"] == 0
    assert s1["tile_splats"] > 0.9 * s1["splats"], s1       # the discs (18 px here) land inside the 64-pixel window
    assert sa["tile_splats"] == s1["tile_splats"], (sa, s1)  # and the device picked the tile kernel by itself
    for k in ("samples", "redistributed", "passthrough", "splats", "attempts"):
        assert s0[k] == s1[k] == sa[k], (k, s0, s1)
    assert abs(so["splats"] - s1["splats"]) <= 1e-3 * so["splats"] + 2
    for a in (0, 1):
        bo, wo = o.buffers(a)
        (b0, w0), (b1, w1) = res["0"][1][a], res["1"][1][a]
        assert np.abs(b1 - b0).sum() / np.abs(b0).sum() <= 1e-5   # same splats, other summation order
        assert np.abs(w1 - w0).sum() / np.abs(w0).sum() <= 1e-5
        assert np.abs(b1 - bo).sum() / np.abs(bo).sum() <= 1e-3   # the oracle
        np.testing.assert_allclose(w1.sum(dtype=np.float64), wo.sum(dtype=np.float64), rtol=1e-4)
    np.testing.assert_array_equal(res["0"][2], res["1"][2])       # closest-filter AOV: not touched by the window
    # discs wider than the window: the device keeps the direct kernel
    monkeypatch.setenv("LB_SPLAT_TILE", "auto")
    W2, H2 = 1920, 1080
    p2 = _params(bidir_sample_mult=8, fstop=1.4, focus_dist=35.0, focal_length_lentil=50.0, **kw)  # config C3's camera: 144-pixel discs
    g = Camera(p2, img, device=0)
    fr2 = workloads.highlight_frame(W2, H2, 1, g.state.tan_fov, "cpu", grid=(2, 1))
    g.filter_begin(W2, H2, [("RGBA", 0, 1)])
    g.filter_accumulate(fr2["px"].cuda(), fr2["py"].cuda(), fr2["rgba"].cuda(), fr2["pos_cs"].cuda(), 1.0)
    st = g.filter_stats()
    assert st["splats"] > 0 and st["tile_splats"] == 0, st
