"""The drop-in claim as a test result: the reference's Arnold nodes (camera, filter, imager; operator + loader in test_operator.py and
the last test here) re-implemented over
liblentil_b200.so (adaptor/lentil_b200_*.cpp, built against the stand-in SDK headers of oracle/shims/) are driven by the SAME
harness entry points (oracle/ref_harness.cpp: AtNodeMethods tables, per-sample CreateRay, per-pixel FilterPixel over an
AtAOVSampleIterator, per-bucket DriverProcessBucket over an AtOutputIterator) as the compiled reference
(oracle/_ref/libref.so), and the framebuffers they leave are compared.

Reference lines: lentil_loader.cpp:20-28 (node tables), lentil_camera.cpp:56-125, lentil_filter.cpp:66-301,
lentil_imager.cpp:66-193."""
import numpy as np
import pytest

from pota_b200 import abi, workloads
from tests.util import branch_frame, po_params, rel_err_vec

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")


def rel_l1(a, b):
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))


def psnr(a, b):
    peak = float(np.abs(b).max())
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r

    if not (r.available() and r.adaptor_available()):
        pytest.skip("oracle/_ref/libref.so or adaptor/_build/libadaptor.so not built")
    return r


@pytest.mark.parametrize("kw", [dict(), dict(aperture_blades_lentil=6), dict(bokeh_enable_image=1), dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0)])
def test_camera_node_create_ray(ref, kw, monkeypatch):
    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = po_params(**kw)
    r, a = ref.RefCamera(p, img), ref.AdaptorCamera(p, img)
    sr, sa = r.state, a.state
    assert sa.aperture_radius == sr.aperture_radius and sa.tan_fov == sr.tan_fov
    if p.camera_type == abi.LB_CAMERA_POLYNOMIAL_OPTICS:  # the thin-lens setup does not touch Camera::sensor_shift (lentil.h:1663-1668)
        assert sa.sensor_shift == sr.sensor_shift
    n = 20000
    w = int(round((n * 16 / 9) ** 0.5))
    ins = workloads.camera_samples(w, -(-n // w), 1, "cpu", 0, n, "linear")
    arrs = [ins[k].numpy() for k in IN_KEYS]
    want, got = r.create_rays(*arrs), a.create_rays(*arrs, nthreads=4)
    # rays whose first try succeeds (the reference's retries draw their lens samples from its process-global xor128, the one
    # stated deviation; `tries` is not observable through the reference's interface, the oracle reports it)
    from oracle import orc

    first = orc.OracleCamera(p, img).create_rays(*arrs, nthreads=4)["tries"] == 0
    assert first.mean() > 0.2
    np.testing.assert_array_equal(want["weight"][:, first], got["weight"][:, first])
    live = first & (want["weight"][0] != 0)
    for k in ("origin", "dir"):
        e = rel_err_vec(got[k][:, live], want[k][:, live])
        assert (e <= 1e-4).mean() >= 0.9999, (k, float(e.max()))
    # without the per-bucket prefetch every CreateRay is a one-ray GPU call: same answers
    monkeypatch.setenv("LB_ADAPTOR_NO_PREFETCH", "1")
    slow = a.create_rays(*[x[:200] for x in arrs])
    for k in ("origin", "dir", "dOdx", "weight"):
        np.testing.assert_array_equal(slow[k], got[k][:, :200], err_msg=k)


@pytest.mark.parametrize("kw,aovs,n_extra", [
    (dict(), [("RGBA", 0, 1)], 0),
    (dict(bokeh_enable_image=1, bidir_add_energy=1.0), [("RGBA", 0, 1), ("light0", 0, 0), ("light1", 0, 0)], 2),
    (dict(), [("RGBA", 0, 1), ("N", 1, 0), ("lentil_debug", 1, 2)], 1),
    (dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0), [("RGBA", 0, 1), ("light0", 0, 0)], 1),
])
def test_filter_and_imager_nodes(ref, kw, aovs, n_extra):
    """filter_pixel per pixel + driver_process_bucket per 16x16 bucket through both plugins: same framebuffers."""
    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **kw)
    r, a = ref.RefCamera(p, img), ref.AdaptorCamera(p, img)
    W, H, spp = 128, 72, 9
    fr = workloads.highlight_frame(W, H, spp, r.state.tan_fov, "cpu", n_extra_aov=n_extra)
    vals = [None] + [v.numpy() for v in fr["aov_values"]] + [None] * (len(aovs) - 1 - n_extra)
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    for cam, threads in ((r, 1), (a, 4)):
        cam.filter_begin(W, H, aovs, spp=spp)
        cam.filter_accumulate(*args, aov_values=vals, nthreads=threads)
    for i, (name, flt, role) in enumerate(aovs):
        want = np.zeros((H, W, 4), np.float32)
        got = np.zeros((H, W, 4), np.float32)
        for y0 in range(0, H, 16):  # the imager is called per bucket
            for x0 in range(0, W, 16):
                w, h = min(16, W - x0), min(16, H - y0)
                want[y0:y0 + h, x0:x0 + w] = r.resolve(i, x0, y0, w, h)
                got[y0:y0 + h, x0:x0 + w] = a.resolve(i, x0, y0, w, h)
        if flt == 0:
            assert rel_l1(got, want) <= 3e-3 and psnr(got, want) >= 50.0, (name, rel_l1(got, want), psnr(got, want))
            bw, ww = r.buffers(i)
            bg, wg = a.buffers(i)
            assert rel_l1(bg, bw) <= 2e-3 and rel_l1(wg, ww) <= 2e-3, name
        else:  # closest filter: per pixel the nearest sample's value
            assert (np.abs(got - want).max(axis=2) > 1e-6).mean() <= 5e-3, name


def test_filter_node_branches_and_camera_matrix(ref):
    """The per-sample branches of filter_pixel through the adaptor's FilterPixel (it forwards world-space P, lentil_raydir,
    transmission, the volume / ignore flags and AiWorldToCameraMatrix; the device applies lentil_filter.cpp:119-164)."""
    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, enable_skydome=1)
    r, a = ref.RefCamera(p), ref.AdaptorCamera(p)
    W, H, spp = 128, 72, 9
    f = branch_frame(r.state.tan_fov, W, H, spp)
    aovs = [("RGBA", 0, 1), ("lentil_debug", 0, 2)]
    more = dict(raydir=f["raydir"], transmission=f["transmission"], flags=f["flags"], world_to_camera=f["world_to_camera"])
    for cam in (r, a):
        cam.filter_begin(W, H, aovs, spp=spp)
        cam.filter_accumulate(f["px"], f["py"], f["rgba"], f["pos"], 1.0 / spp, **more)
    for i in range(2):
        bw, ww = r.buffers(i)
        bg, wg = a.buffers(i)
        assert rel_l1(bg, bw) <= 2e-3 and rel_l1(wg, ww) <= 2e-3, i
        assert psnr(a.resolve(i), r.resolve(i)) >= 50.0


def test_cryptomatte_through_the_nodes(ref):
    p = po_params(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    r, a = ref.RefCamera(p), ref.AdaptorCamera(p)
    W, H, spp = 96, 54, 9
    fr = workloads.highlight_frame(W, H, spp, r.state.tan_fov, "cpu")
    aovs = [("RGBA", 0, 1), ("crypto_material00", 2, 0), ("crypto_object01", 2, 0)]
    cr = workloads.crypto_layers(fr, 4, [1, 2])
    crypto = dict(depth=4, count=cr["count"].numpy(), opacity=cr["opacity"].numpy(), ids={k: v.numpy() for k, v in cr["ids"].items()})
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    for cam in (r, a):
        cam.filter_begin(W, H, aovs, spp=spp)
        cam.filter_accumulate(*args, crypto=crypto)
    for i in (1, 2):
        ir, wr, tr, mr = r.crypto(i, 32)
        ia, wa, ta, ma = a.crypto(i, 32)
        assert ma == mr
        np.testing.assert_array_equal(ia.view(np.uint32), ir.view(np.uint32))  # same id sets per pixel (thin-lens splats land identically)
        np.testing.assert_allclose(wa, wr, rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(ta, tr, rtol=2e-4, atol=1e-6)
        want, got = r.resolve(i, fill=-3.0), a.resolve(i, fill=-3.0)
        assert ((want[..., 0] == got[..., 0]) & (np.abs(want[..., 1] - got[..., 1]) < 1e-4)).mean() > 0.999


def test_operator_driven_frame_through_the_nodes(ref):
    """The whole plugin chain on both sides: NodeLoader -> lentil_operator cooks the scene's outputs (every AOV on its original
    gaussian / closest filter node) -> the camera copies the operator's list (user AOVs + lentil_debug, lentil_time,
    lentil_raydir) -> filter_pixel -> driver_process_bucket.  Reference: lentil_operator.cpp:25-171, lentil.h:988-1117."""
    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    aovs = [("RGBA", 0, 1), ("light0", 0, 0), ("N", 1, 0)]
    W, H, spp = 96, 54, 9
    r, a = ref.RefCamera(p), ref.AdaptorCamera(p)
    fr = workloads.highlight_frame(W, H, spp, r.state.tan_fov, "cpu", n_extra_aov=1)
    vals = [None, fr["aov_values"][0].numpy(), None]
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    for cam, threads in ((r, 1), (a, 4)):
        cam.use_operator(True)
        cam.filter_begin(W, H, aovs, spp=spp)
        cam.filter_accumulate(*args, aov_values=vals, nthreads=threads)
    for i in (0, 1):  # gaussian AOVs
        want, got = r.resolve(i), a.resolve(i)
        assert rel_l1(got, want) <= 3e-3 and psnr(got, want) >= 50.0, (aovs[i][0], rel_l1(got, want), psnr(got, want))
    for i in (2, 3):  # closest: the user's N, and lentil_debug (made by the operator; samples * redistribute of the nearest sample)
        want, got = r.resolve(i), a.resolve(i)
        assert (np.abs(got - want).max(axis=2) > 1e-6).mean() <= 5e-3, i
    assert np.abs(r.resolve(3)).max() > 0  # lentil_debug is live
    for i in (4, 5):  # lentil_time (zero in this scene), lentil_raydir: gaussian helper outputs
        want, got = r.resolve(i), a.resolve(i)
        assert rel_l1(got, want) <= 3e-3 or np.abs(want).max() == 0 and np.abs(got).max() == 0, i
