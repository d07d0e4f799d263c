"""Pins the oracle (oracle/lentil_oracle.cpp, a restatement) against the REFERENCE'S OWN SOURCES compiled
behind shims (oracle/_ref/libref.so: src/lentil.h struct Camera, lentil_camera.cpp camera_create_ray,
lentil_filter.cpp filter_pixel, lentil_imager.cpp driver_process_bucket, lens.h, global.h, imagebokeh.h).

Both sides run in double on the CPU from the same lens pack, so agreement is expected to the last bit
wherever the restatement follows the reference's operation order; the tests assert exact equality and
say where a tolerance is used instead.  Skipped when libref.so has not been built (it needs
/root/reference; `make -C oracle -f ref.mk`).
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import orc, ref
from pota_b200 import abi, workloads
from tests.util import po_params

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so not built")


@pytest.fixture(scope="module")
def libs():
    orc.build()
    return orc.lib(), ref.lib()


def test_integer_rng_primitives(libs):
    O, R = libs
    rs = np.random.default_rng(0)
    for a, b in rs.integers(0, 2**32, size=(200, 2), dtype=np.uint64):
        assert O.orc_tea8(int(a), int(b)) == R.ref_tea8(int(a), int(b))
    so, sr = C.c_uint(12345), C.c_uint(12345)
    for _ in range(1000):
        assert O.orc_rng(C.byref(so)) == R.ref_rng(C.byref(sr))
        assert so.value == sr.value
    n = 64
    a = (C.c_uint32 * n)()
    b = (C.c_uint32 * n)()
    O.orc_xor128_seq(a, n)
    R.ref_xor128_seq(b, n)  # first use of the process-global generator in this process
    assert list(a) == list(b)
    assert a[0] == 3701687786  # Marsaglia's xor128 with the seeds of global.h:23


def test_float_primitives(libs):
    O, R = libs
    xs = np.concatenate([np.linspace(-10, 10, 4001), np.random.default_rng(1).normal(0, 3, 2000)]).astype(np.float32)
    for x in xs:
        assert O.orc_fast_sin(float(x)) == R.ref_fast_sin(float(x))
        assert O.orc_fast_cos(float(x)) == R.ref_fast_cos(float(x))
    for e in range(0, 16):
        for x in (-1.7, 0.3, 2.5):
            assert O.orc_lens_ipow(x, e) == R.ref_lens_ipow(x, e)
    d1, d2 = (C.c_double * 2)(), (C.c_double * 2)()
    for ox, oy in np.random.default_rng(2).uniform(0, 1, size=(2000, 2)):
        for fast in (0, 1):
            O.orc_concentric_disk_sample(C.c_double(ox), C.c_double(oy), fast, d1)
            R.ref_concentric_disk_sample(C.c_double(ox), C.c_double(oy), fast, d2)
            assert list(d1) == list(d2)
    n = 20100
    a, b = (C.c_double * n)(), (C.c_double * n)()
    assert O.orc_logarithmic_values(a, n) == R.ref_logarithmic_values(b, n) == 20001  # lens.h:402: drift-sensitive count
    assert list(a) == list(b)


def test_coordinate_transforms(libs):
    O, R = libs
    rs = np.random.default_rng(3)
    A2, A3 = C.c_double * 2, C.c_double * 3
    for _ in range(2000):
        Rr = float(rs.choice([-1, 1]) * rs.uniform(10, 90))
        pos = rs.uniform(-8, 8, 2)
        dr = rs.uniform(-0.5, 0.5, 2)
        p1, d1, p2, d2 = A3(), A3(), A3(), A3()
        O.orc_sphereToCs(A2(*pos), A2(*dr), C.c_double(-Rr), C.c_double(Rr), p1, d1)
        R.ref_sphereToCs(A2(*pos), A2(*dr), C.c_double(-Rr), C.c_double(Rr), p2, d2)
        assert list(p1) == list(p2) and list(d1) == list(d2)
        q1, e1, q2, e2 = A2(), A2(), A2(), A2()
        view = rs.normal(0, 1, 3)
        O.orc_csToSphere(p1, A3(*view), C.c_double(-Rr), C.c_double(Rr), q1, e1)
        R.ref_csToSphere(p2, A3(*view), C.c_double(-Rr), C.c_double(Rr), q2, e2)
        assert list(q1) == list(q2) and list(e1) == list(e2)
        for cyl_y in (0, 1):
            O.orc_cylinderToCs(A2(*pos), A2(*dr), C.c_double(-Rr), C.c_double(Rr), cyl_y, p1, d1)
            R.ref_cylinderToCs(A2(*pos), A2(*dr), C.c_double(-Rr), C.c_double(Rr), cyl_y, p2, d2)
            assert list(p1) == list(p2) and list(d1) == list(d2)
            O.orc_csToCylinder(p1, A3(*view), C.c_double(-Rr), C.c_double(Rr), cyl_y, q1, e1)
            R.ref_csToCylinder(p2, A3(*view), C.c_double(-Rr), C.c_double(Rr), cyl_y, q2, e2)
            assert list(q1) == list(q2) and list(e1) == list(e2)
        o1, o2 = A3(), A3()
        O.orc_line_plane_intersection(p1, d1, o1)
        R.ref_line_plane_intersection(p2, d2, o2)
        assert list(o1) == list(o2)


@pytest.mark.parametrize("lens_model,fstop,focus", [(5, 2.8, 150.0), (0, 1.4, 50.0), (43, 0.0, 1000.0), (28, 5.6, 80.0)])
def test_setup_solvers(lens_model, fstop, focus):
    p = po_params(lens_model=lens_model, fstop=fstop, focus_dist=focus)
    so, sr = orc.OracleCamera(p).state, ref.RefCamera(p).state
    for k in ("aperture_radius", "sensor_shift", "tan_fov", "focus_distance", "lambda_", "lens_outer_pupil_radius", "lens_inner_pupil_radius",
              "lens_length", "lens_back_focal_length", "lens_aperture_housing_radius", "lens_outer_pupil_curvature_radius", "lens_field_of_view"):
        assert getattr(so, k) == getattr(sr, k), k


def test_generated_bodies(libs):
    """lens_evaluate / lens_pt_sample_aperture / lens_lt_sample_aperture: table evaluation == compiled generated code."""
    O, R = libs
    p = po_params(lens_model=17, fstop=2.0, focus_dist=100.0)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    rs = np.random.default_rng(5)
    A5 = C.c_double * 5
    for _ in range(300):
        x = [rs.uniform(-15, 15), rs.uniform(-10, 10), rs.uniform(-0.2, 0.2), rs.uniform(-0.2, 0.2), 0.55]
        o1, o2 = A5(), A5()
        assert O.orc_lens_evaluate(o._h, A5(*x), o1) == R.ref_lens_evaluate(r._h, A5(*x), o2)
        assert list(o1) == list(o2)
        s1, s2 = A5(x[0], x[1], 0, 0, 0.55), A5(x[0], x[1], 0, 0, 0.55)
        a1, a2 = A5(rs.uniform(-4, 4), rs.uniform(-4, 4), 0, 0, 0), A5()
        for k in range(5):
            a2[k] = a1[k]
        O.orc_lens_pt_sample_aperture(o._h, s1, a1, C.c_double(3.3))
        R.ref_lens_pt_sample_aperture(r._h, s2, a2, C.c_double(3.3))
        assert list(s1) == list(s2) and list(a1) == list(a2)
        scene = (C.c_double * 3)(rs.uniform(-300, 300), rs.uniform(-200, 200), rs.uniform(300, 5000))
        ap = (C.c_double * 2)(rs.uniform(-5, 5), rs.uniform(-5, 5))
        t1, t2, u1, u2 = A5(0, 0, 0, 0, 0.55), A5(0, 0, 0, 0, 0.55), A5(0, 0, 0, 0, 0.55), A5(0, 0, 0, 0, 0.55)
        T1 = O.orc_lens_lt_sample_aperture(o._h, scene, ap, t1, u1, C.c_double(0.55))
        T2 = R.ref_lens_lt_sample_aperture(r._h, scene, ap, t2, u2, C.c_double(0.55))
        assert T1 == T2 and list(t1) == list(t2) and list(u1) == list(u2)


@pytest.mark.parametrize("kw", [dict(), dict(aperture_blades_lentil=6), dict(bokeh_enable_image=1), dict(enable_dof=0), dict(units=abi.LB_UNITS_MM, lens_model=40)])
def test_camera_create_ray(kw):
    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = po_params(**kw)
    o, r = orc.OracleCamera(p, img), ref.RefCamera(p, img)
    n = 20000
    w = int(round((n * 16 / 9) ** 0.5))
    ins = workloads.camera_samples(w, -(-n // w), 1, "cpu", 0, n, "linear")
    arrs = [ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]
    a, b = o.create_rays(*arrs), r.create_rays(*arrs)
    # rays whose first try succeeds never touch the retry RNG (the one stated deviation: counter RNG vs global xor128)
    first = a["tries"] == 0
    assert first.mean() > (0.05 if kw.get("enable_dof") == 0 else 0.5)
    for k in orc.RAY_OUT_FIELDS:
        np.testing.assert_array_equal(a[k][:, first], b[k][:, first], err_msg=k)
    # retried rays: same success statistics
    assert abs((a["weight"][0] == 0).mean() - (b["weight"][0] == 0).mean()) < 0.01


def test_reverse_trace_and_coc(libs):
    O, R = libs
    for kw in (dict(), dict(bokeh_enable_image=1), dict(aperture_blades_lentil=5)):
        img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
        p = po_params(fstop=1.4, focus_dist=35.0, **kw)
        o, r = orc.OracleCamera(p, img), ref.RefCamera(p, img)
        rs = np.random.default_rng(7)
        s1, s2 = (C.c_double * 2)(), (C.c_double * 2)()
        for i in range(400):
            tgt = (C.c_double * 3)(rs.uniform(-250, 250), rs.uniform(-150, 150), rs.uniform(200, 3000))
            px, py, tot = int(rs.integers(0, 1920)), int(rs.integers(0, 1080)), int(rs.integers(0, 5000))
            ok1 = O.orc_trace_ray_bw_po(o._h, tgt, px, py, tot, C.c_float(0.55), s1)
            ok2 = R.ref_trace_ray_bw_po(r._h, tgt, px, py, tot, C.c_float(0.55), s2)
            assert ok1 == ok2 and (not ok1 or list(s1) == list(s2)), (kw, i)
        for z in np.linspace(-500, -1, 200, dtype=np.float32):
            assert O.orc_get_coc_thinlens(o._h, float(z)) == R.ref_get_coc_thinlens(r._h, float(z))


@pytest.mark.parametrize("kw,aovs,n_extra", [
    (dict(), [("RGBA", 0, 1)], 0),
    (dict(bokeh_enable_image=1, bidir_add_energy=1.0), [("RGBA", 0, 1), ("light0", 0, 0), ("light1", 0, 0)], 2),
    (dict(abb_chromatic=0.4), [("RGBA", 0, 1)], 0),
    (dict(), [("RGBA", 0, 1), ("N", 1, 0), ("lentil_debug", 1, 2)], 1),  # closest-filter AOVs (not named Z: that name is the depth AOV)
])
def test_filter_pixel_and_imager(kw, aovs, n_extra):
    """filter_pixel + driver_process_bucket of the reference vs the oracle, single-threaded: identical framebuffers."""
    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **kw)
    o, r = orc.OracleCamera(p, img), ref.RefCamera(p, img)
    W, H, spp = 128, 72, 9
    fr = workloads.highlight_frame(W, H, spp, o.state.tan_fov, "cpu", n_extra_aov=n_extra)
    vals = [None] + [v.numpy() for v in fr["aov_values"]] + [None] * (len(aovs) - 1 - n_extra)
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(*args, aov_values=vals)
    r.filter_begin(W, H, aovs, spp=spp)
    r.filter_accumulate(*args, aov_values=vals)
    assert o.filter_stats()["redistributed"] > 50
    for a in range(len(aovs)):
        bo, wo = o.buffers(a)
        br, wr = r.buffers(a)
        np.testing.assert_array_equal(bo, br, err_msg=f"buffer of {aovs[a][0]}")
        np.testing.assert_array_equal(wo, wr)
        np.testing.assert_array_equal(o.resolve(a), r.resolve(a), err_msg=f"resolve of {aovs[a][0]}")


def test_filter_region(libs):
    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    Wf, Hf, spp = 128, 72, 9
    x0, y0, W, H = 32, 16, 64, 40
    fr = workloads.highlight_frame(Wf, Hf, spp, o.state.tan_fov, "cpu")
    px, py = fr["px"].numpy(), fr["py"].numpy()
    m = (px >= x0) & (px < x0 + W) & (py >= y0) & (py < y0 + H)
    args = (px[m] - x0, py[m] - y0, fr["rgba"].numpy()[m], fr["pos_cs"].numpy()[m], 1.0 / spp)
    aovs = [("RGBA", 0, 1)]
    o.filter_begin(W, H, aovs, xres_full=Wf, yres_full=Hf, region_min=(x0, y0))
    o.filter_accumulate(*args)
    r.filter_begin(W, H, aovs, xres_full=Wf, yres_full=Hf, region_min=(x0, y0), spp=spp)
    r.filter_accumulate(*args)
    np.testing.assert_array_equal(o.buffers(0)[0], r.buffers(0)[0])
    np.testing.assert_array_equal(o.resolve(0), r.resolve(0))


@pytest.mark.parametrize("outer", [1, 2])
def test_cylindrical_outer_pupil(outer):
    """Anamorphic lenses (lens_outer_pupil_geometry cyl-y / cyl-x, lentil.h:387-388): the pack has none, so the
    geometry is overridden on both sides; forward rays and reverse traces must still agree bit for bit."""
    p = po_params(fstop=2.0, focus_dist=80.0)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    o.set_pupil_geometry(outer)
    r.set_pupil_geometry(outer)
    n = 5000
    ins = workloads.camera_samples(100, 50, 1, "cpu", 0, n, "linear")
    arrs = [ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]
    a, b = o.create_rays(*arrs), r.create_rays(*arrs)
    first = a["tries"] == 0
    assert first.mean() > 0.5
    for k in orc.RAY_OUT_FIELDS:
        np.testing.assert_array_equal(a[k][:, first], b[k][:, first], err_msg=k)
    O, R = orc.lib(), ref.lib()
    rs = np.random.default_rng(11)
    s1, s2 = (C.c_double * 2)(), (C.c_double * 2)()
    for i in range(200):
        tgt = (C.c_double * 3)(rs.uniform(-250, 250), rs.uniform(-150, 150), rs.uniform(200, 3000))
        ok1 = O.orc_trace_ray_bw_po(o._h, tgt, 10 + i, 20, i, C.c_float(0.55), s1)
        ok2 = R.ref_trace_ray_bw_po(r._h, tgt, 10 + i, 20, i, C.c_float(0.55), s2)
        assert ok1 == ok2 and (not ok1 or list(s1) == list(s2))


# ---- thin-lens path: trace_ray_fw_thinlens lentil.h:431-569, filter ThinLens branch lentil_filter.cpp:303-447
TL_CASES = [
    dict(),
    dict(fstop=1.4, focus_dist=80.0, focal_length_lentil=50.0),
    dict(abb_coma=0.6, abb_distortion=0.3, optical_vignetting=2.0, bokeh_circle_to_square=0.5, bokeh_anamorphic=0.3),
    dict(abb_spherical=0.3, aperture_blades_lentil=6),
    dict(bokeh_enable_image=1, units=abi.LB_UNITS_M, abb_coma=0.2),
]


def _tl_params(**kw):
    base = dict(camera_type=abi.LB_CAMERA_THINLENS, fstop=2.8, focus_dist=150.0)
    base.update(kw)
    return abi.CameraParams.defaults(**base)


@pytest.mark.parametrize("kw", TL_CASES)
def test_thinlens_camera_create_ray(kw):
    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    p = _tl_params(**kw)
    o, r = orc.OracleCamera(p, img), ref.RefCamera(p, img)
    assert o.state.aperture_radius == r.state.aperture_radius and o.state.tan_fov == r.state.tan_fov
    n = 20000
    w = int(round((n * 16 / 9) ** 0.5))
    ins = workloads.camera_samples(w, -(-n // w), 1, "cpu", 0, n, "linear")
    arrs = [ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]
    a, b = o.create_rays(*arrs), r.create_rays(*arrs)
    first = a["tries"] == 0
    assert first.mean() > 0.2  # optical vignetting rejects many first tries
    for k in orc.RAY_OUT_FIELDS:
        np.testing.assert_array_equal(a[k][:, first], b[k][:, first], err_msg=k)


@pytest.mark.parametrize("kw", TL_CASES)
def test_thinlens_filter_pixel_and_imager(kw):
    img = workloads.disc_bokeh_image(64) if kw.get("bokeh_enable_image") else None
    kw = dict(kw)
    kw.setdefault("fstop", 1.4)
    kw.setdefault("focus_dist", 35.0)
    p = _tl_params(bidir_sample_mult=6, **kw)
    o, r = orc.OracleCamera(p, img), ref.RefCamera(p, img)
    W, H, spp = 128, 72, 9
    z = 0.75 if kw.get("units") == abi.LB_UNITS_M else 75.0
    fr = workloads.highlight_frame(W, H, spp, o.state.tan_fov, "cpu", z_plane=z, pitch=z * 0.072, radius=z * 0.0018, n_extra_aov=1)
    aovs = [("RGBA", 0, 1), ("light0", 0, 0), ("N", 1, 0)]
    vals = [None, fr["aov_values"][0].numpy(), fr["aov_values"][0].numpy()]
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(*args, aov_values=vals)
    r.filter_begin(W, H, aovs, spp=spp)
    r.filter_accumulate(*args, aov_values=vals)
    st = o.filter_stats()
    assert st["redistributed"] > 50 and st["splats"] > 500, st
    for a in range(len(aovs)):
        bo, wo = o.buffers(a)
        br, wr = r.buffers(a)
        np.testing.assert_array_equal(bo, br, err_msg=f"buffer of {aovs[a][0]}")
        np.testing.assert_array_equal(wo, wr)
        np.testing.assert_array_equal(o.resolve(a), r.resolve(a))


CRYPTO_AOVS = [("RGBA", 0, 1), ("crypto_material00", 2, 0), ("crypto_material01", 2, 0), ("crypto_object02", 2, 0), ("crypto_asset00", 2, 0)]


@pytest.mark.parametrize("kw", [dict(), dict(abb_chromatic=0.4), dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0)])
def test_cryptomatte_accumulate_and_ranked_resolve(kw):
    """cryptomatte_construct_cache + add_to_buffer_cryptomatte (lentil.h:779-819) and the ranked resolve
    (lentil_imager.cpp:122-161, including its early row `break`): reference vs oracle, identical tables."""
    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **kw)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    W, H, spp = 96, 54, 9
    fr = workloads.highlight_frame(W, H, spp, o.state.tan_fov, "cpu")
    cr = workloads.crypto_layers(fr, 4, [1, 2, 3, 4])
    crypto = dict(depth=4, count=cr["count"].numpy(), opacity=cr["opacity"].numpy(), ids={a: v.numpy() for a, v in cr["ids"].items()})
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    o.filter_begin(W, H, CRYPTO_AOVS)
    o.filter_accumulate(*args, crypto=crypto)
    r.filter_begin(W, H, CRYPTO_AOVS, spp=spp)
    r.filter_accumulate(*args, crypto=crypto)
    assert o.filter_stats()["redistributed"] > 50
    np.testing.assert_array_equal(o.buffers(0)[0], r.buffers(0)[0])
    sizes = []
    for a in range(1, len(CRYPTO_AOVS)):
        io, wo, to, mo = o.crypto(a, 32)
        ir, wr, tr, mr = r.crypto(a, 32)
        assert mo == mr and mo <= 32
        sizes.append(mo)
        np.testing.assert_array_equal(io.view(np.uint32), ir.view(np.uint32), err_msg=f"ids of {CRYPTO_AOVS[a][0]}")
        np.testing.assert_array_equal(wo, wr)
        np.testing.assert_array_equal(to, tr)
        # whole frame and a bucket; -7 marks what the imager left untouched
        for box in (dict(), dict(x0=16, y0=8, w=40, h=24)):
            ro, rr = o.resolve(a, fill=-7.0, **box), r.resolve(a, fill=-7.0, **box)
            np.testing.assert_array_equal(ro.view(np.uint32), rr.view(np.uint32), err_msg=f"resolve of {CRYPTO_AOVS[a][0]}")
    assert max(sizes) >= 5  # the scene does reach rank 4
    # rank-0 coverage of a resolved pixel: ids' weights over the total weight
    res = o.resolve(1, fill=-7.0)
    assert np.all(res[..., 1] <= 1.0 + 1e-5) and np.all(res[..., 1] >= res[..., 3]) and np.all(res[..., 1] > 0)
    res4 = o.resolve(3, fill=-7.0)  # rank 4: most rows end early
    done = res4[..., 1] != -7.0
    assert done.any() and (~done).any()


def test_reverse_trace_on_reference_debug_positions(libs):
    """The reference's own fixture of camera-space sample positions (tests/po_bidir_debug/
    po_bidir_spheres_debug_position.txt, 3699 rows, input only): trace_ray_bw_po + get_coc_thinlens on every row,
    reference vs oracle, bit for bit."""
    path = "/root/reference/tests/po_bidir_debug/po_bidir_spheres_debug_position.txt"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    import re
    rows = np.array([[float(v) for v in re.findall(r"[-+0-9.eE]+", ln)] for ln in open(path) if ln.strip()], dtype=np.float64)
    assert rows.shape == (3699, 3)
    O, R = libs
    p = po_params(fstop=2.0, focus_dist=100.0)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    s1, s2 = (C.c_double * 2)(), (C.c_double * 2)()
    hits = 0
    for i, P in enumerate(rows):
        # the filter hands trace_ray_bw_po the target -P_cs * 10 (lentil_filter.cpp:271)
        tgt = (C.c_double * 3)(*(-P * 10.0))
        px, py, tot = (i * 7) % 1920, (i * 13) % 1080, i % 50
        ok1 = O.orc_trace_ray_bw_po(o._h, tgt, px, py, tot, C.c_float(0.55), s1)
        ok2 = R.ref_trace_ray_bw_po(r._h, tgt, px, py, tot, C.c_float(0.55), s2)
        assert ok1 == ok2 and (not ok1 or list(s1) == list(s2)), i
        hits += int(bool(ok1))
        assert O.orc_get_coc_thinlens(o._h, float(np.float32(P[2]))) == R.ref_get_coc_thinlens(r._h, float(np.float32(P[2])))
    assert hits > 3000


@pytest.mark.parametrize("lens_model", range(44))
def test_every_lens_forward_and_reverse(lens_model, libs):
    """Lens-pack breadth: setup solvers, camera_create_ray and trace_ray_bw_po of the reference vs the oracle for each of
    the 44 lens ids (the generated bodies differ per lens; the wrappers around them must not care)."""
    O, R = libs
    from pota_b200.lensgen.prescriptions import LENS_IDS
    focal = float(LENS_IDS[lens_model].split("__")[-1].replace("mm", ""))
    p = po_params(lens_model=lens_model, fstop=2.0, focus_dist=120.0, sensor_width=min(36.0, 0.7 * focal))
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    so, sr = o.state, r.state
    assert (so.aperture_radius, so.sensor_shift, so.tan_fov) == (sr.aperture_radius, sr.sensor_shift, sr.tan_fov)
    n = 1500
    ins = workloads.camera_samples(50, 30, 1, "cpu", 0, n, "linear")
    arrs = [ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]
    a, b = o.create_rays(*arrs), r.create_rays(*arrs)
    first = a["tries"] == 0
    assert first.mean() > 0.3
    for k in orc.RAY_OUT_FIELDS:
        np.testing.assert_array_equal(a[k][:, first], b[k][:, first], err_msg=k)
    rs = np.random.default_rng(lens_model)
    s1, s2 = (C.c_double * 2)(), (C.c_double * 2)()
    for i in range(60):
        tgt = (C.c_double * 3)(rs.uniform(-200, 200), rs.uniform(-120, 120), rs.uniform(300, 3000))
        ok1 = O.orc_trace_ray_bw_po(o._h, tgt, 17 + i, 29 + 2 * i, i, C.c_float(0.55), s1)
        ok2 = R.ref_trace_ray_bw_po(r._h, tgt, 17 + i, 29 + 2 * i, i, C.c_float(0.55), s2)
        assert ok1 == ok2 and (not ok1 or list(s1) == list(s2)), i


def test_cryptomatte_render_region(libs):
    """Cryptomatte tables and ranked buckets inside a render region (region_min != 0, lentil.h:1070-1080)."""
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    Wf, Hf, spp = 96, 54, 9
    x0, y0, W, H = 24, 12, 48, 30
    fr = workloads.highlight_frame(Wf, Hf, spp, o.state.tan_fov, "cpu")
    cr = workloads.crypto_layers(fr, 3, [1, 2])
    px, py = fr["px"].numpy(), fr["py"].numpy()
    m = (px >= x0) & (px < x0 + W) & (py >= y0) & (py < y0 + H)
    crypto = dict(depth=3, count=cr["count"].numpy()[m], opacity=cr["opacity"].numpy()[m], ids={a: v.numpy()[m] for a, v in cr["ids"].items()})
    args = (px[m] - x0, py[m] - y0, fr["rgba"].numpy()[m], fr["pos_cs"].numpy()[m], 1.0 / spp)
    aovs = [("RGBA", 0, 1), ("crypto_object00", 2, 0), ("crypto_object01", 2, 0)]
    o.filter_begin(W, H, aovs, xres_full=Wf, yres_full=Hf, region_min=(x0, y0))
    o.filter_accumulate(*args, crypto=crypto)
    r.filter_begin(W, H, aovs, xres_full=Wf, yres_full=Hf, region_min=(x0, y0), spp=spp)
    r.filter_accumulate(*args, crypto=crypto)
    for a in (1, 2):
        io, wo, to, mo = o.crypto(a, 32)
        ir, wr, tr, mr = r.crypto(a, 32)
        assert mo == mr
        np.testing.assert_array_equal(io.view(np.uint32), ir.view(np.uint32))
        np.testing.assert_array_equal(wo, wr)
        np.testing.assert_array_equal(to, tr)
        for box in (dict(), dict(x0=x0 + 8, y0=y0 + 4, w=20, h=10)):
            np.testing.assert_array_equal(o.resolve(a, fill=-7.0, **box).view(np.uint32), r.resolve(a, fill=-7.0, **box).view(np.uint32))


@pytest.mark.parametrize("kw", [dict(), dict(enable_skydome=1), dict(enable_bidir_transmission=1, enable_skydome=1),
                                dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, enable_skydome=1)])
def test_filter_sample_branches_and_camera_matrix(kw):
    """Transmission, volume, lentil_bidir_ignore, skydome / lentil_raydir, small world-space P and a rotated + translated
    camera (lentil_filter.cpp:115-164): reference vs oracle, identical framebuffers."""
    from tests.util import branch_frame

    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **kw)
    o, r = orc.OracleCamera(p), ref.RefCamera(p)
    W, H, spp = 128, 72, 9
    f = branch_frame(o.state.tan_fov, W, H, spp)
    aovs = [("RGBA", 0, 1), ("lentil_debug", 0, 2)]
    args = (f["px"], f["py"], f["rgba"], f["pos"], 1.0 / spp)
    more = dict(raydir=f["raydir"], transmission=f["transmission"], flags=f["flags"], world_to_camera=f["world_to_camera"])
    o.filter_begin(W, H, aovs)
    o.filter_accumulate(*args, **more)
    r.filter_begin(W, H, aovs, spp=spp)
    r.filter_accumulate(*args, **more)
    st = o.filter_stats()
    m = f["masks"]
    assert st["redistributed"] > 50 and st["passthrough"] > st["redistributed"]
    # the skydome patch is redistributed only when enabled; the zero-raydir patch never is
    n_sun = int(m["sun"].sum())
    o2 = orc.OracleCamera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **{**kw, "enable_skydome": 0}))
    o2.filter_begin(W, H, aovs)
    o2.filter_accumulate(*args, **more)
    assert st["redistributed"] - o2.filter_stats()["redistributed"] == (n_sun if kw.get("enable_skydome") else 0)
    for a in range(len(aovs)):
        bo, wo = o.buffers(a)
        br, wr = r.buffers(a)
        np.testing.assert_array_equal(bo, br, err_msg=f"buffer of {aovs[a][0]}")
        np.testing.assert_array_equal(wo, wr)
        np.testing.assert_array_equal(o.resolve(a), r.resolve(a))
