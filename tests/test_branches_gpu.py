"""GPU parity, round 2: the per-sample branches of filter_pixel that the round-1 suite did not reach (transmission,
volume, lentil_bidir_ignore, skydome / lentil_raydir, small world-space P, a rotated + translated camera), the device
primitives as known-answer tests, the three generations of per-lens kernels against each other, the host-bucket
resolve, stream ordering, and parity runs at the sizes BASELINE.json names (C1: 1 M rays; C3: 1920x1080x16 spp).

Everything goes through the C ABI (pota_b200.camera is a ctypes shim); the checker is the CPU oracle, itself pinned
bit for bit to the compiled reference on the same cases (tests/test_oracle_vs_ref.py).
"""
import ctypes as C
import os

import numpy as np
import pytest

from pota_b200 import abi, workloads
from tests.util import branch_frame, po_params, rel_err_vec

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RGBA = ("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)
IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")


def rel_l1(a, b):
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))


def psnr(a, b):
    peak = float(np.abs(b).max())
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


# ---- device primitives ---------------------------------------------------------------------------------------------
def test_device_primitives_known_answers():
    """tea<8>, rng (global.h:32-57) and fast_sin / fast_cos (lens.h:17-37) evaluated ON THE DEVICE, bit for bit against the
    oracle's (which test_oracle_vs_ref pins to the reference's own headers), plus literal known answers."""
    from oracle import orc
    from pota_b200.camera import lib

    O = orc.lib()
    rs = np.random.default_rng(7)
    n = 4096
    v0 = rs.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    v1 = rs.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    v0[:4] = [0, 1, 0xFFFFFFFF, 12345]
    v1[:4] = [0, 0, 0xFFFFFFFF, 0x5EED]
    tea = np.zeros(n, np.uint32)
    st = np.zeros((n, 4), np.uint32)
    fl = np.zeros((n, 4), np.float32)
    tr = np.zeros((n, 2), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib().lb_debug_primitives(0, n, p(v0), p(v1), p(tea), p(st), p(fl), p(tr)) == 0
    for i in range(n):
        t = O.orc_tea8(int(v0[i]), int(v1[i]))
        assert t == int(tea[i]), i
        s = C.c_uint(t)
        for k in range(4):
            f = O.orc_rng(C.byref(s))
            assert s.value == int(st[i, k]) and np.float32(f) == fl[i, k], (i, k)
        x = np.float32((fl[i, 0] - np.float32(0.5)) * np.float32(20.0))
        assert np.float32(O.orc_fast_sin(C.c_float(float(x)))) == tr[i, 0] and np.float32(O.orc_fast_cos(C.c_float(float(x)))) == tr[i, 1], i
    # literal known answers (computed from the reference's global.h, see tests/test_oracle_vs_ref.py::test_integer_rng_primitives)
    assert int(tea[0]) == O.orc_tea8(0, 0) and 0.0 <= fl.min() and fl.max() < 1.0
    s = C.c_uint(12345)
    assert O.orc_rng(C.byref(s)) == np.float32(((12345 * 1664525 + 1013904223) & 0xFFFFFF) / 16777216.0)


# ---- filter_pixel branches -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(enable_skydome=1), dict(enable_bidir_transmission=1, enable_skydome=1),
                                dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, enable_skydome=1)])
def test_filter_sample_branches_and_camera_matrix(kw):
    from oracle import orc
    from pota_b200.camera import Camera

    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **kw)
    ocam, gcam = orc.OracleCamera(p), Camera(p, device=0)
    W, H, spp = 128, 72, 9
    f = branch_frame(ocam.state.tan_fov, W, H, spp)
    aovs = [RGBA, ("lentil_debug", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_LENTIL_DEBUG)]
    ocam.filter_begin(W, H, aovs)
    ocam.filter_accumulate(f["px"], f["py"], f["rgba"], f["pos"], 1.0 / spp, raydir=f["raydir"], transmission=f["transmission"], flags=f["flags"],
                           world_to_camera=f["world_to_camera"], nthreads=8)
    gcam.filter_begin(W, H, aovs)
    t = {k: torch.from_numpy(f[k]).cuda() for k in ("px", "py", "rgba", "pos", "raydir", "transmission")}
    fl = torch.from_numpy(f["flags"].view(np.int32)).cuda()
    gcam.filter_accumulate(t["px"], t["py"], t["rgba"], t["pos"], 1.0 / spp, raydir=t["raydir"], transmission=t["transmission"], flags=fl,
                           world_to_camera=f["world_to_camera"])
    so, sg = ocam.filter_stats(), gcam.filter_stats()
    for k in ("samples", "redistributed", "passthrough"):  # integer decisions: exact
        assert so[k] == sg[k], (k, so, sg)
    assert so["redistributed"] > 50
    for k in ("splats", "attempts"):
        assert abs(so[k] - sg[k]) <= 2e-3 * so[k] + 2, (k, so, sg)
    device_buffers = []
    for a in range(2):
        bo, wo = ocam.buffers(a)
        bg, wg = gcam.buffers(a)
        device_buffers.append(bg)
        assert rel_l1(bg, bo) <= 2e-3 and rel_l1(wg, wo) <= 2e-3, (a, rel_l1(bg, bo), rel_l1(wg, wo))
        assert psnr(gcam.resolve(a).cpu().numpy(), ocam.resolve(a)) >= 55.0
    # the same batch through the host-buffer entry point (pageable numpy memory) gives the same framebuffers
    gcam.filter_begin(W, H, aovs)
    gcam.filter_accumulate_host(f["px"], f["py"], f["rgba"], f["pos"], 1.0 / spp, raydir=f["raydir"], transmission=f["transmission"],
                                flags=f["flags"], world_to_camera=f["world_to_camera"])
    assert gcam.filter_stats()["redistributed"] == so["redistributed"]
    assert rel_l1(gcam.buffers(0)[0], device_buffers[0]) <= 1e-3 and rel_l1(gcam.buffers(1)[0], device_buffers[1]) <= 1e-3


def test_samples_outside_the_region_are_dropped():
    from pota_b200.camera import Camera

    cam = Camera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=4), device=0)
    W, H, spp = 64, 36, 4
    fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, "cuda")
    cam.filter_begin(W, H, [RGBA])
    cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp)
    ref_buf, ref_w = cam.buffers(0)
    bad_px = torch.cat([fr["px"], torch.tensor([-1, W, 5, 5, 2**30], dtype=torch.int32, device="cuda")])
    bad_py = torch.cat([fr["py"], torch.tensor([3, 3, -7, H, 2**30], dtype=torch.int32, device="cuda")])
    extra = torch.ones((5, 4), device="cuda") * 50.0
    pos = torch.cat([fr["pos_cs"], torch.tensor([[0.0, 0.0, -75.0, 75.0]] * 5, device="cuda")])
    cam.filter_begin(W, H, [RGBA])
    cam.filter_accumulate(bad_px, bad_py, torch.cat([fr["rgba"], extra]), pos, 1.0 / spp)
    buf, w = cam.buffers(0)
    np.testing.assert_allclose(w, ref_w, rtol=1e-6)
    np.testing.assert_allclose(buf.sum(dtype=np.float64), ref_buf.sum(dtype=np.float64), rtol=1e-5)


def test_begin_reallocates_by_pixel_count():
    """Two frames with the same number of plane floats but different pixel counts (100x50 with 2 AOVs, then 100x90 with one):
    the closest-filter key planes follow the pixel count (advisor finding r01)."""
    from pota_b200.camera import Camera

    cam = Camera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=4), device=0)
    cam.filter_begin(100, 50, [RGBA, ("N", abi.LB_FILTER_CLOSEST, 0)])
    fr = workloads.highlight_frame(100, 90, 4, cam.state.tan_fov, "cuda", n_extra_aov=1)
    cam.filter_begin(100, 90, [("N", abi.LB_FILTER_CLOSEST, 0)])
    cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 0.25, aov_values=[fr["rgba"]])
    img = cam.resolve(0).cpu().numpy()
    assert img.shape == (90, 100, 4) and np.isfinite(img).all() and (img[..., 3] == 1.0).all()


def test_operations_on_different_streams_execute_in_call_order():
    """Two accumulates and the resolve issued on three different non-blocking streams share the work list and the planes;
    the library orders them on the device (advisor finding r01)."""
    from pota_b200.camera import Camera

    cam = Camera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=8), device=0)
    W, H, spp = 192, 108, 4
    fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, "cuda")
    n = fr["px"].shape[0]
    cam.filter_begin(W, H, [RGBA])
    cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp)
    want = cam.resolve(0).cpu().numpy()
    st_want = cam.filter_stats()
    torch.cuda.synchronize()
    s1, s2, s3 = (torch.cuda.Stream() for _ in range(3))
    for _ in range(3):
        cam.filter_begin(W, H, [RGBA])
        h = n // 2
        cam.filter_accumulate(fr["px"][:h], fr["py"][:h], fr["rgba"][:h], fr["pos_cs"][:h], 1.0 / spp, stream=s1)
        cam.filter_accumulate(fr["px"][h:], fr["py"][h:], fr["rgba"][h:], fr["pos_cs"][h:], 1.0 / spp, stream=s2)
        got = cam.resolve(0, stream=s3)
        torch.cuda.synchronize()
        st = cam.filter_stats()
        assert st["splats"] == st_want["splats"] and st["attempts"] == st_want["attempts"] and st["redistributed"] == st_want["redistributed"]
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-4, atol=1e-5)  # float accumulation order only


# ---- host-bucket resolve -------------------------------------------------------------------------------------------
def test_resolve_host_serves_buckets_like_the_device_resolve():
    """driver_process_bucket per 16x16 bucket with host memory, gaussian / closest / ranked cryptomatte AOVs (whose rows end at
    the first pixel of the BUCKET row with <= rank ids, lentil_imager.cpp:132-134): identical to the per-bucket device resolve."""
    from pota_b200.camera import Camera

    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, fstop=1.4, focus_dist=35.0, focal_length_lentil=50.0, bidir_sample_mult=6)
    cam = Camera(p, device=0)
    W, H, spp = 96, 54, 4
    fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, "cuda", n_extra_aov=1)
    aovs = [RGBA, ("N", abi.LB_FILTER_CLOSEST, 0), ("crypto_material00", abi.LB_FILTER_CRYPTO, 0), ("crypto_material01", abi.LB_FILTER_CRYPTO, 0)]
    cr = workloads.crypto_layers(fr, 3, [2, 3])
    cam.filter_begin(W, H, aovs)
    cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, aov_values=[None, fr["aov_values"][0], None, None],
                          crypto=dict(depth=3, count=cr["count"], opacity=cr["opacity"], ids=cr["ids"]))
    for a in range(len(aovs)):
        for (x0, y0, w, h) in [(0, 0, 16, 16), (80, 38, 16, 16), (16, 32, 16, 16), (0, 0, W, H), (37, 5, 23, 1)]:
            dev = cam.resolve(a, x0, y0, w, h, fill=-7.0).cpu().numpy()
            host = np.full((h, w, 4), -7.0, np.float32)
            cam.resolve_host(a, host, x0, y0, w, h)
            np.testing.assert_array_equal(host, dev, err_msg=f"aov {aovs[a][0]} bucket {(x0, y0, w, h)}")
    # a later accumulate invalidates the cached image
    before = cam.resolve_host(0).copy()
    cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, aov_values=[None, fr["aov_values"][0], None, None],
                          crypto=dict(depth=3, count=cr["count"], opacity=cr["opacity"], ids=cr["ids"]))
    np.testing.assert_array_equal(cam.resolve_host(0), cam.resolve(0).cpu().numpy())
    assert np.isfinite(before).all()


# ---- kernel generations --------------------------------------------------------------------------------------------
def test_kernel_generations_agree(monkeypatch):
    """The wavelength-folded bodies (coefficient table / 550 nm immediates) and the first-generation 5-variate body evaluate
    the same polynomials: rays agree to float rounding, splat statistics are the same."""
    from pota_b200.camera import Camera

    n = 200_000
    ins = workloads.camera_samples(500, 400, 1, "cuda", 0, n, "linear")
    W, H, spp = 192, 108, 4
    rays, stats, imgs = {}, {}, {}
    for gen in ("2", "1", "0"):
        monkeypatch.setenv("LB_KERNEL_GEN", gen)
        cam = Camera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=8), device=0)
        assert cam.kernel_kind == "unrolled"
        rays[gen] = {k: v.cpu().numpy() for k, v in cam.create_rays(*[ins[k] for k in IN_KEYS]).items()}
        fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, "cuda")
        cam.filter_begin(W, H, [RGBA])
        cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp)
        stats[gen] = cam.filter_stats()
        imgs[gen] = cam.resolve(0).cpu().numpy()
    for gen in ("1", "0"):
        np.testing.assert_array_equal(rays[gen]["tries"], rays["2"]["tries"])
        live = rays["2"]["weight"][0] != 0
        for k in ("origin", "dir"):
            assert (rel_err_vec(rays[gen][k], rays["2"][k])[live] <= 2e-5).mean() > 0.9999, (gen, k)
        assert stats[gen]["redistributed"] == stats["2"]["redistributed"]
        assert abs(stats[gen]["splats"] - stats["2"]["splats"]) <= 1e-3 * stats["2"]["splats"]
        assert rel_l1(imgs[gen], imgs["2"]) <= 2e-3
    # another wavelength exercises the coefficient-table kernels under the default generation
    monkeypatch.setenv("LB_KERNEL_GEN", "2")
    from oracle import orc

    p = po_params(fstop=2.8, wavelength=610.0)
    g = Camera(p, device=0).create_rays(*[ins[k] for k in IN_KEYS])
    o = orc.OracleCamera(p).create_rays(*[ins[k].cpu().numpy() for k in IN_KEYS], nthreads=8)
    live = o["weight"][0] != 0
    for k in ("origin", "dir"):
        assert (rel_err_vec(g[k].cpu().numpy(), o[k])[live] <= 1e-4).mean() > 0.9999, k


# ---- BASELINE.json sizes ---------------------------------------------------------------------------------------------
def test_c1_one_million_rays_against_the_oracle():
    """Config C1 at its full size: 1 000 000 samples, seed = tea<8>(i, 0x5EED), lens 5 f/2.8 focus 150 cm."""
    from oracle import orc
    from pota_b200.camera import Camera

    n = 1_000_000
    p = po_params()
    ins = workloads.camera_samples(1000, 1000, 1, "cuda", 0, n, "linear")
    g = Camera(p, device=0).create_rays(*[ins[k] for k in IN_KEYS])
    o = orc.OracleCamera(p).create_rays(*[ins[k].cpu().numpy() for k in IN_KEYS], nthreads=os.cpu_count() or 8)
    np.testing.assert_array_equal(g["tries"].cpu().numpy(), o["tries"])
    np.testing.assert_array_equal(g["weight"].cpu().numpy(), o["weight"])
    live = o["weight"][0] != 0
    assert live.mean() > 0.9
    worst = {}
    for k in ("origin", "dir"):
        e = rel_err_vec(g[k].cpu().numpy(), o[k])[live]
        worst[k] = (float(np.median(e)), float(np.quantile(e, 0.9999)), float((e <= 1e-4).mean()))
        assert (e <= 1e-4).mean() >= 0.9999, (k, worst[k])
    print("C1 parity (median, p99.99, fraction within 1e-4):", worst)


def test_c3_sized_frame_against_the_oracle_and_conservation():
    """Config C3 at its full size (1920x1080, 16 spp, 250x250 image-bokeh kernel).  The oracle runs ~6 k splats/s per host
    thread, so it gets a bounded sub-sample of the SOURCE samples (every 48th highlight sample, ~1.5 M splats, plus the
    background samples of 16 pixel rows); both sides accumulate that sub-sample into full-size framebuffers.  The whole frame
    then runs on the GPU alone and is checked through the size-independent properties."""
    from oracle import orc
    from pota_b200.camera import Camera

    p = po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=10, bokeh_enable_image=1)
    img = workloads.disc_bokeh_image(250)
    gcam, ocam = Camera(p, img, device=0), orc.OracleCamera(p, img)
    W, H, spp = 1920, 1080, 16
    fr = workloads.highlight_frame(W, H, spp, gcam.state.tan_fov, "cuda", grid=(8, 4))
    hit = fr["rgba"][:, 3] > 0
    idx_hit = torch.nonzero(hit).flatten()[::48]
    band = torch.nonzero((~hit) & (fr["py"] >= 500) & (fr["py"] < 516)).flatten()
    idx = torch.sort(torch.cat([idx_hit, band])).values
    sub = {k: fr[k][idx].contiguous() for k in ("px", "py", "rgba", "pos_cs")}
    gcam.filter_begin(W, H, [RGBA])
    gcam.filter_accumulate(sub["px"], sub["py"], sub["rgba"], sub["pos_cs"], 1.0 / spp)
    ocam.filter_begin(W, H, [RGBA])
    ocam.filter_accumulate(*[sub[k].cpu().numpy() for k in ("px", "py", "rgba", "pos_cs")], 1.0 / spp, nthreads=os.cpu_count() or 8)
    so, sg = ocam.filter_stats(), gcam.filter_stats()
    assert so["redistributed"] == sg["redistributed"] == idx_hit.numel() and so["passthrough"] == sg["passthrough"]
    assert abs(so["splats"] - sg["splats"]) <= 1e-3 * so["splats"], (so, sg)
    bo, wo = ocam.buffers(0)
    bg, wg = gcam.buffers(0)
    l1, l1w, db = rel_l1(bg, bo), rel_l1(wg, wo), psnr(bg, bo)
    print(f"C3-size parity on {sg['splats']} splats: buffer rel-L1 {l1:.2e}, weight rel-L1 {l1w:.2e}, PSNR {db:.1f} dB")
    assert l1 <= 2e-3 and l1w <= 2e-3 and db >= 60.0
    # the whole frame, GPU only: weight and energy conservation (every disc's bokeh stays inside the frame)
    gcam.filter_begin(W, H, [RGBA])
    gcam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp)
    st = gcam.filter_stats()
    assert st["samples"] == W * H * spp and st["redistributed"] == int(hit.sum()) and st["splats"] >= 0.999 * st["attempts"]
    buf, wgt = gcam.buffers(0)
    np.testing.assert_allclose(wgt.sum(dtype=np.float64), W * H, rtol=1e-4)
    np.testing.assert_allclose(buf[..., :3].sum(dtype=np.float64), 3 * float(fr["rgba"][:, 0].double().sum()) / spp, rtol=2e-3)
    res = gcam.resolve(0).cpu().numpy()
    assert np.isfinite(res).all()
