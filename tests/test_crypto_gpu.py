"""GPU parity of the cryptomatte redistribution against the CPU oracle, through the C ABI:
cryptomatte_construct_cache + add_to_buffer_cryptomatte (lentil.h:779-819) into fixed-slot per-pixel tables, and the
ranked resolve of lentil_imager.cpp:122-161.

Bounds: which ids a pixel holds and the row-ending `break` are integer facts and compared exactly wherever the
splat positions agree (the thin-lens model: everywhere; polynomial optics: FP32 vs FP64 Newton moves a small
fraction of splats across a pixel edge, as in test_filter_gpu.py); weights by relative L1 <= 2e-3 per id plane.
The resolve kernel itself is checked bit for bit against a numpy ranking of the GPU's own tables.
"""
import numpy as np
import pytest

from pota_b200 import abi, workloads
from tests.util import po_params

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

AOVS = [("RGBA", 0, 1), ("crypto_material00", 2, 0), ("crypto_material01", 2, 0), ("crypto_object02", 2, 0), ("crypto_asset00", 2, 0)]
FREE = np.uint32(0xFFFFFFFF)


def _setup(params, W=96, H=54, spp=9, depth=4, slots=0, host_path=False, aovs=AOVS):
    from oracle import orc
    from pota_b200.camera import Camera

    ocam = orc.OracleCamera(params)
    gcam = Camera(params, None, device=0)
    fr = workloads.highlight_frame(W, H, spp, ocam.state.tan_fov, "cpu")
    crypto_aovs = [k for k, a in enumerate(aovs) if a[1] == abi.LB_FILTER_CRYPTO]
    cr = workloads.crypto_layers(fr, depth, crypto_aovs)
    cnp = dict(depth=depth, count=cr["count"].numpy(), opacity=cr["opacity"].numpy(), ids={a: v.numpy() for a, v in cr["ids"].items()})
    args = (fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    ocam.filter_begin(W, H, aovs)
    ocam.filter_accumulate(*args, crypto=cnp)
    gcam.filter_begin(W, H, aovs, crypto_slots=slots)
    if host_path:
        gcam.filter_accumulate_host(*args, crypto=cnp)
    else:
        cg = dict(depth=depth, count=cr["count"].cuda(), opacity=cr["opacity"].cuda(), ids={a: v.cuda() for a, v in cr["ids"].items()})
        gcam.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), fr["pos_cs"].cuda(), 1.0 / spp, crypto=cg)
    torch.cuda.synchronize()
    return ocam, gcam, cr


def _planes(ids, wts, palette):
    """tables -> dense weight plane per palette id, and the per-pixel id count."""
    bits = ids.view(np.uint32)
    out = np.zeros((len(palette),) + ids.shape[:2], np.float64)
    for k, p in enumerate(palette):
        out[k] = np.where(bits == np.float32(p).view(np.uint32), wts, 0).sum(axis=2)
    return out, (bits != FREE).sum(axis=2)


def _rank_numpy(ids, wts, total, rank, fill):
    """lentil_imager.cpp:122-161 on tables: weight descending, ties in ascending id; rows end at the first pixel with
    <= rank ids."""
    H, W, K = ids.shape
    out = np.full((H, W, 4), fill, np.float32)
    bits = ids.view(np.uint32)
    for j in range(H):
        for i in range(W):
            used = bits[j, i] != FREE
            if used.sum() <= rank:
                break
            e = sorted(zip(ids[j, i][used].tolist(), wts[j, i][used].tolist()), key=lambda kv: (-kv[1], kv[0]))
            o = np.zeros(4, np.float32)
            o[0], o[1] = np.float32(e[rank][0]), np.float32(e[rank][1]) / total[j, i]
            if len(e) > rank + 1:
                o[2], o[3] = np.float32(e[rank + 1][0]), np.float32(e[rank + 1][1]) / total[j, i]
            out[j, i] = o
    return out


def _compare(ocam, gcam, cr, exact_sets):
    so, sg = ocam.filter_stats(), gcam.filter_stats()
    assert so["redistributed"] == sg["redistributed"] > 50 and so["passthrough"] == sg["passthrough"]
    assert sg["crypto_dropped"] == 0
    for a in range(1, len(AOVS)):
        palette = np.unique(cr["ids"][a].numpy())
        palette = np.union1d(palette, np.float32([0.0]))  # samples without depth sub-samples land on id 0 (lentil.h:787,804)
        io, wo, to, mo = ocam.crypto(a, 32)
        ig, wg, tg = gcam.crypto(a)
        assert mo <= ig.shape[2]
        po, no = _planes(io, wo, palette)
        pg, ng = _planes(ig, wg, palette)
        # every weight sits under one of the palette ids, on both sides
        np.testing.assert_allclose(pg.sum(), wg.sum(dtype=np.float64), rtol=1e-6)
        np.testing.assert_allclose(po.sum(), wo.sum(dtype=np.float64), rtol=1e-6)
        same = (no == ng).mean()
        assert same == 1.0 if exact_sets else same > 0.98, ("pixels with the same id count", same)
        for k in range(len(palette)):
            if po[k].sum() > 0:
                err = np.abs(pg[k] - po[k]).sum() / po[k].sum()
                assert err <= (1e-5 if exact_sets else 2e-3), (AOVS[a][0], palette[k], err)
        np.testing.assert_allclose(tg.sum(dtype=np.float64), to.sum(dtype=np.float64), rtol=1e-5)
        assert np.abs(tg - to).sum() / to.sum() <= (1e-5 if exact_sets else 2e-3)
        # the resolve kernel against a numpy ranking of the same tables: bit for bit
        rank = {"00": 0, "01": 2, "02": 4}[AOVS[a][0][-2:]]
        for box in (dict(), dict(x0=16, y0=8, w=40, h=24)):
            rg = gcam.resolve(a, fill=-7.0, **box).cpu().numpy()
            x0, y0 = box.get("x0", 0), box.get("y0", 0)
            sl = (slice(y0, y0 + rg.shape[0]), slice(x0, x0 + rg.shape[1]))
            want = _rank_numpy(ig[sl], wg[sl], tg[sl], rank, -7.0)
            np.testing.assert_array_equal(rg.view(np.uint32), want.view(np.uint32), err_msg=f"resolve {AOVS[a][0]} {box}")
        # ... and against the oracle's resolve: same untouched tails, same ranked ids (two ids whose weights tie to
        # within the accumulation-order error may swap ranks: a handful of pixels)
        ro, rg = ocam.resolve(a, fill=-7.0), gcam.resolve(a, fill=-7.0).cpu().numpy()
        if exact_sets:
            np.testing.assert_array_equal(ro[..., 1] == -7.0, rg[..., 1] == -7.0)
        agree = (ro[..., 0] == rg[..., 0]) & (np.abs(ro[..., 1] - rg[..., 1]) <= 2e-3)
        assert agree.mean() > (0.999 if exact_sets else 0.97), agree.mean()


@pytest.mark.parametrize("kw", [dict(), dict(abb_chromatic=0.4)])
def test_cryptomatte_polynomial_optics(kw, kernel_kind):
    ocam, gcam, cr = _setup(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, **kw))
    _compare(ocam, gcam, cr, exact_sets=False)


def test_cryptomatte_thinlens():
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    ocam, gcam, cr = _setup(p)
    _compare(ocam, gcam, cr, exact_sets=True)


def test_cryptomatte_host_path_and_no_depth_data():
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    ocam, gcam, cr = _setup(p, host_path=True)
    _compare(ocam, gcam, cr, exact_sets=True)
    # a batch without any depth data: every sample contributes {0.0: 1.0} (lentil.h:787,804)
    from pota_b200.camera import Camera
    g2 = Camera(p, None, device=0)
    W, H, spp = 48, 27, 9
    fr = workloads.highlight_frame(W, H, spp, g2.state.tan_fov, "cpu")
    g2.filter_begin(W, H, AOVS[:2])
    g2.filter_accumulate_host(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp)
    ids, wts, tot = g2.crypto(1)
    used = ids.view(np.uint32) != FREE
    assert np.all(used.sum(axis=2) == 1) and np.all(ids[used] == 0.0)
    np.testing.assert_allclose(wts.sum(axis=2), tot, rtol=1e-5)
    res = g2.resolve(1).cpu().numpy()
    np.testing.assert_allclose(res[..., 1], 1.0, rtol=1e-5)
    assert np.all(res[..., 0] == 0.0) and np.all(res[..., 2:] == 0.0)


def test_cryptomatte_slot_overflow_is_counted():
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    ocam, gcam, cr = _setup(p, slots=2)
    st = gcam.filter_stats()
    assert st["crypto_dropped"] > 0
    ids, wts, tot = gcam.crypto(1)
    assert ids.shape[2] == 2
    # what was kept is still a subset of the oracle's tables with the oracle's weights
    io, wo, to, mo = ocam.crypto(1, 32)
    assert mo > 2
    palette = np.union1d(np.unique(cr["ids"][1].numpy()), np.float32([0.0]))
    po, _ = _planes(io, wo, palette)
    pg, _ = _planes(ig := ids, wts, palette)
    kept = pg > 0
    np.testing.assert_allclose(pg[kept], po[kept], rtol=1e-3, atol=1e-7)
    np.testing.assert_allclose(tot, to, rtol=1e-4, atol=1e-7)  # crypto_total_weight counts every splat regardless


def test_cryptomatte_rejects_bad_arguments():
    from pota_b200.camera import Camera, LentilError
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0)
    g = Camera(p, None, device=0)
    with pytest.raises(LentilError):
        g.filter_begin(16, 16, [("RGBA", 0, 1), ("matte", 2, 0)])  # not a crypto_ name
    with pytest.raises(LentilError):
        g.filter_begin(16, 16, AOVS[:2], crypto_slots=1000)
    g.filter_begin(16, 16, AOVS[:2])
    n = 16
    z = torch.zeros((n, 4), device="cuda")
    zi = torch.zeros(n, dtype=torch.int32, device="cuda")
    with pytest.raises(LentilError):
        g.filter_accumulate(zi, zi, z, z, 1.0, crypto=dict(depth=9, opacity=torch.zeros((n, 9), device="cuda"), ids={1: torch.zeros((n, 9), device="cuda")}))


def test_cryptomatte_render_region():
    """Tables and ranked buckets inside a render region: bucket coordinates are frame coordinates (lentil_imager.cpp:116-120)."""
    from oracle import orc
    from pota_b200.camera import Camera

    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    ocam, gcam = orc.OracleCamera(p), Camera(p, None, device=0)
    Wf, Hf, spp = 96, 54, 9
    x0, y0, W, H = 24, 12, 48, 30
    fr = workloads.highlight_frame(Wf, Hf, spp, ocam.state.tan_fov, "cpu")
    cr = workloads.crypto_layers(fr, 3, [1, 2])
    px, py = fr["px"].numpy(), fr["py"].numpy()
    m = (px >= x0) & (px < x0 + W) & (py >= y0) & (py < y0 + H)
    crypto = dict(depth=3, count=cr["count"].numpy()[m], opacity=cr["opacity"].numpy()[m], ids={a: v.numpy()[m] for a, v in cr["ids"].items()})
    args = (px[m] - x0, py[m] - y0, fr["rgba"].numpy()[m], fr["pos_cs"].numpy()[m], 1.0 / spp)
    aovs = [("RGBA", 0, 1), ("crypto_object00", 2, 0), ("crypto_object01", 2, 0)]
    ocam.filter_begin(W, H, aovs, xres_full=Wf, yres_full=Hf, region_min=(x0, y0))
    ocam.filter_accumulate(*args, crypto=crypto)
    gcam.filter_begin(W, H, aovs, xres_full=Wf, yres_full=Hf, region_min=(x0, y0))
    gcam.filter_accumulate_host(*args, crypto=crypto)
    for a in (1, 2):
        palette = np.union1d(np.unique(cr["ids"][a].numpy()), np.float32([0.0]))
        io, wo, to, _ = ocam.crypto(a, 32)
        ig, wg, tg = gcam.crypto(a)
        po, no = _planes(io, wo, palette)
        pg, ng = _planes(ig, wg, palette)
        np.testing.assert_array_equal(no, ng)
        np.testing.assert_allclose(pg, po, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(tg, to, rtol=1e-5, atol=1e-7)
        for box in (dict(), dict(x0=x0 + 8, y0=y0 + 4, w=20, h=10)):
            ro, rg = ocam.resolve(a, fill=-7.0, **box), gcam.resolve(a, fill=-7.0, **box).cpu().numpy()
            np.testing.assert_array_equal(ro[..., 1] == -7.0, rg[..., 1] == -7.0)
            agree = (ro[..., 0] == rg[..., 0]) & (np.abs(ro[..., 1] - rg[..., 1]) <= 2e-3)
            assert agree.mean() > 0.995, agree.mean()
