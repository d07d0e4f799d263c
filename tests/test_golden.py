"""Golden vectors generated from the reference's own sources (tests/golden/make_golden.py, oracle/_ref).

CPU: the oracle must reproduce them exactly.  GPU (-m gpu): the CUDA path is compared with the same
files, through the C ABI, at the tolerances of test_camera_gpu.py / test_filter_gpu.py.
"""
import glob
import os

import numpy as np
import pytest

from oracle import orc
from pota_b200 import abi, workloads
from tests.util import rel_err_vec

GOLD = os.path.join(os.path.dirname(__file__), "golden")
IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")
RAY_FILES = sorted(glob.glob(os.path.join(GOLD, "rays_*.npz")))
FILTER_FILES = sorted(glob.glob(os.path.join(GOLD, "filter_*.npz")))
CRYPTO_FILES = sorted(glob.glob(os.path.join(GOLD, "crypto_*.npz")))
SAMPLEDATA_FILES = sorted(glob.glob(os.path.join(GOLD, "sampledata_*.npz")))


def _params(g):
    return abi.CameraParams.from_buffer_copy(g["params"].tobytes())


def _bokeh(p):
    return workloads.disc_bokeh_image(64) if p.bokeh_enable_image else None


def _frame(g, tan_fov):
    W, H, spp, n_extra = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["n_extra"])
    aovs = [(str(n), int(f), int(r)) for n, f, r in zip(g["aov_names"], g["aov_filter"], g["aov_role"])]
    fr = workloads.highlight_frame(W, H, spp, tan_fov, "cpu", n_extra_aov=n_extra)
    vals = ([None] + list(fr["aov_values"][: len(aovs) - 1]) + [None] * len(aovs))[: len(aovs)]
    return W, H, spp, aovs, fr, vals


def test_fixtures_present():
    assert len(RAY_FILES) >= 3 and len(FILTER_FILES) >= 2 and len(CRYPTO_FILES) >= 2 and len(SAMPLEDATA_FILES) >= 2


@pytest.mark.parametrize("path", RAY_FILES, ids=os.path.basename)
def test_oracle_reproduces_reference_rays(path):
    g = np.load(path)
    p = _params(g)
    cam = orc.OracleCamera(p, _bokeh(p))
    s = cam.state
    assert s.aperture_radius == float(g["aperture_radius"]) and s.sensor_shift == float(g["sensor_shift"]) and s.tan_fov == float(g["tan_fov"])
    out = cam.create_rays(*[g[k] for k in IN_KEYS])
    m = g["first_try"]
    np.testing.assert_array_equal(out["tries"] == 0, m)
    for k in orc.RAY_OUT_FIELDS:
        np.testing.assert_array_equal(out[k][:, m], g[k][:, m], err_msg=k)


@pytest.mark.parametrize("path", FILTER_FILES, ids=os.path.basename)
def test_oracle_reproduces_reference_framebuffers(path):
    g = np.load(path)
    p = _params(g)
    cam = orc.OracleCamera(p, _bokeh(p))
    W, H, spp, aovs, fr, vals = _frame(g, cam.state.tan_fov)
    cam.filter_begin(W, H, aovs)
    cam.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp,
                          aov_values=[None if v is None else v.numpy() for v in vals])
    for a in range(len(aovs)):
        buf, wgt = cam.buffers(a)
        np.testing.assert_array_equal(buf, g[f"buffer{a}"])
        np.testing.assert_array_equal(wgt, g["weight"])
        np.testing.assert_array_equal(cam.resolve(a), g[f"resolved{a}"])


@pytest.mark.parametrize("path", SAMPLEDATA_FILES, ids=os.path.basename)
def test_oracle_reproduces_reference_on_sampledata_rows(path):
    """Source samples = literal rows of the reference's tests/cuda/sampledata.txt (its CUDA prototype's input, the one
    fixture of real renderer samples it ships); framebuffers = the compiled reference's output for them."""
    g = np.load(path)
    cam = orc.OracleCamera(_params(g))
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    cam.filter_begin(W, H, [("RGBA", 0, 1)])
    cam.filter_accumulate(g["px"], g["py"], g["rgba"], g["pos_cs"], 1.0 / spp)
    st = cam.filter_stats()
    assert st["redistributed"] == g["px"].shape[0] and st["splats"] > 20 * st["redistributed"]  # every row is a highlight out of focus
    buf, wgt = cam.buffers(0)
    np.testing.assert_array_equal(buf, g["buffer0"])
    np.testing.assert_array_equal(wgt, g["weight"])
    np.testing.assert_array_equal(cam.resolve(0), g["resolved0"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", SAMPLEDATA_FILES, ids=os.path.basename)
def test_gpu_against_reference_on_sampledata_rows(path, kernel_kind):
    import torch

    from pota_b200.camera import Camera

    g = np.load(path)
    p = _params(g)
    cam = Camera(p, None, device=0)
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    cam.filter_begin(W, H, [("RGBA", 0, 1)])
    cam.filter_accumulate(*[torch.from_numpy(g[k]).cuda() for k in ("px", "py", "rgba", "pos_cs")], 1.0 / spp)
    torch.cuda.synchronize()
    st = cam.filter_stats()
    assert st["redistributed"] == g["px"].shape[0]
    buf, wgt = cam.buffers(0)
    exact = p.camera_type == abi.LB_CAMERA_THINLENS
    l1 = np.abs(buf - g["buffer0"]).sum() / np.abs(g["buffer0"]).sum()
    assert l1 <= (1e-3 if exact else 5e-3), l1
    np.testing.assert_allclose(buf.sum(dtype=np.float64), g["buffer0"].sum(dtype=np.float64), rtol=2e-3)
    np.testing.assert_allclose(wgt.sum(dtype=np.float64), g["weight"].sum(dtype=np.float64), rtol=2e-3)
    res = cam.resolve(0).cpu().numpy()
    peak = float(np.abs(g["resolved0"]).max())
    mse = float(((res.astype(np.float64) - g["resolved0"].astype(np.float64)) ** 2).mean())
    assert mse == 0 or 10.0 * np.log10(peak * peak / mse) >= 50.0


def _crypto_inputs(g, tan_fov):
    W, H, spp, depth = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["depth"])
    aovs = [(str(n), 0 if str(n) == "RGBA" else 2, 1 if str(n) == "RGBA" else 0) for n in g["aov_names"]]
    fr = workloads.highlight_frame(W, H, spp, tan_fov, "cpu")
    cr = workloads.crypto_layers(fr, depth, [a for a in range(len(aovs)) if aovs[a][1] == 2])
    return W, H, spp, aovs, fr, cr


@pytest.mark.parametrize("path", CRYPTO_FILES, ids=os.path.basename)
def test_oracle_reproduces_reference_cryptomatte(path):
    g = np.load(path)
    cam = orc.OracleCamera(_params(g))
    W, H, spp, aovs, fr, cr = _crypto_inputs(g, cam.state.tan_fov)
    cam.filter_begin(W, H, aovs)
    cam.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp,
                          crypto=dict(depth=cr["depth"], count=cr["count"].numpy(), opacity=cr["opacity"].numpy(), ids={a: v.numpy() for a, v in cr["ids"].items()}))
    for a in range(1, len(aovs)):
        ids, wts, tot, mx = cam.crypto(a, g[f"ids{a}"].shape[2])
        np.testing.assert_array_equal(ids.view(np.uint32), g[f"ids{a}"].view(np.uint32))
        np.testing.assert_array_equal(wts, g[f"weights{a}"])
        np.testing.assert_array_equal(tot, g[f"total{a}"])
        np.testing.assert_array_equal(cam.resolve(a, fill=-7.0).view(np.uint32), g[f"resolved{a}"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("path", CRYPTO_FILES, ids=os.path.basename)
def test_gpu_cryptomatte_against_reference_golden(path, kernel_kind):
    import torch

    from pota_b200.camera import Camera
    from tests.test_crypto_gpu import _planes

    g = np.load(path)
    p = _params(g)
    cam = Camera(p, None, device=0)
    W, H, spp, aovs, fr, cr = _crypto_inputs(g, cam.state.tan_fov)
    cam.filter_begin(W, H, aovs)
    cam.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), fr["pos_cs"].cuda(), 1.0 / spp,
                          crypto=dict(depth=cr["depth"], count=cr["count"].cuda(), opacity=cr["opacity"].cuda(), ids={a: v.cuda() for a, v in cr["ids"].items()}))
    torch.cuda.synchronize()
    assert cam.filter_stats()["crypto_dropped"] == 0
    exact = p.camera_type == abi.LB_CAMERA_THINLENS  # FP32 thin-lens splat positions are the reference's own
    for a in range(1, len(aovs)):
        palette = np.union1d(np.unique(cr["ids"][a].numpy()), np.float32([0.0]))
        ids, wts, tot = cam.crypto(a)
        pg, ng = _planes(ids, wts, palette)
        po, no = _planes(g[f"ids{a}"], g[f"weights{a}"], palette)
        assert (ng == no).mean() >= (1.0 if exact else 0.98)
        for k in range(len(palette)):
            if po[k].sum() > 0:
                assert np.abs(pg[k] - po[k]).sum() / po[k].sum() <= (1e-5 if exact else 5e-3), (aovs[a][0], palette[k])
        np.testing.assert_allclose(tot.sum(dtype=np.float64), g[f"total{a}"].sum(dtype=np.float64), rtol=1e-5)
        ro, rg = g[f"resolved{a}"], cam.resolve(a, fill=-7.0).cpu().numpy()
        agree = (ro[..., 0] == rg[..., 0]) & (np.abs(ro[..., 1] - rg[..., 1]) <= 2e-3)
        assert agree.mean() > (0.999 if exact else 0.97), agree.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("path", RAY_FILES, ids=os.path.basename)
def test_gpu_rays_against_reference_golden(path, kernel_kind):
    import torch

    from pota_b200.camera import Camera

    g = np.load(path)
    p = _params(g)
    cam = Camera(p, _bokeh(p), device=0)
    s = cam.state
    assert s.aperture_radius == float(g["aperture_radius"]) and s.sensor_shift == float(g["sensor_shift"]) and s.tan_fov == float(g["tan_fov"])
    out = cam.create_rays(*[torch.from_numpy(g[k]).cuda() for k in IN_KEYS])
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    m = g["first_try"] & (got["tries"] == 0)
    assert m.sum() >= 0.999 * g["first_try"].sum()
    np.testing.assert_array_equal(got["weight"][:, m], g["weight"][:, m])
    for k in ("origin", "dir"):
        e = rel_err_vec(got[k][:, m], g[k][:, m])
        assert (e <= 1e-4).mean() >= 0.999, (k, e.max())
    for k, base in (("dOdx", "origin"), ("dOdy", "origin"), ("dDdx", "dir"), ("dDdy", "dir")):
        unit = np.spacing(np.abs(g[base][:, m]).max(axis=0).astype(np.float32)) / 1e-3
        d = np.abs(got[k][:, m] - g[k][:, m]).max(axis=0) / unit
        assert np.median(d) <= 1.0 and np.quantile(d, 0.99) <= 4.0, (k, np.median(d), np.quantile(d, 0.99))


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILTER_FILES, ids=os.path.basename)
def test_gpu_framebuffers_against_reference_golden(path, kernel_kind):
    import torch

    from pota_b200.camera import Camera

    g = np.load(path)
    p = _params(g)
    cam = Camera(p, _bokeh(p), device=0)
    W, H, spp, aovs, fr, vals = _frame(g, cam.state.tan_fov)
    cam.filter_begin(W, H, aovs)
    cam.filter_accumulate(fr["px"].cuda(), fr["py"].cuda(), fr["rgba"].cuda(), fr["pos_cs"].cuda(), 1.0 / spp,
                          aov_values=[None if v is None else v.cuda() for v in vals])
    torch.cuda.synchronize()
    for a, (name, flt, role) in enumerate(aovs):
        buf, wgt = cam.buffers(a)
        ref_buf = g[f"buffer{a}"]
        if flt == abi.LB_FILTER_GAUSSIAN:
            l1 = np.abs(buf - ref_buf).sum() / np.abs(ref_buf).sum()
            assert l1 <= 5e-3, (name, l1)  # small frame, few splats per pixel: single splats crossing a pixel edge weigh more
            np.testing.assert_allclose(buf.sum(dtype=np.float64), ref_buf.sum(dtype=np.float64), rtol=2e-3)
        else:
            assert (np.abs(buf - ref_buf).max(axis=2) > 1e-6).mean() <= 1e-2, name
    np.testing.assert_allclose(wgt.sum(dtype=np.float64), g["weight"].sum(dtype=np.float64), rtol=2e-3)
