"""Generates tests/golden/*.npz from the REFERENCE'S OWN SOURCES (oracle/_ref/libref.so, built by
`make -C oracle -f ref.mk` where /root/reference exists).  The fixtures travel with the repo, so the GPU
box — which has no /root/reference — still checks the oracle and the CUDA path against reference output.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import orc, ref  # noqa: E402
from pota_b200 import abi, workloads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")

RAY_CASES = {
    "rays_takumar50_f2.8": dict(lens_model=5, fstop=2.8, focus_dist=150.0),
    "rays_angenieux49_f1.4_bokeh": dict(lens_model=0, fstop=1.4, focus_dist=50.0, bokeh_enable_image=1),
    "rays_zeiss65_f5.6_blades_mm": dict(lens_model=40, fstop=5.6, focus_dist=500.0, aperture_blades_lentil=7, units=abi.LB_UNITS_MM),
}
FILTER_CASES = {
    "filter_takumar50_multi_aov": (dict(lens_model=5, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6, bidir_add_energy=0.5),
                                   [("RGBA", 0, 1), ("light0", 0, 0), ("light1", 0, 0), ("N", 1, 0)], 3),
    "filter_cooke50_bokeh_chromatic": (dict(lens_model=19, fstop=2.0, focus_dist=40.0, bidir_sample_mult=5, bokeh_enable_image=1, abb_chromatic=0.3),
                                       [("RGBA", 0, 1)], 0),
}
CRYPTO_AOVS = [("RGBA", 0, 1), ("crypto_material00", 2, 0), ("crypto_material01", 2, 0), ("crypto_object02", 2, 0)]
CRYPTO_CASES = {
    "crypto_takumar50_f1.4": dict(camera_type=abi.LB_CAMERA_POLYNOMIAL_OPTICS, lens_model=5, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6),
    "crypto_thinlens50_f1.4": dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6),
}
CRYPTO_DEPTH, CRYPTO_SLOTS = 4, 16
# the reference's own fixture of real source samples (input only): /root/reference/tests/cuda/sampledata.txt, rows
# "r g b a depth x_cs y_cs z_cs" written by its CUDA prototype run of tests/cuda/lightgrid.ass (SURVEY.md §8c fixture 1)
SAMPLEDATA = "/root/reference/tests/cuda/sampledata.txt"
SAMPLEDATA_ROWS = 9 * 128
SAMPLEDATA_CASES = {
    "sampledata_thinlens50_f1.4": dict(camera_type=abi.LB_CAMERA_THINLENS, focal_length_lentil=50.0, fstop=1.4, focus_dist=35.0, bidir_sample_mult=4),
    "sampledata_takumar50_f1.4": dict(camera_type=abi.LB_CAMERA_POLYNOMIAL_OPTICS, lens_model=5, fstop=1.4, focus_dist=35.0, bidir_sample_mult=2),
}


def sampledata_frame(rows, tan_fov, W, H, spp):
    """Source samples from sampledata rows: runs of `spp` consecutive rows share the pixel the run's first row
    projects to through a pinhole (filter_pixel needs AA^2 samples per call, lentil_filter.cpp:79-87)."""
    n = rows.shape[0] // spp * spp
    rows = rows[:n]
    rgba = rows[:, 0:4].astype(np.float32)
    pos = np.concatenate([rows[:, 5:8], rows[:, 4:5]], axis=1).astype(np.float32)  # xyz camera space, w = depth
    first = pos[::spp]
    sx = first[:, 0] / (-first[:, 2] * tan_fov)
    sy = first[:, 1] / (-first[:, 2] * tan_fov)
    px = np.clip(np.floor((sx + 1.0) * 0.5 * W), 0, W - 1).astype(np.int32)
    py = np.clip(np.floor((1.0 - sy * (W / H)) * 0.5 * H), 0, H - 1).astype(np.int32)
    for g in range(1, px.shape[0]):  # consecutive runs must differ in pixel, or the iterator would hand them over as one
        if px[g] == px[g - 1] and py[g] == py[g - 1]:
            px[g] = (px[g] + 1) % W
    return np.repeat(px, spp), np.repeat(py, spp), rgba, pos


def params(**kw):
    return abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_POLYNOMIAL_OPTICS, **kw)


def main():
    assert ref.available(), "build oracle/_ref first: make -C oracle -f ref.mk"
    img = workloads.disc_bokeh_image(64)
    n = 2048
    for name, kw in RAY_CASES.items():
        p = params(**kw)
        bok = img if kw.get("bokeh_enable_image") else None
        r, o = ref.RefCamera(p, bok), orc.OracleCamera(p, bok)
        w = int(round((n * 16 / 9) ** 0.5))
        ins = workloads.camera_samples(w, -(-n // w), 1, "cpu", 0, n, "linear")
        arrs = [ins[k].numpy() for k in IN_KEYS]
        out = r.create_rays(*arrs)
        first_try = o.create_rays(*arrs)["tries"] == 0  # the reference's interface does not expose `tries`
        s = r.state
        np.savez_compressed(os.path.join(HERE, name + ".npz"), params=np.frombuffer(bytes(p), np.uint8), first_try=first_try,
                            aperture_radius=s.aperture_radius, sensor_shift=s.sensor_shift, tan_fov=s.tan_fov,
                            **{k: a for k, a in zip(IN_KEYS, arrs)}, **{k: out[k] for k in orc.RAY_OUT_FIELDS})
        print(name, "first-try rays", int(first_try.sum()), "of", n)
    for name, (kw, aovs, n_extra) in FILTER_CASES.items():
        p = params(**kw)
        bok = img if kw.get("bokeh_enable_image") else None
        r = ref.RefCamera(p, bok)
        W, H, spp = 96, 54, 9
        fr = workloads.highlight_frame(W, H, spp, r.state.tan_fov, "cpu", n_extra_aov=n_extra)
        vals = [None] + [v.numpy() for v in fr["aov_values"][: len(aovs) - 1]] + [None] * max(0, len(aovs) - 1 - n_extra)
        vals = vals[: len(aovs)]
        r.filter_begin(W, H, aovs, spp=spp)
        r.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp, aov_values=vals)
        d = dict(params=np.frombuffer(bytes(p), np.uint8), W=W, H=H, spp=spp, n_extra=n_extra,
                 aov_names=np.array([a[0] for a in aovs]), aov_filter=np.array([a[1] for a in aovs]), aov_role=np.array([a[2] for a in aovs]))
        for a in range(len(aovs)):
            buf, wgt = r.buffers(a)
            d[f"buffer{a}"] = buf
            d[f"resolved{a}"] = r.resolve(a)
            d["weight"] = wgt
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "energy", float(d["buffer0"][..., :3].sum()))
    for name, kw in CRYPTO_CASES.items():
        p = abi.CameraParams.defaults(**kw)
        r = ref.RefCamera(p)
        W, H, spp = 96, 54, 9
        fr = workloads.highlight_frame(W, H, spp, r.state.tan_fov, "cpu")
        cr = workloads.crypto_layers(fr, CRYPTO_DEPTH, [1, 2, 3])
        crypto = dict(depth=CRYPTO_DEPTH, count=cr["count"].numpy(), opacity=cr["opacity"].numpy(), ids={a: v.numpy() for a, v in cr["ids"].items()})
        r.filter_begin(W, H, CRYPTO_AOVS, spp=spp)
        r.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / spp, crypto=crypto)
        d = dict(params=np.frombuffer(bytes(p), np.uint8), W=W, H=H, spp=spp, depth=CRYPTO_DEPTH, aov_names=np.array([a[0] for a in CRYPTO_AOVS]))
        for a in (1, 2, 3):
            ids, wts, tot, mx = r.crypto(a, CRYPTO_SLOTS)
            assert mx <= CRYPTO_SLOTS
            d[f"ids{a}"], d[f"weights{a}"], d[f"total{a}"] = ids, wts, tot
            d[f"resolved{a}"] = r.resolve(a, fill=-7.0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "largest id table", mx)
    rows = np.loadtxt(SAMPLEDATA, max_rows=SAMPLEDATA_ROWS)
    for name, kw in SAMPLEDATA_CASES.items():
        p = abi.CameraParams.defaults(**kw)
        r = ref.RefCamera(p)
        W, H, spp = 480, 270, 9
        px, py, rgba, pos = sampledata_frame(rows, r.state.tan_fov, W, H, spp)
        aovs = [("RGBA", 0, 1)]
        r.filter_begin(W, H, aovs, spp=spp)
        r.filter_accumulate(px, py, rgba, pos, 1.0 / spp)
        buf, wgt = r.buffers(0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), params=np.frombuffer(bytes(p), np.uint8), W=W, H=H, spp=spp,
                            px=px, py=py, rgba=rgba, pos_cs=pos, buffer0=buf, weight=wgt, resolved0=r.resolve(0))
        print(name, "rows", rgba.shape[0], "energy", float(buf[..., :3].sum()), "weight", float(wgt.sum()))


if __name__ == "__main__":
    main()
