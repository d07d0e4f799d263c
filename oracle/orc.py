"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from pota_b200/.  All arrays are numpy host arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pota_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(quiet: bool = True):
    subprocess.run(["make", "-s", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_camera_create.argtypes = [C.POINTER(abi.CameraParams), C.POINTER(abi.BokehImage), C.POINTER(C.c_void_p)]
        L.orc_camera_destroy.argtypes = [C.c_void_p]
        L.orc_camera_get_state.argtypes = [C.c_void_p, C.POINTER(abi.CameraState)]
        L.orc_camera_set_state.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_camera_set_pupil_geometry.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_camera_create_rays.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(abi.RayIn), C.POINTER(abi.RayOut), C.c_int]
        L.orc_filter_begin.argtypes = [C.c_void_p, C.POINTER(abi.FrameDesc), C.c_int, C.POINTER(abi.AovDesc)]
        L.orc_filter_accumulate.argtypes = [C.c_void_p, C.POINTER(abi.Samples), C.c_int]
        L.orc_filter_get_stats.argtypes = [C.c_void_p, C.POINTER(abi.FilterStats)]
        L.orc_imager_resolve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_filter_buffers.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orc_filter_crypto.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_camera_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 4)]
        L.orc_tea8.restype = C.c_uint
        L.orc_tea8.argtypes = [C.c_uint, C.c_uint]
        L.orc_rng.restype = C.c_float
        L.orc_rng.argtypes = [C.POINTER(C.c_uint)]
        L.orc_fast_sin.restype = C.c_float
        L.orc_fast_sin.argtypes = [C.c_float]
        L.orc_fast_cos.restype = C.c_float
        L.orc_fast_cos.argtypes = [C.c_float]
        L.orc_lens_ipow.restype = C.c_double
        L.orc_lens_ipow.argtypes = [C.c_double, C.c_int]
        L.orc_get_coc_thinlens.restype = C.c_float
        L.orc_get_coc_thinlens.argtypes = [C.c_void_p, C.c_float]
        L.orc_lens_evaluate.restype = C.c_double
        L.orc_lens_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lens_pt_sample_aperture.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_lens_lt_sample_aperture.restype = C.c_double
        L.orc_lens_lt_sample_aperture.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_trace_ray_bw_po.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.orc_bokeh_sample.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bokeh_image(pixels: np.ndarray | None):
    """pixels: float32 [H, W, C] or None -> (BokehImage | None, keepalive)."""
    if pixels is None:
        return None, None
    px = np.ascontiguousarray(pixels, dtype=np.float32)
    img = abi.BokehImage(px.shape[1], px.shape[0], px.shape[2], px.ctypes.data_as(C.POINTER(C.c_float)))
    return img, px


RAY_OUT_FIELDS = ("origin", "dir", "dOdx", "dOdy", "dDdx", "dDdy", "weight")


class OracleCamera:
    def __init__(self, params: abi.CameraParams, bokeh: np.ndarray | None = None):
        self._h = C.c_void_p()
        img, self._keep = bokeh_image(bokeh)
        rc = lib().orc_camera_create(C.byref(params), C.byref(img) if img is not None else None, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"orc_camera_create failed: {rc}")
        self.params = params

    def close(self):
        if self._h:
            lib().orc_camera_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def state(self) -> abi.CameraState:
        s = abi.CameraState()
        lib().orc_camera_get_state(self._h, C.byref(s))
        return s

    def set_state(self, aperture_radius: float, sensor_shift: float):
        lib().orc_camera_set_state(self._h, aperture_radius, sensor_shift)

    def set_pupil_geometry(self, outer: int, inner: int = 0):
        lib().orc_camera_set_pupil_geometry(self._h, outer, inner)

    def create_rays(self, sx, sy, dsx, dsy, lensx, lensy, ray_id_base: int = 0, nthreads: int = 1):
        n = sx.shape[0]
        ins = [np.ascontiguousarray(a, dtype=np.float32) for a in (sx, sy, dsx, dsy, lensx, lensy)]
        rin = abi.RayIn(*[_ptr(a) for a in ins])
        out = {k: np.zeros((3, n), np.float32) for k in RAY_OUT_FIELDS}
        out["tries"] = np.zeros(n, np.int32)
        rout = abi.RayOut(*[_ptr(out[k]) for k in RAY_OUT_FIELDS], _ptr(out["tries"]))
        rc = lib().orc_camera_create_rays(self._h, n, ray_id_base, C.byref(rin), C.byref(rout), nthreads)
        assert rc == 0, rc
        return out

    def counters(self):
        c = (C.c_uint64 * 4)()
        lib().orc_camera_counters(self._h, C.byref(c))
        return dict(fw_newton_its=c[0], fw_traces=c[1], bw_newton_its=c[2], bw_attempts=c[3])

    # ---- filter / imager ----
    def filter_begin(self, xres, yres, aovs, xres_full=None, yres_full=None, region_min=(0, 0)):
        """aovs: list of (name, filter, role)."""
        self._frame = abi.FrameDesc(xres, yres, xres_full or xres, yres_full or yres, region_min[0], region_min[1])
        arr = (abi.AovDesc * len(aovs))()
        for i, (name, flt, role) in enumerate(aovs):
            arr[i].name = name.encode()
            arr[i].filter = flt
            arr[i].role = role
        self._naov = len(aovs)
        rc = lib().orc_filter_begin(self._h, C.byref(self._frame), len(aovs), arr)
        assert rc == 0, rc

    def filter_accumulate(self, px, py, rgba, pos_cs, inv_density, aov_values=None, raydir=None, transmission=None, flags=None, nthreads=1, crypto=None,
                          world_to_camera=None):
        S, _keep = abi.host_samples(self._naov, px, py, rgba, pos_cs, inv_density, aov_values, raydir, transmission, flags, crypto, world_to_camera)
        rc = lib().orc_filter_accumulate(self._h, C.byref(S), nthreads)
        assert rc == 0, rc

    def filter_stats(self):
        s = abi.FilterStats()
        lib().orc_filter_get_stats(self._h, C.byref(s))
        return {k: getattr(s, k) for k, _ in s._fields_}

    _resolve_fn = "orc_imager_resolve"
    _crypto_fn = "orc_filter_crypto"

    def resolve(self, aov, x0=None, y0=None, w=None, h=None, fill=0.0):
        """fill: what the bucket holds beforehand (cryptomatte rows can end early and leave it, lentil_imager.cpp:132-134)."""
        f = self._frame
        x0 = f.region_min_x if x0 is None else x0
        y0 = f.region_min_y if y0 is None else y0
        w = f.xres if w is None else w
        h = f.yres if h is None else h
        out = np.full((h, w, 4), fill, np.float32)
        rc = getattr(lib(), self._resolve_fn)(self._h, aov, x0, y0, w, h, _ptr(out))
        assert rc == 0, rc
        return out

    def crypto(self, aov, slots=16):
        """(ids [yres][xres][slots], weights, total_weight [yres][xres], largest map size)."""
        f = self._frame
        ids = np.zeros((f.yres, f.xres, slots), np.float32)
        wts = np.zeros((f.yres, f.xres, slots), np.float32)
        tot = np.zeros((f.yres, f.xres), np.float32)
        mx = getattr(lib(), self._crypto_fn)(self._h, aov, slots, _ptr(ids), _ptr(wts), _ptr(tot))
        assert mx >= 0, mx
        return ids, wts, tot, mx

    def buffers(self, aov):
        b, w = C.c_void_p(), C.c_void_p()
        lib().orc_filter_buffers(self._h, aov, C.byref(b), C.byref(w))
        f = self._frame
        buf = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=(f.yres, f.xres, 4)).copy()
        wgt = np.ctypeslib.as_array(C.cast(w, C.POINTER(C.c_float)), shape=(f.yres, f.xres)).copy()
        return buf, wgt
