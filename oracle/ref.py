"""ctypes binding of oracle/_ref/libref.so — the REFERENCE'S OWN sources compiled behind shims
(oracle/ref.mk).  TEST INFRASTRUCTURE ONLY: pins the oracle and generates tests/golden/.

Same calling conventions as oracle/orc.py.  Differences dictated by the reference's interface:
filter_pixel derives the inverse sample density from the number of samples it is handed per pixel
(lentil_filter.cpp:79-87), so samples of one pixel must be contiguous and spp must be a perfect
square >= 9; `tries` is not observable; the forward retry RNG is the process-global xor128.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from pota_b200 import abi

from . import orc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref.so")
_LIB = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _declare_scenario_api(L):
    """argtypes of the entry points both builds of ref_harness.cpp export (the reference's nodes / the adaptor's nodes)"""
    vp = C.c_void_p
    L.ref_camera_create.argtypes = [C.POINTER(abi.CameraParams), C.POINTER(abi.BokehImage), C.POINTER(vp)]
    L.ref_camera_destroy.argtypes = [vp]
    L.ref_camera_get_state.argtypes = [vp, C.POINTER(abi.CameraState)]
    L.ref_camera_set_state.argtypes = [vp, C.c_double, C.c_double]
    L.ref_camera_set_pupil_geometry.argtypes = [vp, C.c_int, C.c_int]
    L.ref_camera_create_rays.argtypes = [vp, C.c_size_t, C.c_uint64, C.POINTER(abi.RayIn), C.POINTER(abi.RayOut), C.c_int]
    L.ref_filter_begin.argtypes = [vp, C.POINTER(abi.FrameDesc), C.c_int, C.POINTER(abi.AovDesc), C.c_int]
    L.ref_filter_accumulate.argtypes = [vp, C.POINTER(abi.Samples), C.c_int]
    L.ref_imager_resolve.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.ref_filter_buffers.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.ref_filter_crypto.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.ref_camera_use_operator.argtypes = [vp, C.c_int]
    L.ref_node_loader.argtypes = [C.c_char_p, C.c_size_t]
    L.ref_node_interface.argtypes = [C.c_char_p, C.c_size_t]
    L.ref_operator_cook.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    return L


def node_loader(L) -> list[str]:
    """what NodeLoader(i, ...) answers for i = 0, 1, ... (lentil_loader.cpp:20-28), one line per node"""
    buf = C.create_string_buffer(1 << 14)
    L.ref_node_loader(buf, len(buf))
    return buf.value.decode().splitlines()


def node_interface(L) -> list[str]:
    """what every node of the library declares to the renderer: parameters (type, name, default, enum strings), metadata, the
    filter's required AOVs / width / output types, the imager's render hints (ref_node_interface in oracle/ref_harness.cpp)"""
    buf = C.create_string_buffer(1 << 16)
    rc = L.ref_node_interface(buf, len(buf))
    if rc != 0:
        raise RuntimeError(f"ref_node_interface: {rc}")
    return buf.value.decode().splitlines()


def operator_cook(L, scene: str, cooks: int = 1) -> list[str]:
    """operator_init + operator_cook of the library's lentil_operator node on the scene (see ref_operator_cook in
    oracle/ref_harness.cpp for the scene lines) -> everything the cook leaves behind, as text lines"""
    buf = C.create_string_buffer(1 << 18)
    rc = L.ref_operator_cook(scene.encode(), cooks, buf, len(buf))
    if rc != 0:
        raise RuntimeError(f"ref_operator_cook: {rc}")
    return buf.value.decode().splitlines()


ADAPTOR_LIB_PATH = os.path.join(os.path.dirname(HERE), "adaptor", "_build", "libadaptor.so")
_ADAPTOR = None


def adaptor_available() -> bool:
    return os.path.exists(ADAPTOR_LIB_PATH)


def adaptor_lib():
    """adaptor/_build/libadaptor.so: the same harness (ref_harness.cpp -DLB_ADAPTOR) over the adaptor's Arnold nodes, which call
    liblentil_b200.so -- needs a GPU."""
    global _ADAPTOR
    if _ADAPTOR is None:
        _ADAPTOR = _declare_scenario_api(C.CDLL(ADAPTOR_LIB_PATH))
    return _ADAPTOR


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ref_camera_create.argtypes = [C.POINTER(abi.CameraParams), C.POINTER(abi.BokehImage), C.POINTER(vp)]
        L.ref_camera_destroy.argtypes = [vp]
        L.ref_camera_get_state.argtypes = [vp, C.POINTER(abi.CameraState)]
        L.ref_camera_set_state.argtypes = [vp, C.c_double, C.c_double]
        L.ref_camera_set_pupil_geometry.argtypes = [vp, C.c_int, C.c_int]
        L.ref_camera_create_rays.argtypes = [vp, C.c_size_t, C.c_uint64, C.POINTER(abi.RayIn), C.POINTER(abi.RayOut), C.c_int]
        L.ref_filter_begin.argtypes = [vp, C.POINTER(abi.FrameDesc), C.c_int, C.POINTER(abi.AovDesc), C.c_int]
        L.ref_filter_accumulate.argtypes = [vp, C.POINTER(abi.Samples), C.c_int]
        L.ref_imager_resolve.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.ref_filter_buffers.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
        L.ref_filter_crypto.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
        L.ref_tea8.restype = C.c_uint
        L.ref_tea8.argtypes = [C.c_uint, C.c_uint]
        L.ref_rng.restype = C.c_float
        L.ref_rng.argtypes = [C.POINTER(C.c_uint)]
        L.ref_fast_sin.restype = C.c_float
        L.ref_fast_sin.argtypes = [C.c_float]
        L.ref_fast_cos.restype = C.c_float
        L.ref_fast_cos.argtypes = [C.c_float]
        L.ref_lens_ipow.restype = C.c_double
        L.ref_lens_ipow.argtypes = [C.c_double, C.c_int]
        L.ref_get_coc_thinlens.restype = C.c_float
        L.ref_get_coc_thinlens.argtypes = [vp, C.c_float]
        L.ref_lens_evaluate.restype = C.c_double
        L.ref_lens_evaluate.argtypes = [vp, vp, vp]
        L.ref_lens_pt_sample_aperture.argtypes = [vp, vp, vp, C.c_double]
        L.ref_lens_lt_sample_aperture.restype = C.c_double
        L.ref_lens_lt_sample_aperture.argtypes = [vp, vp, vp, vp, vp, C.c_double]
        L.ref_trace_ray_bw_po.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp]
        L.ref_bokeh_sample.argtypes = [vp, C.c_float, C.c_float, vp]
        _declare_scenario_api(L)
        _LIB = L
    return _LIB


class RefCamera(orc.OracleCamera):
    """struct Camera of the reference + its camera/filter/imager node callbacks."""

    _L = staticmethod(lambda: lib())

    def __init__(self, params: abi.CameraParams, bokeh: np.ndarray | None = None):
        self._h = C.c_void_p()
        img, self._keep = orc.bokeh_image(bokeh)
        rc = self._L().ref_camera_create(C.byref(params), C.byref(img) if img is not None else None, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"ref_camera_create failed: {rc}")
        self.params = params

    def close(self):
        if self._h:
            self._L().ref_camera_destroy(self._h)
            self._h = C.c_void_p()

    @property
    def state(self) -> abi.CameraState:
        s = abi.CameraState()
        self._L().ref_camera_get_state(self._h, C.byref(s))
        return s

    def set_state(self, aperture_radius, sensor_shift):
        self._L().ref_camera_set_state(self._h, aperture_radius, sensor_shift)

    def set_pupil_geometry(self, outer: int, inner: int = 0):
        self._L().ref_camera_set_pupil_geometry(self._h, outer, inner)

    def create_rays(self, sx, sy, dsx, dsy, lensx, lensy, ray_id_base: int = 0, nthreads: int = 1):
        n = sx.shape[0]
        ins = [np.ascontiguousarray(a, dtype=np.float32) for a in (sx, sy, dsx, dsy, lensx, lensy)]
        rin = abi.RayIn(*[orc._ptr(a) for a in ins])
        out = {k: np.zeros((3, n), np.float32) for k in orc.RAY_OUT_FIELDS}
        out["tries"] = np.zeros(n, np.int32)
        rout = abi.RayOut(*[orc._ptr(out[k]) for k in orc.RAY_OUT_FIELDS], orc._ptr(out["tries"]))
        rc = self._L().ref_camera_create_rays(self._h, n, ray_id_base, C.byref(rin), C.byref(rout), nthreads)
        assert rc == 0, rc
        return out

    def counters(self):
        raise NotImplementedError("the reference keeps no counters")

    def use_operator(self, on: bool = True):
        """from the next filter_begin on, the AOV list is what the library's own lentil_operator node cooks from the scene's
        outputs (lentil_operator.cpp:25-171): the given AOVs, then lentil_debug, lentil_time, lentil_raydir as AOVs n, n+1, n+2"""
        self._L().ref_camera_use_operator(self._h, int(on))

    def filter_begin(self, xres, yres, aovs, xres_full=None, yres_full=None, region_min=(0, 0), spp=9):
        aa = int(round(math.sqrt(spp)))
        assert aa * aa == spp and aa >= 3, "the reference only redistributes at AA >= 3 with AA^2 samples per pixel"
        self._frame = abi.FrameDesc(xres, yres, xres_full or xres, yres_full or yres, region_min[0], region_min[1])
        arr = (abi.AovDesc * len(aovs))()
        for i, (name, flt, role) in enumerate(aovs):
            arr[i].name = name.encode()
            arr[i].filter = flt
            arr[i].role = role
        self._naov = len(aovs)
        rc = self._L().ref_filter_begin(self._h, C.byref(self._frame), len(aovs), arr, aa)
        assert rc == 0, rc

    def filter_accumulate(self, px, py, rgba, pos_cs, inv_density, aov_values=None, raydir=None, transmission=None, flags=None, nthreads=1, crypto=None,
                          world_to_camera=None):
        S, _keep = abi.host_samples(self._naov, px, py, rgba, pos_cs, inv_density, aov_values, raydir, transmission, flags, crypto, world_to_camera)
        rc = self._L().ref_filter_accumulate(self._h, C.byref(S), nthreads)
        assert rc == 0, rc

    def filter_stats(self):
        raise NotImplementedError("the reference keeps no counters")

    def resolve(self, aov, x0=None, y0=None, w=None, h=None, fill=0.0):
        f = self._frame
        x0 = f.region_min_x if x0 is None else x0
        y0 = f.region_min_y if y0 is None else y0
        w = f.xres if w is None else w
        h = f.yres if h is None else h
        out = np.full((h, w, 4), fill, np.float32)
        rc = self._L().ref_imager_resolve(self._h, aov, x0, y0, w, h, orc._ptr(out))
        assert rc == 0, rc
        return out

    def crypto(self, aov, slots=16):
        f = self._frame
        ids = np.zeros((f.yres, f.xres, slots), np.float32)
        wts = np.zeros((f.yres, f.xres, slots), np.float32)
        tot = np.zeros((f.yres, f.xres), np.float32)
        mx = self._L().ref_filter_crypto(self._h, aov, slots, orc._ptr(ids), orc._ptr(wts), orc._ptr(tot))
        assert mx >= 0, mx
        return ids, wts, tot, mx

    def buffers(self, aov):
        b, w = C.c_void_p(), C.c_void_p()
        rc = self._L().ref_filter_buffers(self._h, aov, C.byref(b), C.byref(w))
        assert rc == 0, rc
        f = self._frame
        buf = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=(f.yres, f.xres, 4)).copy()
        wgt = np.ctypeslib.as_array(C.cast(w, C.POINTER(C.c_float)), shape=(f.yres, f.xres)).copy()
        return buf, wgt


class AdaptorCamera(RefCamera):
    """The adaptor's Arnold nodes (adaptor/lentil_b200_*.cpp over liblentil_b200.so) behind the same harness entry points."""

    _L = staticmethod(lambda: adaptor_lib())
