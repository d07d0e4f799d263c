"""Host-side mirror of the reference's node interface over the C ABI (include/lentil_b200.h).

The product is liblentil_b200.so (C++ host + CUDA kernels).  This module only marshals arguments:
torch owns device memory and streams, ctypes calls the C entry points.  Naming follows the
reference's nodes: `lentil_camera` (camera_create_ray, camera_reverse_ray), `lentil_filter`
(filter_pixel) and `imager_lentil` (driver_process_bucket) — /root/reference/src/lentil_loader.cpp:6-9.

There is no CPU path: constructing a Camera without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("LB_LIBRARY") or os.path.join(_HERE, "liblentil_b200.so")  # LB_LIBRARY: tuning variants
_lib = None

RAY_OUT_FIELDS = ("origin", "dir", "dOdx", "dOdy", "dDdx", "dDdy", "weight")

# every symbol include/lentil_b200.h declares
EXPORTS = [
    "lb_camera_params_default", "lb_lens_count", "lb_lens_name", "lb_last_error", "lb_version",
    "lb_camera_create", "lb_camera_update", "lb_camera_destroy", "lb_camera_get_state", "lb_camera_set_state",
    "lb_camera_create_rays", "lb_camera_create_rays_host", "lb_camera_reverse_rays", "lb_camera_lens_work",
    "lb_filter_begin", "lb_filter_accumulate", "lb_filter_accumulate_host", "lb_filter_get_stats",
    "lb_filter_newton_iterations", "lb_imager_resolve", "lb_imager_resolve_host", "lb_filter_buffers", "lb_filter_buffers_host", "lb_filter_crypto_host",
    "lb_bench_fp32_peak", "lb_bench_red_peak", "lb_bench_splat_accum", "lb_camera_set_pupil_geometry", "lb_camera_kernel_kind", "lb_comm_unique_id", "lb_comm_init", "lb_filter_set_sample_base", "lb_filter_reduce", "lb_comm_destroy",
    "lb_filter_reduce_scatter", "lb_filter_slab", "lb_imager_resolve_gather", "lb_debug_primitives", "lb_imager_resolve_peer", "lb_imager_peer_image",
]  # fmt: skip


class LentilError(RuntimeError):
    pass


def lib():
    """Load liblentil_b200.so; fails loudly if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise LentilError(f"{_LIB_PATH} is missing: run `python -m pota_b200.build` (there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        vp, i, sz, u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64
        L.lb_last_error.restype = C.c_char_p
        L.lb_version.restype = C.c_char_p
        L.lb_lens_name.restype = C.c_char_p
        L.lb_lens_name.argtypes = [i]
        L.lb_camera_params_default.argtypes = [C.POINTER(abi.CameraParams)]
        L.lb_camera_create.argtypes = [C.POINTER(abi.CameraParams), C.POINTER(abi.BokehImage), i, C.POINTER(vp)]
        L.lb_camera_update.argtypes = [vp, C.POINTER(abi.CameraParams), C.POINTER(abi.BokehImage)]
        L.lb_camera_destroy.argtypes = [vp]
        L.lb_camera_destroy.restype = None
        L.lb_camera_get_state.argtypes = [vp, C.POINTER(abi.CameraState)]
        L.lb_camera_set_state.argtypes = [vp, C.c_double, C.c_double]
        L.lb_camera_lens_work.argtypes = [vp, C.POINTER(abi.LensWork)]
        L.lb_camera_create_rays.argtypes = [vp, sz, u64, C.POINTER(abi.RayIn), C.POINTER(abi.RayOut), vp]
        L.lb_camera_create_rays_host.argtypes = [vp, sz, u64, C.POINTER(abi.RayIn), C.POINTER(abi.RayOut)]
        L.lb_camera_reverse_rays.argtypes = [vp, sz, vp, vp, vp]
        L.lb_filter_begin.argtypes = [vp, C.POINTER(abi.FrameDesc), i, C.POINTER(abi.AovDesc)]
        L.lb_filter_accumulate.argtypes = [vp, C.POINTER(abi.Samples), vp]
        L.lb_filter_accumulate_host.argtypes = [vp, C.POINTER(abi.Samples)]
        L.lb_filter_get_stats.argtypes = [vp, C.POINTER(abi.FilterStats)]
        L.lb_filter_newton_iterations.argtypes = [vp, C.POINTER(u64)]
        L.lb_imager_resolve.argtypes = [vp, i, i, i, i, i, vp, vp]
        L.lb_imager_resolve_host.argtypes = [vp, i, i, i, i, i, vp]
        L.lb_filter_buffers.argtypes = [vp, i, C.POINTER(vp), C.POINTER(vp)]
        L.lb_filter_buffers_host.argtypes = [vp, i, vp, vp]
        L.lb_filter_crypto_host.argtypes = [vp, i, vp, vp, C.POINTER(C.c_int)]
        L.lb_bench_fp32_peak.argtypes = [i, C.POINTER(C.c_double)]
        L.lb_camera_kernel_kind.argtypes = [vp]
        L.lb_bench_red_peak.argtypes = [i, i, C.POINTER(C.c_double)]
        L.lb_bench_splat_accum.argtypes = [i, i, i, C.POINTER(C.c_double)]
        L.lb_camera_set_pupil_geometry.argtypes = [vp, i, i]
        L.lb_comm_unique_id.argtypes = [C.c_char_p]
        L.lb_comm_init.argtypes = [vp, i, i, C.c_char_p]
        L.lb_filter_set_sample_base.argtypes = [vp, u64]
        L.lb_filter_reduce.argtypes = [vp, i, vp]
        L.lb_comm_destroy.argtypes = [vp]
        L.lb_filter_reduce_scatter.argtypes = [vp, vp]
        L.lb_filter_slab.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
        L.lb_imager_resolve_gather.argtypes = [vp, i, vp, i, vp]
        L.lb_imager_resolve_peer.argtypes = [vp, i, C.POINTER(C.c_int), C.POINTER(vp), i, vp]
        L.lb_imager_peer_image.argtypes = [vp, i, C.POINTER(vp)]
        L.lb_debug_primitives.argtypes = [i, sz, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise LentilError(f"lentil_b200 error {rc}: {lib().lb_last_error().decode()}")


def lens_names() -> list[str]:
    L = lib()
    return [L.lb_lens_name(k).decode() for k in range(L.lb_lens_count())]


def _dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _hptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())  # (pinned) CPU torch tensor


def _stream_ptr(stream):
    if stream is None:
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))


class Camera:
    """lentil_camera node: node_initialize + node_update on construction."""

    def __init__(self, params: abi.CameraParams, bokeh: np.ndarray | None = None, device: int = 0):
        self._h = C.c_void_p()
        self.device = device
        self.params = params
        self._bokeh_keep = None
        img = None
        if bokeh is not None:
            px = np.ascontiguousarray(bokeh, dtype=np.float32)
            self._bokeh_keep = px
            img = abi.BokehImage(px.shape[1], px.shape[0], px.shape[2], px.ctypes.data_as(C.POINTER(C.c_float)))
        _check(lib().lb_camera_create(C.byref(params), C.byref(img) if img is not None else None, device, C.byref(self._h)))
        self._frame = None
        self._aovs = []
        self._rank = 0
        self._world = 1

    def close(self):
        if getattr(self, "_h", None):
            lib().lb_camera_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ---------------------------------------------------------------------------------
    @property
    def state(self) -> abi.CameraState:
        s = abi.CameraState()
        _check(lib().lb_camera_get_state(self._h, C.byref(s)))
        return s

    def set_state(self, aperture_radius: float, sensor_shift: float):
        _check(lib().lb_camera_set_state(self._h, aperture_radius, sensor_shift))

    def set_pupil_geometry(self, outer: int, inner: int = 0):
        _check(lib().lb_camera_set_pupil_geometry(self._h, outer, inner))

    @property
    def kernel_kind(self) -> str:
        return "unrolled" if lib().lb_camera_kernel_kind(self._h) else "table"

    @property
    def lens_work(self) -> abi.LensWork:
        w = abi.LensWork()
        _check(lib().lb_camera_lens_work(self._h, C.byref(w)))
        return w

    # -- camera_create_ray ---------------------------------------------------------------------
    def create_rays(self, sx, sy, dsx, dsy, lensx, lensy, ray_id_base: int = 0, out: dict | None = None, stream=None,
                    want=RAY_OUT_FIELDS, want_tries: bool = True) -> dict:
        """Device tensors in (float32 [n] each) -> dict of device tensors ([3, n] planes, tries int32 [n])."""
        import torch

        n = sx.numel()
        dev = sx.device
        if out is None:
            out = {k: torch.empty((3, n), dtype=torch.float32, device=dev) for k in want}
            if want_tries:
                out["tries"] = torch.empty(n, dtype=torch.int32, device=dev)
        rin = abi.RayIn(*[_dptr(t) for t in (sx, sy, dsx, dsy, lensx, lensy)])
        rout = abi.RayOut(*[_dptr(out.get(k)) for k in RAY_OUT_FIELDS], _dptr(out.get("tries")))
        _check(lib().lb_camera_create_rays(self._h, n, ray_id_base, C.byref(rin), C.byref(rout), _stream_ptr(stream)))
        return out

    def create_rays_host(self, sx, sy, dsx, dsy, lensx, lensy, out: dict, ray_id_base: int = 0) -> dict:
        """Host arrays in (numpy or pinned CPU torch tensors), host arrays out; copies are pipelined inside."""
        n = int(sx.shape[0])
        rin = abi.RayIn(*[_hptr(t) for t in (sx, sy, dsx, dsy, lensx, lensy)])
        rout = abi.RayOut(*[_hptr(out.get(k)) for k in RAY_OUT_FIELDS], _hptr(out.get("tries")))
        _check(lib().lb_camera_create_rays_host(self._h, n, ray_id_base, C.byref(rin), C.byref(rout)))
        return out

    def reverse_rays(self, Po, stream=None):
        """camera_reverse_ray: Po float32 [n, 4] (device) -> Ps float32 [n, 2]."""
        import torch

        Ps = torch.empty((Po.shape[0], 2), dtype=torch.float32, device=Po.device)
        _check(lib().lb_camera_reverse_rays(self._h, Po.shape[0], _dptr(Po), _dptr(Ps), _stream_ptr(stream)))
        return Ps

    # -- lentil_filter / imager_lentil ----------------------------------------------------------
    def filter_begin(self, xres: int, yres: int, aovs, xres_full=None, yres_full=None, region_min=(0, 0), crypto_slots: int = 0):
        """aovs: list of (name, LB_FILTER_*, LB_AOV_*).  setup_filter: allocates zeroed device framebuffers."""
        self._frame = abi.FrameDesc(xres, yres, xres_full or xres, yres_full or yres, region_min[0], region_min[1], crypto_slots)
        arr = (abi.AovDesc * len(aovs))()
        for k, (name, flt, role) in enumerate(aovs):
            arr[k].name = name.encode()
            arr[k].filter = flt
            arr[k].role = role
        self._aovs = list(aovs)
        _check(lib().lb_filter_begin(self._h, C.byref(self._frame), len(aovs), arr))

    def _samples(self, px, py, rgba, pos_cs, inv_density, aov_values, raydir, transmission, flags, ptr, crypto=None, world_to_camera=None):
        n = int(px.shape[0])
        av = (C.c_void_p * max(len(self._aovs), 1))()
        for k in range(len(self._aovs)):
            a = None if aov_values is None else aov_values[k]
            av[k] = None if a is None else ptr(a).value
        S = abi.Samples(n, ptr(px), ptr(py), ptr(rgba), ptr(pos_cs), ptr(raydir), ptr(transmission), ptr(flags), av, inv_density)
        keep = [av]
        if crypto is not None:  # dict(depth, opacity [n, depth], ids {aov index: [n, depth]}, count [n] uint8 or None)
            ci = (C.c_void_p * max(len(self._aovs), 1))()
            for k, a in crypto.get("ids", {}).items():
                assert tuple(a.shape) == (n, int(crypto["depth"])), "crypto ids must be [n, depth]"
                ci[int(k)] = ptr(a).value
            S.crypto_depth = int(crypto["depth"])
            S.crypto_count = ptr(crypto.get("count"))
            S.crypto_opacity = ptr(crypto.get("opacity"))
            S.crypto_ids = ci
            keep.append(ci)
        if world_to_camera is not None:  # AtMatrix, float32 [4, 4] row-major on the HOST (lentil_filter.cpp:139-142)
            m = np.ascontiguousarray(np.asarray(world_to_camera, np.float32).reshape(4, 4))
            S.world_to_camera = C.c_void_p(m.ctypes.data)
            keep.append(m)
        return S, keep

    def filter_accumulate(self, px, py, rgba, pos_cs, inv_density, aov_values=None, raydir=None, transmission=None, flags=None, stream=None,
                          crypto=None, world_to_camera=None):
        """filter_pixel (RGBA branch) for a device-resident batch of samples.  world_to_camera: AtMatrix [4, 4] on the host;
        pos_cs / raydir are then world space (the device applies lentil_filter.cpp:121-142)."""
        S, keep = self._samples(px, py, rgba, pos_cs, inv_density, aov_values, raydir, transmission, flags, _dptr, crypto, world_to_camera)
        _check(lib().lb_filter_accumulate(self._h, C.byref(S), _stream_ptr(stream)))

    def filter_accumulate_host(self, px, py, rgba, pos_cs, inv_density, aov_values=None, raydir=None, transmission=None, flags=None, crypto=None,
                               world_to_camera=None):
        S, keep = self._samples(px, py, rgba, pos_cs, inv_density, aov_values, raydir, transmission, flags, _hptr, crypto, world_to_camera)
        _check(lib().lb_filter_accumulate_host(self._h, C.byref(S)))

    def filter_stats(self) -> dict:
        s = abi.FilterStats()
        _check(lib().lb_filter_get_stats(self._h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in s._fields_}
        its = C.c_uint64()
        _check(lib().lb_filter_newton_iterations(self._h, C.byref(its)))
        d["newton_its"] = its.value
        return d

    def resolve(self, aov: int, x0=None, y0=None, w=None, h=None, stream=None, fill: float = 0.0):
        """driver_process_bucket for one bucket (default: the whole region) -> device tensor [h, w, 4].
        fill: what the bucket holds beforehand; cryptomatte rows can end early and keep it (lentil_imager.cpp:132-134)."""
        import torch

        f = self._frame
        x0 = f.region_min_x if x0 is None else x0
        y0 = f.region_min_y if y0 is None else y0
        w = f.xres if w is None else w
        h = f.yres if h is None else h
        out = torch.full((h, w, 4), fill, dtype=torch.float32, device=f"cuda:{self.device}")
        _check(lib().lb_imager_resolve(self._h, aov, x0, y0, w, h, _dptr(out), _stream_ptr(stream)))
        return out

    def resolve_host(self, aov: int, out: np.ndarray | None = None, x0=None, y0=None, w=None, h=None):
        """driver_process_bucket with a host bucket (default: the whole region); `out` is [h, w, 4] float32."""
        f = self._frame
        x0 = f.region_min_x if x0 is None else x0
        y0 = f.region_min_y if y0 is None else y0
        w = f.xres if w is None else w
        h = f.yres if h is None else h
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        _check(lib().lb_imager_resolve_host(self._h, aov, x0, y0, w, h, _hptr(out)))
        return out

    def buffers(self, aov: int):
        """Raw accumulators as numpy copies: AOVData::buffer [yres, xres, 4], filter_weight_buffer [yres, xres]."""
        f = self._frame
        buf = np.empty((f.yres, f.xres, 4), np.float32)
        wgt = np.empty((f.yres, f.xres), np.float32)
        _check(lib().lb_filter_buffers_host(self._h, aov, _hptr(buf), _hptr(wgt)))
        return buf, wgt

    def crypto(self, aov: int):
        """AOVData::crypto_hash_map of a cryptomatte AOV as numpy tables: (ids [yres, xres, slots] float32 with NaN
        (all bits set) in unused slots, weights [yres, xres, slots], crypto_total_weight [yres, xres])."""
        f = self._frame
        slots = C.c_int()
        _check(lib().lb_filter_crypto_host(self._h, aov, None, None, C.byref(slots)))
        ids = np.empty((f.yres, f.xres, slots.value), np.float32)
        wts = np.empty((f.yres, f.xres, slots.value), np.float32)
        _check(lib().lb_filter_crypto_host(self._h, aov, _hptr(ids), _hptr(wts), None))
        return ids, wts, self.buffers(aov)[0][..., 0].copy()

    def buffer_pointers(self, aov: int):
        """Device addresses of the raw accumulators (owned by the camera)."""
        b, w = C.c_void_p(), C.c_void_p()
        _check(lib().lb_filter_buffers(self._h, aov, C.byref(b), C.byref(w)))
        return b.value, w.value

    # -- multi-GPU ---------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().lb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, world_size: int, rank: int, unique_id: bytes):
        _check(lib().lb_comm_init(self._h, world_size, rank, unique_id))
        self._rank = rank
        self._world = world_size

    def filter_set_sample_base(self, base: int):
        _check(lib().lb_filter_set_sample_base(self._h, base))

    def filter_reduce(self, root: int = -1, stream=None):
        _check(lib().lb_filter_reduce(self._h, root, _stream_ptr(stream)))

    def filter_reduce_scatter(self, stream=None):
        """Sum-reduce that leaves every rank owning one pixel slab of every plane (ncclReduceScatter per plane)."""
        _check(lib().lb_filter_reduce_scatter(self._h, _stream_ptr(stream)))

    def filter_slab(self):
        """(first pixel, pixel count) of the slab this rank owns after filter_reduce_scatter."""
        lo, n = C.c_size_t(), C.c_size_t()
        _check(lib().lb_filter_slab(self._h, C.byref(lo), C.byref(n)))
        return lo.value, n.value

    def resolve_gather(self, aov: int, root: int = 0, stream=None, out=None):
        """Every rank resolves its slab; the slabs are gathered on `root` (< 0: on every rank).  Returns the [yres, xres, 4]
        device tensor on the ranks that receive the image, None elsewhere."""
        import torch

        f = self._frame
        rank_gets = out is not None or root < 0 or self._rank == root
        if out is None and rank_gets:
            out = torch.empty((f.yres, f.xres, 4), dtype=torch.float32, device=f"cuda:{self.device}")
        _check(lib().lb_imager_resolve_gather(self._h, aov, _dptr(out), root, _stream_ptr(stream)))
        return out if rank_gets else None

    def resolve_peer(self, aovs, root: int = 0, stream=None, copy: bool = False):
        """lb_imager_resolve_peer: combine + resolve of the listed AOVs in one kernel over peer memory (collective).  Returns the
        [yres, xres, 4] device tensors on the receiving ranks (views of the library's image block unless copy=True), None elsewhere."""
        import torch

        f = self._frame
        aovs = list(aovs)
        idx = (C.c_int * len(aovs))(*aovs)
        _check(lib().lb_imager_resolve_peer(self._h, len(aovs), idx, None, root, _stream_ptr(stream)))
        if not (root < 0 or self._rank == root or self._world == 1):
            return None
        outs = []
        for a in aovs:
            ptr = C.c_void_p()
            _check(lib().lb_imager_peer_image(self._h, a, C.byref(ptr)))

            class _View:  # zero-copy: the image stays in the library's block
                __cuda_array_interface__ = {"shape": (f.yres, f.xres, 4), "typestr": "<f4", "data": (ptr.value, False), "version": 2}

            t = torch.as_tensor(_View(), device=f"cuda:{self.device}")
            outs.append(t.clone() if copy else t)
        return outs

    def comm_destroy(self):
        _check(lib().lb_comm_destroy(self._h))
