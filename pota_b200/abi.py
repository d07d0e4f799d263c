"""ctypes mirror of include/lentil_b200.h (POD types only; no library is loaded here)."""
from __future__ import annotations

import ctypes as C

LB_OK = 0
LB_ERR_INVALID, LB_ERR_NO_DEVICE, LB_ERR_CUDA, LB_ERR_LENS, LB_ERR_STATE, LB_ERR_IMAGE, LB_ERR_COMM = -1, -2, -3, -4, -5, -6, -7
LB_UNITS_MM, LB_UNITS_CM, LB_UNITS_DM, LB_UNITS_M = 0, 1, 2, 3
LB_CAMERA_THINLENS, LB_CAMERA_POLYNOMIAL_OPTICS = 0, 1
LB_FILTER_GAUSSIAN, LB_FILTER_CLOSEST, LB_FILTER_CRYPTO = 0, 1, 2
LB_CRYPTO_MAX_DEPTH, LB_CRYPTO_DEFAULT_SLOTS, LB_CRYPTO_MAX_SLOTS = 8, 16, 64
LB_AOV_PLAIN, LB_AOV_RGBA, LB_AOV_LENTIL_DEBUG = 0, 1, 2
LB_SAMPLE_VOLUME, LB_SAMPLE_IGNORE = 1, 2

_f, _i = C.c_float, C.c_int32
_pf = C.POINTER(C.c_float)


class CameraParams(C.Structure):
    """lb_camera_params == node parameters of lentil_camera (lentil_camera.cpp:19-52)."""

    _fields_ = [
        ("camera_type", _i),
        ("bidir_sample_mult", _i),
        ("units", _i),
        ("sensor_width", _f),
        ("enable_dof", _i),
        ("fstop", _f),
        ("focus_dist", _f),
        ("aperture_blades_lentil", _i),
        ("exp", _f),
        ("lens_model", _i),
        ("wavelength", _f),
        ("extra_sensor_shift", _f),
        ("focal_length_lentil", _f),
        ("optical_vignetting", _f),
        ("abb_spherical", _f),
        ("abb_distortion", _f),
        ("abb_coma", _f),
        ("abb_chromatic", _f),
        ("abb_chromatic_type", _i),
        ("bokeh_circle_to_square", _f),
        ("bokeh_anamorphic", _f),
        ("bokeh_enable_image", _i),
        ("vignetting_retries", _i),
        ("bidir_add_energy", _f),
        ("bidir_add_energy_minimum_luminance", _f),
        ("bidir_add_energy_transition", _f),
        ("enable_bidir_transmission", _i),
        ("enable_skydome", _i),
    ]

    @classmethod
    def defaults(cls, **kw) -> "CameraParams":
        """The reference's C++ defaults (lentil_camera.cpp:19-52)."""
        p = cls(
            camera_type=LB_CAMERA_THINLENS,
            bidir_sample_mult=5,
            units=LB_UNITS_CM,
            sensor_width=36.0,
            enable_dof=1,
            fstop=0.0,
            focus_dist=150.0,
            aperture_blades_lentil=0,
            exp=1.0,
            lens_model=16,  # cooke__speed_panchro__1920__40mm
            wavelength=550.0,
            extra_sensor_shift=0.0,
            focal_length_lentil=35.0,
            optical_vignetting=0.0,
            abb_spherical=0.5,
            abb_distortion=0.0,
            abb_coma=0.0,
            abb_chromatic=0.0,
            abb_chromatic_type=0,
            bokeh_circle_to_square=0.0,
            bokeh_anamorphic=0.0,
            bokeh_enable_image=0,
            vignetting_retries=15,
            bidir_add_energy=0.0,
            bidir_add_energy_minimum_luminance=2.0,
            bidir_add_energy_transition=1.0,
            enable_bidir_transmission=0,
            enable_skydome=0,
        )
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p


class BokehImage(C.Structure):
    _fields_ = [("width", _i), ("height", _i), ("channels", _i), ("pixels", _pf)]


class CameraState(C.Structure):
    _fields_ = [
        (n, C.c_double)
        for n in (
            "aperture_radius sensor_shift tan_fov focus_distance lambda_ lens_outer_pupil_radius lens_inner_pupil_radius "
            "lens_length lens_back_focal_length lens_effective_focal_length lens_aperture_pos lens_aperture_housing_radius "
            "lens_inner_pupil_curvature_radius lens_outer_pupil_curvature_radius lens_field_of_view lens_fstop "
            "lens_aperture_radius_at_fstop"
        ).split()
    ] + [("outer_pupil_geometry", _i), ("inner_pupil_geometry", _i), ("focus_check_ok", _i), ("focus_check_distance", C.c_double)]


class RayIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")]


class RayOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("origin", "dir", "dOdx", "dOdy", "dDdx", "dDdy", "weight", "tries")]


class LensWork(C.Structure):
    _fields_ = [(n, _i) for n in ("terms_eval", "terms_ap", "terms_ap_jac", "terms_out_jac")] + [
        (n, C.c_double) for n in ("F_eval", "F_ap", "F_apxy", "F_apJ", "F_out4", "F_outJ", "F_T")
    ]


class AovDesc(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("filter", _i), ("role", _i)]


class FrameDesc(C.Structure):
    _fields_ = [(n, _i) for n in ("xres", "yres", "xres_without_region", "yres_without_region", "region_min_x", "region_min_y", "crypto_slots")]


class Samples(C.Structure):
    _fields_ = [
        ("n", C.c_size_t),
        ("px", C.c_void_p),
        ("py", C.c_void_p),
        ("rgba", C.c_void_p),
        ("pos_cs", C.c_void_p),
        ("raydir", C.c_void_p),
        ("transmission", C.c_void_p),
        ("flags", C.c_void_p),
        ("aov_values", C.POINTER(C.c_void_p)),
        ("inv_density", _f),
        ("crypto_depth", _i),
        ("crypto_count", C.c_void_p),
        ("crypto_opacity", C.c_void_p),
        ("crypto_ids", C.POINTER(C.c_void_p)),
        ("world_to_camera", C.c_void_p),
    ]


class FilterStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "redistributed", "splats", "attempts", "passthrough", "crypto_dropped", "tile_splats")]


def host_samples(n_aov, px, py, rgba, pos_cs, inv_density, aov_values=None, raydir=None, transmission=None, flags=None, crypto=None,
                 world_to_camera=None):
    """lb_samples over HOST numpy arrays -> (Samples, keepalive list).

    crypto: dict(depth=D, opacity=[n][D] float32, ids={aov index: [n][D] float32}, count=[n] uint8 or None).
    """
    import numpy as np

    keep = []

    def ptr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return C.c_void_p(a.ctypes.data)

    n = int(np.asarray(px).shape[0])
    av = (C.c_void_p * max(n_aov, 1))()
    for i in range(n_aov):
        a = None if aov_values is None else aov_values[i]
        if a is not None:
            av[i] = ptr(a, np.float32)
    S = Samples(n, ptr(px, np.int32), ptr(py, np.int32), ptr(rgba, np.float32), ptr(pos_cs, np.float32), ptr(raydir, np.float32),
                ptr(transmission, np.float32), ptr(flags, np.uint32), av, inv_density)
    keep.append(av)
    if crypto is not None:
        D = int(crypto["depth"])
        ci = (C.c_void_p * max(n_aov, 1))()
        for i, a in crypto.get("ids", {}).items():
            assert np.asarray(a).shape == (n, D)
            ci[int(i)] = ptr(a, np.float32)
        S.crypto_depth = D
        S.crypto_count = ptr(crypto.get("count"), np.uint8)
        S.crypto_opacity = ptr(crypto.get("opacity"), np.float32)
        S.crypto_ids = ci
        keep.append(ci)
    if world_to_camera is not None:
        S.world_to_camera = ptr(np.asarray(world_to_camera, np.float32).reshape(4, 4), np.float32)
    return S, keep
