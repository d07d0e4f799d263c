"""Shared helpers of the parity tests."""
import numpy as np

from pota_b200 import abi

VEC_FIELDS = ("origin", "dir")
DERIV_FIELDS = ("dOdx", "dOdy", "dDdx", "dDdy")


def po_params(**kw):
    base = dict(camera_type=abi.LB_CAMERA_POLYNOMIAL_OPTICS, lens_model=5, fstop=2.8, focus_dist=150.0)
    base.update(kw)
    return abi.CameraParams.defaults(**base)


def rel_err_vec(a, ref):
    """max over components of |a-ref| / max(|ref| over the vector): a and ref are [3, n]."""
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-30)
    return np.abs(a - ref).max(axis=0) / scale


def rel_err_comp(a, ref, floor):
    """per-component relative error with an absolute floor on the denominator."""
    return np.abs(a - ref) / np.maximum(np.abs(ref), floor)
