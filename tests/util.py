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


def branch_frame(tan_fov, W=128, H=72, spp=9, seed=3):
    """Source samples that walk every per-sample branch of filter_pixel (lentil_filter.cpp:115-164) in one frame, seen
    through a camera that is rotated and translated in the world (non-identity AiWorldToCameraMatrix):

      * the emissive discs of workloads.highlight_frame (redistributed);
      * a band of them behind glass: transmission > 0 (energy subtracted, passed through unless enable_bidir_transmission);
      * a band inside a volume (LB_SAMPLE_VOLUME) and a band with lentil_bidir_ignore (LB_SAMPLE_IGNORE);
      * background samples (Z = AI_INFINITE): most black, a bright "sun" patch that carries lentil_raydir (redistributed
        from raydir * 99999999 when enable_skydome, passed through otherwise) and a patch whose raydir is zero;
      * samples with a finite depth whose world-space P is (almost) the origin: AiV3IsSmall in WORLD space.

    Returns numpy arrays: px, py, rgba, pos (WORLD space xyz + Z), raydir (world), transmission, flags, world_to_camera.
    """
    import torch

    from pota_b200 import workloads

    fr = workloads.highlight_frame(W, H, spp, tan_fov, "cpu")
    px, py = fr["px"].numpy(), fr["py"].numpy()
    rgba, pos_cs = fr["rgba"].numpy().copy(), fr["pos_cs"].numpy().copy()
    n = px.shape[0]
    rs = np.random.default_rng(seed)
    hit = rgba[:, 3] > 0
    transmission = np.zeros((n, 4), np.float32)
    band = hit & (px >= W // 4) & (px < W // 2)
    transmission[band, :3] = (rgba[band, :3] * rs.uniform(0.1, 0.6, (int(band.sum()), 1))).astype(np.float32)
    transmission[band, 3] = 1.0
    flags = np.zeros(n, np.uint32)
    flags[hit & (py < H // 5)] |= abi.LB_SAMPLE_VOLUME
    flags[hit & (py >= 4 * H // 5)] |= abi.LB_SAMPLE_IGNORE
    # skydome: camera-space view direction of the pixel, stored in WORLD space below
    aspect = W / H
    sx = 2.0 * (px + 0.5) / W - 1.0
    sy = (1.0 - 2.0 * (py + 0.5) / H) / aspect
    d = np.stack([sx * tan_fov, sy * tan_fov, -np.ones(n)], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    raydir_cs = np.zeros((n, 4), np.float32)
    sun = (~hit) & (px >= 3 * W // 4) & (px < 3 * W // 4 + 6) & (py >= H // 2) & (py < H // 2 + 6)
    rgba[sun] = np.array([40.0, 35.0, 20.0, 1.0], np.float32)
    raydir_cs[sun, :3] = d[sun]  # only the sun patch carries a direction: every other background sample keeps raydir = 0
    no_dir = (~hit) & (px < 4) & (py < 4)
    rgba[no_dir] = np.array([30.0, 30.0, 30.0, 1.0], np.float32)
    raydir_cs[no_dir] = 0.0
    # rigid camera: p_cam = p_world * R + t  (AiM4PointByMatrixMult, row vectors)
    a, b = 0.4, -0.25
    Ry = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), np.sin(b)], [0, -np.sin(b), np.cos(b)]])
    R = (Ry @ Rx).astype(np.float32).astype(np.float64)
    t = np.array([12.5, -3.0, 40.0])
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = R
    M[3, :3] = t
    pos = pos_cs.copy()
    pos[:, :3] = ((pos_cs[:, :3].astype(np.float64) - t) @ R.T).astype(np.float32)
    pos[~hit, :3] = 0.0  # background: no position data
    raydir = raydir_cs.copy()
    raydir[:, :3] = (raydir_cs[:, :3].astype(np.float64) @ R.T).astype(np.float32)
    raydir[no_dir] = 0.0
    # finite depth but world-space P at the origin: the world-space smallness test must catch these
    small = hit & (np.cumsum(hit) % 7 == 3)
    pos[small, :3] = rs.uniform(-5e-5, 5e-5, (int(small.sum()), 3)).astype(np.float32)
    return dict(px=px, py=py, rgba=rgba, pos=pos, raydir=raydir, transmission=transmission, flags=flags, world_to_camera=M,
                masks=dict(hit=hit, band=band, sun=sun, no_dir=no_dir, small=small))
