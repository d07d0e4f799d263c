"""Synthetic inputs of the BASELINE.json configs, built with torch so the same code
runs on the CPU (tests, oracle) and on the GPU (bench).  The jitter / lens samples come from the
reference's own counter RNG, tea<8> + LCG (/root/reference/src/global.h:32-57), restated here on
64-bit integer tensors (bit-exact, checked against the oracle in tests/test_workloads.py).
"""
from __future__ import annotations

import math

import torch

M32 = 0xFFFFFFFF


def tea8(v0: torch.Tensor, v1: torch.Tensor) -> torch.Tensor:
    """tea<8>(val0, val1) on int64 tensors holding uint32 values."""
    v0 = v0.to(torch.int64) & M32
    v1 = v1.to(torch.int64) & M32
    s0 = 0
    for _ in range(8):
        s0 = (s0 + 0x9E3779B9) & M32
        v0 = (v0 + ((((v1 << 4) + 0xA341316C) & M32) ^ ((v1 + s0) & M32) ^ (((v1 >> 5) + 0xC8013EA4) & M32))) & M32
        v1 = (v1 + ((((v0 << 4) + 0xAD90777D) & M32) ^ ((v0 + s0) & M32) ^ (((v0 >> 5) + 0x7E95761E) & M32))) & M32
    return v0


def lcg(state: torch.Tensor):
    """rng(previous): returns (new_state, float32 in [0,1))."""
    state = (state * 1664525 + 1013904223) & M32
    return state, (state & 0x00FFFFFF).to(torch.float32) / 16777216.0


def camera_samples(width: int, height: int, spp: int, device="cpu", first: int = 0, count: int | None = None, seed_mode: str = "pixel"):
    """AtCameraInput arrays for a width x height x spp frame (configs C1/C2).

    Sample index k = (py*width + px)*spp + s.  seed_mode "pixel": seed = tea<8>(py*width+px, s) (C2);
    "linear": seed = tea<8>(k, 0x5EED) (C1, spp = 1).  Returns dict of float32 [n] tensors.
    """
    total = width * height * spp
    count = total - first if count is None else count
    k = torch.arange(first, first + count, dtype=torch.int64, device=device)
    pix = k // spp
    s = k % spp
    px = pix % width
    py = pix // width
    seed = tea8(pix, s) if seed_mode == "pixel" else tea8(k, torch.full_like(k, 0x5EED))
    seed, jx = lcg(seed)
    seed, jy = lcg(seed)
    seed, lensx = lcg(seed)
    seed, lensy = lcg(seed)
    aspect = width / height
    sx = 2.0 * (px.to(torch.float32) + jx) / width - 1.0
    sy = (1.0 - 2.0 * (py.to(torch.float32) + jy) / height) / aspect
    d = torch.full_like(sx, 2.0 / width)
    return dict(sx=sx.contiguous(), sy=sy.contiguous(), dsx=d, dsy=d.clone(), lensx=lensx.contiguous(), lensy=lensy.contiguous())


def tile_partition(width: int, height: int, spp: int, rank: int, world: int, tile: int = 64, device="cpu") -> torch.Tensor:
    """Sample indices (int64, k = (py*width + px)*spp + s) of the frame's samples that `rank` of `world` accumulates:
    tile x tile pixel tiles dealt to the ranks by a HASH of the tile coordinates ("image tiles x sample
    ranges, round-robin to GPUs").  A hash instead of a regular round-robin pattern: highlights on a regular pitch (the
    light grid of config C3) alias with any regular pattern -- (tx + ty) % 8 with 4 x 4 tiles gave one rank 3.4x the mean
    highlight load, the hash 1.05x with 1 x 1 tiles.  Ordered tile by tile, row-major inside a tile, the spp samples of a
    pixel adjacent (the classify kernel merges runs of equal pixels).  The union over ranks is every sample exactly once."""
    py, px = torch.meshgrid(torch.arange(height, dtype=torch.int64, device=device), torch.arange(width, dtype=torch.int64, device=device), indexing="ij")
    tx, ty = px // tile, py // tile
    h = (tx * 0x9E3779B1 + ty * 0x85EBCA77) & M32
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & M32
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & M32
    h = h ^ (h >> 15)
    mine = (h % world) == rank
    px, py, tx, ty = px[mine], py[mine], tx[mine], ty[mine]
    tiles_x = -(-width // tile)
    key = ((ty * tiles_x + tx) * tile + (py % tile)) * tile + (px % tile)
    order = torch.argsort(key)
    pix = (py * width + px)[order]
    return (pix[:, None] * spp + torch.arange(spp, dtype=torch.int64, device=device)[None, :]).reshape(-1)


def highlight_frame(width: int, height: int, spp: int, tan_fov: float, device="cpu", first: int = 0, count: int | None = None,
                    z_plane: float = 75.0, pitch: float = 5.4, radius: float = 0.133, grid=(16, 12), radiance: float = 84.1589,
                    n_extra_aov: int = 0, samples: torch.Tensor | None = None):
    """Synthetic HDR frame of config C3, modelled on /root/reference/tests/cuda/lightgrid.ass: a black
    background at Z = inf plus a grid of emissive discs on the plane z_cs = -z_plane.  One sample per
    (pixel, s); the sample position comes from a pinhole ray through the jittered pixel position.

    `samples`: explicit sample indices (e.g. tile_partition) instead of the range [first, first + count).
    Returns dict(px, py int32 [n]; rgba, pos_cs float32 [n,4]; aov_values list of [n,4]).
    """
    total = width * height * spp
    if samples is not None:
        k = samples.to(device=device, dtype=torch.int64)
        count = int(k.numel())
    else:
        count = total - first if count is None else count
        k = torch.arange(first, first + count, dtype=torch.int64, device=device)
    pix = k // spp
    s = k % spp
    px = pix % width
    py = pix // width
    seed = tea8(pix, s + 0x1000)
    seed, jx = lcg(seed)
    seed, jy = lcg(seed)
    aspect = width / height
    sx = 2.0 * (px.to(torch.float32) + jx) / width - 1.0
    sy = (1.0 - 2.0 * (py.to(torch.float32) + jy) / height) / aspect
    x = sx * (tan_fov * z_plane)
    y = sy * (tan_fov * z_plane)
    # nearest disc centre of the grid (centred on the axis)
    gx, gy = grid
    cx = (torch.clamp(torch.round(x / pitch + (gx - 1) / 2.0), 0, gx - 1) - (gx - 1) / 2.0) * pitch
    cy = (torch.clamp(torch.round(y / pitch + (gy - 1) / 2.0), 0, gy - 1) - (gy - 1) / 2.0) * pitch
    hit = ((x - cx) ** 2 + (y - cy) ** 2) <= radius * radius
    n = count
    rgba = torch.zeros((n, 4), dtype=torch.float32, device=device)
    rgba[hit, 0:3] = radiance
    rgba[hit, 3] = 1.0
    pos = torch.zeros((n, 4), dtype=torch.float32, device=device)
    pos[:, 3] = 1.0e30  # AI_INFINITE: background is never redistributed (lentil_filter.cpp:130-133)
    pos[hit, 0] = x[hit]
    pos[hit, 1] = y[hit]
    pos[hit, 2] = -z_plane
    pos[hit, 3] = torch.sqrt(x[hit] ** 2 + y[hit] ** 2 + z_plane**2)
    out = dict(px=px.to(torch.int32).contiguous(), py=py.to(torch.int32).contiguous(), rgba=rgba, pos_cs=pos)
    # per-light AOVs (config C5): light l owns the discs with (column + row) % n_extra == l
    extra = []
    if n_extra_aov:
        col = torch.round(cx / pitch + (gx - 1) / 2.0).to(torch.int64)
        row = torch.round(cy / pitch + (gy - 1) / 2.0).to(torch.int64)
        owner = (col + row) % n_extra_aov
        for l in range(n_extra_aov):
            a = torch.zeros((n, 4), dtype=torch.float32, device=device)
            m = hit & (owner == l)
            a[m] = rgba[m]
            extra.append(a)
    out["aov_values"] = extra
    return out


def crypto_layers(frame: dict, depth: int, aov_indices, n_ids: int = 10, cell: int = 16, seed: int = 0):
    """Synthetic cryptomatte depth sub-samples for the samples of `frame` (highlight_frame): what
    cryptomatte_construct_cache (lentil.h:779-811) reads per sample.  Ids come from a palette of n_ids hash
    floats (one of them 0.0, some negative) chosen per cell x cell pixel block and layer, so neighbouring
    pixels share ids like objects do; opacities are multiples of 1/16 (AiColorToGrey of an equal-channel
    colour is then exact); about one sample in eight has no depth sub-sample at all.

    Returns dict(depth, count uint8 [n], opacity float32 [n, depth], ids {aov index: float32 [n, depth]}).
    """
    px, py = frame["px"].to(torch.int64), frame["py"].to(torch.int64)
    n, dev = px.shape[0], px.device
    pal_bits = tea8(torch.arange(n_ids, dtype=torch.int64, device=dev), torch.full((n_ids,), 77 + seed, dtype=torch.int64, device=dev))
    # keep the exponent in a sane range like Cryptomatte's hash_to_float does (no NaN / Inf / denormals)
    pal_bits = (pal_bits & 0x807FFFFF) | (((pal_bits >> 23) & 0x3F) + 96 << 23)
    palette = pal_bits.to(torch.int32).view(torch.float32).clone()
    palette[0] = 0.0
    k = torch.arange(n, dtype=torch.int64, device=dev)
    h = tea8(k, torch.full_like(k, 0x2000 + seed))
    count = (h % (depth + 1)).to(torch.uint8)
    count[(h >> 8) % 8 == 0] = 0
    d = torch.arange(depth, dtype=torch.int64, device=dev)[None, :]
    hd = tea8(k[:, None] * depth + d, torch.full((n, depth), 0x3000 + seed, dtype=torch.int64, device=dev))
    opacity = ((hd % 17).to(torch.float32) / 16.0).contiguous()
    cellid = (px // cell) + (py // cell) * 37
    ids = {}
    for j, a in enumerate(aov_indices):
        which = (cellid[:, None] * 3 + d * 5 + j * 7 + ((hd >> 6) % 2)) % n_ids
        ids[a] = palette[which].contiguous()
    return dict(depth=depth, count=count.contiguous(), opacity=opacity, ids=ids)


def disc_bokeh_image(size: int = 64):
    """A small procedural bokeh kernel (ring + hot centre), float32 [size, size, 3] in [0,1]."""
    import numpy as np

    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    c = (size - 1) / 2.0
    r = np.sqrt((xx - c) ** 2 + (yy - c) ** 2) / c
    lum = np.where(r <= 1.0, 0.35 + 0.65 * np.exp(-((r - 0.85) ** 2) / 0.01) + 0.3 * np.exp(-(r**2) / 0.02), 0.0).astype(np.float32)
    return np.repeat(lum[:, :, None], 3, axis=2).astype(np.float32)


def camera_ray_flops(work, newton_its_per_trace: float, traces_per_ray: float = 3.0) -> float:
    """per forward trace k*(F_ap + 16) + F_eval + 60 flop."""
    return traces_per_ray * (newton_its_per_trace * (work.F_ap + 16.0) + work.F_eval + 60.0)


def reverse_attempt_flops(work, newton_its: float) -> float:
    """per reverse attempt k*(F_apxy + F_apJ + F_out4 + F_outJ + 120) + F_T flop."""
    return newton_its * (work.F_apxy + work.F_apJ + work.F_out4 + work.F_outJ + 120.0) + work.F_T
