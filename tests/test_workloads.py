"""Input generators: the torch restatement of tea<8>/rng is bit-exact against the oracle's (and, via
test_oracle_vs_ref.py, the reference's); KATs computed by an independent pure-Python evaluation."""
import ctypes as C

import numpy as np
import torch

from pota_b200 import workloads


def _tea8_py(v0, v1):
    s0 = 0
    for _ in range(8):
        s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
        v0 = (v0 + ((((v1 << 4) + 0xA341316C) & 0xFFFFFFFF) ^ ((v1 + s0) & 0xFFFFFFFF) ^ (((v1 >> 5) + 0xC8013EA4) & 0xFFFFFFFF))) & 0xFFFFFFFF
        v1 = (v1 + ((((v0 << 4) + 0xAD90777D) & 0xFFFFFFFF) ^ ((v0 + s0) & 0xFFFFFFFF) ^ (((v0 >> 5) + 0x7E95761E) & 0xFFFFFFFF))) & 0xFFFFFFFF
    return v0


def test_tea8_and_lcg_match_oracle(oracle_lib):
    rs = np.random.default_rng(0)
    a = rs.integers(0, 2**32, 5000, dtype=np.uint64).astype(np.int64)
    b = rs.integers(0, 2**32, 5000, dtype=np.uint64).astype(np.int64)
    t = workloads.tea8(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    for i in range(0, 5000, 37):
        assert int(t[i]) == oracle_lib.orc_tea8(int(a[i]), int(b[i])) == _tea8_py(int(a[i]), int(b[i]))
    state = torch.from_numpy(a[:64].copy())
    for _ in range(5):
        prev = state.clone()
        state, f = workloads.lcg(state)
        for i in range(64):
            s = C.c_uint(int(prev[i]))
            assert float(f[i]) == oracle_lib.orc_rng(C.byref(s)) and s.value == int(state[i])
            assert 0.0 <= float(f[i]) < 1.0


def test_camera_samples_shapes_and_ranges():
    s = workloads.camera_samples(64, 36, 4)
    n = 64 * 36 * 4
    for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy"):
        assert s[k].shape == (n,) and s[k].dtype == torch.float32
    assert s["sx"].min() >= -1 and s["sx"].max() <= 1 and abs(float(s["sy"].max()) - 36 / 64) < 0.02
    assert 0 <= s["lensx"].min() and s["lensx"].max() < 1
    # ragged range == slice of the full frame
    part = workloads.camera_samples(64, 36, 4, first=1000, count=333)
    for k in s:
        assert torch.equal(part[k], s[k][1000:1333])


def test_highlight_frame_layout():
    fr = workloads.highlight_frame(96, 54, 4, 0.36, n_extra_aov=3)
    n = 96 * 54 * 4
    assert fr["rgba"].shape == (n, 4) and fr["pos_cs"].shape == (n, 4) and fr["px"].dtype == torch.int32
    hit = fr["rgba"][:, 3] > 0
    assert 0 < hit.sum() < n // 10
    assert torch.all(fr["pos_cs"][~hit, 3] == 1.0e30) and torch.all(fr["pos_cs"][hit, 2] == -75.0)
    total = sum(fr["aov_values"])
    assert torch.equal(total, fr["rgba"])  # the per-light AOVs partition the beauty


def test_tile_partition_covers_every_sample_once_and_balances():
    import torch

    W, H, spp = 200, 120, 3
    for world in (1, 2, 3, 8):
        parts = [workloads.tile_partition(W, H, spp, r, world, tile=16) for r in range(world)]
        allk = torch.cat(parts)
        assert allk.numel() == W * H * spp and torch.equal(torch.sort(allk).values, torch.arange(W * H * spp))
        # tiles are dealt by a hash of their coordinates: with many small tiles every rank gets close to its share
        sizes = [p.numel() for p in [workloads.tile_partition(W, H, spp, r, world, tile=2) for r in range(world)]]
        assert sum(sizes) == W * H * spp and max(sizes) <= 1.2 * (W * H * spp / world)
        for p in parts:  # the spp samples of a pixel are adjacent
            assert torch.equal(p.reshape(-1, spp) // spp, (p.reshape(-1, spp)[:, :1] // spp).expand(-1, spp))
    # a frame built from a partition holds exactly the samples of the range-built frame
    k = workloads.tile_partition(64, 36, 4, 1, 2, tile=16)
    a = workloads.highlight_frame(64, 36, 4, 0.36, "cpu", samples=k)
    b = workloads.highlight_frame(64, 36, 4, 0.36, "cpu")
    for key in ("px", "py", "rgba", "pos_cs"):
        assert torch.equal(a[key], b[key][k])
