"""N > 1 host logic on the CPU (gloo, world_size 2): the sample-range sharding bench.py uses and the
sum-combine of per-rank partial framebuffers reproduce the single-process result.  The per-rank compute
here is the oracle (test infrastructure) — the GPU path uses the same partition with lb_filter_reduce."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import orc
from pota_b200 import abi, workloads
from tests.util import po_params

W, H, SPP = 64, 36, 4


def _shard(total, rank, world):
    return total * rank // world, total * (rank + 1) // world


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam = orc.OracleCamera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=4))
    lo, hi = _shard(W * H * SPP, rank, world)
    fr = workloads.highlight_frame(W, H, SPP, cam.state.tan_fov, "cpu", lo, hi - lo)
    cam.filter_begin(W, H, [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)])
    cam.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / SPP)
    buf, wgt = cam.buffers(0)
    block = torch.from_numpy(np.concatenate([buf.ravel(), wgt.ravel()]))  # the contiguous reduction unit of the product
    dist.reduce(block, dst=0, op=dist.ReduceOp.SUM)
    # camera rays: weak scaling, every rank traces its own sample range with globally unique ray ids
    n = 5000
    ins = workloads.camera_samples(100, 100, world, "cpu", rank * n, n, "pixel")
    rays = cam.create_rays(*[ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")], ray_id_base=rank * n)
    gathered = [torch.zeros(3, n) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(rays["dir"]), gathered, dst=0)
    if rank == 0:
        q.put((block.numpy(), torch.cat(gathered, dim=1).numpy()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_partition_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    block, dirs = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cam = orc.OracleCamera(po_params(fstop=1.4, focus_dist=35.0, bidir_sample_mult=4))
    fr = workloads.highlight_frame(W, H, SPP, cam.state.tan_fov, "cpu")
    cam.filter_begin(W, H, [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)])
    cam.filter_accumulate(fr["px"].numpy(), fr["py"].numpy(), fr["rgba"].numpy(), fr["pos_cs"].numpy(), 1.0 / SPP)
    buf, wgt = cam.buffers(0)
    single = np.concatenate([buf.ravel(), wgt.ravel()])
    np.testing.assert_allclose(block, single, rtol=1e-5, atol=1e-6)  # float sum order differs across the cut
    n = 5000
    ins = workloads.camera_samples(100, 100, world, "cpu", 0, world * n, "pixel")
    rays = cam.create_rays(*[ins[k].numpy() for k in ("sx", "sy", "dsx", "dsy", "lensx", "lensy")], ray_id_base=0)
    np.testing.assert_array_equal(dirs, rays["dir"])


def test_shard_covers_every_sample_once():
    for total in (0, 1, 7, 33_177_600):
        for world in (1, 2, 3, 8):
            cuts = [_shard(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
