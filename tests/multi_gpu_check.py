"""Multi-GPU check of lb_filter_reduce, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank accumulates its sample range of one frame; after the NCCL combine rank 0 must hold the same
framebuffers (gaussian AOVs, weight, closest AOV, cryptomatte id tables) as a single-GPU run over all samples.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pota_b200 import abi, workloads  # noqa: E402
from pota_b200.camera import Camera  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = abi.CameraParams.defaults(camera_type=1, lens_model=5, fstop=1.4, focus_dist=35.0, bidir_sample_mult=6)
    W, H, spp = 320, 180, 4
    aovs = [("RGBA", 0, 1), ("light0", 0, 0), ("N", 1, 0), ("crypto_material00", 2, 0), ("crypto_object01", 2, 0)]
    total = W * H * spp

    def run(cam, lo, hi, dev):
        fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, dev, lo, hi - lo, n_extra_aov=2)
        cam.filter_begin(W, H, aovs)
        cam.filter_set_sample_base(lo)
        # crypto layers are generated for the whole frame so that every rank sees the single-GPU run's values
        full = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, dev)
        cr = workloads.crypto_layers(full, 3, [3, 4])
        crypto = dict(depth=3, count=cr["count"][lo:hi].contiguous(), opacity=cr["opacity"][lo:hi].contiguous(),
                      ids={a: v[lo:hi].contiguous() for a, v in cr["ids"].items()})
        cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp,
                              aov_values=[None, fr["aov_values"][0], fr["aov_values"][1], None, None], crypto=crypto)
        return fr, crypto

    dev = torch.device("cuda", local)
    cam = Camera(p, device=local)
    lo, hi = total * rank // world, total * (rank + 1) // world
    keep = run(cam, lo, hi, dev)
    uid = [Camera.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    cam.comm_init(world, rank, uid[0])
    for root in (0, -1):  # reduce to rank 0, then all-reduce on fresh partials
        if root == -1:
            keep = run(cam, lo, hi, dev)
        cam.filter_reduce(root=root)
        torch.cuda.synchronize()
        if rank == 0 or root == -1:
            single = Camera(p, device=local)
            keep2 = run(single, 0, total, dev)
            torch.cuda.synchronize()
            for a, (name, flt, role) in enumerate(aovs):
                got, gw = cam.buffers(a)
                ref, rw = single.buffers(a)
                if flt == 0:
                    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5, err_msg=f"root={root} {name}")
                    np.testing.assert_allclose(gw, rw, rtol=2e-5, atol=1e-6)
                elif flt == 2:
                    # same {id -> weight} table per pixel (slot positions may differ), same crypto_total_weight
                    np.testing.assert_allclose(got[..., 0], ref[..., 0], rtol=2e-5, atol=1e-6, err_msg=f"root={root} total weight {name}")
                    ig, wg, _ = cam.crypto(a)
                    ir, wr, _ = single.crypto(a)
                    og, orr = np.argsort(ig.view(np.uint32), axis=2), np.argsort(ir.view(np.uint32), axis=2)
                    np.testing.assert_array_equal(np.take_along_axis(ig.view(np.uint32), og, 2), np.take_along_axis(ir.view(np.uint32), orr, 2))
                    np.testing.assert_allclose(np.take_along_axis(wg, og, 2), np.take_along_axis(wr, orr, 2), rtol=2e-5, atol=1e-6)
                    rg, rr = cam.resolve(a, fill=-7.0).cpu().numpy(), single.resolve(a, fill=-7.0).cpu().numpy()
                    assert ((rg[..., 0] == rr[..., 0]) & (np.abs(rg[..., 1] - rr[..., 1]) < 1e-4)).mean() > 0.999
                else:
                    assert (np.abs(got - ref).max(axis=2) > 1e-6).mean() < 1e-3, f"root={root} closest {name}"
            assert cam.filter_stats()["samples"] == hi - lo  # (a rank whose slice holds no highlight has 0 splats)
            assert cam.filter_stats()["crypto_dropped"] == 0
    # ---- the scalable combine: round-robin tiles of source samples, reduce-scatter, per-rank resolve, gather -------------
    aovs2 = [("RGBA", 0, 1), ("light0", 0, 0), ("light1", 0, 0), ("N", 1, 0)]

    def run2(c, k):
        fr = workloads.highlight_frame(W, H, spp, c.state.tan_fov, dev, n_extra_aov=2, samples=k)
        c.filter_begin(W, H, aovs2)
        # closest AOVs need globally unique sample indices: the position of the rank's first sample in the concatenation of all shares
        c.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, aov_values=[None, fr["aov_values"][0], fr["aov_values"][1], fr["rgba"]])
        return fr

    shares = [workloads.tile_partition(W, H, spp, r, world, tile=32, device=dev) for r in range(world)]
    base = sum(int(s.numel()) for s in shares[:rank])
    fr2 = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, dev, n_extra_aov=2, samples=shares[rank])
    cam.filter_begin(W, H, aovs2)
    cam.filter_set_sample_base(base)
    cam.filter_accumulate(fr2["px"], fr2["py"], fr2["rgba"], fr2["pos_cs"], 1.0 / spp, aov_values=[None, fr2["aov_values"][0], fr2["aov_values"][1], fr2["rgba"]])
    # the fused combine + resolve over peer memory runs on the untouched partial planes, before the reduce-scatter consumes them
    single2 = Camera(p, device=local)
    run2(single2, torch.cat(shares))
    want2 = [single2.resolve(a).cpu().numpy() for a in range(len(aovs2))]
    for root in (0, -1, world - 1):
        for rep in range(2):  # twice: the second call reuses the peer mappings
            imgs = cam.resolve_peer(range(len(aovs2)), root=root, copy=True)
        torch.cuda.synchronize()
        if root < 0 or rank == root:
            for a, (name, flt, role) in enumerate(aovs2):
                got = imgs[a].cpu().numpy()
                if flt == 0:
                    np.testing.assert_allclose(got, want2[a], rtol=1e-4, atol=2e-5, err_msg=f"resolve_peer root={root} {name}")
                else:
                    assert (np.abs(got - want2[a]).max(axis=2) > 1e-6).mean() < 1e-3, f"resolve_peer root={root} closest {name}"
        else:
            assert imgs is None
    # a frame of another size: buffers are reallocated on every rank, the mappings renewed
    W3, H3 = 256, 144
    shares3 = [workloads.tile_partition(W3, H3, spp, r, world, tile=16, device=dev) for r in range(world)]
    cam3_aovs = [("RGBA", 0, 1), ("N", 1, 0)]

    def run3(c, k, base):
        fr = workloads.highlight_frame(W3, H3, spp, c.state.tan_fov, dev, samples=k)
        c.filter_begin(W3, H3, cam3_aovs)
        c.filter_set_sample_base(base)
        c.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, aov_values=[None, fr["rgba"]])
        return fr

    keep3 = run3(cam, shares3[rank], sum(int(s.numel()) for s in shares3[:rank]))
    imgs3 = cam.resolve_peer([0, 1], root=0)
    torch.cuda.synchronize()
    if rank == 0:
        single3 = Camera(p, device=local)
        run3(single3, torch.cat(shares3), 0)
        np.testing.assert_allclose(imgs3[0].cpu().numpy(), single3.resolve(0).cpu().numpy(), rtol=1e-4, atol=2e-5, err_msg="resolve_peer after a frame-size change")
        assert (np.abs(imgs3[1].cpu().numpy() - single3.resolve(1).cpu().numpy()).max(axis=2) > 1e-6).mean() < 1e-3
    dist.barrier()
    # back to the first frame for the reduce-scatter path
    cam.filter_begin(W, H, aovs2)
    cam.filter_set_sample_base(base)
    cam.filter_accumulate(fr2["px"], fr2["py"], fr2["rgba"], fr2["pos_cs"], 1.0 / spp, aov_values=[None, fr2["aov_values"][0], fr2["aov_values"][1], fr2["rgba"]])
    cam.filter_reduce_scatter()
    lo_px, n_px = cam.filter_slab()
    assert n_px > 0 and (rank > 0 or lo_px == 0)
    try:
        cam.resolve(0)
        raise AssertionError("lb_imager_resolve must refuse after lb_filter_reduce_scatter")
    except Exception as e:  # noqa: BLE001
        assert "slab" in str(e)
    for root in (0, -1):
        imgs = [cam.resolve_gather(a, root=root) for a in range(len(aovs2))]
        torch.cuda.synchronize()
        if rank == 0 or root == -1:
            single = Camera(p, device=local)
            run2(single, torch.cat(shares))
            for a, (name, flt, role) in enumerate(aovs2):
                want = single.resolve(a).cpu().numpy()
                got = imgs[a].cpu().numpy()
                if flt == 0:
                    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5, err_msg=f"resolve_gather root={root} {name}")
                else:
                    assert (np.abs(got - want).max(axis=2) > 1e-6).mean() < 1e-3, f"resolve_gather root={root} closest {name}"
        else:
            assert all(i is None for i in imgs)
    dist.barrier()
    cam.comm_destroy()
    if rank == 0:
        print(f"multi_gpu_check ok: world={world}, {total} samples, reduce + all-reduce + peer-memory combine + reduce-scatter/gather match the single-GPU framebuffers")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
