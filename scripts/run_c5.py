"""Config C5 (BASELINE.json configs[4]): 7680x4320 multi-AOV redistribution (beauty + 8 per-light RGBA AOVs + a
closest-filter AOV).  Source samples are dealt to the ranks in round-robin pixel tiles; every rank accumulates full-frame
partials; the combine is either lb_filter_reduce_scatter + per-rank resolve + gather on rank 0 (default) or the round-1
path lb_filter_reduce to rank 0 + resolve there (--combine reduce), so both can be measured on the same box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 scripts/run_c5.py [--spp 16]
    python scripts/run_c5.py --spp 4          (single GPU)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pota_b200 import abi, workloads  # noqa: E402
from pota_b200.camera import Camera  # noqa: E402

W, H = 7680, 4320


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--combine", default="peer", choices=["peer", "scatter", "reduce", "both"], help="both: the peer combine, then reduce-scatter + gather, on the same partials of every step")
    ap.add_argument("--tile", type=int, default=4)
    ap.add_argument("--emulate-world", type=int, default=0, help="single GPU: accumulate rank 0's share of an N-rank run (no combine partner)")
    ap.add_argument("--order", default="row", choices=["row", "bucket"], help="sample order of a batch: row-major, or 128x128 render buckets")
    ap.add_argument("--thin", action="store_true", help="ThinLens camera_type instead of PolynomialOptics")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS if a.thin else abi.LB_CAMERA_POLYNOMIAL_OPTICS, lens_model=5, fstop=1.4,
                                  focus_dist=35.0, bidir_sample_mult=10, bokeh_enable_image=1, focal_length_lentil=50.0)
    cam = Camera(p, bokeh=workloads.disc_bokeh_image(250), device=local)
    aovs = [("RGBA", 0, 1)] + [(f"light{k}", 0, 0) for k in range(8)] + [("N", 1, 0)]
    part_world = a.emulate_world if (world == 1 and a.emulate_world > 1) else world
    mine = workloads.tile_partition(W, H, a.spp, rank, part_world, tile=a.tile, device=dev)  # int64 sample indices of this rank
    if a.order == "bucket":  # as a renderer delivers them: bucket after bucket
        pix = mine // a.spp
        key = ((pix // W) // 128) * ((W + 127) // 128) + (pix % W) // 128
        mine = mine[torch.sort(key, stable=True).indices]
    n_mine = int(mine.numel())
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = n_mine
    if world > 1:
        dist.all_reduce(sizes)
    base = int(sizes[:rank].sum().item())  # globally unique sample indices for the closest-filter AOV
    lo, hi = 0, n_mine
    cam.filter_begin(W, H, aovs)
    if world > 1:
        uid = [Camera.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        cam.comm_init(world, rank, uid[0])
    chunk = 1920 * 1080 * 16  # source samples generated and accumulated per call
    stream = torch.cuda.current_stream()
    times = []
    peer_sum, scatter_sum = [], []
    for step in range(a.steps + 1):
        cam.filter_begin(W, H, aovs)
        cam.filter_set_sample_base(base)
        t_acc = t_red = 0.0
        for c0 in range(lo, hi, chunk):
            m = min(chunk, hi - c0)
            fr = workloads.highlight_frame(W, H, a.spp, cam.state.tan_fov, dev, grid=(8, 4), n_extra_aov=8, samples=mine[c0:c0 + m])
            vals = [None] + fr["aov_values"] + [fr["aov_values"][0]]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / a.spp, aov_values=vals, stream=stream)
            e1.record(stream)
            torch.cuda.synchronize()
            t_acc += e0.elapsed_time(e1)
            del fr, vals
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if world > 1:
            dist.barrier()
        t_peer = 0.0
        if a.combine == "both":  # the peer combine leaves the partial planes as they are: the NCCL path below runs on the same frame
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            imgs = cam.resolve_peer(range(len(aovs)), root=0, stream=stream)
            p1.record(stream)
            torch.cuda.synchronize()
            t_peer = p0.elapsed_time(p1)
            if rank == 0 and step == a.steps:
                peer_sum = [float(i[..., :3].double().sum()) for i in imgs]
            del imgs
            if world > 1:
                dist.barrier()
        e0.record(stream)
        if a.combine == "peer":  # one kernel: sum over NVLink peer memory + resolve + store on rank 0, all AOVs
            e1.record(stream)
            imgs = cam.resolve_peer(range(len(aovs)), root=0, stream=stream)
        elif a.combine in ("scatter", "both"):
            cam.filter_reduce_scatter(stream=stream)
            e1.record(stream)
            out = torch.empty((H, W, 4), dtype=torch.float32, device=dev) if rank == 0 else None
            imgs = []
            for k in range(len(aovs)):  # one reusable 531 MB image
                imgs.append(cam.resolve_gather(k, root=0, stream=stream, out=out))
                if a.combine == "both" and rank == 0 and step == a.steps:
                    scatter_sum.append(float(out[..., :3].double().sum()))
        else:
            cam.filter_reduce(root=0, stream=stream)
            e1.record(stream)
            imgs = [cam.resolve(k, stream=stream) for k in range(len(aovs))] if rank == 0 else []
        e2.record(stream)
        torch.cuda.synchronize()
        t_red, t_res = e0.elapsed_time(e1), e1.elapsed_time(e2)
        st = cam.filter_stats()
        if step > 0:
            times.append((t_acc, t_red, t_res, st["splats"], t_peer))
        del imgs
    t = torch.tensor([[x[0], x[1], x[2], x[3], x[4]] for x in times], dtype=torch.float64, device=dev).mean(0)
    tmax = t.clone()
    tsum = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        splats = float(tsum[3])
        total_ms = float(tmax[0] + tmax[1] + tmax[2])
        block_gb = W * H * (4 * len(aovs) + 1) * 4 / 1e9
        print("C5 " + json.dumps({"tag": a.tag, "order": a.order, "camera": "thinlens" if a.thin else "po", "emulate_world": a.emulate_world, "config": f"C5 {W}x{H}x{a.spp}spp, {len(aovs)} AOVs (9 gaussian RGBA + 1 closest), {world} GPU(s)", "combine": a.combine,
                          "partition": f"hashed {a.tile}x{a.tile} pixel tiles", "accumulate_ms": float(tmax[0]),
                          "reduce_ms": float(tmax[1]), "resolve_gather_ms": float(tmax[2]), "splats": splats, "splats_per_s": splats / (total_ms * 1e-3),
                          "framebuffer_block_GB": block_gb, "reduce_GBps_per_rank": block_gb / (float(tmax[1]) * 1e-3) if world > 1 and float(tmax[1]) > 0 else None,
                          "peer_combine_resolve_gather_ms": float(tmax[4]) if a.combine == "both" else None,
                          "peer_vs_scatter_image_sums_rel_diff": (max(abs(x - y) / max(abs(y), 1e-30) for x, y in zip(peer_sum, scatter_sum)) if peer_sum and scatter_sum else None)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
