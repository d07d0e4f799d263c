#!/usr/bin/env python
"""bench.py — headline benchmark of the lentil hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Prints ONE JSON line.  Headline metric (BASELINE.json configs[1]): camera rays/s for a 3840x2160 x 64 spp
frame, one lens, fixed f-stop/focus.  A "step" is one pass of camera_create_ray over that frame
(530 841 600 rays = 1 592 524 800 forward traces).  `value` is measured with inputs and outputs
resident in HBM; `e2e` drives the same frame through the host-buffer C-ABI call (pinned host memory,
H2D and D2H inside the timed region).

The other half of BASELINE.json's metric, redistributed splats/s (configs[2]: 1920x1080, 16 spp,
image-bokeh kernel), is measured in the same run and reported as FLAT keys inside the objects the driver
keeps: `roofline.splat_*` (device-resident value, executed FMA-slot fraction, reduction traffic),
`e2e.splat_*` (host buffers through lb_filter_accumulate_host + lb_imager_resolve_host),
`cpu_baseline.splat_*` (the compiled reference's filter_pixel on the host cores) -- and once more in the
`summary` string at the END of the line, which survives in the driver's stdout tail.

N > 1: one process per GPU.  Camera rays shard with no collective (every rank traces its own frame's
worth of samples, weak scaling).  The splat frame is ONE frame for all ranks (strong scaling): source
samples are dealt to the ranks in hashed pixel tiles, every rank accumulates a full-frame
partial, lb_filter_reduce_scatter (one ncclReduceScatter per plane over NVLink) leaves every rank owning a
pixel slab, which it resolves itself; lb_imager_resolve_gather collects the slabs on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from pota_b200 import abi, workloads  # noqa: E402

FRAME_W, FRAME_H, FRAME_SPP = 3840, 2160, 64
SPLAT_W, SPLAT_H, SPLAT_SPP = 1920, 1080, 16
SPLAT_GRID = (8, 4)  # discs whose bokeh stays inside the frame (out-of-frame splats burn 5x attempts, lentil_filter.cpp:282-287)
LENS_MODEL = 5  # asahi__takumar__1969__50mm stand-in: the pack's double-Gauss 50 mm
CHUNK_RAYS = FRAME_W * FRAME_H * 4  # 33 177 600 rays per call on the e2e path (4 spp of the frame)
IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")
CPU_SAMPLE_PER_THREAD = 200_000  # rays per host thread in the CPU legs (~1-2 s with the compiled reference)
CPU_SPLAT_SOURCES_PER_THREAD = 100  # redistributed source samples per host thread in the CPU splat leg (2000 splats each, ~10 s in all)
SPLAT_TILE = 1  # tile edge (pixels) of the hashed multi-GPU source-sample partition: pixel granularity balances ~10 px highlights to 1.05x the mean


def camera_params():
    return abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_POLYNOMIAL_OPTICS, lens_model=LENS_MODEL, fstop=2.8, focus_dist=150.0)


def splat_params():
    return abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_POLYNOMIAL_OPTICS, lens_model=LENS_MODEL, fstop=1.4, focus_dist=35.0,
                                     bidir_sample_mult=10, bokeh_enable_image=1)


class ClockSampler:
    """SM clock + throttle reasons sampled every 20 ms through NVML while the timed region runs
    (nvidia-smi -lms as the fallback: its start-up alone can outlast a sub-second timed region)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.proc = None
        self._stop = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _poll_nvml(self):
        n = self._nvml
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def _read_smi(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.sm.append(float(c[0]))
                self.max_mhz = max(self.max_mhz or 0.0, float(c[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def __enter__(self):
        if self._nvml is not None:
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
def cpu_camera_baseline(nthreads: int, n_rays: int):
    """Oracle (CPU restatement of the reference) on a bounded sample of the same frame."""
    from oracle import orc, ref

    # one sample per pixel of a coarser 16:9 grid of ~n_rays pixels: covers the sensor exactly like the frame does
    w = int(np.ceil((n_rays * 16 / 9) ** 0.5))
    h = -(-n_rays // w)
    ins = workloads.camera_samples(w, h, 1, "cpu", 0, n_rays, "linear")
    arrs = [ins[k].numpy() for k in IN_KEYS]
    # iteration statistics for the roofline's algorithmic flop count: the oracle keeps counters
    ocam = orc.OracleCamera(camera_params())
    m = min(n_rays, 100_000)
    ocam.create_rays(*[a[:m] for a in arrs], nthreads=nthreads)
    c = ocam.counters()
    # timed leg: the reference's own sources when oracle/_ref was built (kind "reference"), else the oracle ("port")
    kind = "reference" if ref.available() else "port"
    cam = ref.RefCamera(camera_params()) if kind == "reference" else ocam
    t0 = time.perf_counter()
    out = cam.create_rays(*arrs, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return dict(rays_per_s=n_rays / dt, seconds=dt, newton_its_per_trace=c["fw_newton_its"] / max(c["fw_traces"], 1),
                traces_per_ray=c["fw_traces"] / m, dead_fraction=float((out["weight"][0] == 0).mean()), kind=kind)


def cpu_splat_baseline(nthreads: int):
    """filter_pixel + driver_process_bucket of the reference's own sources (oracle/_ref, kind "reference"; the oracle port
    otherwise) on a bounded sub-sample of the C3 frame's SOURCE samples: whole pixels (16 samples each, as Arnold hands them
    over) that hold highlight samples, until CPU_SPLAT_SOURCES_PER_THREAD redistributed samples per host thread are collected;
    each is redistributed into up to 2000 splats.  Threads share the framebuffers the way Arnold's bucket threads do
    (lentil.h:828-829).  The reference keeps no counters: the splat count is read back from a `lentil_debug` AOV, which
    accumulates samples * (inv_density / samples) per splat (lentil_filter.cpp:209-211,295-298), i.e. inv_density per splat."""
    import torch

    from oracle import orc, ref

    kind = "reference" if ref.available() else "port"
    img = workloads.disc_bokeh_image(250)
    p = splat_params()
    cam = ref.RefCamera(p, img) if kind == "reference" else orc.OracleCamera(p, img)
    W, H, spp = SPLAT_W, SPLAT_H, SPLAT_SPP
    want = CPU_SPLAT_SOURCES_PER_THREAD * nthreads
    fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, "cpu", first=0, count=W * H * spp // 4, grid=SPLAT_GRID)  # top quarter: one disc row
    hit = fr["rgba"][:, 3] > 0
    pix = torch.unique(torch.nonzero(hit).flatten() // spp)
    per_pixel = hit.reshape(-1, spp).sum(dim=1)[pix]
    take = int(torch.searchsorted(torch.cumsum(per_pixel, 0), want).item()) + 1
    pix = pix[:take]
    n_src = int(per_pixel[:take].sum())
    idx = (pix[:, None] * spp + torch.arange(spp)[None, :]).reshape(-1)
    a = [fr[k][idx].contiguous().numpy() for k in ("px", "py", "rgba", "pos_cs")]
    aovs = [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA), ("lentil_debug", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_LENTIL_DEBUG)]
    if kind == "reference":
        cam.filter_begin(W, H, aovs, spp=spp)
    else:
        cam.filter_begin(W, H, aovs)
    t0 = time.perf_counter()
    cam.filter_accumulate(*a, 1.0 / spp, nthreads=nthreads)
    img_out = cam.resolve(0)
    dt = time.perf_counter() - t0
    splats = float(cam.buffers(1)[0][..., 0].sum(dtype=np.float64)) * spp
    return dict(kind=kind, seconds=dt, sources=n_src, samples=int(idx.numel()), splats=splats, splats_per_s=splats / dt,
                finite=bool(np.isfinite(img_out).all()),
                sample=f"{n_src} redistributed source samples ({int(idx.numel())} samples of {take} whole pixels of the same 1920x1080x16spp frame), "
                       f"{splats:.0f} splats, {nthreads} threads sharing the framebuffers; "
                       + ("reference = /root/reference/src filter_pixel + driver_process_bucket compiled behind oracle/shims (oracle/_ref)" if kind == "reference"
                          else "port = oracle/lentil_oracle.cpp"))


def kernel_source_hash() -> str:
    """Hash of everything that decides what the dominant kernel's SASS is: profiles/kernel_traffic.json records the hash its
    ncu capture was taken at, so a stale DRAM-traffic figure is detected instead of silently reported."""
    import hashlib

    h = hashlib.sha256()
    for rel in ("pota_b200/csrc/camera_kernels.cuh", "pota_b200/csrc/lens_device.cuh", "pota_b200/csrc/lens_table.h", "pota_b200/lensgen/emit_cuda.py",
                "pota_b200/lensgen/emit_folded.py", f"pota_b200/lenses/{LENS_ID}.json"):
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


LENS_ID = "asahi__takumar__1969__50mm"
WORKLOAD = (f"camera_create_ray {FRAME_W}x{FRAME_H}x{FRAME_SPP}spp, {LENS_ID} (pack double-Gauss 50mm), f/2.8, focus 150cm")
SPLAT_WORKLOAD = (f"bidirectional redistribution {SPLAT_W}x{SPLAT_H}x{SPLAT_SPP}spp synthetic highlight frame, 250x250 image-bokeh kernel, "
                  f"{SPLAT_GRID[0]}x{SPLAT_GRID[1]} emissive discs at z=-75, f/1.4 focus 35, bidir_sample_mult 10, 1 RGBA AOV")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of both paths on the host cores (rank 0 only)."""
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    n = CPU_SAMPLE_PER_THREAD * nthreads  # bounded sample per step (a few seconds of CPU work)
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_camera_baseline(nthreads, n)
        if i >= args.warmup:
            times.append(r["seconds"])
    v = n / float(np.mean(times))
    sp = cpu_splat_baseline(nthreads)  # once: ~10 s of CPU work
    sample = f"{n} rays per step on a coarser 16:9 pixel grid covering the same sensor, {nthreads} threads"
    line = {
        "impl": "reference", "metric": "camera_rays_per_s", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "splat_workload": SPLAT_WORKLOAD, "splat_sample": sp["sample"]},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": nthreads, "kind": r["kind"], "sample": sample,
                         "splat_value": sp["splats_per_s"], "splat_unit": "splats/s", "splat_kind": sp["kind"], "splat_seconds": sp["seconds"],
                         "splat_sample": sp["sample"]},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "splat_value": sp["splats_per_s"], "splat_unit": "splats/s"},
        "roofline": {"splat_value": sp["splats_per_s"], "splat_unit": "splats/s"},
        "summary": f"reference CPU, {nthreads} threads: camera {v:.4g} rays/s; redistribution {sp['splats_per_s']:.4g} splats/s ({sp['kind']})",
    }
    print(json.dumps(line), flush=True)


def parity_sample(cam, dev):
    """A small parity statement for the bench line (the tests are the gate): 50 000 rays of the bench camera against the oracle."""
    from oracle import orc

    n = 50_000
    ins = workloads.camera_samples(300, 169, 1, "cpu", 0, n, "linear")
    o = orc.OracleCamera(camera_params()).create_rays(*[ins[k].numpy() for k in IN_KEYS], nthreads=os.cpu_count() or 1)
    g = {k: v.cpu().numpy() for k, v in cam.create_rays(*[ins[k].to(dev) for k in IN_KEYS]).items()}
    live = (o["weight"][0] != 0) & (o["tries"] == g["tries"])
    out = {"rays": n, "tries_equal": float((o["tries"] == g["tries"]).mean()), "tolerance": 1e-4}
    for k in ("origin", "dir"):
        e = np.abs(g[k][:, live] - o[k][:, live]).max(axis=0) / np.maximum(np.abs(o[k][:, live]).max(axis=0), 1e-30)
        out[f"{k}_within_tol"] = float((e <= 1e-4).mean())
        out[f"{k}_median_rel_err"] = float(np.median(e))
    # differentials: tolerance of record = finite-difference quantisation units, ulp(|vector|_inf) / 1e-3 (tests/test_camera_gpu.py)
    med, p99 = [], []
    for k, base in (("dOdx", "origin"), ("dOdy", "origin"), ("dDdx", "dir"), ("dDdy", "dir")):
        unit = np.spacing(np.abs(o[base][:, live]).max(axis=0).astype(np.float32)) / 1e-3
        d = np.abs(g[k][:, live] - o[k][:, live]).max(axis=0) / unit
        med.append(float(np.median(d)))
        p99.append(float(np.quantile(d, 0.99)))
    out["differentials_median_units"] = max(med)
    out["differentials_p99_units"] = max(p99)
    out["differentials_bound"] = "median <= 1, p99 <= 4 quantisation units (unit = float32 ulp of the traced vector / 1e-3 step); FD baseline 16x"
    return out


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frame-scale", type=float, default=1.0, help="debug: scale the spp of both workloads")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-splat", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-thinlens", action="store_true")
    ap.add_argument("--skip-crypto", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from pota_b200.camera import RAY_OUT_FIELDS, Camera, lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    spp = max(1, int(round(FRAME_SPP * args.frame_scale)))
    n_rays = FRAME_W * FRAME_H * spp
    cam = Camera(camera_params(), device=local_rank)
    work = cam.lens_work

    # ---- device-resident frame: inputs generated on the device chunk by chunk ----------------------
    ins = {k: torch.empty(n_rays, dtype=torch.float32, device=dev) for k in IN_KEYS}
    first_sample = rank * n_rays  # every rank traces its own frame's worth of samples (weak scaling)
    gen = FRAME_W * FRAME_H
    for s0 in range(0, n_rays, gen):
        m = min(gen, n_rays - s0)
        c = workloads.camera_samples(FRAME_W, FRAME_H, spp * world, dev, first_sample + s0, m, "pixel")
        for k in IN_KEYS:
            ins[k][s0:s0 + m] = c[k]
        del c
    out = {k: torch.empty((3, n_rays), dtype=torch.float32, device=dev) for k in RAY_OUT_FIELDS}
    stream = torch.cuda.current_stream()

    def step_device():
        cam.create_rays(*[ins[k] for k in IN_KEYS], ray_id_base=first_sample, out=out, stream=stream)

    for _ in range(args.warmup):
        step_device()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        ev[0].record(stream)
        for i in range(args.steps):
            step_device()
            ev[i + 1].record(stream)
        barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_s = max_over_ranks(sum(step_ms) * 1e-3)
    value = world * n_rays * args.steps / total_s
    kernel_ms = float(np.mean(step_ms))
    dead = float((out["weight"][0, : 1 << 22] == 0).float().mean().item())
    clock_summary = clocks.summary()

    # ---- e2e: same frame through the host-buffer C-ABI call -------------------------------------------
    e2e = None
    if not args.skip_e2e:
        chunk = min(CHUNK_RAYS, n_rays)
        h_in = {k: torch.empty(chunk, dtype=torch.float32).pin_memory() for k in IN_KEYS}
        h_out = {k: torch.empty((3, chunk), dtype=torch.float32).pin_memory() for k in RAY_OUT_FIELDS}
        for k in IN_KEYS:
            h_in[k].copy_(ins[k][:chunk])
        calls = -(-n_rays // chunk)
        # warm-up (allocates the staging ring) doubling as a check: the host path must reproduce the device path bit for bit
        cam.create_rays_host(*[h_in[k] for k in IN_KEYS], out=h_out, ray_id_base=first_sample)
        checked = all(bool(torch.equal(h_out[k].to(dev), out[k][:, :chunk])) for k in RAY_OUT_FIELDS)
        assert checked, "host path and device path disagree"
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 2))
        for _ in range(e2e_steps):
            for cidx in range(calls):
                m = min(chunk, n_rays - cidx * chunk)
                cam.create_rays_host(*[h_in[k][:m] for k in IN_KEYS], out=h_out, ray_id_base=first_sample + cidx * chunk)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n_rays * e2e_steps / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": n_rays * 24, "d2h_bytes_per_step": n_rays * 76,
               "host_bytes_delivered_per_step": n_rays * 84,
               "d2h_note": "84 B/ray land in the caller's buffers; 76 cross the link: the three weight channels are one number, planes 1-2 are filled on the host",
               "ms_per_step": e2e_s / e2e_steps * 1e3, "calls_per_step": calls, "host_memory": "pinned", "matches_device_path": checked}
        # the same call with PAGEABLE caller memory (what a CPU renderer's own buffers are): one call of the frame
        p_in = {k: h_in[k].numpy().copy() for k in IN_KEYS}
        p_out = {k: np.empty((3, chunk), np.float32) for k in RAY_OUT_FIELDS}
        cam.create_rays_host(*[p_in[k] for k in IN_KEYS], out=p_out, ray_id_base=first_sample)  # touches the pages
        barrier()
        t0 = time.perf_counter()
        cam.create_rays_host(*[p_in[k] for k in IN_KEYS], out=p_out, ray_id_base=first_sample)
        barrier()
        e2e["pageable_value"] = world * chunk / max_over_ranks(time.perf_counter() - t0)
        e2e["pageable_note"] = f"one call of {chunk} rays with pageable numpy buffers (the driver stages the copies), same unit"
        del p_in, p_out
        # what bounds it: the host link.  Plain pinned-memory copies of this box, device-to-host alone and with a
        # host-to-device copy running beside it (the e2e path moves 84 B out and 24 B in per ray, full duplex).
        # All ranks copy at the same time (barrier first): at N > 1 the aggregate is what the host side sustains.
        nb = 256 << 20
        hb0, hb1 = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
        db0, db1 = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def copy_rate(duplex):
            best = 0.0
            for _ in range(3):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(s1):
                    a.record(s1)
                    for _ in range(4):
                        hb0.copy_(db0, non_blocking=True)
                    b.record(s1)
                if duplex:
                    with torch.cuda.stream(s2):
                        for _ in range(2):
                            db1.copy_(hb1, non_blocking=True)
                torch.cuda.synchronize()
                best = max(best, 4 * nb / (a.elapsed_time(b) * 1e-3) / 1e9)
            return best

        d2h_alone, d2h_duplex = copy_rate(False), copy_rate(True)
        achieved = n_rays * 76 * e2e_steps / e2e_s / 1e9
        e2e.update({"link_d2h_copy_gbs": d2h_alone, "link_d2h_copy_gbs_with_h2d_beside": d2h_duplex, "link_d2h_achieved_gbs_per_gpu": achieved,
                    "link_frac_of_copy_rate": achieved / max(d2h_duplex, 1e-9),
                    "link_aggregate_d2h_copy_gbs": sum_over_ranks(d2h_duplex), "link_aggregate_d2h_achieved_gbs": world * achieved,
                    "link_note": "pinned 256 MiB cudaMemcpyAsync on this box, all ranks copying at the same time; the e2e path is bound by the "
                                 "device-to-host direction of the host link"})
        try:
            e2e["host_cpus_visible"] = len(os.sched_getaffinity(0))
        except AttributeError:
            pass
        del hb0, hb1, db0, db1
        del h_in, h_out
    del ins, out
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    import ctypes

    peak = ctypes.c_double()
    lib().lb_bench_fp32_peak(local_rank, ctypes.byref(peak))
    cpu = None
    k_its, traces_per_ray = 4.0, 3.0
    if rank == 0 and not args.skip_cpu:
        nthreads = os.cpu_count() or 1
        cpu = cpu_camera_baseline(nthreads, CPU_SAMPLE_PER_THREAD * nthreads)
        k_its, traces_per_ray = cpu["newton_its_per_trace"], cpu["traces_per_ray"]
    flop_per_ray = workloads.camera_ray_flops(work, k_its, traces_per_ray)
    algorithmic_tflops = flop_per_ray * n_rays / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # EXECUTED work: what the unrolled kernel issues to the FMA pipe for the polynomials -- the FMUL2/FFMA2 of the
    # wavelength-folded bodies; a packed instruction is two lane-operations serving two rays, i.e. one per ray.  (The
    # algorithmic count of SURVEY.md §8d charges every term its full monomial and every wavelength power: the kernel shares
    # monomials, folds the wavelength and skips the dead pred_dx/pred_dy, so that ratio exceeds 1 and is kept as frac_algorithmic.)
    from pota_b200.lensgen import emit_cuda
    from pota_b200.lensgen.pack import load_pack

    st = emit_cuda.lens_unit(load_pack()[camera_params().lens_model])[1]
    slots_per_ray = traces_per_ray * (k_its * sum(st["ap_jac2_imm"]) + sum(st["out5_2_imm"]))
    executed_tflops = 2.0 * slots_per_ray * n_rays / (kernel_ms * 1e-3) / 1e12  # an FMA-pipe slot is rated 2 flop in the peak
    traffic, traffic_note = None, "no ncu capture recorded"
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            kt = json.load(f)["k_create_rays"]
        if kt.get("kernel_source_hash") == kernel_source_hash():
            traffic = kt["dram_bytes_per_ray"] * n_rays
            traffic_note = f"{kt['dram_bytes_per_ray']:.2f} B/ray (dram read+write) from {kt['source']}; algorithmic 108 B/ray"
        else:
            traffic_note = f"stale: kernel sources changed since {kt['source']} (recorded {kt.get('kernel_source_hash')}, now {kernel_source_hash()})"
            print("bench.py: " + traffic_note, file=sys.stderr)
    except (OSError, KeyError, ValueError) as e:
        traffic_note = f"profiles/kernel_traffic.json unreadable: {e}"
    hbm_gbs = n_rays * 108 / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "fp32", "achieved": executed_tflops, "peak": peak.value, "unit": "TFLOP/s", "frac": executed_tflops / max(peak.value, 1e-9),
                "traffic": traffic, "traffic_note": traffic_note,
                "frac_note": "executed FMA-pipe lane-slots of the polynomial bodies x 2 flop / measured FFMA peak",
                "achieved_algorithmic": algorithmic_tflops, "frac_algorithmic": algorithmic_tflops / max(peak.value, 1e-9),
                "fma_slots_per_ray": slots_per_ray, "flop_per_ray_algorithmic": flop_per_ray, "newton_its_per_trace": k_its, "traces_per_ray": traces_per_ray,
                "peak_source": "lb_bench_fp32_peak (register FFMA chains) measured in this run; nominal 148*128*2*1.965 GHz = 74.5",
                "hbm_achieved_gbs": hbm_gbs, "hbm_peak_gbs": hbm_peak, "hbm_frac": hbm_gbs / hbm_peak, "hbm_bytes_per_ray": 108,
                "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
    parity = None
    if rank == 0 and not args.skip_cpu:
        parity = parity_sample(cam, dev)

    # ---- the other half of the metric: redistribution (configs[2]) --------------------------------------
    splat = None
    launches = args.steps
    if not args.skip_splat:
        splat = bench_splat(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks, peak.value)
    cpu_splat = None
    if rank == 0 and not args.skip_cpu and not args.skip_splat:
        cpu_splat = cpu_splat_baseline(os.cpu_count() or 1)
    thin = None
    if not args.skip_thinlens:
        thin = bench_thinlens(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks, hbm_peak)
    crypto = None
    if not args.skip_crypto:
        crypto = bench_crypto(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks)

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu_obj = None
        if cpu is not None:
            cpu_obj = {"value": cpu["rays_per_s"], "unit": "rays/s", "cores": cores, "kind": cpu["kind"],
                       "sample": f"{CPU_SAMPLE_PER_THREAD * cores} rays on a coarser 16:9 pixel grid covering the same sensor, all host threads; "
                                 + ("reference = /root/reference/src compiled behind oracle/shims (oracle/_ref)" if cpu["kind"] == "reference"
                                    else "port = oracle/lentil_oracle.cpp (FP64 restatement)")}
            if cpu_splat is not None:
                cpu_obj.update({"splat_value": cpu_splat["splats_per_s"], "splat_unit": "splats/s", "splat_kind": cpu_splat["kind"],
                                "splat_seconds": cpu_splat["seconds"], "splat_sample": cpu_splat["sample"]})
        config = {"workload": f"camera_create_ray {FRAME_W}x{FRAME_H}x{spp}spp per GPU ({n_rays} rays = {3 * n_rays} forward traces), "
                              f"{LENS_ID} (pack double-Gauss 50mm), f/2.8, focus 150cm",
                  "l2": "inputs (12.7 GB) and outputs (44.6 GB) exceed L2, no flush needed", "kernel": cam.kernel_kind, "dead_ray_fraction": dead}
        if splat is not None:  # flat keys inside the objects the driver keeps
            config.update({"splat_workload": SPLAT_WORKLOAD, "splat_partition": splat["partition"]})
            roofline.update({f"splat_{k}": v for k, v in splat["flat"].items()})
            if e2e is not None and splat.get("e2e"):
                e2e.update({f"splat_{k}": v for k, v in splat["e2e"].items()})
        summary = f"N={world}: camera {value:.4g} rays/s"
        if e2e is not None:
            summary += f", e2e {e2e['value']:.4g}"
        if splat is not None:
            summary += f"; redistribution {splat['flat']['value']:.4g} splats/s ({splat['flat']['ms']:.2f} ms/frame, strong scaling)"
            if splat.get("e2e"):
                summary += f", e2e {splat['e2e']['value']:.4g}"
        if cpu_obj is not None:
            summary += f"; CPU {cpu_obj['kind']} {cores} threads: {cpu_obj['value']:.4g} rays/s"
            if "splat_value" in cpu_obj:
                summary += f", {cpu_obj['splat_value']:.4g} splats/s"
        line = {
            "metric": "camera_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": launches, "clocks": clock_summary, "roofline": roofline,
            "cpu_baseline": cpu_obj, "parity": parity,
            "splat": None if splat is None else splat["detail"], "thinlens": thin, "cryptomatte": crypto,
            "splat_value": None if splat is None else splat["flat"]["value"], "splat_unit": "splats/s",
            "splat_ms_per_step": None if splat is None else splat["flat"]["ms"],
            "splat_e2e_value": None if splat is None or not splat.get("e2e") else splat["e2e"]["value"],
            "splat_cpu_baseline_value": None if cpu_splat is None else cpu_splat["splats_per_s"],
            "summary": summary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_splat(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks, fp32_peak):
    """configs[2]: ONE 1920x1080x16spp frame redistributed by all ranks (strong scaling)."""
    import ctypes

    import torch
    import torch.distributed as dist

    from pota_b200.camera import Camera, lib
    from pota_b200.lensgen import emit_folded
    from pota_b200.lensgen.pack import load_pack

    spp = max(1, int(round(SPLAT_SPP * args.frame_scale)))
    img = workloads.disc_bokeh_image(250)
    cam = Camera(splat_params(), bokeh=img, device=local_rank)
    total = SPLAT_W * SPLAT_H * spp
    # source samples: hashed pixel tiles over the ranks, the 16 samples of a pixel together
    if world > 1:
        mine = workloads.tile_partition(SPLAT_W, SPLAT_H, spp, rank, world, tile=SPLAT_TILE, device=dev)
        fr = workloads.highlight_frame(SPLAT_W, SPLAT_H, spp, cam.state.tan_fov, dev, grid=SPLAT_GRID, samples=mine)
        partition = (f"hashed {SPLAT_TILE}x{SPLAT_TILE} pixel tiles of source samples over {world} ranks; " +
                     ("combine + resolve + gather on rank 0 in one kernel over NVLink peer memory (lb_imager_resolve_peer)"
                      if os.environ.get("LB_BENCH_COMBINE", "peer") == "peer" else "ncclReduceScatter per plane, per-rank resolve, gather on rank 0"))
    else:
        fr = workloads.highlight_frame(SPLAT_W, SPLAT_H, spp, cam.state.tan_fov, dev, grid=SPLAT_GRID)
        partition = "single GPU: whole frame, no collective"
    aovs = [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)]
    cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
    if world > 1:
        uid = [Camera.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        cam.comm_init(world, rank, uid[0])
    stream = torch.cuda.current_stream()
    out_img = torch.empty((SPLAT_H, SPLAT_W, 4), dtype=torch.float32, device=dev)

    # N > 1 combine: "peer" = lb_imager_resolve_peer, ONE kernel that sums the ranks' partial planes over NVLink peer memory,
    # resolves and stores the image on rank 0; "scatter" = ncclReduceScatter + per-rank resolve + ncclSend/Recv gather
    combine = os.environ.get("LB_BENCH_COMBINE", "peer")

    def step():
        cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
        cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, stream=stream)
        if world > 1 and combine == "peer":
            img = cam.resolve_peer([0], root=0, stream=stream)
            return None if img is None else img[0]
        if world > 1:
            cam.filter_reduce_scatter(stream=stream)
        return cam.resolve_gather(0, root=0, stream=stream, out=out_img)

    steps = max(1, min(args.steps, 5))
    if world > 1 and combine == "peer":
        try:  # the failure is collective (every rank gets the error), so every rank falls back together
            step()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"bench.py: peer-memory combine unavailable ({e}); using reduce-scatter + gather", file=sys.stderr)
            combine = "scatter"
            partition = partition.replace("combine + resolve + gather on rank 0 in one kernel over NVLink peer memory (lb_imager_resolve_peer)",
                                          "ncclReduceScatter per plane, per-rank resolve, gather on rank 0")
    for _ in range(2):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        last_img = step()
    e1.record(stream)
    barrier()
    if rank == 0 and last_img is not None:
        out_img = last_img
    s = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / steps
    st = cam.filter_stats()
    splats = sum_over_ranks(float(st["splats"]))
    attempts = sum_over_ranks(float(st["attempts"]))
    its = sum_over_ranks(float(st["newton_its"]))
    passthrough = sum_over_ranks(float(st["passthrough"]))
    image_energy = float(out_img[..., :3].sum().item()) if rank == 0 else 0.0
    # accumulate alone (no begin / reduce / resolve): where the time of a step goes
    cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
    barrier()
    e0.record(stream)
    cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, stream=stream)
    e1.record(stream)
    barrier()
    acc_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    w = cam.lens_work
    flop = its * (w.F_apxy + w.F_apJ + w.F_out4 + w.F_outJ + 120.0) + attempts * w.F_T
    m = emit_folded.folded_evaluators(load_pack()[LENS_MODEL], imm_lambda=emit_folded.LAMBDA_550)[1]["lt_all_mirror_imm"]
    slots_per_it = 2 * (m["fmul2"] + m["ffma2"]) + m["fmul"] + m["ffma"]  # FMA-pipe lane-slots of the 14 polynomials of one Newton trip
    executed = 2.0 * its * slots_per_it / s / 1e12 / world
    red_peak = ctypes.c_double()
    lib().lb_bench_red_peak(local_rank, 41, ctypes.byref(red_peak))  # 33 MB RGBA plane + 8 MB weight plane, L2-resident
    red_bytes = (splats + passthrough) * 20.0  # 16 B vector reduction + 4 B weight reduction per add and gaussian RGBA AOV
    flat = {"value": splats / s, "unit": "splats/s", "ms": s * 1e3, "accumulate_ms": acc_s * 1e3, "scaling": "strong",
            "frac": executed / max(fp32_peak, 1e-9), "achieved_tflops_per_gpu": executed,
            "frac_note": "executed FMA-pipe lane-slots of the mirror-packed lt_all body x Newton trips x 2 flop / measured FFMA peak, per GPU",
            "frac_algorithmic": flop / s / 1e12 / world / max(fp32_peak, 1e-9),
            "red_gbs_per_gpu": red_bytes / s / 1e9 / world, "red_peak_gbs": red_peak.value, "red_frac": red_bytes / s / 1e9 / world / max(red_peak.value, 1e-9),
            "splats_per_step": splats, "attempts_per_step": attempts, "newton_its_per_attempt": its / max(attempts, 1.0),
            "fma_slots_per_newton_trip": slots_per_it}
    # the accumulate alone, three ways (lb_bench_splat_accum): L2 reductions into L2-resident / HBM-resident planes, and a
    # shared-memory tile with float atomics -- the evidence behind "the accumulate lives in L2" 
    if world == 1:
        v = ctypes.c_double()
        for key, mode, mb in (("accum_l2_red_gsplats", 0, 41), ("accum_hbm_red_gsplats", 0, 1600), ("accum_smem_tile_cas_gsplats", 1, 41)):
            if lib().lb_bench_splat_accum(local_rank, mode, mb, ctypes.byref(v)) == 0:
                flat[key] = v.value
        flat["accum_note"] = ("1e9 splats/s of the accumulate alone (RGBA + weight, 20 B/splat, random pixels): red.global.add.v4.f32 + .f32 into 41 MB "
                              "(L2-resident) / 1600 MB planes, and a 96x96 shared-memory tile with float atomics flushed by vector reductions")
    detail = {"metric": "redistributed_splats_per_s", "value": splats / s, "unit": "splats/s", "ms_per_step": s * 1e3, "scaling": "strong",
              "config": {"workload": SPLAT_WORKLOAD, "partition": partition}, "source_samples": total, "image_energy": image_energy,
              "step": "lb_filter_begin + lb_filter_accumulate + " + ("lb_imager_resolve_peer (combine + resolve + gather in one kernel over NVLink peer memory)"
                                                                      if world > 1 and combine == "peer" else "(lb_filter_reduce_scatter +) lb_imager_resolve_gather")}
    # ---- e2e: the same frame with HOST buffers through lb_filter_accumulate_host + lb_imager_resolve_host (N = 1: the path a CPU
    # renderer's filter / imager nodes would drive; at N > 1 every rank feeds its own share and rank 0 reads the gathered image) ----
    e2e = None
    if not args.skip_e2e:
        host = {k: fr[k].cpu().pin_memory() for k in ("px", "py", "rgba", "pos_cs")}
        res = np.empty((SPLAT_H, SPLAT_W, 4), np.float32)
        res_dev_host = torch.empty((SPLAT_H, SPLAT_W, 4), dtype=torch.float32).pin_memory()

        def step_host():
            cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
            cam.filter_accumulate_host(host["px"], host["py"], host["rgba"], host["pos_cs"], 1.0 / spp)
            if world > 1:
                if combine == "peer":
                    img = cam.resolve_peer([0], root=0, stream=stream)
                else:
                    cam.filter_reduce_scatter(stream=stream)
                    img = [cam.resolve_gather(0, root=0, stream=stream, out=out_img)]
                if rank == 0:
                    res_dev_host.copy_(img[0], non_blocking=True)
                    stream.synchronize()
            else:
                cam.resolve_host(0, res)

        step_host()
        barrier()
        t0 = time.perf_counter()
        hsteps = max(1, min(args.steps, 3))
        for _ in range(hsteps):
            step_host()
        barrier()
        hs = max_over_ranks(time.perf_counter() - t0) / hsteps
        n_mine = int(fr["px"].shape[0])
        e2e = {"value": splats / hs, "unit": "splats/s", "ms": hs * 1e3, "h2d_bytes": sum_over_ranks(float(n_mine * 40)), "d2h_bytes": SPLAT_W * SPLAT_H * 16,
               "note": "px,py,rgba,pos_cs of every source sample host->device (40 B/sample) in 2 Mi-sample chunks beside the kernels, resolved RGBA image device->host"}
        del host
    return {"flat": flat, "detail": detail, "e2e": e2e, "partition": partition}


def bench_thinlens(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks, hbm_peak):
    """Thin-lens camera model (the reference's default camera_type): rays/s against the HBM
    roofline (108 B/ray) and splats/s against the L2 reduction roofline on the same two frames."""
    import ctypes

    import torch

    from pota_b200.camera import RAY_OUT_FIELDS, Camera, lib

    spp = max(1, int(round(16 * args.frame_scale)))
    n = FRAME_W * FRAME_H * spp
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, fstop=2.8, focus_dist=150.0, focal_length_lentil=50.0)
    cam = Camera(p, device=local_rank)
    ins = workloads.camera_samples(FRAME_W, FRAME_H, spp * world, dev, rank * n, n, "pixel")
    out = {k: torch.empty((3, n), dtype=torch.float32, device=dev) for k in RAY_OUT_FIELDS}
    stream = torch.cuda.current_stream()
    args_in = [ins[k] for k in IN_KEYS]
    for _ in range(2):
        cam.create_rays(*args_in, ray_id_base=rank * n, out=out, stream=stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, args.steps)
    e0.record(stream)
    for _ in range(steps):
        cam.create_rays(*args_in, ray_id_base=rank * n, out=out, stream=stream)
    e1.record(stream)
    barrier()
    s = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / steps
    rays = {"value": world * n / s, "unit": "rays/s", "ms_per_step": s * 1e3, "rays_per_step_per_gpu": n,
            "roofline": {"bound": "hbm", "achieved": n * 108 / s / 1e9, "peak": hbm_peak, "unit": "GB/s per GPU", "frac": n * 108 / s / 1e9 / hbm_peak,
                         "bytes_per_ray": 108}}
    del ins, out, args_in
    torch.cuda.empty_cache()
    # bidirectional, same synthetic frame as the PO splat bench
    sspp = max(1, int(round(SPLAT_SPP * args.frame_scale)))
    p2 = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, fstop=1.4, focus_dist=35.0, focal_length_lentil=50.0, bidir_sample_mult=10,
                                   bokeh_enable_image=1)
    cam2 = Camera(p2, bokeh=workloads.disc_bokeh_image(250), device=local_rank)
    total = SPLAT_W * SPLAT_H * sspp
    lo, hi = total * rank // world, total * (rank + 1) // world
    fr = workloads.highlight_frame(SPLAT_W, SPLAT_H, sspp, cam2.state.tan_fov, dev, lo, hi - lo, grid=SPLAT_GRID)
    aovs = [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)]

    def step():
        cam2.filter_begin(SPLAT_W, SPLAT_H, aovs)
        cam2.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / sspp, stream=stream)

    step()
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    s2 = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / steps
    st = cam2.filter_stats()
    splats = sum_over_ranks(float(st["splats"]))
    adds = splats + sum_over_ranks(float(st["passthrough"]))
    red_peak = ctypes.c_double()
    lib().lb_bench_red_peak(local_rank, 41, ctypes.byref(red_peak))
    return {"rays": rays,
            "splat": {"value": splats / s2, "unit": "splats/s", "ms_per_step": s2 * 1e3, "splats_per_step": splats, "accumulate_only": "no reduce/resolve in this leg",
                      "roofline": {"bound": "l2_reduction", "achieved": adds * 20.0 / s2 / 1e9 / world, "peak": red_peak.value, "unit": "GB/s per GPU",
                                   "frac": adds * 20.0 / s2 / 1e9 / world / max(red_peak.value, 1e-9)}},
            "config": {"workload": f"ThinLens camera_type, focal 50 mm: rays {FRAME_W}x{FRAME_H}x{spp}spp per GPU f/2.8 focus 150; "
                                   f"splats {SPLAT_W}x{SPLAT_H}x{sspp}spp highlight frame f/1.4 focus 35, 250x250 image-bokeh kernel"}}


def bench_crypto(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks):
    """Cryptomatte redistribution on the thin-lens splat frame: RGBA + 3 ranked cryptomatte AOVs
    with 4 depth sub-samples per sample, against the same frame with RGBA alone.  The extra cost is table traffic:
    per splat and cryptomatte AOV one scalar reduction (crypto_total_weight) plus, per cached id, one slot probe
    (4 B read or compare-and-swap) and one scalar reduction (4 B)."""
    import torch

    from pota_b200.camera import Camera

    sspp = max(1, int(round(SPLAT_SPP * args.frame_scale)))
    p = abi.CameraParams.defaults(camera_type=abi.LB_CAMERA_THINLENS, fstop=1.4, focus_dist=35.0, focal_length_lentil=50.0, bidir_sample_mult=10,
                                  bokeh_enable_image=1)
    cam = Camera(p, bokeh=workloads.disc_bokeh_image(250), device=local_rank)
    total = SPLAT_W * SPLAT_H * sspp
    lo, hi = total * rank // world, total * (rank + 1) // world
    fr = workloads.highlight_frame(SPLAT_W, SPLAT_H, sspp, cam.state.tan_fov, dev, lo, hi - lo, grid=SPLAT_GRID)
    depth = 4
    aov_sets = {"rgba_only": [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)],
                "rgba_plus_3_crypto": [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA), ("crypto_material00", abi.LB_FILTER_CRYPTO, 0),
                                       ("crypto_object00", abi.LB_FILTER_CRYPTO, 0), ("crypto_asset00", abi.LB_FILTER_CRYPTO, 0)]}
    cr = workloads.crypto_layers(fr, depth, [1, 2, 3], cell=64)
    crypto = dict(depth=depth, count=cr["count"], opacity=cr["opacity"], ids=cr["ids"])
    stream = torch.cuda.current_stream()
    steps = max(1, min(args.steps, 3))
    out = {}
    for name, aovs in aov_sets.items():
        def step():
            cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
            cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / sspp, stream=stream, crypto=crypto if len(aovs) > 1 else None)

        step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        s = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / steps
        st = cam.filter_stats()
        splats = sum_over_ranks(float(st["splats"]))
        out[name] = {"value": splats / s, "unit": "splats/s", "ms_per_step": s * 1e3, "splats_per_step": splats,
                     "crypto_dropped": sum_over_ranks(float(st["crypto_dropped"]))}
    res = None
    if len(aov_sets["rgba_plus_3_crypto"]) > 1:
        res_t0, res_t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res_t0.record(stream)
        img = cam.resolve(1, stream=stream)
        res_t1.record(stream)
        torch.cuda.synchronize()
        res = {"ms": res_t0.elapsed_time(res_t1), "pixels": SPLAT_W * SPLAT_H, "mean_rank0_coverage": float(img[..., 1].mean().item())}
    out["slowdown"] = out["rgba_only"]["value"] / max(out["rgba_plus_3_crypto"]["value"], 1e-9)
    out["ranked_resolve"] = res
    out["config"] = {"workload": f"ThinLens splats {SPLAT_W}x{SPLAT_H}x{sspp}spp highlight frame f/1.4 focus 35, 250x250 image-bokeh kernel; "
                                 f"3 cryptomatte AOVs, {depth} depth sub-samples per sample, 16 id slots per pixel, ids constant over 64x64 pixel blocks"}
    return out


if __name__ == "__main__":
    main()
