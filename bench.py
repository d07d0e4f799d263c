#!/usr/bin/env python
"""bench.py — headline benchmark of the lentil hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Prints ONE JSON line.  Headline metric (BASELINE.json configs[1]): camera rays/s for a 3840x2160 x 64 spp
frame, one lens, fixed f-stop/focus.  A "step" is one pass of camera_create_ray over that frame
(530 841 600 rays = 1 592 524 800 forward traces).  `value` is measured with inputs and outputs
resident in HBM; `e2e` drives the same frame through the host-buffer C-ABI call (pinned host memory,
H2D and D2H inside the timed region).  The secondary object `splat` reports the redistribution path
(configs[2]: 1920x1080, 16 spp, image-bokeh kernel) in splats/s.

N > 1: one process per GPU.  Camera rays shard with no collective (every rank traces its own frame's
worth of samples, weak scaling); the splat frame is split by sample range and combined with
lb_filter_reduce (NCCL over NVLink).
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


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
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
    sample = f"{n} rays per step on a coarser 16:9 pixel grid covering the same sensor, {nthreads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "camera_rays_per_s", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "camera_create_ray 3840x2160x64spp, asahi__takumar__1969__50mm (pack double-Gauss 50mm), f/2.8, focus 150cm",
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": nthreads, "kind": r["kind"], "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


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

        def step_host():
            for cidx in range(calls):
                m = min(chunk, n_rays - cidx * chunk)
                cam.create_rays_host(*[h_in[k][:m] for k in IN_KEYS], out={k: h_out[k] for k in RAY_OUT_FIELDS} if m == chunk else
                                     {k: h_out[k] for k in RAY_OUT_FIELDS}, ray_id_base=first_sample + cidx * chunk)

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
        # what bounds it: the host link.  Plain pinned-memory copies of this box, device-to-host alone and with a
        # host-to-device copy running beside it (the e2e path moves 84 B out and 24 B in per ray, full duplex)
        nb = 256 << 20
        hb0, hb1 = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
        db0, db1 = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def copy_rate(duplex):
            best = 0.0
            for _ in range(3):
                torch.cuda.synchronize()
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
        e2e["link"] = {"d2h_copy_gbs": d2h_alone, "d2h_copy_gbs_with_h2d_beside": d2h_duplex, "d2h_achieved_gbs_per_gpu": achieved,
                       "frac_of_copy_rate": achieved / max(d2h_duplex, 1e-9),
                       "note": "pinned 256 MiB cudaMemcpyAsync on this box; the e2e path is bound by the device-to-host direction"}
        del hb0, hb1, db0, db1
        del h_in, h_out
    del ins, out
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peak = __import__("ctypes").c_double()
    lib().lb_bench_fp32_peak(local_rank, __import__("ctypes").byref(peak))
    cpu = None
    k_its, traces_per_ray = 4.0, 3.0
    if rank == 0 and not args.skip_cpu:
        nthreads = os.cpu_count() or 1
        cpu = cpu_camera_baseline(nthreads, CPU_SAMPLE_PER_THREAD * nthreads)
        k_its, traces_per_ray = cpu["newton_its_per_trace"], cpu["traces_per_ray"]
    flop_per_ray = workloads.camera_ray_flops(work, k_its, traces_per_ray)
    achieved_tflops = flop_per_ray * n_rays / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # what the unrolled kernel actually issues for the polynomials: FMUL(2)/FFMA(2) slots of the shared-monomial bodies
    # (the algorithmic count above charges every term its full monomial, SURVEY.md §8d, so `frac` can exceed 1)
    executed = None
    try:
        from pota_b200.lensgen import emit_cuda
        from pota_b200.lensgen.pack import load_pack

        st = emit_cuda.lens_unit(load_pack()[camera_params().lens_model])[1]
        slots = traces_per_ray * (k_its * sum(st["ap_jac"]) + sum(st["out5"]))
        slot_rate = slots * n_rays / (kernel_ms * 1e-3)
        executed = {"fma_pipe_slots_per_ray": slots, "slots_per_s": slot_rate, "peak_slots_per_s": peak.value * 1e12 / 2.0,
                    "frac": slot_rate / max(peak.value * 1e12 / 2.0, 1.0),
                    "note": "polynomial FMUL/FFMA lane-slots only (packed FMUL2/FFMA2 count two); ncu of the same kernel: FMA pipe 80-83 % active"}
    except Exception as e:  # the lens pack tooling is optional at run time
        executed = {"unavailable": str(e)}
    roofline = {"bound": "fp32", "achieved": achieved_tflops, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved_tflops / max(peak.value, 1e-9),
                "executed": executed,
                "traffic": 106.3 * n_rays, "traffic_source": "ncu --set full of this kernel (profiles/r01_k1_create_rays_ncu.txt): dram read+write = 106.3 B/ray vs 108 B/ray algorithmic",
                "peak_source": "lb_bench_fp32_peak (register FFMA chains) measured in this run; nominal 148*128*2*1.965 GHz = 74.5",
                "flop_per_ray": flop_per_ray, "newton_its_per_trace": k_its, "traces_per_ray": traces_per_ray,
                "hbm": {"achieved": n_rays * 108 / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": n_rays * 108 / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_ray": 108,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}

    # ---- secondary: redistribution (configs[2]) --------------------------------------------------------
    splat = None
    launches = args.steps
    if not args.skip_splat:
        splat = bench_splat(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks)
    thin = None
    if not args.skip_thinlens:
        thin = bench_thinlens(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks, hbm_peak)

    crypto = None
    if not args.skip_crypto:
        crypto = bench_crypto(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks)

    if rank == 0:
        line = {
            "metric": "camera_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"camera_create_ray {FRAME_W}x{FRAME_H}x{spp}spp per GPU ({n_rays} rays = {3 * n_rays} forward traces), "
                                   "asahi__takumar__1969__50mm (pack double-Gauss 50mm), f/2.8, focus 150cm",
                       "l2": "inputs (12.7 GB) and outputs (44.6 GB) exceed L2, no flush needed", "kernel": cam.kernel_kind,
                       "dead_ray_fraction": dead},
            "e2e": e2e, "gpu_launches": launches, "clocks": clock_summary, "roofline": roofline,
            "cpu_baseline": None if cpu is None else {"value": cpu["rays_per_s"], "unit": "rays/s", "cores": os.cpu_count(), "kind": cpu["kind"],
                                                     "sample": f"{CPU_SAMPLE_PER_THREAD * (os.cpu_count() or 1)} rays on a coarser 16:9 pixel grid covering the same sensor, all host threads; "
                                                               + ("reference = /root/reference/src compiled behind oracle/shims (oracle/_ref)" if cpu["kind"] == "reference"
                                                                  else "port = oracle/lentil_oracle.cpp (FP64 restatement)")},
            "splat": splat, "thinlens": thin, "cryptomatte": crypto,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_splat(args, rank, world, local_rank, dev, barrier, max_over_ranks, sum_over_ranks):
    import torch
    import torch.distributed as dist

    from pota_b200.camera import Camera

    spp = max(1, int(round(SPLAT_SPP * args.frame_scale)))
    cam = Camera(splat_params(), bokeh=workloads.disc_bokeh_image(250), device=local_rank)
    total = SPLAT_W * SPLAT_H * spp
    lo, hi = total * rank // world, total * (rank + 1) // world  # strong scaling: the frame's samples are split by range
    fr = workloads.highlight_frame(SPLAT_W, SPLAT_H, spp, cam.state.tan_fov, dev, lo, hi - lo, grid=SPLAT_GRID)
    aovs = [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)]
    cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
    if world > 1:
        uid = [Camera.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        cam.comm_init(world, rank, uid[0])
    stream = torch.cuda.current_stream()

    def step():
        cam.filter_begin(SPLAT_W, SPLAT_H, aovs)
        cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp, stream=stream)
        if world > 1:
            cam.filter_reduce(root=0, stream=stream)
        return cam.resolve(0, stream=stream)

    steps = max(1, min(args.steps, 3))
    step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        img = step()
    e1.record(stream)
    barrier()
    s = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / steps
    st = cam.filter_stats()
    splats = sum_over_ranks(float(st["splats"]))
    attempts = sum_over_ranks(float(st["attempts"]))
    its = sum_over_ranks(float(st["newton_its"]))
    w = cam.lens_work
    flop = its * (w.F_apxy + w.F_apJ + w.F_out4 + w.F_outJ + 120.0) + attempts * w.F_T
    import ctypes

    from pota_b200.camera import lib

    fp32_peak, red_peak = ctypes.c_double(), ctypes.c_double()
    lib().lb_bench_fp32_peak(local_rank, ctypes.byref(fp32_peak))
    lib().lb_bench_red_peak(local_rank, 41, ctypes.byref(red_peak))  # 33 MB RGBA plane + 8 MB weight plane, L2-resident
    # two rooflines: the reverse Newton trace is FP32 bound, the accumulate is reduction (L2 atomic) bound
    passthrough = sum_over_ranks(float(st["passthrough"]))
    red_bytes = (splats + passthrough) * 20.0  # 16 B vector reduction + 4 B weight reduction per add and gaussian RGBA AOV
    roof = {"trace": {"bound": "fp32", "achieved": flop / s / 1e12 / world, "peak": fp32_peak.value, "unit": "TFLOP/s per GPU",
                      "frac": flop / s / 1e12 / world / max(fp32_peak.value, 1e-9)},
            "accumulate": {"bound": "l2_reduction", "achieved": red_bytes / s / 1e9 / world, "peak": red_peak.value, "unit": "GB/s per GPU",
                           "frac": red_bytes / s / 1e9 / world / max(red_peak.value, 1e-9),
                           "peak_source": "lb_bench_red_peak: red.global.add.v4.f32 at random pixels of a 41 MB plane"}}
    return {"roofline": roof, "metric": "redistributed_splats_per_s", "value": splats / s, "unit": "splats/s", "ms_per_step": s * 1e3, "scaling": "strong",
            "config": {"workload": f"bidirectional redistribution {SPLAT_W}x{SPLAT_H}x{spp}spp synthetic highlight frame, 250x250 image-bokeh kernel, "
                                   f"{SPLAT_GRID[0]}x{SPLAT_GRID[1]} emissive discs at z=-75, f/1.4 focus 35, bidir_sample_mult 10, 1 RGBA AOV"},
            "splats_per_step": splats, "attempts_per_step": attempts, "newton_its_per_attempt": its / max(attempts, 1.0),
            "source_samples": SPLAT_W * SPLAT_H * spp, "tflops_algorithmic": flop / s / 1e12,
            "image_energy": float(img[..., :3].sum().item())}


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
