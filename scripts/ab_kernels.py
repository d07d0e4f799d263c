"""A/B timing of kernel variants on one GPU: K1 (camera rays, config C2 at reduced spp) and K2 (PO splats, config C3).

    LB_LIBRARY=<variant .so> [LB_NO_FOLD=1] python scripts/ab_kernels.py [--tag NAME] [--spp 16] [--lens 5] [--thin]

Prints one line per leg; the environment selects the library variant (pota_b200/build.py LB_LIB_OUT / LB_BUILD_TAG)
and, inside one library, the first-generation bodies (LB_NO_FOLD=1).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pota_b200 import abi, workloads  # noqa: E402
from pota_b200.camera import RAY_OUT_FIELDS, Camera  # noqa: E402

IN_KEYS = ("sx", "sy", "dsx", "dsy", "lensx", "lensy")


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--lens", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--thin", action="store_true")
    ap.add_argument("--skip-k1", action="store_true")
    ap.add_argument("--skip-k2", action="store_true")
    ap.add_argument("--focus", type=float, default=35.0, help="K2 leg: focus distance (35 = config C3: 144-pixel bokeh discs at 1080p)")
    ap.add_argument("--disc-radius", type=float, default=0.133, help="K2 leg: radius of the emissive discs (0.133 = config C3)")
    ap.add_argument("--k2-size", default="1920x1080")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = {"tag": a.tag, "lib": os.path.basename(os.environ.get("LB_LIBRARY", "default")), "no_fold": os.environ.get("LB_NO_FOLD", "0"), "lens": a.lens}
    ctype = abi.LB_CAMERA_THINLENS if a.thin else abi.LB_CAMERA_POLYNOMIAL_OPTICS
    if not a.skip_k1:
        n = 3840 * 2160 * a.spp
        cam = Camera(abi.CameraParams.defaults(camera_type=ctype, lens_model=a.lens, fstop=2.8, focus_dist=150.0, focal_length_lentil=50.0), device=0)
        ins = workloads.camera_samples(3840, 2160, a.spp, dev, 0, n, "pixel")
        out = {k: torch.empty((3, n), dtype=torch.float32, device=dev) for k in RAY_OUT_FIELDS}
        args = [ins[k] for k in IN_KEYS]
        s = timed(lambda: cam.create_rays(*args, out=out), a.reps)
        res["k1_rays_per_s"] = n / s
        res["k1_ms"] = s * 1e3
        res["k1_dead"] = float((out["weight"][0] == 0).float().mean())
        res["k1_checksum"] = float(out["dir"].double().sum())
        del ins, out, args
        cam.close()
        torch.cuda.empty_cache()
    if not a.skip_k2:
        W, H = (int(v) for v in a.k2_size.split("x"))
        spp = 16
        res["k2_frame"] = f"{W}x{H}x{spp} focus {a.focus} disc radius {a.disc_radius}"
        cam = Camera(abi.CameraParams.defaults(camera_type=ctype, lens_model=a.lens, fstop=1.4, focus_dist=a.focus, bidir_sample_mult=10,
                                               bokeh_enable_image=1, focal_length_lentil=50.0),
                     bokeh=workloads.disc_bokeh_image(250), device=0)
        fr = workloads.highlight_frame(W, H, spp, cam.state.tan_fov, dev, grid=(8, 4), radius=a.disc_radius)
        aovs = [("RGBA", abi.LB_FILTER_GAUSSIAN, abi.LB_AOV_RGBA)]

        def step():
            cam.filter_begin(W, H, aovs)
            cam.filter_accumulate(fr["px"], fr["py"], fr["rgba"], fr["pos_cs"], 1.0 / spp)

        s = timed(step, 3)
        st = cam.filter_stats()
        # individual repetitions + the SM clock: the persistent kernel's makespan depends on how the last work items fall
        reps = []
        clk = []
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:
            h = None
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            if h is not None:
                clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            torch.cuda.synchronize()
            reps.append(round(e0.elapsed_time(e1), 2))
        res["k2_reps_ms"] = reps
        res["k2_sm_mhz"] = clk
        s = min(s, min(reps) * 1e-3)
        res["k2_splats_per_s"] = st["splats"] / s
        res["k2_ms"] = s * 1e3
        res["k2_splats"] = st["splats"]
        res["k2_tile_splats"] = st["tile_splats"]
        res["k2_redistributed"] = st["redistributed"]
        res["k2_attempts"] = st["attempts"]
        res["k2_its"] = st["newton_its"]
        img = cam.resolve(0)
        res["k2_energy"] = float(img[..., :3].double().sum())
    print("AB " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
