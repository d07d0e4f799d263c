"""Build liblentil_b200.so (hand-written CUDA for sm_100a + the C-ABI host code) in-tree with nvcc.

    python -m pota_b200.build [--force] [--no-unrolled] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to pota_b200/csrc/build/, the library to
pota_b200/liblentil_b200.so (git-ignored, travels to the GPU box with the tree).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen" + os.environ.get("LB_BUILD_TAG", ""))
OBJ = os.path.join(CSRC, "build" + os.environ.get("LB_BUILD_TAG", ""))
LIB = os.environ.get("LB_LIB_OUT") or os.path.join(HERE, "liblentil_b200.so")  # LB_LIB_OUT: tuning variants
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", CSRC]

# translation unit -> extra flags.  The two -fmad=false units take integer/branch decisions from
# float/double expressions that must be evaluated exactly as the reference's host code does.
UNITS = {
    "camera_kernels.cu": [],
    "filter_kernels.cu": [],
    "setup_kernels.cu": ["-fmad=false"],
    "filter_classify.cu": ["-fmad=false"],
    "thinlens_kernels.cu": ["-fmad=false"],
    "lentil_host.cu": [],
    "filter_host.cu": ["-I", "/usr/include"],
    "microbench.cu": [],
}


def _sha(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("command failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)
    return r


def generate(unrolled: bool, verbose: bool = False):
    """Regenerate csrc/gen/ from the committed lens pack."""
    from .lensgen import emit

    os.makedirs(GEN, exist_ok=True)
    emit.emit_tables(os.path.join(GEN, "lens_pack_data.inc"))
    gen_units = []
    if unrolled and os.path.exists(os.path.join(HERE, "lensgen", "emit_cuda.py")):
        from .lensgen import emit_cuda

        gen_units = emit_cuda.emit_cuda(GEN)
    return gen_units


def build(force: bool = False, unrolled: bool = True, verbose: bool = False, jobs: int | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    gen_units = generate(unrolled, verbose)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(GEN, f) for f in os.listdir(GEN) if f.endswith((".h", ".cuh", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "lentil_b200.h"))
    hdr_sha = _sha(headers)
    manifest_path = os.path.join(OBJ, "manifest.json")
    manifest = {}
    if os.path.exists(manifest_path) and not force:
        with open(manifest_path) as f:
            manifest = json.load(f)
    units = {os.path.join(CSRC, k): v for k, v in UNITS.items()}
    if gen_units:
        for g in gen_units:
            units[g] = ["-DLB_HAVE_UNROLLED", *os.environ.get("LB_EXTRA_NVCC_FLAGS", "").split()]
    else:
        units[os.path.join(CSRC, "unrolled_none.cu")] = []
    todo, objs = [], []
    for src, extra in units.items():
        obj = os.path.join(OBJ, os.path.basename(src).replace(".cu", ".o"))
        objs.append(obj)
        key = _sha([src]) + hdr_sha + " ".join(extra)
        if manifest.get(obj) != key or not os.path.exists(obj):
            todo.append((src, obj, extra, key))

    def compile_one(item):
        src, obj, extra, key = item
        _run([NVCC, *ARCH, *COMMON, *extra, "-c", src, "-o", obj], verbose)
        return obj, key

    if todo:
        with ThreadPoolExecutor(jobs or os.cpu_count() or 4) as ex:
            for obj, key in ex.map(compile_one, todo):
                manifest[obj] = key
    link_key = hashlib.sha256("".join(manifest.get(o, "") for o in objs).encode()).hexdigest()
    if todo or manifest.get("__lib__") != link_key or not os.path.exists(LIB):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl"], verbose)
        manifest["__lib__"] = link_key
    with open(manifest_path, "w") as f:
        json.dump(manifest, f)
    return LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--no-unrolled", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    print(build(a.force, not a.no_unrolled, a.verbose, a.jobs))


if __name__ == "__main__":
    main()
