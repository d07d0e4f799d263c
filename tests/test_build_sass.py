"""What the sm_100a build of the library looks like, checked without a GPU (cuobjdump on pota_b200/liblentil_b200.so): the
properties the measured numbers of DESIGN.md §4 rest on.  Resident CTAs per SM follow from the register counts (65 536 registers,
128-thread CTAs): K1 <= 128 registers = 4 CTAs, the polynomial-optics splat kernels <= 96 = 5 CTAs, for every lens of the pack;
the forward kernel runs on the packed FFMA2 / FMUL2 instructions; the splat kernel's mirror-packed bodies use the operand-swap
form (.LO_HI); float accumulation is issued as REDG reductions, never as ATOMG atomics with a discarded result."""
import os
import re
import shutil
import subprocess

import pytest

from pota_b200 import camera

LIB = camera._LIB_PATH
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="cuobjdump or the built library missing")


@pytest.fixture(scope="module")
def resources():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    res, arch = {}, set(re.findall(r"arch = (\S+)", out))
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+)", out):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return res, arch


def _kernels(res, pattern):
    return {k: v for k, v in res.items() if re.search(pattern, k)}


def test_built_for_sm_100a_only(resources):
    _, arch = resources
    assert arch == {"sm_100a"}


def test_register_budgets_of_every_lens(resources):
    res, _ = resources
    k1 = _kernels(res, r"k_create_rays_w550_\d+E")
    k1f = _kernels(res, r"k_create_rays_\d+E")
    k2 = _kernels(res, r"k_filter_splat_w550_\d+E")
    k2f = _kernels(res, r"k_filter_splat_\d+E")
    k2c = _kernels(res, r"k_filter_splat_chroma_\d+E")
    assert len(k1) == len(k1f) == len(k2) == len(k2f) == len(k2c) == 44  # one unrolled unit per lens of the pack
    assert max(r for r, _ in k1.values()) <= 128 and max(r for r, _ in k1f.values()) <= 128   # 4 CTAs of 128 threads per SM
    assert max(r for r, _ in k2.values()) <= 96 and max(r for r, _ in k2f.values()) <= 96 and max(r for r, _ in k2c.values()) <= 96  # 5 CTAs
    # the bench lens: what profiles/r02_final_k{1,2}_*_ncu.txt were captured on
    (r1, s1), = _kernels(res, r"k_create_rays_w550_5E").values()
    (r2, s2), = _kernels(res, r"k_filter_splat_w550_5E").values()
    assert 112 < r1 <= 128 and 80 < r2 <= 96 and s1 <= 64 and s2 <= 64  # 125 / 96 with nvcc 12.9, no spill


def test_shared_kernels(resources):
    res, _ = resources
    cls = _kernels(res, r"k_filter_classifyILb[01]E")
    thin = _kernels(res, r"k_filter_splat_thinlensILb[01]E")
    assert len(cls) == 2 and len(thin) == 2  # frames with cryptomatte AOVs run their own instantiations
    assert max(r for r, _ in cls.values()) <= 64
    (plain,) = [v for k, v in thin.items() if "ILb0E" in k]
    assert plain[0] <= 80  # 6 CTAs per SM: the RGBA thin-lens splat as profiled


def _sass(res, pattern):
    (name,) = _kernels(res, pattern).keys()
    return subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True).stdout


def test_instruction_forms_of_the_bench_kernels(resources):
    res, _ = resources
    k1 = _sass(res, r"k_create_rays_w550_5E")
    n_ffma2, n_fmul2 = len(re.findall(r"\bFFMA2\b", k1)), len(re.findall(r"\bFMUL2\b", k1))
    assert n_ffma2 >= 200 and n_fmul2 >= 80, (n_ffma2, n_fmul2)  # two rays per thread on the packed FP32 instructions
    assert "LDG.E.CONSTANT" not in k1 or len(re.findall(r"LDG", k1)) < 40  # coefficients are immediates, not table loads
    k2 = _sass(res, r"k_filter_splat_w550_5E")
    assert len(re.findall(r"\bFFMA2\b", k2)) >= 80 and ".LO_HI" in k2  # mirror packing: one half holds the x<->y mirrored monomial
    assert re.search(r"REDG\.E\.ADD\.F32", k2) and not re.search(r"ATOMG\.E\.ADD\.F32", k2)
    assert re.search(r"REDG\.E\.MIN\.64", k2)  # closest-filter depth keys
