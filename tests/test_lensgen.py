"""Lens pack integrity and the CUDA code generator, checked on the CPU: the emitted straight-line code
(shared monomial DAG + FFMA chains) is executed as Python and compared with direct polynomial evaluation."""
import json
import os
import re

import numpy as np
import pytest

from pota_b200.lensgen import emit, emit_cuda, pack
from pota_b200.lensgen.fit import poly_eval
from pota_b200.lensgen.prescriptions import LENS_DB_DIRS, LENS_IDS


def test_pack_is_complete_and_float32_exact():
    lenses = pack.load_pack()
    assert [l["lens_id"] for l in lenses] == LENS_IDS and [l["db_dir"] for l in lenses] == LENS_DB_DIRS
    for l in lenses:
        assert set(l["polys"]) == set(pack.POLY_NAMES)
        for name, terms in l["polys"].items():
            assert 1 <= len(terms) <= 64
            for c, e in terms:
                assert float(np.float32(c)) == c and c != 0.0, (l["lens_id"], name)
                assert len(e) == 5 and all(0 <= x <= 15 for x in e) and sum(e) <= l["max_degree"]
        c = l["constants"]
        assert c["lens_outer_pupil_radius"] < abs(c["lens_outer_pupil_curvature_radius"])
        assert abs(c["lens_effective_focal_length"] - l["focal_mm"]) < 1e-3
        assert c["lens_outer_pupil_geometry"] == "spherical"
    # the pack spans the degree / term-count range of SURVEY.md §8d config C4
    assert {l["max_degree"] for l in lenses} == {5, 7, 9, 11}
    assert min(l["max_terms"] for l in lenses) == 12 and max(l["max_terms"] for l in lenses) == 64


def _run_group(lines, b):
    """Execute one emitted device function body (C statements) as numpy float64 Python."""
    env = {"fmaf": lambda a, x, y: a * x + y, "b": b,
           # packed bodies: both halves carry the same numpy vector here
           "make_float2": lambda a, b_: a, "__fmul2_rn": lambda a, x: a * x, "__ffma2_rn": lambda a, x, y: a * x + y, "__fadd2_rn": lambda a, x: a + x}
    out = {"ap": [None] * 2, "J": [None] * 4, "out": [None] * 4, "K": [None] * 4}
    env.update(out)
    for ln in lines:
        ln = ln.strip()
        if not ln or ln.startswith("LB_DEV") or ln == "}":
            continue
        ln = re.sub(r"^(const )?float2? ", "", ln).rstrip(";")
        ln = re.sub(r"(\d)f\b", r"\1", ln)  # 1.5f -> 1.5
        if ln.startswith("b0 = "):
            for k in range(5):
                env[f"b{k}"] = b[k]
            continue
        exec(ln, {}, env)
    return env


@pytest.mark.parametrize("lens_index", [5, 28, 43])
def test_emitted_code_matches_direct_evaluation(lens_index):
    lens = pack.load_pack()[lens_index]
    P = {n: [(c, tuple(e)) for c, e in t] for n, t in lens["polys"].items()}
    rs = np.random.default_rng(lens_index)
    X = np.stack([rs.uniform(-10, 10, 64), rs.uniform(-8, 8, 64), rs.uniform(-0.25, 0.25, 64), rs.uniform(-0.25, 0.25, 64), rs.uniform(0.4, 0.7, 64)])
    d = lambda n, v: [(c, tuple(e)) for c, e in emit.derivative(P[n], v)]
    polys = [P["ap_x"], P["ap_y"], d("ap_x", 2), d("ap_x", 3), d("ap_y", 2), d("ap_y", 3), P["out_x"], P["out_y"], P["out_dx"], P["out_dy"],
             d("out_dx", 0), d("out_dx", 1), d("out_dy", 0), d("out_dy", 1)]
    outs = ["ap[0]", "ap[1]", "J[0]", "J[1]", "J[2]", "J[3]", "out[0]", "out[1]", "out[2]", "out[3]", "K[0]", "K[1]", "K[2]", "K[3]"]
    lines, muls, ffma = emit_cuda.emit_group("lt_all", "...", polys, outs)
    env = _run_group(lines[1:], [X[k] for k in range(5)])
    got = [env["ap"][0], env["ap"][1]] + env["J"] + env["out"] + env["K"]
    for terms, g in zip(polys, got):
        want = poly_eval(terms, X)
        # float32-printed coefficients (%.9g round-trips float32) -> agreement to double rounding
        np.testing.assert_allclose(g, want, rtol=1e-6, atol=1e-7)  # coefficients are printed with 9 significant digits (float32 literals)
    assert ffma == sum(len([1 for c, e in t if sum(e) > 0]) for t in polys)
    assert muls < 1.6 * ffma  # the DAG shares monomials across the 14 polynomials
    # the packed (FFMA2) bodies of the forward kernel: same polynomials, monomials consumed as they are formed
    for order in ("lex", "degree", "grouped"):
        emit_cuda.PACKED_ORDER = order
        try:
            lines, muls2, ffma2 = emit_cuda.emit_group("ap_jac2", "...", polys[:6], outs[:6], packed=True)
        finally:
            emit_cuda.PACKED_ORDER = "lex"
        env = _run_group(lines[1:], [X[k] for k in range(5)])
        for terms, g in zip(polys[:6], [env["ap"][0], env["ap"][1]] + env["J"]):
            np.testing.assert_allclose(g, poly_eval(terms, X), rtol=1e-6, atol=1e-7)
        assert ffma2 == sum(len([1 for c, e in t if sum(e) > 0]) for t in polys[:6])


def test_upstream_headers_are_emitted_for_every_lens(tmp_path):
    emit.emit_upstream(str(tmp_path))
    for d in LENS_DB_DIRS:
        for f in ("pt_evaluate.h", "pt_sample_aperture.h", "lt_sample_aperture.h", "lens_constants.h"):
            p = tmp_path / "polynomial-optics" / "database" / "lenses" / d / "code" / f
            assert p.exists() and p.read_text().startswith("case ")


# ---- second-generation bodies (wavelength folded, K1 two-ray packed / K2 mirror packed) ------------------------
@pytest.fixture(scope="module")
def folded_host_lib(tmp_path_factory):
    """The generated EvalFA / EvalFB source lines compiled by g++ against a float2 stand-in (no GPU needed)."""
    import ctypes
    import subprocess

    from pota_b200.lensgen import emit_folded

    d = tmp_path_factory.mktemp("folded")
    libs = {}
    for name, imm in (("table", None), ("imm550", emit_folded.LAMBDA_550)):  # coefficient-table bodies / 550 nm immediates
        src = d / f"folded_{name}.cpp"
        src.write_text(emit_folded.host_test_source(FOLDED_TEST_LENSES, imm_lambda=imm))
        so = d / f"libfolded_{name}.so"
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
        libs[name] = ctypes.CDLL(str(so))
    return libs


FOLDED_TEST_LENSES = [5, 12, 28, 43]


@pytest.mark.parametrize("lens_index", FOLDED_TEST_LENSES)
@pytest.mark.parametrize("lam,kind", [(0.55, "table"), (0.42, "table"), (0.7, "table"), (550.0 * 0.001, "imm550")])
def test_folded_bodies_match_direct_evaluation(folded_host_lib, lens_index, lam, kind):
    """Every output of the mirror-packed lt_all and of the two-ray packed ap_jac2 / out5_2, against the pack's polynomials
    evaluated term by term in double at the same wavelength."""
    import ctypes

    L = folded_host_lib[kind]
    lens = pack.load_pack()[lens_index]
    P = {n: [(c, tuple(e)) for c, e in t] for n, t in lens["polys"].items()}
    d = lambda n, v: [(c, tuple(e)) for c, e in emit.derivative(P[n], v)]  # noqa: E731
    lt = [P["ap_x"], P["ap_y"], d("ap_x", 2), d("ap_x", 3), d("ap_y", 2), d("ap_y", 3), P["out_x"], P["out_y"], P["out_dx"], P["out_dy"],
          d("out_dx", 0), d("out_dx", 1), d("out_dy", 0), d("out_dy", 1)]
    rs = np.random.default_rng(100 + lens_index)
    n = 48
    X = np.stack([rs.uniform(-10, 10, n), rs.uniform(-8, 8, n), rs.uniform(-0.25, 0.25, n), rs.uniform(-0.25, 0.25, n), np.full(n, lam)])
    X[:, 0] = [3.0, -2.0, 0.0, 0.1, lam]  # an axis-aligned point: mirror halves differ strongly
    X32 = X.astype(np.float32).astype(np.float64)
    X32[4] = lam

    def scale(terms):  # sum of |term|: what float32 rounding is relative to
        s = np.zeros(n)
        for c, e in terms:
            s += np.abs(c * np.prod([X32[k] ** e[k] for k in range(5)], axis=0))
        return s + 1e-30

    fp = ctypes.POINTER(ctypes.c_float)
    for i in range(n):
        b = np.array(list(X32[:4, i]) + [lam], np.float32)
        out = np.zeros(14, np.float32)
        assert L.ft_lt_all(lens_index, ctypes.c_double(lam), b.ctypes.data_as(fp), out.ctypes.data_as(fp)) == 0
        for j, terms in enumerate(lt):
            want = poly_eval(terms, X32[:, i:i + 1])[0]
            assert abs(out[j] - want) <= 4e-6 * scale(terms)[i], (lens_index, j, i, out[j], want)
        t = np.zeros(1, np.float32)
        assert L.ft_transmittance(lens_index, ctypes.c_double(lam), b.ctypes.data_as(fp), t.ctypes.data_as(fp)) == 0
        assert abs(t[0] - poly_eval(P["out_t"], X32[:, i:i + 1])[0]) <= 4e-6 * scale(P["out_t"])[i]
    # two-ray packed bodies: points i and i+1 in the two halves
    for i in range(0, n, 2):
        b2 = np.zeros((5, 2), np.float32)
        b2[:4] = X32[:4, i:i + 2]
        b2[4] = lam
        aj = np.zeros((6, 2), np.float32)
        assert L.ft_ap_jac2(lens_index, ctypes.c_double(lam), b2.ctypes.data_as(fp), aj.ctypes.data_as(fp)) == 0
        o5 = np.zeros((5, 2), np.float32)
        assert L.ft_out5_2(lens_index, ctypes.c_double(lam), b2.ctypes.data_as(fp), o5.ctypes.data_as(fp)) == 0
        for h in range(2):
            for j, terms in enumerate(lt[:6]):
                want = poly_eval(terms, X32[:, i + h:i + h + 1])[0]
                assert abs(aj[j, h] - want) <= 4e-6 * scale(terms)[i + h], (lens_index, j, i, h)
            for j, terms in enumerate([P["out_x"], P["out_y"], P["out_dx"], P["out_dy"], P["out_t"]]):
                want = poly_eval(terms, X32[:, i + h:i + h + 1])[0]
                assert abs(o5[j, h] - want) <= 4e-6 * scale(terms)[i + h], (lens_index, j, i, h)


def test_folded_bodies_issue_fewer_operations():
    """What the second generation is for: the mirror-packed lt_all issues about half the instructions of the 5-variate
    scalar body, and the folded K1 bodies fewer FMA-pipe operations than the 5-variate packed ones."""
    from pota_b200.lensgen import emit_folded

    for k in (5, 43):
        _, st = emit_cuda.lens_unit(pack.load_pack()[k])
        assert sum(st["lt_all_mirror"]) < 0.6 * sum(st["lt_all"])
        assert sum(st["ap_jac2_folded"]) < 0.92 * sum(st["ap_jac"]) and sum(st["out5_2_folded"]) < 0.92 * sum(st["out5"])
        _, fst, _, cb = emit_folded.folded_evaluators(pack.load_pack()[k])
        m = fst["lt_all_mirror"]
        # FMA-pipe lane operations (packed instructions count two) stay at or below the scalar body's
        assert 2 * (m["fmul2"] + m["ffma2"]) + m["fmul"] + m["ffma"] <= sum(st["lt_all"])
