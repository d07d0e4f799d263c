"""Per-lens unrolled device code (csrc/gen/lens_<k>.cu) from the committed lens pack.

The reference compiles every lens's polynomials into the plugin as flat `c * lens_ipow(x, a) * ...`
sums (/root/reference/src/lentil.h:1261-1263 + the absent polynomial-optics headers).  Here each lens
becomes one translation unit whose evaluator computes every distinct monomial of a polynomial GROUP
once (a group = all polynomials the solver needs at one point: e.g. the 14 of one lt_sample_aperture
iteration) through a shared multiplication DAG, and sums each polynomial as a chain of FFMAs with the
coefficient as an immediate operand.  No tables, no loops, no memory traffic for coefficients.
"""
from __future__ import annotations

import os

from . import emit_folded
from .emit import derivative
from .pack import load_pack

VARS = ["b0", "b1", "b2", "b3", "b4"]


def _fmt(c: float) -> str:
    s = "%.9g" % c
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


class Dag:
    """Greedy multiplication DAG over exponent tuples."""

    def __init__(self, lines, prefix, packed=False):
        self.lines = lines
        self.prefix = prefix
        self.packed = packed
        self.avail = {}
        self.count = 0
        self.muls = 0
        for i in range(5):
            t = [0] * 5
            t[i] = 1
            self.avail[tuple(t)] = VARS[i]

    @staticmethod
    def _div(a, m):
        return all(x <= y for x, y in zip(a, m))

    def get(self, m):
        if m in self.avail:
            return self.avail[m]
        best = None
        for a in self.avail:
            if self._div(a, m):
                r = tuple(y - x for x, y in zip(a, m))
                if r in self.avail:
                    if best is None or abs(sum(a) - sum(r)) < best[2]:
                        best = (a, r, abs(sum(a) - sum(r)))  # balanced products shorten dependency chains
        if best is None:
            a = max((a for a in self.avail if self._div(a, m)), key=lambda t: (sum(t), t))
            r = tuple(y - x for x, y in zip(a, m))
            self.get(r)
        else:
            a, r = best[0], best[1]
        name = "%s%d" % (self.prefix, self.count)
        self.count += 1
        self.muls += 1
        if self.packed:
            self.lines.append("    const float2 %s = __fmul2_rn(%s, %s);" % (name, self.avail[a], self.avail[r]))
        else:
            self.lines.append("    const float %s = %s * %s;" % (name, self.avail[a], self.avail[r]))
        self.avail[m] = name
        return name


def _c2(c: float) -> str:
    return "make_float2(%s, %s)" % (_fmt(c), _fmt(c))  # folds to the broadcast-immediate operand of FFMA2/FMUL2


# Order in which the packed bodies form and consume their monomials.  Measured on lens 5 (K1, registers / spill bytes at
# 4 blocks per SM): grouped (all monomials, then all sums) 202 / 248, degree-sorted streaming 164 / 200,
# lexicographic streaming 154 / 40 -- exponent-tuple order finishes whole sub-trees of the multiplication DAG early.
PACKED_ORDER = os.environ.get("LB_PACKED_ORDER", "lex")  # grouped | degree | lex
# same choice for the scalar bodies (K2's lt_all: 154 registers grouped, 137 lex; at 5 blocks per SM 96 registers with
# 96 bytes of spill grouped, none lex)
SCALAR_ORDER = os.environ.get("LB_SCALAR_ORDER", "lex")


def _emit_streamed(lines, polys, outputs, packed=True, order=None, acc=2):
    """Body with every monomial consumed right after it is formed (short live ranges: the packed values take two
    registers each).  Accumulation order per polynomial therefore follows the monomial order, not the term order."""
    order = order or PACKED_ORDER
    key = (lambda t: (sum(t), t)) if order == "degree" else (lambda t: t)
    ty = "float2" if packed else "float"
    cst = _c2 if packed else _fmt
    fma = "__ffma2_rn" if packed else "fmaf"
    monos = sorted({tuple(e) for terms in polys for _, e in terms if sum(e) > 0}, key=key)
    uses = {}
    for j, terms in enumerate(polys):
        for c, e in terms:
            if sum(e) > 0:
                uses.setdefault(tuple(e), []).append((j, c))
    consts = [sum(c for c, e in terms if sum(e) == 0) for terms in polys]
    nterms = [sum(1 for c, e in terms if sum(e) > 0) for terms in polys]
    n_acc = [acc if n >= 12 else 1 for n in nterms]
    started = [[False, False] for _ in polys]
    count = [0] * len(polys)
    names = [out.replace("[", "").replace("]", "").replace("*", "") for out in outputs]
    ffma = 0

    def consume(m, var):
        nonlocal ffma
        for j, c in uses.get(m, []):
            k = count[j] % n_acc[j]
            count[j] += 1
            acc = "%s_%d" % (names[j], k)
            if not started[j][k]:
                started[j][k] = True
                if k == 0 and consts[j] != 0.0:
                    lines.append("    %s %s = %s(%s, %s, %s);" % (ty, acc, fma, cst(c), var, cst(consts[j])))
                elif packed:
                    lines.append("    float2 %s = __fmul2_rn(%s, %s);" % (acc, _c2(c), var))
                else:
                    lines.append("    float %s = %s * %s;" % (acc, _fmt(c), var))
            else:
                lines.append("    %s = %s(%s, %s, %s);" % (acc, fma, cst(c), var, acc))
            ffma += 1

    body = []
    dag = Dag(body, "m", packed)
    for i in range(5):
        t = [0] * 5
        t[i] = 1
        consume(tuple(t), VARS[i])
    for m in monos:
        if sum(m) == 1:
            continue
        before = len(body)
        var = dag.get(m)
        lines.extend(body[before:])
        consume(m, var)
        # intermediates created on the way may be monomials of the group as well: they are consumed when their turn comes
    for j, out in enumerate(outputs):
        accs = ["%s_%d" % (names[j], k) for k in range(2) if started[j][k]]
        if not accs:
            lines.append("    %s = %s;" % (out, cst(consts[j])))
        elif len(accs) == 1:
            lines.append("    %s = %s;" % (out, accs[0]))
        elif packed:
            lines.append("    %s = __fadd2_rn(%s, %s);" % (out, accs[0], accs[1]))
        else:
            lines.append("    %s = %s + %s;" % (out, accs[0], accs[1]))
    lines.append("  }")
    return lines, dag.muls, ffma


def emit_group(name, signature, polys, outputs, packed=False, acc=2):
    """One device function evaluating `polys` (list of term lists) at point b, writing `outputs` (lvalues).

    packed: every value is a float2 holding the same quantity of TWO independent evaluation points, every operation the
    packed FP32 instruction of sm_100a (FMUL2 / FFMA2 / FADD2 through __fmul2_rn / __ffma2_rn / __fadd2_rn) in the same
    order as the scalar body, so each half is bit-identical to the scalar result at half the issue slots."""
    ty = "float2" if packed else "float"
    lines = ["  LB_DEV void %s(%s) const {" % (name, signature), "    const %s b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3], b4 = b[4];" % ty]
    if packed and PACKED_ORDER != "grouped":
        return _emit_streamed(lines, polys, outputs, acc=acc)
    if not packed and SCALAR_ORDER != "grouped":
        return _emit_streamed(lines, polys, outputs, packed=False, order=SCALAR_ORDER)
    dag = Dag(lines, "m", packed)
    monos = sorted({tuple(e) for terms in polys for _, e in terms if sum(e) > 0}, key=lambda t: (sum(t), t))
    for m in monos:
        dag.get(m)
    ffma = 0
    for terms, out in zip(polys, outputs):
        const = sum(c for c, e in terms if sum(e) == 0)
        rest = [(c, tuple(e)) for c, e in terms if sum(e) > 0]
        # two interleaved accumulators halve the FFMA dependency chain of long polynomials
        n_acc = 2 if len(rest) >= 12 else 1
        acc = []
        for k in range(n_acc):
            part = rest[k::n_acc]
            if not part:
                continue
            var = "%s_%d" % (out.replace("[", "").replace("]", "").replace("*", ""), k)
            c0, e0 = part[0]
            if packed:
                if k == 0 and const != 0.0:
                    lines.append("    float2 %s = __ffma2_rn(%s, %s, %s);" % (var, _c2(c0), dag.avail[e0], _c2(const)))
                else:
                    lines.append("    float2 %s = __fmul2_rn(%s, %s);" % (var, _c2(c0), dag.avail[e0]))
                for c, e in part[1:]:
                    lines.append("    %s = __ffma2_rn(%s, %s, %s);" % (var, _c2(c), dag.avail[e], var))
            elif k == 0 and const != 0.0:
                lines.append("    float %s = fmaf(%s, %s, %s);" % (var, _fmt(c0), dag.avail[e0], _fmt(const)))
                for c, e in part[1:]:
                    lines.append("    %s = fmaf(%s, %s, %s);" % (var, _fmt(c), dag.avail[e], var))
            else:
                lines.append("    float %s = %s * %s;" % (var, _fmt(c0), dag.avail[e0]))
                for c, e in part[1:]:
                    lines.append("    %s = fmaf(%s, %s, %s);" % (var, _fmt(c), dag.avail[e], var))
            ffma += len(part)
            acc.append(var)
        if not acc:
            lines.append("    %s = %s;" % (out, _c2(const) if packed else _fmt(const)))
        elif len(acc) == 1:
            lines.append("    %s = %s;" % (out, acc[0]))
        elif packed:
            lines.append("    %s = __fadd2_rn(%s, %s);" % (out, acc[0], acc[1]))
        else:
            lines.append("    %s = %s + %s;" % (out, acc[0], acc[1]))
    lines.append("  }")
    return lines, dag.muls, ffma


def lens_unit(lens) -> tuple[str, dict]:
    """One translation unit per lens.  Kernels:
      k_create_rays_<k>          K1, wavelength-folded two-ray packed bodies, coefficients from a __grid_constant__ table
      k_create_rays_w550_<k>     K1, the same bodies folded at the default wavelength (550 nm) into immediates
      k_filter_splat_<k>         K2, wavelength-folded mirror-packed body, coefficient table
      k_filter_splat_w550_<k>    K2, folded at 550 nm into immediates
      k_filter_splat_chroma_<k>  K2, first-generation 5-variate scalar body: chromatic aberration gives every lane its own
                                 wavelength (lentil_filter.cpp:254-267)
    """
    k = lens["index"]
    P = {n: [(float(c), list(e)) for c, e in terms] for n, terms in lens["polys"].items()}

    def d(name, var):
        return [(float(c), list(e)) for c, e in derivative(P[name], var)]

    dap = [d("ap_x", 2), d("ap_x", 3), d("ap_y", 2), d("ap_y", 3)]
    dout = [d("out_dx", 0), d("out_dx", 1), d("out_dy", 0), d("out_dy", 1)]
    src = ["// generated by pota_b200.lensgen.emit_cuda from pota_b200/lenses/%s.json -- do not edit" % lens["lens_id"],
           '#include "../camera_kernels.cuh"', '#include "../filter_kernels.cuh"', '#include "../unrolled_dispatch.h"', "",
           "namespace lb {", "namespace {", "", "struct Eval%d {  // 5-variate bodies (wavelength as a variable)" % k]
    stats = {}
    # op statistics of the first-generation K1 bodies (kept for the comparison in tests / DESIGN; not emitted any more)
    for name, polys, outs in (("ap_jac", [P["ap_x"], P["ap_y"]] + dap, ["ap[0]", "ap[1]", "J[0]", "J[1]", "J[2]", "J[3]"]),
                              ("out5", [P["out_x"], P["out_y"], P["out_dx"], P["out_dy"], P["out_t"]], ["out[0]", "out[1]", "out[2]", "out[3]", "T"])):
        _, mul, ffma = emit_group(name, "", polys, outs)
        stats[name] = (mul, ffma)
    body, mul, ffma = emit_group("transmittance_", "const float b[5], float &T", [P["out_t"]], ["T"])
    src += body
    src += ["  LB_DEV float transmittance(const float b[5]) const { float T; transmittance_(b, T); return T; }"]
    stats["transmittance"] = (mul, ffma)
    body, mul, ffma = emit_group("lt_all", "const float b[5], float ap[2], float J[4], float out[4], float K[4]",
                                 [P["ap_x"], P["ap_y"]] + dap + [P["out_x"], P["out_y"], P["out_dx"], P["out_dy"]] + dout,
                                 ["ap[0]", "ap[1]", "J[0]", "J[1]", "J[2]", "J[3]", "out[0]", "out[1]", "out[2]", "out[3]", "K[0]", "K[1]", "K[2]", "K[3]"])
    src += body
    stats["lt_all"] = (mul, ffma)
    src += ["};", ""]
    # second generation: wavelength folded into the coefficients (emit_folded.py); K1 two-ray packed, K2 mirror packed
    for imm in (None, emit_folded.LAMBDA_550):
        fsrc, fstats, _, _ = emit_folded.folded_evaluators(lens, imm_lambda=imm)
        src += fsrc
        stats.update({g: (v if isinstance(v, tuple) else (v["fmul"] + v["fmul2"], v["ffma"] + v["ffma2"])) for g, v in fstats.items()})
    splat = "splat_persistent"
    k1_sig = "(const __grid_constant__ CamConsts<float> cam, const __grid_constant__ RayIO io, size_t n, uint64_t ray_id_base"
    k2_sig = ("(const __grid_constant__ CamConsts<float> cam, const __grid_constant__ FilterConsts fc, const __grid_constant__ AovSet aovs,\n"
              "    const __grid_constant__ SampleIO s, const WorkItem *__restrict__ work, FilterCounters *__restrict__ counters, uint64_t sample_base")
    k1_body = ["  const size_t j = (size_t)blockIdx.x * 128 + threadIdx.x;  // rays j and j + ceil(n/2)", "  if (j >= (n + 1) / 2) return;"]
    src += ["__global__ void __launch_bounds__(128%s)" % K1F_MIN_BLOCKS,
            "k_create_rays_%d%s, const __grid_constant__ FoldA%d C) {" % (k, k1_sig, k)] + k1_body + [
            "  camera_create_ray_pair(EvalFA%d{C}, cam, io, j, n, ray_id_base);" % k, "}", "",
            "__global__ void __launch_bounds__(128%s)" % K1_MIN_BLOCKS,
            "k_create_rays_w550_%d%s) {" % (k, k1_sig)] + k1_body + [
            "  camera_create_ray_pair(EvalIA%d{}, cam, io, j, n, ray_id_base);" % k, "}", "",
            "__global__ void __launch_bounds__(128%s)" % K2F_MIN_BLOCKS,
            "k_filter_splat_%d%s, const __grid_constant__ FoldB%d C) {" % (k, k2_sig, k),
            "  splat_persistent(EvalFB%d{C}, cam, fc, aovs, s, work, counters, sample_base);" % k, "}", "",
            "__global__ void __launch_bounds__(128%s)" % K2F_MIN_BLOCKS,
            "k_filter_splat_w550_%d%s) {" % (k, k2_sig),
            "  splat_persistent(EvalIB%d{}, cam, fc, aovs, s, work, counters, sample_base);" % k, "}", "",
            "__global__ void __launch_bounds__(128%s)" % K2_MIN_BLOCKS,
            "k_filter_splat_chroma_%d%s) {" % (k, k2_sig),
            "  %s(Eval%d{}, cam, fc, aovs, s, work, counters, sample_base);" % (splat, k), "}", "",
            "constexpr double kLambda550 = 550.0 * 0.001;  // `wavelength * 0.001` of the default parameter (lentil.h:1213)",
            "}  // namespace", "",
            "cudaError_t launch_fw_lens_%d(const CamConsts<float> &cam, const RayIO &io, size_t n, uint64_t ray_id_base, cudaStream_t stream) {" % k,
            "  const unsigned grid = (unsigned)(((n + 1) / 2 + 127) / 128);",
            "  if (cam.lambda_exact == kLambda550 && kernel_generation() == 2) {",
            "    k_create_rays_w550_%d<<<grid, 128, 0, stream>>>(cam, io, n, ray_id_base);" % k,
            "  } else {  // wavelength folded into the coefficients on the host, once per launch",
            "    FoldA%d C;" % k,
            "    fold_a%d(cam.lambda_exact, C);" % k,
            "    k_create_rays_%d<<<grid, 128, 0, stream>>>(cam, io, n, ray_id_base, C);" % k,
            "  }",
            "  return cudaGetLastError();", "}",
            "cudaError_t launch_bw_lens_%d(const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s, const WorkItem *work," % k,
            "                             FilterCounters *counters, uint64_t sample_base, int grid, cudaStream_t stream) {",
            "  if (fc.abb_chromatic > 0.0f || kernel_generation() == 0) {",
            "    k_filter_splat_chroma_%d<<<grid, 128, 0, stream>>>(cam, fc, aovs, s, work, counters, sample_base);" % k,
            "  } else if (cam.lambda_exact == kLambda550 && kernel_generation() == 2) {",
            "    k_filter_splat_w550_%d<<<grid, 128, 0, stream>>>(cam, fc, aovs, s, work, counters, sample_base);" % k,
            "  } else {",
            "    FoldB%d C;" % k,
            "    fold_b%d(cam.lambda_exact, C);" % k,
            "    k_filter_splat_%d<<<grid, 128, 0, stream>>>(cam, fc, aovs, s, work, counters, sample_base, C);" % k,
            "  }",
            "  return cudaGetLastError();", "}", "", "}  // namespace lb", ""]
    return "\n".join(src), stats


# tuning knobs: resident blocks per SM the register allocator must allow ("" = compiler's choice).
# Measured on B200 (lens 5, C2/C3): the scalar K1 was issue-bound and flat from 3 to 5 blocks (116 -> 96 registers); the
# packed two-rays-per-thread K1 is FMA-pipe/latency-bound and wants warps: 4.94e9 rays/s at 3 blocks (156 registers),
# 5.24e9 at 4 (128 registers, 8 bytes of spill).  K2 (latency-bound): 6.36e8 splats/s at 3 blocks, 6.70e8 at 4, and with
# the lexicographic bodies 7.01e8 at 4, 7.43e8 at 5 (96 registers, no spill), 7.44e8 at 6 (80, spills).
K1_MIN_BLOCKS = ", " + os.environ.get("LB_K1_MINBLOCKS", "4")
K2_MIN_BLOCKS = ", " + os.environ.get("LB_K2_MINBLOCKS", "5")
K1F_MIN_BLOCKS = ", " + os.environ.get("LB_K1F_MINBLOCKS", "4")
K2F_MIN_BLOCKS = ", " + os.environ.get("LB_K2F_MINBLOCKS", "5")


def emit_cuda(out_dir: str, only=None):
    """Write the per-lens units + the dispatch table; returns the list of .cu paths."""
    os.makedirs(out_dir, exist_ok=True)
    pack = load_pack()
    sel = os.environ.get("LB_UNROLLED_LENSES")
    if only is None and sel:
        only = [int(x) for x in sel.split(",") if x.strip()]
    paths, ids, all_stats = [], [], {}
    for lens in pack:
        if only is not None and lens["index"] not in only:
            continue
        text, stats = lens_unit(lens)
        p = os.path.join(out_dir, "lens_%d.cu" % lens["index"])
        _write_if_changed(p, text)
        paths.append(p)
        ids.append(lens["index"])
        all_stats[lens["lens_id"]] = stats
    disp = ["// generated by pota_b200.lensgen.emit_cuda -- do not edit", '#include "../unrolled_dispatch.h"', "namespace lb {"]
    for k in ids:
        disp.append("cudaError_t launch_fw_lens_%d(const CamConsts<float> &, const RayIO &, size_t, uint64_t, cudaStream_t);" % k)
        disp.append("cudaError_t launch_bw_lens_%d(const CamConsts<float> &, const FilterConsts &, const AovSet &, const SampleIO &, const WorkItem *,"
                    " FilterCounters *, uint64_t, int, cudaStream_t);" % k)
    disp.append("FwLauncher unrolled_fw_launcher(int m) {\n  switch (m) {")
    disp += ["    case %d: return launch_fw_lens_%d;" % (k, k) for k in ids]
    disp.append("    default: return nullptr;\n  }\n}")
    disp.append("BwLauncher unrolled_bw_launcher(int m) {\n  switch (m) {")
    disp += ["    case %d: return launch_bw_lens_%d;" % (k, k) for k in ids]
    disp.append("    default: return nullptr;\n  }\n}\n}  // namespace lb\n")
    p = os.path.join(out_dir, "unrolled_dispatch.cu")
    _write_if_changed(p, "\n".join(disp))
    paths.append(p)
    # stale units of a previous selection
    keep = {os.path.basename(x) for x in paths}
    for f in os.listdir(out_dir):
        if f.startswith("lens_") and f.endswith(".cu") and f not in keep:
            os.remove(os.path.join(out_dir, f))
    with open(os.path.join(out_dir, "unrolled_stats.txt"), "w") as f:
        f.write("# lens: group (multiplies, ffma)\n")
        for lid, st in all_stats.items():
            f.write(lid + ": " + ", ".join("%s (%d, %d)" % (g, a, b) for g, (a, b) in st.items()) + "\n")
    return paths


def _write_if_changed(path, text):
    if os.path.exists(path):
        with open(path) as f:
            if f.read() == text:
                return
    with open(path, "w") as f:
        f.write(text)
