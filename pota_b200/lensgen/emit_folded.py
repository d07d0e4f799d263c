"""Wavelength-folded per-lens evaluators (second generation of csrc/gen/lens_<k>.cu bodies).

The generated bodies of polynomial-optics are 5-variate sums in (x, y, dx, dy, lambda)
(/root/reference/src/lentil.h:1259-1262 and SURVEY.md Appendix A).  The wavelength is a per-camera
constant (`lambda = wavelength * 0.001`, lentil.h:1213) and warp-uniform, so every polynomial collapses to a
4-variate one whose coefficients are  sum_k c_k * lambda^e_k  -- computed once per launch on the host in double
and handed to the kernel BY VALUE (`__grid_constant__` struct -> constant bank -> uniform registers).  Terms that
differ only in their lambda exponent merge, and the monomial DAG loses a dimension: 14-18 % fewer FMA-pipe
operations on every lens of the pack (profiles/r02_poly_opcounts.txt).

Two evaluators are emitted per lens:

* ``EvalFA<k>``  (K1, camera_create_ray): the two-rays-per-thread packed bodies ``ap_jac2`` / ``out5_2`` over the
  folded polynomials; a coefficient is one 4-byte constant broadcast into both halves of an FFMA2/FMUL2
  (`UR.F32` operand form).
* ``EvalFB<k>``  (K2, lt_sample_aperture): ONE evaluation point per lane, packed over the lens's mirror symmetry.
  With X = {x, y} and D = {dx, dy}, a product of packed variables holds a monomial in its low half and the
  monomial with x<->y, dx<->dy exchanged in its high half, and the operand-swap form of FMUL2/FFMA2 (`.LO_HI`)
  is free.  The polynomial pairs (ap_x, ap_y), (d ap_x/d dx, d ap_y/d dy), (d ap_x/d dy, d ap_y/d dx), (out_x, out_y),
  (out_dx, out_dy), ... are mirror images of each other up to their coefficients, so one FFMA2 with the coefficient
  pair {c_P, c_Q} advances both; terms present in only one polynomial of a pair run as scalar FFMAs on the
  half they need.  Nothing about the symmetry is ASSUMED: every term of every polynomial is evaluated with
  its own coefficient, the packing only decides which instruction carries it.  Half the issue slots of the
  scalar body at the same register count (one evaluation point per lane, unlike the two-slot experiment of r01).
"""
from __future__ import annotations

from .emit import derivative

ZERO = (0, 0, 0, 0)
XYDE = ["x", "y", "dx", "dy"]


def mir(t):
    return (t[1], t[0], t[3], t[2])


def canon(t):
    return max(t, mir(t))


def _f(c: float) -> str:
    s = "%.9g" % c
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


def fold(terms5):
    """[(c, e5)] -> ordered {e4: [(c, lambda_exponent), ...]}"""
    out = {}
    for c, e in terms5:
        out.setdefault(tuple(e[:4]), []).append((float(c), int(e[4])))
    return out


class Coefs:
    """Folded coefficient table of one kernel: scalars `s[]` and pairs `p[]`, each a sum of c * lambda^e parts."""

    def __init__(self):
        self.s, self.p = [], []
        self._s_index, self._p_index = {}, {}

    def scalar(self, parts) -> int:
        key = tuple(parts)
        if key not in self._s_index:
            self._s_index[key] = len(self.s)
            self.s.append(list(parts))
        return self._s_index[key]

    def pair(self, parts_lo, parts_hi) -> int:
        key = (tuple(parts_lo), tuple(parts_hi))
        if key not in self._p_index:
            self._p_index[key] = len(self.p)
            self.p.append((list(parts_lo), list(parts_hi)))
        return self._p_index[key]

    def values(self, lam: float):
        """(s, p) evaluated at lambda -- what the generated host fold function computes (tests)."""
        ev = lambda parts: sum(c * lam**e for c, e in parts)  # noqa: E731
        return [ev(x) for x in self.s], [(ev(a), ev(b)) for a, b in self.p]


# ---------------------------------------------------------------------------------------------------------
# K1: two-ray packed bodies over the folded polynomials (same streamed lexicographic order as emit_cuda)
# ---------------------------------------------------------------------------------------------------------
class _Dag4:
    def __init__(self, lines, prefix, packed):
        self.lines, self.prefix, self.packed = lines, prefix, packed
        self.avail = {}
        self.count = self.muls = 0
        for i in range(4):
            t = [0] * 4
            t[i] = 1
            self.avail[tuple(t)] = "b%d" % i

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
                if r in self.avail and (best is None or abs(sum(a) - sum(r)) < best[2]):
                    best = (a, r, abs(sum(a) - sum(r)))
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


def emit_group_folded(name, signature, polys5, outputs, coefs: Coefs, packed=True, cname="C", imm_lambda=None):
    """Device function evaluating the folded `polys5` at b[0..3].  packed: two-ray float2 form.
    imm_lambda: fold at this wavelength into immediates instead of reading the coefficient table (A/B experiment)."""
    ty = "float2" if packed else "float"
    fma = "__ffma2_rn" if packed else "fmaf"
    lines = ["  LB_DEV void %s(%s) const {" % (name, signature), "    const %s b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3];" % ty]
    folded = [fold(p) for p in polys5]

    def cst(parts):
        if imm_lambda is not None:
            v = _f(sum(c * imm_lambda**e for c, e in parts))
            return "make_float2(%s, %s)" % (v, v) if packed else v
        j = coefs.scalar(parts)
        return ("make_float2(%s.s[%d], %s.s[%d])" % (cname, j, cname, j)) if packed else "%s.s[%d]" % (cname, j)

    monos = sorted({e for f in folded for e in f if sum(e) > 0})
    uses = {}
    for j, f in enumerate(folded):
        for e, parts in f.items():
            if sum(e) > 0:
                uses.setdefault(e, []).append((j, parts))
    nterms = [sum(1 for e in f if sum(e) > 0) for f in folded]
    n_acc = [2 if n >= 12 else 1 for n in nterms]
    started = [[False, False] for _ in folded]
    count = [0] * len(folded)
    names = [o.replace("[", "").replace("]", "").replace("*", "") for o in outputs]
    ffma = 0

    def consume(m, var):
        nonlocal ffma
        for j, parts in uses.get(m, []):
            k = count[j] % n_acc[j]
            count[j] += 1
            acc = "%s_%d" % (names[j], k)
            if not started[j][k]:
                started[j][k] = True
                if k == 0 and ZERO in folded[j]:
                    lines.append("    %s %s = %s(%s, %s, %s);" % (ty, acc, fma, cst(parts), var, cst(folded[j][ZERO])))
                elif packed:
                    lines.append("    float2 %s = __fmul2_rn(%s, %s);" % (acc, cst(parts), var))
                else:
                    lines.append("    float %s = %s * %s;" % (acc, cst(parts), var))
            else:
                lines.append("    %s = %s(%s, %s, %s);" % (acc, fma, cst(parts), var, acc))
            ffma += 1

    body = []
    dag = _Dag4(body, "m", packed)
    for i in range(4):
        t = [0] * 4
        t[i] = 1
        consume(tuple(t), "b%d" % i)
    for m in monos:
        if sum(m) == 1:
            continue
        before = len(body)
        var = dag.get(m)
        lines.extend(body[before:])
        consume(m, var)
    for j, out in enumerate(outputs):
        accs = ["%s_%d" % (names[j], k) for k in range(2) if started[j][k]]
        if not accs:
            lines.append("    %s = %s;" % (out, cst(folded[j].get(ZERO, [(0.0, 0)]))))
        elif len(accs) == 1:
            lines.append("    %s = %s;" % (out, accs[0]))
        elif packed:
            lines.append("    %s = __fadd2_rn(%s, %s);" % (out, accs[0], accs[1]))
        else:
            lines.append("    %s = %s + %s;" % (out, accs[0], accs[1]))
    lines.append("  }")
    return lines, dag.muls, ffma


# ---------------------------------------------------------------------------------------------------------
# K2: mirror-packed body
# ---------------------------------------------------------------------------------------------------------
class _Node:
    __slots__ = ("key", "a", "r", "selfm", "need", "var", "single")

    def __init__(self, key, a, r):
        self.key, self.a, self.r = key, a, r
        self.selfm = key == mir(key)
        self.need = set()  # monomials (tuples) of this node somebody reads: subset of {key, mir(key)}
        self.var = None
        self.single = None  # the one monomial computed when only one half is needed


class MirrorDag:
    """Multiplication DAG over canonical exponent tuples; a node stands for the packed value {m, mir(m)}."""

    def __init__(self):
        self.nodes = {}
        self.order = []
        self.avail = set()
        for key in [(1, 0, 0, 0), (0, 0, 1, 0)]:
            n = _Node(key, None, None)
            self.nodes[key] = n
            self.avail.add(key)
            self.avail.add(mir(key))

    def plan(self, t):
        """make sure the node of tuple t exists (recursively); returns it"""
        key = canon(t)
        if key in self.nodes:
            return self.nodes[key]
        m = key
        best = None
        for a in self.avail:
            if all(x <= y for x, y in zip(a, m)):
                r = tuple(y - x for x, y in zip(a, m))
                if r in self.avail and (best is None or abs(sum(a) - sum(r)) < best[2]):
                    best = (a, r, abs(sum(a) - sum(r)))
        if best is None:
            a = max((a for a in self.avail if all(x <= y for x, y in zip(a, m))), key=lambda t: (sum(t), t))
            r = tuple(y - x for x, y in zip(a, m))
            self.plan(r)
        else:
            a, r = best[0], best[1]
        n = _Node(m, a, r)
        self.nodes[m] = n
        self.order.append(m)
        self.avail.add(m)
        self.avail.add(mir(m))
        return n

    def need(self, mono):
        self.nodes[canon(mono)].need.add(mono)

    def propagate(self):
        for key in reversed(self.order):
            n = self.nodes[key]
            for q in list(n.need):
                a, r = (n.a, n.r) if q == key else (mir(n.a), mir(n.r))
                self.need(a)
                self.need(r)

    # ---- expressions ------------------------------------------------------------------------------
    def mono(self, t):
        """scalar expression of monomial t"""
        n = self.nodes[canon(t)]
        if n.a is None:  # base pair
            return n.var + (".x" if t == n.key else ".y")
        if n.selfm:
            return n.var
        if n.single is not None:
            assert n.single == t, (n.key, n.single, t)
            return n.var
        return n.var + (".x" if t == n.key else ".y")

    def pair(self, t):
        """float2 expression {mono(t), mono(mir t)}"""
        n = self.nodes[canon(t)]
        if n.selfm:
            return "make_float2(%s, %s)" % (n.var, n.var)
        assert n.single is None, (n.key, t)
        return n.var if t == n.key else "make_float2(%s.y, %s.x)" % (n.var, n.var)

    def emit_node(self, key, lines, counter):
        n = self.nodes[key]
        if n.selfm:
            n.var = "s%d" % counter[0]
            counter[0] += 1
            lines.append("    const float %s = %s * %s;" % (n.var, self.mono(n.a), self.mono(n.r)))
            return 1, 0
        if len(n.need) == 2:
            n.var = "m%d" % counter[0]
            counter[0] += 1
            lines.append("    const float2 %s = __fmul2_rn(%s, %s);" % (n.var, self.pair(n.a), self.pair(n.r)))
            return 0, 1
        (q,) = tuple(n.need)
        n.single = q
        n.var = "h%d" % counter[0]
        counter[0] += 1
        a, r = (n.a, n.r) if q == key else (mir(n.a), mir(n.r))
        lines.append("    const float %s = %s * %s;" % (n.var, self.mono(a), self.mono(r)))
        return 1, 0


class _Chain:
    """accumulator chain (one or two interleaved accumulators)"""

    def __init__(self, name, packed, n_terms, init=None):
        self.name, self.packed, self.init = name, packed, init
        self.n_acc = 2 if n_terms >= 12 else 1
        self.started = [False, False]
        self.count = 0

    def add(self, coef, var, lines):
        k = self.count % self.n_acc
        self.count += 1
        acc = "%s_%d" % (self.name, k)
        ty, fma = ("float2", "__ffma2_rn") if self.packed else ("float", "fmaf")
        if not self.started[k]:
            self.started[k] = True
            if k == 0 and self.init is not None:
                lines.append("    %s %s = %s(%s, %s, %s);" % (ty, acc, fma, coef, var, self.init))
            elif self.packed:
                lines.append("    float2 %s = __fmul2_rn(%s, %s);" % (acc, coef, var))
            else:
                lines.append("    float %s = %s * %s;" % (acc, coef, var))
        else:
            lines.append("    %s = %s(%s, %s, %s);" % (acc, fma, coef, var, acc))

    def result(self, lines):
        accs = ["%s_%d" % (self.name, k) for k in range(2) if self.started[k]]
        if not accs:
            return self.init
        if len(accs) == 1:
            return accs[0]
        r = self.name + "_r"
        if self.packed:
            lines.append("    const float2 %s = __fadd2_rn(%s, %s);" % (r, accs[0], accs[1]))
        else:
            lines.append("    const float %s = %s + %s;" % (r, accs[0], accs[1]))
        return r


def emit_mirror_group(name, signature, pairs, coefs: Coefs, cname="C", imm_lambda=None):
    """pairs: list of (termsP5, termsQ5, outP, outQ).  Emits one device function evaluating all of them at
    (b[0], b[1], b[2], b[3]) = (x, y, dx, dy).  Returns (lines, stats).
    imm_lambda: fold at this wavelength into immediate operands instead of reading the coefficient table."""

    def val(parts):
        return sum(c * imm_lambda**e for c, e in parts)

    def pair_coef(cp, cq):
        if imm_lambda is not None:
            return "make_float2(%s, %s)" % (_f(val(cp or [])), _f(val(cq or [])))
        if cp == cq:  # a lens fitted symmetrically: one constant word broadcast into both halves
            j = coefs.scalar(cp)
            return "make_float2(%s.s[%d], %s.s[%d])" % (cname, j, cname, j)
        return "%s.p[%d]" % (cname, coefs.pair(cp or [], cq or []))

    def scalar_coef(parts):
        if imm_lambda is not None:
            return _f(val(parts))
        return "%s.s[%d]" % (cname, coefs.scalar(parts))

    lines = ["  LB_DEV void %s(%s) const {" % (name, signature),
             "    const float2 X = make_float2(b[0], b[1]), D = make_float2(b[2], b[3]);"]
    dag = MirrorDag()
    dag.nodes[(1, 0, 0, 0)].var = "X"
    dag.nodes[(0, 0, 1, 0)].var = "D"
    folded = [(fold(p), fold(q)) for p, q, _, _ in pairs]
    # per pair and orientation t: P's coefficient of monomial t, Q's coefficient of monomial mir(t)
    orient = []
    for fp, fq in folded:
        o = {}
        for t, parts in fp.items():
            if sum(t) > 0:
                o.setdefault(t, [None, None])[0] = parts
        for u, parts in fq.items():
            if sum(u) > 0:
                o.setdefault(mir(u), [None, None])[1] = parts
        orient.append(o)
    all_t = sorted({canon(t) for o in orient for t in o})
    for key in all_t:
        dag.plan(key)
    for o in orient:
        for t, (cp, cq) in o.items():
            if cp is not None:
                dag.need(t)
            if cq is not None:
                dag.need(mir(t))
    dag.propagate()

    chains = []
    for k, ((fp, fq), o) in enumerate(zip(folded, orient)):
        both = sum(1 for cp, cq in o.values() if cp is not None and cq is not None)
        only_p = sum(1 for cp, cq in o.values() if cp is not None and cq is None)
        only_q = sum(1 for cp, cq in o.values() if cp is None and cq is not None)
        cP, cQ = fp.get(ZERO), fq.get(ZERO)
        init2 = None
        if cP is not None or cQ is not None:
            init2 = pair_coef(cP, cQ)
        chains.append((_Chain("q%d" % k, True, both, init2), _Chain("q%dp" % k, False, only_p), _Chain("q%dq" % k, False, only_q)))
    stats = dict(fmul=0, fmul2=0, ffma=0, ffma2=0)

    def consume(key):
        ts = [key] if key == mir(key) else [key, mir(key)]
        for k, o in enumerate(orient):
            c2, cp_chain, cq_chain = chains[k]
            for t in ts:
                if t not in o:
                    continue
                cp, cq = o[t]
                if cp is not None and cq is not None:
                    c2.add(pair_coef(cp, cq), dag.pair(t), lines)
                    stats["ffma2"] += 1
                elif cp is not None:
                    cp_chain.add(scalar_coef(cp), dag.mono(t), lines)
                    stats["ffma"] += 1
                else:
                    cq_chain.add(scalar_coef(cq), dag.mono(mir(t)), lines)
                    stats["ffma"] += 1

    counter = [0]
    consume((1, 0, 0, 0))
    consume((0, 0, 1, 0))
    for key in dag.order:
        n = dag.nodes[key]
        if not n.need:
            continue  # planned on the way but read by nobody
        a, b = dag.emit_node(key, lines, counter)
        stats["fmul"] += a
        stats["fmul2"] += b
        consume(key)
    for k, (_, _, outP, outQ) in enumerate(pairs):
        c2, cpc, cqc = chains[k]
        r2 = c2.result(lines)
        rp = cpc.result(lines)
        rq = cqc.result(lines)
        for out, half, extra in ((outP, "x", rp), (outQ, "y", rq)):
            terms = []
            if r2 is not None:
                terms.append("%s.%s" % (r2, half))
            if extra is not None:
                terms.append(extra)
            lines.append("    %s = %s;" % (out, " + ".join(terms) if terms else "0.0f"))
    lines.append("  }")
    return lines, stats


# ---------------------------------------------------------------------------------------------------------
def host_fold_code(tag, coefs: Coefs, struct_name):
    """C++ (host) that fills `struct_name` from a wavelength: tables of parts + one loop."""
    cs, es, off = [], [], [0]
    order = [parts for parts in coefs.s]
    for lo, hi in coefs.p:
        order.append(lo)
        order.append(hi)
    for parts in order:
        for c, e in parts:
            cs.append(c)
            es.append(e)
        off.append(len(cs))
    L = []
    ns, np_ = len(coefs.s), len(coefs.p)
    L.append("struct %s {" % struct_name)
    L.append("  float2 p[%d];" % max(np_, 1))
    L.append("  float s[%d];" % max(ns, 1))
    L.append("};")
    L.append("static const double kFoldC_%s[] = {%s};" % (tag, ", ".join(repr(float(c)) for c in cs) or "0.0"))
    L.append("static const unsigned char kFoldE_%s[] = {%s};" % (tag, ", ".join(str(e) for e in es) or "0"))
    L.append("static const unsigned short kFoldO_%s[] = {%s};" % (tag, ", ".join(str(o) for o in off)))
    L.append("static void fold_%s(double lambda, %s &K) {" % (tag, struct_name))
    L.append("  double pw[16]; pw[0] = 1.0; for (int i = 1; i < 16; ++i) pw[i] = pw[i - 1] * lambda;")
    L.append("  auto ev = [&](int j) { double v = 0.0; for (int i = kFoldO_%s[j]; i < kFoldO_%s[j + 1]; ++i) v += kFoldC_%s[i] * pw[kFoldE_%s[i]]; return (float)v; };" % (tag, tag, tag, tag))
    L.append("  for (int j = 0; j < %d; ++j) K.s[j] = ev(j);" % ns)
    L.append("  for (int j = 0; j < %d; ++j) K.p[j] = make_float2(ev(%d + 2 * j), ev(%d + 2 * j + 1));" % (np_, ns, ns))
    L.append("}")
    return L


def lens_polys(lens):
    P = {n: [(float(c), list(e)) for c, e in terms] for n, terms in lens["polys"].items()}

    def d(name, var):
        return [(float(c), list(e)) for c, e in derivative(P[name], var)]

    dap = [d("ap_x", 2), d("ap_x", 3), d("ap_y", 2), d("ap_y", 3)]
    dout = [d("out_dx", 0), d("out_dx", 1), d("out_dy", 0), d("out_dy", 1)]
    return P, dap, dout


LAMBDA_550 = 550.0 * 0.001  # `wavelength * 0.001` of the default parameter (lentil_camera.cpp:30, lentil.h:1213), in double


def folded_evaluators(lens, imm_lambda=None):
    """Source lines of the second-generation evaluators of one lens + their coefficient structs and host fold functions,
    and op statistics.  imm_lambda None: EvalFA<k> / EvalFB<k> read the folded coefficients from a __grid_constant__ table
    (any wavelength).  imm_lambda given: EvalIA<k> / EvalIB<k>, the same bodies with the coefficients folded at that
    wavelength into immediate operands (no constant loads at all)."""
    k = lens["index"]
    P, dap, dout = lens_polys(lens)
    src, stats = [], {}
    tag = "F" if imm_lambda is None else "I"
    sfx = "_folded" if imm_lambda is None else "_imm"
    # ---- K1 ----
    ca = Coefs()
    group = emit_group_horner if POLY_FORM == "horner" else emit_group_folded
    mirror_group = emit_mirror_group_horner if POLY_FORM == "horner" else emit_mirror_group
    body_ap, mul, ffma = group("ap_jac2", "const float2 b[5], float2 ap[2], float2 J[4]", [P["ap_x"], P["ap_y"]] + dap,
                               ["ap[0]", "ap[1]", "J[0]", "J[1]", "J[2]", "J[3]"], ca, imm_lambda=imm_lambda)
    stats["ap_jac2" + sfx] = (mul, ffma)
    body_o5, mul, ffma = group("out5_2", "const float2 b[5], float2 out[4], float2 &T",
                               [P["out_x"], P["out_y"], P["out_dx"], P["out_dy"], P["out_t"]],
                               ["out[0]", "out[1]", "out[2]", "out[3]", "T"], ca, imm_lambda=imm_lambda)
    stats["out5_2" + sfx] = (mul, ffma)
    if imm_lambda is None:
        src += host_fold_code("a%d" % k, ca, "FoldA%d" % k)
        src += ["struct EvalFA%d {" % k, "  const FoldA%d &C;" % k]
    else:
        src += ["struct EvalIA%d {" % k]
    src += body_ap + body_o5 + ["};", ""]
    # ---- K2 ----
    cb = Coefs()
    pairs = [(P["ap_x"], P["ap_y"], "ap[0]", "ap[1]"), (dap[0], dap[3], "J[0]", "J[3]"), (dap[1], dap[2], "J[1]", "J[2]"),
             (P["out_x"], P["out_y"], "out[0]", "out[1]"), (P["out_dx"], P["out_dy"], "out[2]", "out[3]"),
             (dout[0], dout[3], "K[0]", "K[3]"), (dout[1], dout[2], "K[1]", "K[2]")]
    body_lt, st = mirror_group("lt_all", "const float b[5], float ap[2], float J[4], float out[4], float K[4]", pairs, cb,
                               imm_lambda=imm_lambda)
    stats["lt_all_mirror" + ("" if imm_lambda is None else "_imm")] = st
    body_t, mul, ffma = group("transmittance_", "const float b[5], float &T", [P["out_t"]], ["T"], cb, packed=False,
                              imm_lambda=imm_lambda)
    stats["transmittance" + sfx] = (mul, ffma)
    if imm_lambda is None:
        src += host_fold_code("b%d" % k, cb, "FoldB%d" % k)
        src += ["struct EvalFB%d {" % k, "  const FoldB%d &C;" % k]
    else:
        src += ["struct EvalIB%d {" % k]
    src += body_lt + body_t
    src += ["  LB_DEV float transmittance(const float b[5]) const { float T; transmittance_(b, T); return T; }", "};", ""]
    return src, stats, ca, cb


# ---------------------------------------------------------------------------------------------------------
# host build of the generated bodies (tests/test_lensgen.py): the same source lines compiled by g++ against a
# float2 stand-in, so the generator is checked against the lens pack without a GPU
# ---------------------------------------------------------------------------------------------------------
HOST_SHIM = r"""
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#define LB_DEV inline
"""


def host_test_source(lens_indices, imm_lambda=None):
    from .pack import load_pack

    pack = load_pack()
    src = ["// generated by pota_b200.lensgen.emit_folded.host_test_source -- test infrastructure", HOST_SHIM]
    cases_lt, cases_aj, cases_o5, cases_t = [], [], [], []
    for k in lens_indices:
        lines, _, _, _ = folded_evaluators(pack[k], imm_lambda=imm_lambda)
        src += lines
        if imm_lambda is None:
            mk_a = "FoldA%d C; fold_a%d(lambda, C); EvalFA%d ev{C};" % (k, k, k)
            mk_b = "FoldB%d C; fold_b%d(lambda, C); EvalFB%d ev{C};" % (k, k, k)
        else:
            mk_a = "(void)lambda; EvalIA%d ev{};" % k
            mk_b = "(void)lambda; EvalIB%d ev{};" % k
        cases_lt.append("    case %d: { %s ev.lt_all(b, out, out + 2, out + 6, out + 10); return 0; }" % (k, mk_b))
        cases_t.append("    case %d: { %s *out = ev.transmittance(b); return 0; }" % (k, mk_b))
        cases_aj.append("    case %d: { %s ev.ap_jac2(b, out, out + 2); return 0; }" % (k, mk_a))
        cases_o5.append("    case %d: { %s ev.out5_2(b, out, out[4]); return 0; }" % (k, mk_a))
    src += ['extern "C" {',
            "// out: ap[2], J[4], out[4], K[4]",
            "int ft_lt_all(int lens, double lambda, const float *b, float *out) {", "  switch (lens) {"] + cases_lt + ["  }", "  return -1;", "}",
            "int ft_transmittance(int lens, double lambda, const float *b, float *out) {", "  switch (lens) {"] + cases_t + ["  }", "  return -1;", "}",
            "// b: float2[5] (two evaluation points), out: float2 ap[2], J[4]",
            "int ft_ap_jac2(int lens, double lambda, const float2 *b, float2 *out) {", "  switch (lens) {"] + cases_aj + ["  }", "  return -1;", "}",
            "// out: float2 out[4], T",
            "int ft_out5_2(int lens, double lambda, const float2 *b, float2 *out) {", "  switch (lens) {"] + cases_o5 + ["  }", "  return -1;", "}",
            "}"]
    return "\n".join(src) + "\n"


# ---------------------------------------------------------------------------------------------------------
# Horner form (r02b).  The bodies above form every distinct monomial of a polynomial GROUP once (a shared multiplication
# DAG) and then spend one FMA per term: ~1.65 FMA-pipe operations per term on the pack.  A greedy multivariate Horner scheme
# per polynomial -- pick the variable most terms contain, P = v^k * A + B, recurse -- needs no monomials at all: every step
# is ONE fused multiply-add, ~1.3 operations per term (profiles/r02_poly_opcounts.txt), at shorter live ranges.  The mirror
# packing of K2 carries over unchanged: with the packed variables {x, y}, {y, x}, {dx, dy}, {dy, dx} the scheme of P evaluates
# its mirror partner Q in the other half, provided their supports mirror each other (checked per pair; pairs that do not
# are evaluated as two scalar schemes).
# ---------------------------------------------------------------------------------------------------------
import os as _os

POLY_FORM = _os.environ.get("LB_POLY_FORM", "horner")  # horner | dag


def _horner_cost(exps):
    """operation count of the greedy scheme for a list of exponent tuples (no code emitted)"""
    if not exps:
        return 0
    if len(exps) == 1:
        return sum(1 for e in exps[0] if e > 0)
    v = max(range(4), key=lambda i: sum(1 for t in exps if t[i] > 0))
    with_v = [t for t in exps if t[v] > 0]
    rest = [t for t in exps if t[v] == 0]
    k = min(t[v] for t in with_v)
    a = [tuple(e - (k if i == v else 0) for i, e in enumerate(t)) for t in with_v]
    ca = 0 if (len(a) == 1 and sum(a[0]) == 0) else _horner_cost(a)
    return ca + _horner_cost(rest) + 1


class HornerEmitter:
    """Emits one polynomial as a nested Horner scheme.  packed: float2 arithmetic on FFMA2/FMUL2."""

    def __init__(self, lines, packed, var_expr, power_expr, coef_expr, prefix, stats):
        self.lines, self.packed, self.var_expr, self.power_expr, self.coef_expr = lines, packed, var_expr, power_expr, coef_expr
        self.prefix, self.stats, self.n = prefix, stats, 0

    def _tmp(self):
        self.n += 1
        return "%s%d" % (self.prefix, self.n)

    def _mul(self, a, b):
        t = self._tmp()
        if self.packed:
            self.lines.append("    const float2 %s = __fmul2_rn(%s, %s);" % (t, a, b))
            self.stats["fmul2"] += 1
        else:
            self.lines.append("    const float %s = %s * %s;" % (t, a, b))
            self.stats["fmul"] += 1
        return t

    def _fma(self, a, b, c):
        t = self._tmp()
        if self.packed:
            self.lines.append("    const float2 %s = __ffma2_rn(%s, %s, %s);" % (t, a, b, c))
            self.stats["ffma2"] += 1
        else:
            self.lines.append("    const float %s = fmaf(%s, %s, %s);" % (t, a, b, c))
            self.stats["ffma"] += 1
        return t

    def _pow(self, v, k):
        return self.var_expr(v) if k == 1 else self.power_expr(v, k)

    def emit(self, terms):
        """terms: [(payload, exps)] -> expression holding the value (a temporary or a coefficient expression)"""
        if len(terms) == 1:
            c, e = terms[0]
            expr = self.coef_expr(c)
            for v in range(4):
                if e[v] > 0:
                    expr = self._mul(expr, self._pow(v, e[v]))
            return expr
        # the variable whose choice gives the cheapest scheme one level down (ties: most terms)
        best = None
        for v in range(4):
            with_v = [t for t in terms if t[1][v] > 0]
            if not with_v:
                continue
            rest = [t for t in terms if t[1][v] == 0]
            k = min(t[1][v] for t in with_v)
            a = [tuple(x - (k if i == v else 0) for i, x in enumerate(t[1])) for t in with_v]
            cost = (0 if (len(a) == 1 and sum(a[0]) == 0) else _horner_cost(a)) + _horner_cost([t[1] for t in rest]) + 1
            if best is None or (cost, -len(with_v)) < (best[0], -best[1]):
                best = (cost, len(with_v), v, k)
        _, _, v, k = best
        with_v = [(c, tuple(x - (k if i == v else 0) for i, x in enumerate(e))) for c, e in terms if e[v] > 0]
        rest = [(c, e) for c, e in terms if e[v] == 0]
        a = self.emit(with_v)
        if not rest:
            return self._mul(a, self._pow(v, k))
        b = self.emit(rest)
        return self._fma(a, self._pow(v, k), b)


class _Powers:
    """x^k, y^k, dx^k, dy^k of one evaluation: packed pairs {x^k, y^k} and {dx^k, dy^k} built on demand (K2, mirror) or
    per-variable values (K1 two-ray packed / scalar)."""

    def __init__(self, lines, mode, stats):
        self.lines, self.mode, self.stats, self.have = lines, mode, stats, {}

    def get(self, base, k):  # base: 0 = X pair (or variable index in plain mode), 2 = D pair
        if k == 1:
            return self.base_name(base)
        if (base, k) in self.have:
            return self.have[(base, k)]
        lo, hi = k // 2, k - k // 2
        a, b = self.get(base, lo), self.get(base, hi)
        name = "p%d_%d" % (base, k)
        if self.mode == "scalar":
            self.lines.append("    const float %s = %s * %s;" % (name, a, b))
            self.stats["fmul"] += 1
        else:
            self.lines.append("    const float2 %s = __fmul2_rn(%s, %s);" % (name, a, b))
            self.stats["fmul2"] += 1
        self.have[(base, k)] = name
        return name

    def base_name(self, base):
        if self.mode == "mirror":
            return "X" if base == 0 else "D"
        return "b%d" % base


def emit_group_horner(name, signature, polys5, outputs, coefs: Coefs, packed=True, cname="C", imm_lambda=None):
    """K1 (two-ray packed) or scalar bodies in Horner form over the folded polynomials."""
    ty = "float2" if packed else "float"
    lines = ["  LB_DEV void %s(%s) const {" % (name, signature), "    const %s b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3];" % ty]
    stats = dict(fmul=0, fmul2=0, ffma=0, ffma2=0)
    powers = _Powers(lines, "packed" if packed else "scalar", stats)

    def cst(parts):
        if imm_lambda is not None:
            v = _f(sum(c * imm_lambda**e for c, e in parts))
            return "make_float2(%s, %s)" % (v, v) if packed else v
        j = coefs.scalar(parts)
        return ("make_float2(%s.s[%d], %s.s[%d])" % (cname, j, cname, j)) if packed else "%s.s[%d]" % (cname, j)

    for j, (poly, out) in enumerate(zip(polys5, outputs)):
        f = fold(poly)
        em = HornerEmitter(lines, packed, lambda v: "b%d" % v, lambda v, k: powers.get(v, k), cst, "h%d_" % j, stats)
        terms = [(parts, e) for e, parts in f.items()]
        lines.append("    %s = %s;" % (out, em.emit(terms) if terms else cst([(0.0, 0)])))
    lines.append("  }")
    return lines, stats["fmul"] + stats["fmul2"], stats["ffma"] + stats["ffma2"]


def emit_mirror_group_horner(name, signature, pairs, coefs: Coefs, cname="C", imm_lambda=None):
    """K2: lt_all in Horner form, pairs with mirrored supports on packed instructions, the others as two scalar schemes."""
    lines = ["  LB_DEV void %s(%s) const {" % (name, signature),
             "    const float2 X = make_float2(b[0], b[1]), D = make_float2(b[2], b[3]);"]
    stats = dict(fmul=0, fmul2=0, ffma=0, ffma2=0)
    powers = _Powers(lines, "mirror", stats)

    def val(parts):
        return sum(c * imm_lambda**e for c, e in parts)

    def swap(x):
        return "make_float2(%s.y, %s.x)" % (x, x)

    def pvar(v):  # packed variable seen from P: x -> {x, y}, y -> {y, x}, dx -> {dx, dy}, dy -> {dy, dx}
        base = "X" if v < 2 else "D"
        return base if v % 2 == 0 else swap(base)

    def ppow(v, k):
        pk = powers.get(0 if v < 2 else 2, k)
        return pk if v % 2 == 0 else swap(pk)

    def svar(v):  # scalar variable
        return ("X" if v < 2 else "D") + (".x" if v % 2 == 0 else ".y")

    def spow(v, k):
        return powers.get(0 if v < 2 else 2, k) + (".x" if v % 2 == 0 else ".y")

    def pair_coef(cpq):
        cp, cq = cpq
        if imm_lambda is not None:
            return "make_float2(%s, %s)" % (_f(val(cp)), _f(val(cq)))
        if cp == cq:
            j = coefs.scalar(cp)
            return "make_float2(%s.s[%d], %s.s[%d])" % (cname, j, cname, j)
        return "%s.p[%d]" % (cname, coefs.pair(cp, cq))

    def scalar_coef(parts):
        if imm_lambda is not None:
            return _f(val(parts))
        return "%s.s[%d]" % (cname, coefs.scalar(parts))

    for k, (p5, q5, outP, outQ) in enumerate(pairs):
        fp, fq = fold(p5), fold(q5)
        if set(fp) == {mir(u) for u in fq} and fp:  # supports mirror each other: one packed scheme serves both
            em = HornerEmitter(lines, True, pvar, ppow, pair_coef, "g%d_" % k, stats)
            r = em.emit([((parts, fq[mir(e)]), e) for e, parts in fp.items()])
            lines.append("    %s = %s.x;" % (outP, r) if not r.startswith("make_float2") and not r.startswith(cname) else
                         "    { const float2 r_ = %s; %s = r_.x; %s = r_.y; }" % (r, outP, outQ))
            if not r.startswith("make_float2") and not r.startswith(cname):
                lines.append("    %s = %s.y;" % (outQ, r))
        else:
            for f, out, tag in ((fp, outP, "p"), (fq, outQ, "q")):
                em = HornerEmitter(lines, False, svar, spow, scalar_coef, "g%d%s_" % (k, tag), stats)
                terms = [(parts, e) for e, parts in f.items()]
                lines.append("    %s = %s;" % (out, em.emit(terms) if terms else "0.0f"))
    lines.append("  }")
    return lines, stats
