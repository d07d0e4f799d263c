"""Sparse polynomial fit (orthogonal matching pursuit) of the lens maps.

Stands in for polynomial-optics' `fit` tool (absent from /root/reference).  Produces, per lens,
nine sparse 5-variate polynomials in (x, y, dx, dy, lambda):
    out[0..3], out_transmittance           -- sensor -> outer pupil   (consumed by lentil.h:1257-1266)
    ap_x, ap_y, ap_dx, ap_dy               -- sensor -> aperture      (consumed by lentil.h:1272-1313)
A polynomial is a list of (coeff, (ex, ey, edx, edy, elambda)) in selection order.
"""
from __future__ import annotations

import itertools

import numpy as np


def candidates(max_degree: int, parity_x: int, parity_y: int):
    """All exponent tuples of total degree <= max_degree with the symmetry of a rotationally
    symmetric system: (ex+edx) % 2 == parity_x and (ey+edy) % 2 == parity_y."""
    out = []
    for a, b, c, d in itertools.product(range(max_degree + 1), repeat=4):
        s = a + b + c + d
        if s > max_degree or (a + c) % 2 != parity_x or (b + d) % 2 != parity_y:
            continue
        for e in range(0, max_degree - s + 1):
            if e > 3:
                break
            out.append((a, b, c, d, e))
    out.sort(key=lambda t: (sum(t), t))
    return out


def design(X: np.ndarray, exps) -> np.ndarray:
    """Monomial design matrix [N, len(exps)] for X[5, N]."""
    max_e = max(max(t) for t in exps) if exps else 0
    pw = np.ones((5, max_e + 1, X.shape[1]))
    for k in range(1, max_e + 1):
        pw[:, k] = pw[:, k - 1] * X
    A = np.empty((X.shape[1], len(exps)))
    for j, (a, b, c, d, e) in enumerate(exps):
        A[:, j] = pw[0, a] * pw[1, b] * pw[2, c] * pw[3, d] * pw[4, e]
    return A


def omp(A: np.ndarray, y: np.ndarray, max_terms: int, rel_tol: float = 1e-7):
    """Orthogonal matching pursuit; returns (selected column indices in order, coefficients)."""
    norms = np.linalg.norm(A, axis=0)
    norms[norms == 0] = 1.0
    An = A / norms
    sel: list[int] = []
    r = y.copy()
    y_norm = np.linalg.norm(y) + 1e-300
    coef = np.zeros(0)
    for _ in range(min(max_terms, A.shape[1])):
        score = np.abs(An.T @ r)
        score[sel] = -1.0
        j = int(np.argmax(score))
        sel.append(j)
        coef, *_ = np.linalg.lstsq(An[:, sel], y, rcond=None)
        r = y - An[:, sel] @ coef
        if np.linalg.norm(r) / y_norm < rel_tol:
            break
    return sel, coef / norms[sel]


def fit_poly(X, y, max_degree, max_terms, parity):
    exps = candidates(max_degree, *parity)
    A = design(X, exps)
    sel, coef = omp(A, y, max_terms)
    terms = [(float(c), tuple(int(v) for v in exps[j])) for j, c in zip(sel, coef)]
    resid = y - A[:, sel] @ coef
    return terms, float(np.sqrt(np.mean(resid**2)))


def mirror(terms):
    """x<->y mirrored polynomial: out_y(x,y,dx,dy) = out_x(y,x,dy,dx)."""
    return [(c, (e[1], e[0], e[3], e[2], e[4])) for c, e in terms]


def poly_eval(terms, X):
    acc = np.zeros(X.shape[1])
    for c, e in terms:
        m = np.full(X.shape[1], c)
        for v, p in enumerate(e):
            if p:
                m = m * X[v] ** p
        acc += m
    return acc


def poly_derivative(terms, var: int):
    """d/d(var) of a sparse polynomial, dropping vanished terms, keeping term order."""
    out = []
    for c, e in terms:
        if e[var] == 0:
            continue
        ne = list(e)
        ne[var] -= 1
        out.append((c * e[var], tuple(ne)))
    return out


def fit_lens(lens, max_degree: int, max_terms: int, n_rays: int = 6000, sensor_half: float | None = None, seed: int = 7):
    """Fit all nine polynomials of one lens.  Returns dict name -> terms, plus rms residuals."""
    if sensor_half is None:
        sensor_half = min(21.0, 0.42 * lens.efl)
    X, ap, sph, T = lens.sample(n_rays, sensor_half, seed)
    polys, rms = {}, {}
    # x-like outputs are fitted, y-like ones mirrored (exact rotational symmetry)
    Xm = np.stack([X[1], X[0], X[3], X[2], X[4]])  # mirrored samples double the data for the x fit
    X2 = np.concatenate([X, Xm], axis=1)
    for name, yx, yy in (
        ("out_x", sph[0], sph[1]),
        ("ap_x", ap[0], ap[1]),
        ("ap_dx", ap[2], ap[3]),
    ):
        terms, r = fit_poly(X2, np.concatenate([yx, yy]), max_degree, max_terms, (1, 0))
        polys[name] = terms
        polys[name.replace("x", "y")] = mirror(terms)
        rms[name] = r
    # the outer-pupil tangent frame of csToSphere (lens.h:139-145) is not x<->y symmetric:
    # the two direction components are fitted independently
    polys["out_dx"], rms["out_dx"] = fit_poly(X, sph[2], max_degree, max_terms, (1, 0))
    polys["out_dy"], rms["out_dy"] = fit_poly(X, sph[3], max_degree, max_terms, (0, 1))
    t_terms, r = fit_poly(X, T, max_degree, max(6, max_terms // 2), (0, 0))
    polys["out_t"] = t_terms
    rms["out_t"] = r
    return polys, rms
