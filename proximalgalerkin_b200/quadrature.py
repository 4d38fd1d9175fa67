"""Quadrature rules on the reference simplex, delivered to the GPU path as data.

The reference gets its rules from basix through ``Measure("dx", metadata={"quadrature_degree": 6})``
(examples/01_obstacle_problem/obstacle_pg.py:106-115).  basix is not available here, so rules are
plain tables: fully symmetric interior rules (``quadrature_tables.py``, polished to double precision)
where one exists at the requested degree, otherwise a collapsed Gauss-Jacobi product rule.  A caller
that has basix can pass ``basix.make_quadrature`` output straight to ``fem.functionspace(..., rule=)``;
nothing downstream depends on where the points came from.
"""
import numpy as np

from .quadrature_tables import TABLES

_TDIM = {"triangle": 2, "tetrahedron": 3}


def _gauss_jacobi(m, a):
    """m-point Gauss rule for the weight (1 - x)^a on [0, 1] (Golub-Welsch)."""
    k = np.arange(m, dtype=np.float64)
    b = 0.0
    # recurrence of Jacobi polynomials P^(a, b) on [-1, 1]
    den = (2 * k + a + b) * (2 * k + a + b + 2.0)
    diag = np.where(den == 0, (b - a) / (a + b + 2.0), (b * b - a * a) / np.where(den == 0, 1.0, den))
    kk = k[1:]
    off = (
        2.0
        / (2 * kk + a + b)
        * np.sqrt(kk * (kk + a) * (kk + b) * (kk + a + b) / ((2 * kk + a + b - 1.0) * (2 * kk + a + b + 1.0)))
    )
    T =np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
    x, V = np.linalg.eigh(T)
    mu0 = 2.0 ** (a + b + 1.0) / (a + b + 1.0)  # integral of (1-x)^a on [-1, 1] for b = 0
    w = mu0 * V[0, :] ** 2
    # map to [0, 1]: x -> (x + 1) / 2, weight (1 - t)^a picks up 2^-(a + 1)
    return 0.5 * (x + 1.0), w / 2.0 ** (a + 1.0)


def gauss_jacobi(cell, degree):
    m = (degree + 2) // 2
    tdim = _TDIM[cell]
    rules = [_gauss_jacobi(m, float(tdim - 1 - d)) for d in range(tdim)]
    pts, wts = [], []
    if tdim == 2:
        (x0, w0), (x1, w1) = rules
        for i in range(m):
            for j in range(m):
                pts.append((x0[i], x1[j] * (1.0 - x0[i])))
                wts.append(w0[i] * w1[j])
    else:
        (x0, w0), (x1, w1), (x2, w2) = rules
        for i in range(m):
            for j in range(m):
                for k in range(m):
                    pts.append((x0[i], x1[j] * (1.0 - x0[i]), x2[k] * (1.0 - x0[i]) * (1.0 - x1[j])))
                    wts.append(w0[i] * w1[j] * w2[k])
    return np.array(pts), np.array(wts)


def make_quadrature(cell, degree, scheme="default"):
    """(points [nq, tdim], weights [nq]) exact to ``degree`` on the reference ``cell``."""
    if scheme in ("default", "symmetric"):
        cands = sorted(d for (c, d) in TABLES if c == cell and d >= degree)
        if cands:
            p, w = TABLES[(cell, cands[0])]
            return np.array(p, dtype=np.float64), np.array(w, dtype=np.float64)
        if scheme == "symmetric":
            raise ValueError(f"no symmetric table for {cell} degree {degree}")
    return gauss_jacobi(cell, degree)
