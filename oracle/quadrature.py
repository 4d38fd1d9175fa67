"""Quadrature rules on the reference simplex (oracle; see oracle/__init__.py).

Stands in for ``basix.make_quadrature`` (reference call site: the ``dx`` measure with
``metadata={"quadrature_degree": 6}`` at examples/01_obstacle_problem/obstacle_pg.py:106-115 and
the quadrature element at :107).  Two schemes:

* ``"symmetric"`` -- tables from oracle/quadrature_tables.py (Dunavant / Keast orbits, polished to
  double precision by oracle/gen_quadrature.py): 12 points (triangle) / 24 points (tetrahedron) at
  degree 6.
* ``"gauss_jacobi"`` -- collapsed Gauss-Jacobi product rule, any degree (basix's own scheme of that
  name follows the same published construction: Stroud conical product).

Both are exact to the requested degree; the nonlinear ``exp`` terms depend on the rule, so rules
are always passed to the GPU path as data.
"""
import numpy as np
from scipy.special import roots_jacobi

from .quadrature_tables import TABLES

CELL_TDIM = {"triangle": 2, "tetrahedron": 3}


def gauss_jacobi(cell: str, degree: int):
    m = (degree + 2) // 2
    tdim = CELL_TDIM[cell]
    if tdim == 2:
        x0, w0 = roots_jacobi(m, 1.0, 0.0)
        x1, w1 = roots_jacobi(m, 0.0, 0.0)
        x0 = 0.5 * (x0 + 1.0)
        x1 = 0.5 * (x1 + 1.0)
        pts, wts = [], []
        for i in range(m):
            for j in range(m):
                pts.append((x0[i], x1[j] * (1.0 - x0[i])))
                wts.append(w0[i] * w1[j] * 0.125)
        return np.array(pts), np.array(wts)
    x0, w0 = roots_jacobi(m, 2.0, 0.0)
    x1, w1 = roots_jacobi(m, 1.0, 0.0)
    x2, w2 = roots_jacobi(m, 0.0, 0.0)
    x0, x1, x2 = 0.5 * (x0 + 1.0), 0.5 * (x1 + 1.0), 0.5 * (x2 + 1.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            for k in range(m):
                pts.append((x0[i], x1[j] * (1.0 - x0[i]), x2[k] * (1.0 - x0[i]) * (1.0 - x1[j])))
                wts.append(w0[i] * w1[j] * w2[k] * 0.125 * 0.25 * 0.5)
    return np.array(pts), np.array(wts)


def make_quadrature(cell: str, degree: int, scheme: str = "default"):
    """Return (points [nq, tdim], weights [nq]) exact for polynomials up to ``degree``."""
    if scheme in ("default", "symmetric"):
        cands = sorted(d for (c, d) in TABLES if c == cell and d >= degree)
        if cands:
            p, w = TABLES[(cell, cands[0])]
            return np.array(p, dtype=np.float64), np.array(w, dtype=np.float64)
        if scheme == "symmetric":
            raise ValueError(f"no symmetric table for {cell} degree {degree}")
    return gauss_jacobi(cell, degree)
