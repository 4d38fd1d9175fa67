"""Control flow of the bt line-search Newton loop (proximalgalerkin_b200/linesearch.py) against the restated
PETSc algorithm (oracle/snes.py) through a numpy backend built on the oracle's residual and Jacobian.  On the
GPU the same loop runs over the library's entry points (tests/test_gpu_parity.py)."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle import mesh as omesh
from oracle import obstacle as oobs
from oracle import snes as osnes
from proximalgalerkin_b200 import linesearch as ls


class NumpyBackend(ls.Backend):
    """Backend over the oracle: vectors are 1-element lists holding numpy arrays."""

    def __init__(self, residual, jacobian, n):
        self.res, self.jac, self.n = residual, jacobian, n
        self.J = None
        self.calls = 0

    def vector(self):
        return [np.zeros(self.n)]

    def residual(self, x, F):
        F[0] = self.res(x[0])
        self.J = self.jac(x[0])
        self.calls += 1
        return float(np.linalg.norm(F[0]))

    def solve(self, F, y):
        y[0] = spla.splu(self.J.tocsc()).solve(F[0])
        return 1, 2

    def mult(self, y, Jy):
        Jy[0] = self.J @ y[0]

    def waxpy(self, w, a, y, x):
        w[0] = x[0] + a * y[0]

    def copy(self, dst, src):
        dst[0] = src[0].copy()

    def dot(self, a, b):
        return float(np.dot(a[0], b[0]))

    def norm(self, a):
        return float(np.linalg.norm(a[0]))

    def rellength(self, y, x):
        return float(np.max(np.abs(y[0]) / np.maximum(np.abs(x[0]), 1.0)))


def _problem(n=5, alpha=3.0, seed=2):
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    rng = np.random.default_rng(seed)
    xk = 0.3 * rng.standard_normal(orc.num_rows)
    x0 = np.zeros(orc.num_rows)
    x0[1::2] = -4.0  # far from the solution: exp(psi) has to grow by orders of magnitude, the full step overshoots
    return orc, (lambda z: orc.assemble_residual(z, xk, alpha)), (lambda z: orc.jacobian(z, alpha)), x0


def test_bt_matches_restated_petsc_loop():
    orc, res, jac, x0 = _problem()
    xo, reason_o, its_o, hist_o = osnes.newton_ls(res, jac, x0, linesearch="bt", rtol=1e-10, max_it=60)
    be = NumpyBackend(res, jac, orc.num_rows)
    nb = ls.NewtonBT(be, rtol=1e-10, max_it=60)
    x = [x0.copy()]
    hist = [nb.begin(x)]
    lams = []
    while not nb.reason:
        nb.step(x)
        hist.append(nb.fnorm)
        lams.append(nb.last_lambda)
    assert (nb.reason, nb.its) == (reason_o, its_o)
    assert reason_o > 0
    assert min(lams) < 1.0, "the case must make the line search backtrack"
    assert np.allclose(hist[: len(hist_o)], hist_o, rtol=1e-9, atol=0.0)
    assert np.linalg.norm(x[0] - xo) <= 1e-10 * np.linalg.norm(xo)


def test_bt_is_the_full_step_when_it_suffices():
    orc, res, jac, _ = _problem()
    x0 = np.zeros(orc.num_rows)
    xo, reason_o, its_o, hist_o = osnes.newton_ls_none(res, jac, x0, rtol=1e-8, max_it=50)
    be = NumpyBackend(res, jac, orc.num_rows)
    nb = ls.NewtonBT(be, rtol=1e-8, max_it=50)
    x = [x0.copy()]
    reason, its = nb.solve(x)
    # from this start the full Newton step always satisfies the sufficient-decrease test: bt == none
    assert (reason, its) == (reason_o, its_o)
    assert np.linalg.norm(x[0] - xo) <= 1e-12 * np.linalg.norm(xo)
    assert be.calls == its + 1  # one residual evaluation per step, no extra trial points


def test_bt_halves_through_overflow():
    """A start where the full step overflows exp(psi): the function is inf at lambda = 1 and the search halves
    until it is finite instead of giving up."""
    orc, res, jac, x0 = _problem()
    x0 = x0.copy()
    x0[1::2] = -60.0  # exp(-60) ~ 1e-26: the Newton step for psi is astronomically large
    be = NumpyBackend(res, jac, orc.num_rows)
    nb = ls.NewtonBT(be, rtol=1e-8, max_it=5)
    x = [x0.copy()]
    with np.errstate(over="ignore", invalid="ignore"):
        nb.begin(x)
        f0 = nb.fnorm
        nb.step(x)
    assert nb.reason in (0, ls.DIVERGED_LINE_SEARCH) or nb.reason > 0
    assert np.all(np.isfinite(x[0]))
    if nb.reason != ls.DIVERGED_LINE_SEARCH:
        assert np.isfinite(nb.fnorm) and nb.fnorm <= f0


import pytest  # noqa: E402


@pytest.mark.parametrize("maxstep", [1e8, 1.0])
def test_l2_matches_restated_petsc_loop(maxstep):
    """snes_linesearch_type l2 (SNESLineSearchApply_L2, one secant iteration): the host loop against the oracle's
    restatement, from a start where the full step is too long -- the first secant steps are 0.25 and 0.98; with
    maxstep = 1 (snes_linesearch_maxlambda 1, fracture_dolfinx.py:135) the later ones above 1 are cut to 1."""
    orc, res, jac, x0 = _problem()
    x0 = x0.copy()
    x0[1::2] = -2.0
    xo, reason_o, its_o, hist_o = osnes.newton_ls(res, jac, x0, linesearch="l2", rtol=1e-10, max_it=80, maxstep=maxstep)
    be = NumpyBackend(res, jac, orc.num_rows)
    nb = ls.NewtonBT(be, rtol=1e-10, max_it=80, linesearch="l2", maxstep=maxstep)
    x = [x0.copy()]
    hist, lams = [nb.begin(x)], []
    while not nb.reason:
        nb.step(x)
        hist.append(nb.fnorm)
        lams.append(nb.last_lambda)
    assert (nb.reason, nb.its) == (reason_o, its_o)
    assert reason_o > 0
    assert lams[0] < 0.5 and max(lams) <= max(1.0, min(maxstep, 2.0)), lams
    assert np.allclose(hist[: len(hist_o)], hist_o, rtol=1e-9, atol=0.0)
    assert np.linalg.norm(x[0] - xo) <= 1e-10 * np.linalg.norm(xo)
    assert be.calls == 1 + 3 * nb.its  # midpoint, end point and the accepted point per step


def test_l2_known_answers_on_a_quadratic():
    """phi(lambda) = ||F(x - lambda y)||^2 is exactly quadratic for a linear F, so the three-point differences are
    exact and one secant iteration lands on the minimiser of phi along y, whatever the bracket."""
    A = np.array([[3.0, 1.0], [1.0, 2.0]])
    b = np.array([1.0, -1.0])
    res = lambda z: A @ z - b  # noqa: E731
    x = np.array([2.0, 1.0])
    y = np.array([1.0, 0.5])   # not the Newton direction: the minimiser along it is not lambda = 1
    F0 = res(x)
    lam_star = float((A @ y) @ F0 / ((A @ y) @ (A @ y)))
    xn, Fn, gnorm, lam, ok = osnes.linesearch_l2(res, x, y, np.linalg.norm(F0))
    assert ok and abs(lam - lam_star) < 1e-12 and np.allclose(xn, x - lam_star * y)
    # maxstep = 1 (snes_linesearch_maxlambda 1, fracture_dolfinx.py:135): a secant step beyond 1 is not taken
    y2 = 0.25 * y  # minimiser at 4 lam_star > 1
    assert 4 * lam_star > 1.0
    xn, Fn, gnorm, lam, ok = osnes.linesearch_l2(res, x, y2, np.linalg.norm(F0), maxstep=1.0)
    assert ok and lam == 1.0
    import scipy.sparse as sp

    be = NumpyBackend(res, lambda z: sp.csr_matrix(A), 2)
    w, G = be.vector(), be.vector()
    g, lam_p, ok, _ = ls.linesearch_l2(be, [x], [y], w, G, float(np.linalg.norm(F0)))
    assert ok and abs(lam_p - lam_star) < 1e-12 and abs(g - gnorm_at(res, x, y, lam_star)) < 1e-12
    g, lam_p, ok, _ = ls.linesearch_l2(be, [x], [y2], w, G, float(np.linalg.norm(F0)), maxstep=1.0)
    assert ok and lam_p == 1.0


def gnorm_at(res, x, y, lam):
    return float(np.linalg.norm(res(x - lam * y)))


def test_l2_option_parsing():
    import proximalgalerkin_b200 as lvpp
    from proximalgalerkin_b200 import _capi
    import pytest

    o = lvpp.newton_options({"snes_linesearch_type": "l2", "snes_linesearch_maxlambda": 1})
    assert o.snes_linesearch == _capi.LINESEARCH_L2 and o.linesearch_maxstep == 1.0
    assert lvpp.newton_options({"snes_linesearch_type": "bt"}).linesearch_maxstep == 1e8
    with pytest.raises(NotImplementedError):
        lvpp.newton_options({"snes_linesearch_type": "l2"}, generic=True)  # the mixed-form engine knows none and bt
    with pytest.raises(NotImplementedError):
        lvpp.newton_options({"snes_linesearch_type": "cp"})
