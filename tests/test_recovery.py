"""alpha_scheme="adaptive": the failure-recovering proximal step control (SURVEY.md 8f N2;
proximalgalerkin_b200/recovery.py) against the oracle's restatement of fracture_dolfinx.py:215-283.

CPU tests: the product's driver (obstacle_pg.solve_problem) runs unchanged over a numpy stand-in for the device
problem, so that its control flow -- which solve counts as failed, what is restored, how alpha moves -- is compared
with the restated loop attempt by attempt.  The GPU parity test is in tests/test_gpu_linesearch.py.
"""
import types

import numpy as np
import pytest

from oracle import lvpp_driver, mesh as omesh, obstacle as oobs, snes
from proximalgalerkin_b200 import obstacle_pg, recovery


def test_adaptive_alpha_rule_known_answers():
    c = recovery.AdaptiveAlpha()
    assert (c.alpha, c.k, c.nfail) == (1.0, 1, 0)  # fracture_dolfinx.py:215-219
    c.accepted(4)      # <= 4 Newton steps: alpha *= r (:276-277)
    assert (c.alpha, c.k) == (2.0, 2)
    c.accepted(5)      # 5 .. 9: unchanged
    assert (c.alpha, c.k) == (2.0, 3)
    c.accepted(10)     # >= 10: alpha /= r (:278-279)
    assert (c.alpha, c.k) == (1.0, 4)
    c.failed()         # alpha /= 2, same k (:242-249)
    assert (c.alpha, c.k, c.nfail) == (0.5, 4, 1)
    assert recovery.AdaptiveAlpha.is_failure(-3, 7) and recovery.AdaptiveAlpha.is_failure(3, 0)  # :234-240
    assert not recovery.AdaptiveAlpha.is_failure(3, 1)
    c = recovery.AdaptiveAlpha(alpha_max=3.0)
    c.accepted(1), c.accepted(1)
    assert c.alpha == 3.0
    c = recovery.AdaptiveAlpha(nfail_max=2)
    c.failed()
    with pytest.raises(recovery.GaveUp):  # nfail >= NFAIL_MAX (:255-260)
        c.failed()


class _Solver:
    reason, its = 0, 0

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its

    def getLinearSolveIterations(self):
        return self.its


class _NumpyProblem:
    """NonlinearProblem-shaped stand-in: the oracle's Newton loop on the host arrays the driver mutates."""

    def __init__(self, orc, sol, sol_k, alpha, max_it, fail_on=(), raises=False):
        self.orc, self.sol, self.sol_k, self.alpha, self.max_it = orc, sol, sol_k, alpha, max_it
        self.fail_on, self.raises, self.calls = set(fail_on), raises, 0
        self.solver = _Solver()
        dev = types.SimpleNamespace()
        dev.x = types.SimpleNamespace(arr=None)
        dev.x.set = lambda a: setattr(dev.x, "arr", np.array(a))
        dev.observables = lambda x: orc.observables(x.arr, sol_k.x.array, alpha.value)
        self.device_problem = dev

    def newton(self, x0, xk, a):
        self.calls += 1
        if self.calls in self.fail_on:
            return x0, -3, 2
        xn, reason, n, _ = snes.newton_ls_none(lambda z: self.orc.assemble_residual(z, xk, a), lambda z: self.orc.jacobian(z, a),
                                               x0, rtol=1e-6, max_it=self.max_it)
        return xn, reason, n

    def solve(self):
        xn, reason, n = self.newton(self.sol.x.array.copy(), self.sol_k.x.array.copy(), self.alpha.value)
        self.solver.reason, self.solver.its = reason, n
        if reason > 0:  # u is overwritten only on convergence (src/lvpp/problem.py:121-123)
            self.sol.x.array[:] = xn
        elif self.raises:
            raise obstacle_pg.NotConvergedError("SNES did not converge", reason, n)


def _patched(monkeypatch, orc, max_it, fail_on=(), raises=False):
    def fake_setup(msh, polynomial_order=1, **kw):
        sol = types.SimpleNamespace(x=types.SimpleNamespace(array=np.zeros(orc.num_rows)))
        sol_k = types.SimpleNamespace(x=types.SimpleNamespace(array=np.zeros(orc.num_rows)))
        alpha = types.SimpleNamespace(value=1.0)
        return dict(sol=sol, sol_k=sol_k, alpha=alpha, problem=_NumpyProblem(orc, sol, sol_k, alpha, max_it, fail_on, raises))

    monkeypatch.setattr(obstacle_pg, "setup", fake_setup)
    return types.SimpleNamespace(rank=0)


@pytest.mark.parametrize("max_it,fail_on,raises", [(100, (), False), (100, (2, 3, 6), False), (100, (1,), True), (4, (), False), (3, (), False)])
def test_solve_problem_adaptive_follows_restated_loop(monkeypatch, max_it, fail_on, raises):
    orc = oobs.ObstacleOracle(omesh.rectangle(12, 12))
    msh = _patched(monkeypatch, orc, max_it, fail_on, raises)
    sol, total, h = obstacle_pg.solve_problem(msh, 1, 40, "adaptive", 1e5, 1e-5, adaptive=dict(nfail_max=12))
    ref = _NumpyProblem(orc, None, None, None, max_it, fail_on)
    xo, ho = lvpp_driver.solve_obstacle_adaptive(orc, max_outer=40, alpha_max=1e5, tol_exit=1e-5, nfail_max=12,
                                                 snes_max_it=max_it, newton=ref.newton)
    assert [tuple(a) for a in h["attempts"]] == [tuple(a) for a in ho["attempts"]]
    assert h["newton_steps"] == ho["newton_steps"] and h["alpha"] == ho["alpha"]
    assert bool(h.get("gave_up", False)) == ho["gave_up"]
    assert np.allclose(h["primal_increment"], ho["primal_increment"], rtol=1e-12, atol=0)
    assert np.array_equal(sol.x.array, xo)
    if fail_on:
        assert ho["nfail"] == len(fail_on) == sum(a[3] < 0 for a in ho["attempts"])
    if max_it == 4:  # the first proximal step needs 5 Newton steps at alpha = 1: one real failure, recovered at alpha = 1/2
        assert ho["attempts"][0][2:] == (4, -5) and ho["attempts"][1][:2] == (1, 0.5) and ho["nfail"] == 1
    if max_it == 3:  # every alpha >= 2^-11 needs more than 3 Newton steps from the zero start: the loop gives up
        assert ho["gave_up"] and ho["newton_steps"] == []
    else:
        assert not ho["gave_up"] and ho["primal_increment"][-1] < 1e-5


def test_adaptive_reaches_the_fixed_schedule_solution():
    """Same discrete problem, different alpha path: the LVPP limit does not depend on the schedule."""
    orc = oobs.ObstacleOracle(omesh.rectangle(12, 12))
    xa, ha = lvpp_driver.solve_obstacle_adaptive(orc, tol_exit=1e-8)
    xs, hs = lvpp_driver.solve_obstacle(orc, 100, "double_exponential", 1e2, 1e-8)
    u_a, u_s = xa[0::2], xs[0::2]
    assert np.linalg.norm(u_a - u_s) < 1e-5 * np.linalg.norm(u_s)
    assert ha["nfail"] == 0 and ha["alpha"][0] == 1.0 and max(ha["alpha"]) > 1e2  # alpha keeps doubling: no clamp in the loop
