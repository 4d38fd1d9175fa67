"""LvppStepper (the device-resident outer / Newton loop bench.py times) over a numpy stand-in for the device: the control
flow -- alpha schedule, SNES reasons, observables / stopping test, sol_k <- sol, the `same_iterate` hint, reset() -- is
checked on the CPU against the oracle's restated loop.  The GPU parity of the same loop is in tests/test_gpu_mg.py."""
import types

import numpy as np
import scipy.sparse.linalg as spla
import torch

from oracle import lvpp_driver, mesh as omesh, obstacle as oobs
from proximalgalerkin_b200 import obstacle_pg
from proximalgalerkin_b200.problem import newton_options


class _Vec:
    def __init__(self, n):
        self.tensor = torch.zeros(n, dtype=torch.float64)


class _NumpyDevice:
    """What LvppStepper touches of problem.DeviceProblem, with exact (sparse LU) Newton steps of the oracle."""

    def __init__(self, orc):
        self.orc, self.n, self.device = orc, orc.num_rows, "cpu"
        self.x = _Vec(self.n)
        self.alpha, self.xk, self.F = 1.0, np.zeros(self.n), None
        self.same_iterate_calls, self.begin_calls, self.assemblies = 0, 0, 0
        self._last_evaluated = None

    def set_alpha(self, a):
        self.alpha = float(a)

    def set_previous(self, xk):
        self.xk = xk.tensor.numpy().copy()

    def newton_begin(self, x, same_iterate=False):
        xa = x.tensor.numpy()
        self.begin_calls += 1
        if same_iterate:  # the hint must only be given at the iterate of the last evaluation
            assert self._last_evaluated is not None and np.array_equal(xa, self._last_evaluated)
            self.same_iterate_calls += 1
        else:
            self.assemblies += 1
        self.F = self.orc.assemble_residual(xa, self.xk, self.alpha)
        self._last_evaluated = xa.copy()
        return float(np.linalg.norm(self.F))

    def newton_step(self, x, opts):
        xa = x.tensor.numpy()
        y = spla.splu(self.orc.jacobian(xa, self.alpha).tocsc()).solve(self.F)
        xa -= y  # in place: the tensor shares the memory
        self.F = self.orc.assemble_residual(xa, self.xk, self.alpha)
        self._last_evaluated = xa.copy()
        self.assemblies += 1
        return (float(np.linalg.norm(self.F)), float(np.linalg.norm(y)), float(np.linalg.norm(xa))), 1, 2

    def observables(self, x):
        return self.orc.observables(x.tensor.numpy(), self.xk, self.alpha)


def _stepper(orc, scheme, alpha_max, tol, monkeypatch, max_outer=500):
    dev = _NumpyDevice(orc)
    opts = dict(obstacle_pg.PETSC_OPTIONS)
    s = {"problem": types.SimpleNamespace(device_problem=dev), "options": opts}
    monkeypatch.setattr(obstacle_pg, "DeviceVector", lambda n, device: _Vec(n))
    st = obstacle_pg.LvppStepper(types.SimpleNamespace(rank=0), 1, scheme, alpha_max, tol, max_outer=max_outer, setup_objects=s)
    assert st.opts.snes_rtol == newton_options(opts).snes_rtol == 1e-6
    return st, dev


def test_stepper_follows_the_restated_outer_loop_and_hints_same_iterate(monkeypatch):
    orc = oobs.ObstacleOracle(omesh.rectangle(14, 14))
    xo, ho = lvpp_driver.solve_obstacle(orc, max_outer=100, alpha_scheme="double_exponential", alpha_max=1e2, tol_exit=1e-4)
    st, dev = _stepper(orc, "double_exponential", 1e2, 1e-4, monkeypatch)
    steps = 0
    while st.step():
        steps += 1
    assert st.finished and st.history["newton_steps"] == ho["newton_steps"] and st.history["reason"] == ho["reason"]
    assert np.allclose(st.history["alpha"], ho["alpha"]) and np.allclose(st.history["primal_increment"], ho["primal_increment"], rtol=1e-9)
    assert np.allclose(st.x.tensor.numpy(), xo, rtol=0, atol=1e-11)
    assert st.total_newton == sum(ho["newton_steps"]) == steps + 1
    outer = len(ho["newton_steps"])
    # every proximal step but the first begins at the iterate the previous one ended at: D(psi) is kept there
    assert dev.begin_calls == outer and dev.same_iterate_calls == outer - 1
    # reset(): a fresh solve from the zero iterate reproduces the history; its first begin is a full evaluation
    st.reset()
    assert not st.finished and st.k == 0 and float(st.x.tensor.abs().max()) == 0.0
    while st.step():
        pass
    assert st.history["newton_steps"] == ho["newton_steps"] and st.solves_completed == 2
    assert dev.begin_calls == 2 * outer and dev.same_iterate_calls == 2 * (outer - 1)


def test_stepper_constant_schedule_is_the_script_default(monkeypatch):
    orc = oobs.ObstacleOracle(omesh.rectangle(10, 10))
    xo, ho = lvpp_driver.solve_obstacle(orc, max_outer=30, alpha_scheme="constant", alpha_max=1e5, tol_exit=1e-6)
    st, dev = _stepper(orc, "constant", 1e5, 1e-6, monkeypatch, max_outer=30)  # (--max-iter; the script's default is 100)
    while st.step():
        pass
    assert st.history["newton_steps"] == ho["newton_steps"] and set(st.history["alpha"]) == {1.0}
    assert np.allclose(st.x.tensor.numpy(), xo, rtol=0, atol=1e-11)
