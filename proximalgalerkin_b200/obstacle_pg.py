"""The LVPP obstacle driver of the reference, on the B200 path.

``solve_problem`` follows examples/01_obstacle_problem/obstacle_pg.py:53-264 statement by statement
(spaces :68-70, BCs :76-83, obstacle :92-111, forms :116-125, options :128-139, outer loop :154-227)
with this package's ``fem`` / ``NonlinearProblem`` in place of dolfinx's, and a mesh object in place
of the XDMF file name.  ``LvppStepper`` runs the same loop device-resident, one Newton step per call
(the granularity bench.py times).
"""
import numpy as np

from . import fem, recovery
from .problem import NonlinearProblem, NotConvergedError, derivative, newton_options, obstacle_residual, DeviceVector


def alpha_update(rule, k, alpha_value, alpha_k, alpha_max, C=1.0, r=1.5, q=1.5):
    """obstacle_pg.py:175-186.  Returns (alpha.value, alpha_k)."""
    if rule == "constant":
        alpha_value = C
    elif rule == "double_exponential":
        try:
            alpha_value = max(C * r ** (q**k) - alpha_k, C)
        except OverflowError:
            pass
        alpha_k = alpha_value
        alpha_value = min(alpha_value, alpha_max)
    else:
        alpha_value = C * r**k
    return alpha_value, alpha_k


PETSC_OPTIONS = {  # obstacle_pg.py:128-139
    "ksp_type": "preonly",
    "pc_type": "lu",
    "pc_factor_mat_solver_type": "mumps",
    "ksp_error_if_not_converged": True,
    "snes_error_if_not_converged": True,
    "snes_linesearch_type": "none",
    "snes_rtol": 1e-6,
    "snes_max_it": 100,
}


def setup(msh, polynomial_order=1, quadrature_degree=6, obstacle="phi_set", f_value=0.0, petsc_options=None,
          obstacle_period=None, obstacle_origin=0.0, rule=None):
    """Everything obstacle_pg.py builds before the outer loop.  Returns a dict of the objects.
    ``rule``: explicit quadrature rule (points, weights) instead of the table of ``quadrature_degree`` -- e.g.
    basix's, exported by tools/export_from_dolfinx.py."""
    V = fem.functionspace(msh, ("Lagrange", polynomial_order), quadrature_degree=quadrature_degree, rule=rule)
    alpha = fem.Constant(msh, 1.0)
    f = fem.Constant(msh, f_value)
    dofs = fem.locate_dofs_boundary(V.sub(0))
    bcs = fem.dirichletbc(value=0.0, dofs=dofs, V=V.sub(0))
    sol = fem.Function(V)
    sol_k = fem.Function(V)
    phi = fem.QuadratureFunction(V, name="phi")
    if obstacle == "phi_set":
        phi.interpolate_phi_set(obstacle_period, obstacle_origin)  # closed form evaluated on the device
    elif callable(obstacle):
        phi.interpolate(obstacle)
    else:
        raise ValueError(obstacle)
    F = obstacle_residual(sol, sol_k, alpha, f, phi)
    J = derivative(F, sol)
    opts = dict(PETSC_OPTIONS)
    opts.update(petsc_options or {})
    problem = NonlinearProblem(F, u=sol, bcs=[bcs], J=J, petsc_options=opts, petsc_options_prefix="obstacle_")
    return dict(V=V, alpha=alpha, f=f, bcs=bcs, sol=sol, sol_k=sol_k, phi=phi, F=F, J=J, problem=problem, options=opts)


def solve_problem(msh, polynomial_order=1, maximum_number_of_outer_loop_iterations=100, alpha_scheme="constant",
                  alpha_max=1e5, tol_exit=1e-6, obstacle="phi_set", petsc_options=None, verbose=False, adaptive=None,
                  output_dir=None):
    """Returns (sol, total Newton steps, history dict) -- the reference returns (sol, sum(Newton_steps))
    and writes the history to CSV (obstacle_pg.py:245-264).

    ``alpha_scheme="adaptive"`` (not in obstacle_pg.py) replaces the fixed schedule by the
    failure-recovering control of fracture_dolfinx.py:215-283: alpha doubles after a proximal step of <= 4 Newton steps,
    halves after one of >= 10, and a failed Newton solve halves alpha, restores ``sol`` from ``sol_k`` and repeats
    the step.  ``adaptive`` passes keyword arguments to ``recovery.AdaptiveAlpha`` (alpha_0, r, nfail_max)."""
    s = setup(msh, polynomial_order, obstacle=obstacle, petsc_options=petsc_options)
    sol, sol_k, alpha, problem = s["sol"], s["sol_k"], s["alpha"], s["problem"]
    dev = problem.device_problem
    sol.x.array[:] = 0.0
    sol_k.x.array[:] = sol.x.array[:]
    alpha_k = 1
    hist = {k: [] for k in ("energy", "complementarity", "feasibility", "dual_feasibility", "newton_steps",
                            "alpha", "primal_increment", "latent_increment", "reason", "krylov_iterations")}
    ctl = None
    if alpha_scheme == "adaptive":  # the recovery loop of fracture_dolfinx.py:215-283 (recovery.py)
        ctl = recovery.AdaptiveAlpha(alpha_max=alpha_max, **(adaptive or {}))
        hist["attempts"] = ctl.attempts
    k = 0
    while k < maximum_number_of_outer_loop_iterations:
        if ctl is None:
            alpha.value, alpha_k = alpha_update(alpha_scheme, k, alpha.value, alpha_k, alpha_max)
            problem.solve()
        else:
            alpha.value = ctl.alpha
            try:
                problem.solve()
            except NotConvergedError:  # *_error_if_not_converged: the reason is on the solver either way
                pass  # (a device / communication failure is an LvppError and propagates)
        reason = problem.solver.getConvergedReason()
        n = problem.solver.getIterationNumber()
        if ctl is not None:
            ctl.record(n, reason)
            if ctl.is_failure(reason, n):
                if verbose and msh.rank == 0:
                    print(f"Failed to converge ({reason}), k={ctl.k} alpha={alpha.value}")
                sol.x.array[:] = sol_k.x.array[:]  # back to the last accepted proximal iterate (:250-253)
                try:
                    ctl.failed()
                except recovery.GaveUp:
                    hist["gave_up"] = True
                    break
                continue
        # observables (obstacle_pg.py:196-201): evaluated on the device at the iterate just computed
        dev.x.set(sol.x.array)
        obs = dev.observables(dev.x)
        increment = float(np.sqrt(obs[4]))
        hist["energy"].append(obs[0])
        hist["complementarity"].append(abs(obs[1]))
        hist["feasibility"].append(obs[2])
        hist["dual_feasibility"].append(obs[3])
        hist["newton_steps"].append(n)
        hist["alpha"].append(alpha.value)
        hist["primal_increment"].append(increment)
        hist["latent_increment"].append(float(np.sqrt(obs[5])))
        hist["reason"].append(reason)
        hist["krylov_iterations"].append(problem.solver.getLinearSolveIterations())
        if verbose and msh.rank == 0:
            print(f"OUTER LOOP {k + 1} alpha: {alpha.value}  Newton steps: {n}  Converged: {reason}  "
                  f"Increment size: {increment}")
        if increment < tol_exit:
            break
        if ctl is not None:
            ctl.accepted(n)
        sol_k.x.array[:] = sol.x.array[:]
        k += 1
    if output_dir is not None and msh.rank == 0:
        # obstacle_pg.py:229-259: the solution for ParaView (VTX there, .vtu here; P1 fields) and the history table as CSV
        from pathlib import Path

        from . import io

        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        if polynomial_order == 1 and msh.nranks == 1:
            io.write_solution(out / "solution.vtu", s["V"], sol.x.array)
        io.write_history_csv(out / f"pg_{alpha_scheme}_p{polynomial_order}.csv", hist, dofs=s["V"].num_rows // 2)
    return sol, sum(hist["newton_steps"]), hist


class LvppStepper:
    """The same outer/Newton loop with every vector resident in HBM, advanced one Newton step at a
    time.  ``step()`` returns True while the LVPP iteration is still running."""

    def __init__(self, msh, polynomial_order=1, alpha_scheme="double_exponential", alpha_max=1e2, tol_exit=1e-4,
                 max_outer=500, obstacle="phi_set", petsc_options=None, setup_objects=None, obstacle_period=None,
                 obstacle_origin=0.0, adaptive=None):
        s = setup_objects or setup(msh, polynomial_order, obstacle=obstacle, petsc_options=petsc_options,
                                   obstacle_period=obstacle_period, obstacle_origin=obstacle_origin)
        self.s = s
        self.msh = msh
        self.dev = s["problem"].device_problem
        self.opts = newton_options(s["options"])
        self.alpha_scheme, self.alpha_max, self.tol_exit, self.max_outer = alpha_scheme, alpha_max, tol_exit, max_outer
        dev = self.dev
        self.x = dev.x
        self.xk = DeviceVector(dev.n, dev.device)
        self.total_newton = 0
        self.total_krylov = 0
        self.solves_completed = 0
        self._adaptive = adaptive
        self.nb, self.last_lambda = None, 1.0
        if self.opts.snes_linesearch != 0:
            from . import linesearch

            o = self.opts
            self.nb = linesearch.NewtonBT(linesearch.DeviceBackend(dev, o), rtol=o.snes_rtol, atol=o.snes_atol,
                                          stol=o.snes_stol, max_it=o.snes_max_it, divtol=o.snes_divtol,
                                          linesearch="l2" if o.snes_linesearch == 2 else "bt",
                                          maxstep=getattr(o, "linesearch_maxstep", 1e8))
        self.reset()

    def reset(self):
        """Start a fresh LVPP solve from the zero iterate (obstacle_pg.py:157-158) on the same handle: bench.py times
        K Newton steps whatever K is, so a solve that finishes inside the timed region is followed by the next one."""
        self.x.tensor.zero_()
        self.xk.tensor.zero_()
        self._x_is_last_evaluated = False
        self.k = 0  # outer iteration
        self.alpha_value, self.alpha_k = 1.0, 1
        self.newton_its = 0  # within the current outer iteration
        self.krylov_outer = 0  # Krylov iterations within the current outer iteration
        self.finished = False
        self.history = {k: [] for k in ("newton_steps", "alpha", "primal_increment", "reason", "krylov_iterations")}
        # alpha_scheme "adaptive": failure-recovering control of fracture_dolfinx.py:215-283 (recovery.py)
        self.ctl = (recovery.AdaptiveAlpha(alpha_max=self.alpha_max, **(self._adaptive or {}))
                    if self.alpha_scheme == "adaptive" else None)
        if self.ctl is not None:
            self.history["attempts"] = self.ctl.attempts
        self._begin_outer()

    def _begin_outer(self):
        if self.ctl is not None:
            self.alpha_value = self.ctl.alpha
        else:
            self.alpha_value, self.alpha_k = alpha_update(self.alpha_scheme, self.k, self.alpha_value, self.alpha_k, self.alpha_max)
        self.dev.set_alpha(self.alpha_value)
        self.dev.set_previous(self.xk)
        if self.nb is not None:  # snes_linesearch_type bt: host loop over the library's entry points
            self.fnorm0 = self.nb.begin(self.x)
        else:
            # after an accepted proximal step x is the iterate of the last residual evaluation: D(psi) is kept
            self.fnorm0 = self.dev.newton_begin(self.x, same_iterate=self._x_is_last_evaluated)
            self._x_is_last_evaluated = False
        self.ttol = self.fnorm0 * self.opts.snes_rtol
        self.newton_its = 0
        self.krylov_outer = 0

    def _snes_reason(self, it, xnorm, snorm, fnorm):
        o = self.opts
        if not np.isfinite(fnorm):
            return -4
        if fnorm < o.snes_atol:
            return 2
        if it:
            if fnorm <= self.ttol:
                return 3
            if snorm < o.snes_stol * xnorm:
                return 4
            if o.snes_divtol > 0 and fnorm > o.snes_divtol * self.fnorm0:
                return -9
        return 0

    def step(self):
        """One Newton step (Krylov solve, update, residual + Jacobian at the new iterate); closes the
        outer iteration (observables, stopping test, sol_k <- sol, alpha update) when SNES converges."""
        if self.finished:
            return False
        if self.nb is not None:
            lin0 = self.nb.linear_its
            reason = self.nb.step(self.x)
            self.last_lambda = self.nb.last_lambda
            self.newton_its = self.nb.its
            self.total_newton += 1
            self.total_krylov += self.nb.linear_its - lin0
            self.krylov_outer += self.nb.linear_its - lin0
        else:
            (fnorm, ynorm, xnorm), kits, kreason = self.dev.newton_step(self.x, self.opts)
            self.newton_its += 1
            self.total_newton += 1
            self.total_krylov += kits
            self.krylov_outer += kits
            reason = -3 if kreason < 0 else self._snes_reason(self.newton_its, xnorm, ynorm, fnorm)
            if reason == 0 and self.newton_its >= self.opts.snes_max_it:
                reason = -5
        if reason == 0:
            return True
        if self.ctl is not None:
            self.ctl.record(self.newton_its, reason)
            if self.ctl.is_failure(reason, self.newton_its):
                # halve alpha, go back to the last accepted proximal iterate, repeat the step (:241-262)
                self.x.tensor.copy_(self.xk.tensor)
                try:
                    self.ctl.failed()
                except recovery.GaveUp:
                    self.finished = True
                    self.history["gave_up"] = True
                    return False
                self._begin_outer()
                return True
        if reason < 0:
            raise NotConvergedError(f"SNES did not converge: reason {reason}", reason, self.newton_its)
        obs = self.dev.observables(self.x)
        increment = float(np.sqrt(obs[4]))
        self.history["newton_steps"].append(self.newton_its)
        self.history["alpha"].append(self.alpha_value)
        self.history["primal_increment"].append(increment)
        self.history["reason"].append(reason)
        self.history["krylov_iterations"].append(self.krylov_outer)
        self.k += 1
        if increment < self.tol_exit or self.k >= self.max_outer:
            self.finished = True
            self.solves_completed += 1
            return False
        if self.ctl is not None:
            self.ctl.accepted(self.newton_its)
        self.xk.tensor.copy_(self.x.tensor)
        self._x_is_last_evaluated = self.nb is None  # the library's last residual / Jacobian evaluation was at this x
        self._begin_outer()
        return True
