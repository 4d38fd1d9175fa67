"""The proximal outer loop of the obstacle example (oracle; see oracle/__init__.py).

Restates examples/01_obstacle_problem/obstacle_pg.py:154-227: alpha schedule (:175-186, constants
C=1, r=1.5, q=1.5 at :161-163; alpha_k is stored *before* clamping, :182-183), Newton solve (:190),
observables (:196-201), stop on the H1 increment (:203,222), sol_k <- sol (:226).
"""
import numpy as np

from . import snes


def alpha_schedule(rule, k, alpha_prev, alpha_max, C=1.0, r=1.5, q=1.5, alpha_current=1.0):
    """One evaluation of obstacle_pg.py:175-186.  Returns (alpha, alpha_k_new)."""
    if rule == "constant":
        return C, alpha_prev
    if rule == "double_exponential":
        alpha = alpha_current
        try:
            alpha = max(C * r ** (q**k) - alpha_prev, C)
        except OverflowError:
            pass
        alpha_k = alpha
        return min(alpha, alpha_max), alpha_k
    return C * r**k, alpha_prev


def solve_obstacle(problem, max_outer=100, alpha_scheme="constant", alpha_max=1e5, tol_exit=1e-6,
                   snes_rtol=1e-6, snes_max_it=100, linear_solve=None, verbose=False):
    """Returns (x, history dict).  ``problem`` is an oracle.obstacle.ObstacleOracle."""
    x = np.zeros(problem.num_rows)
    xk = x.copy()
    alpha_k = 1
    alpha = 1.0
    hist = {k: [] for k in ("energy", "complementarity", "feasibility", "dual_feasibility", "newton_steps",
                            "alpha", "primal_increment", "latent_increment", "reason", "fnorms")}
    for k in range(max_outer):
        alpha, alpha_k = alpha_schedule(alpha_scheme, k, alpha_k, alpha_max, alpha_current=alpha)
        xn, reason, n, fn = snes.newton_ls_none(
            lambda z: problem.assemble_residual(z, xk, alpha),
            lambda z: problem.jacobian(z, alpha),
            x, rtol=snes_rtol, max_it=snes_max_it, linear_solve=linear_solve,
        )
        if reason <= 0:
            raise RuntimeError(f"SNES did not converge: reason {reason}")  # snes_error_if_not_converged
        x = xn
        obs = problem.observables(x, xk, alpha)
        increment = np.sqrt(obs[4])
        hist["energy"].append(obs[0])
        hist["complementarity"].append(abs(obs[1]))
        hist["feasibility"].append(obs[2])
        hist["dual_feasibility"].append(obs[3])
        hist["newton_steps"].append(n)
        hist["alpha"].append(alpha)
        hist["primal_increment"].append(increment)
        hist["latent_increment"].append(np.sqrt(obs[5]))
        hist["reason"].append(reason)
        hist["fnorms"].append(fn)
        if verbose:
            print(f"outer {k+1} alpha {alpha:.6g} newton {n} reason {reason} incr {increment:.3e}")
        if increment < tol_exit:
            break
        xk = x.copy()
    return x, hist
