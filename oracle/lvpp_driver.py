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


def solve_gradient_constraint(problem, max_iterations=25, alpha_scheme="doubling", alpha_0=1.0, alpha_c=1.0,
                              stopping_tol=1e-8, verbose=False):
    """examples/06_gradient_constraints/gradient_constraint_dolfinx.py:113-132,171-205: alpha schedule with i
    from 0 (:172-177), SNES newtonls / line search none / atol = rtol = stol = 1e-9 / max_it 20 (:118-131), stop on
    ||u - u0||_L2 < tol (:185,199), w0 <- sol (:205).  ``problem``: oracle.forms.GradientConstraintOracle."""
    x = np.zeros(problem.num_rows)
    problem.w0 = x.copy()
    hist = {"newton_steps": [], "alpha": [], "l2_diff": [], "reason": []}
    for i in range(max_iterations):
        if alpha_scheme == "linear":
            problem.alpha = alpha_0 + alpha_c * i
        elif alpha_scheme == "doubling":
            problem.alpha = alpha_0 * 2**i
        else:
            problem.alpha = alpha_0
        xn, reason, n, _ = snes.newton_ls(problem.assemble_residual, problem.jacobian, x, "none",
                                          rtol=1e-9, atol=1e-9, stol=1e-9, max_it=20)
        if reason <= 0:
            raise RuntimeError(f"SNES did not converge: reason {reason}")
        x = xn
        diff = np.sqrt(problem.l2_increment_sq(x, problem.w0))
        hist["newton_steps"].append(n)
        hist["alpha"].append(problem.alpha)
        hist["l2_diff"].append(diff)
        hist["reason"].append(reason)
        if verbose:
            print(f"iteration {i + 1}: alpha {problem.alpha} newton {n} |delta u| {diff:.3e}")
        if diff < stopping_tol:
            break
        problem.w0 = x.copy()
    return x, hist


def multiphase_initial_condition(coords, cells, num_species=4, tol=1e-14):
    """u_prev of multiphase_dolfinx.py:92-125: species 0 everywhere, then species 1 / 2 / 3 interpolated on
    the cells whose vertices all lie in three rectangles (locate_entities marks an entity when the
    marker is true at all of its vertices)."""
    x, y = coords[:, 0], coords[:, 1]
    rectangle = (0.2 - tol <= y) & (y <= 0.75 + tol) & (0.2 - tol <= x) & (x <= 0.8 + tol)
    lower_left = (y <= 0.5 + tol) & (0.2 - tol <= y) & (0.2 - tol <= x) & (x <= 0.5 + tol)
    lower_right = (y <= 0.5 + tol) & (0.2 <= y + tol) & (0.5 - tol <= x) & (x <= 0.8 + tol)
    u = np.zeros((coords.shape[0], num_species))
    u[:, 0] = 1.0
    for species, marker in ((1, rectangle), (2, lower_left), (3, lower_right)):
        inside = np.all(marker[cells], axis=1)
        nodes = np.unique(cells[inside])
        u[nodes, :] = 0.0
        u[nodes, species] = 1.0
    return u


def solve_multiphase(problem, num_steps=1, max_iterations=20, alpha_scheme="constant", alpha_0=1.0, alpha_c=1.0,
                     alpha_max=50.0, stopping_tol=1e-5, verbose=False):
    """multiphase_dolfinx.py:174-233: per time step psi <- ln(|u| + 1e-7) + 1 in sol and lvpp_old (:190-196),
    u_old <- 0; LVPP loop with i from 1 (:199-205), SNES newtonls / bt / atol = rtol = 1e-8 / max_it 25 (:128-143),
    lvpp_old <- sol, stop on ||u - u_old||_L2 < tol (:210-225); u_prev <- u (:227).
    ``problem``: oracle.forms.MultiphaseOracle with u_prev set."""
    ns, N = problem.NS, problem.N
    x = np.zeros(problem.num_rows)
    newton_its, lvpp_its = [], []
    for j in range(1, num_steps + 1):
        X = x.reshape(N, 3, ns)
        psi0 = np.log(np.abs(X[:, 0, :]) + 1e-7) + 1.0
        X[:, 2, :] = psi0
        problem.lvpp_old.reshape(N, 3, ns)[:, 2, :] = psi0
        u_old = np.zeros((N, ns))
        total = 0
        for i in range(1, max_iterations + 1):
            if alpha_scheme == "linear":
                problem.alpha = min(alpha_0 + alpha_c * i, alpha_max)
            elif alpha_scheme == "doubling":
                problem.alpha = min(alpha_0 * 2**i, alpha_max)
            else:
                problem.alpha = alpha_0
            xn, reason, n, _ = snes.newton_ls(problem.assemble_residual, problem.jacobian, x, "bt",
                                              rtol=1e-8, atol=1e-8, max_it=25)
            if reason <= 0:
                raise RuntimeError(f"SNES did not converge: reason {reason}")
            x = xn
            total += n
            diff = np.sqrt(problem.l2_increment_sq(x, u_old))
            if verbose:
                print(f"step {j} iteration {i}: alpha {problem.alpha} newton {n} |delta u| {diff:.3e}")
            u_old = x.reshape(N, 3, ns)[:, 0, :].copy()
            problem.lvpp_old = x.copy()
            if diff < stopping_tol:
                break
        problem.u_prev = x.reshape(N, 3, ns)[:, 0, :].copy()
        newton_its.append(total)
        lvpp_its.append(i)
    return x, {"newton_iterations": newton_its, "lvpp_iterations": lvpp_its}


def solve_signorini(problem, max_iterations=25, alpha_scheme="doubling", alpha_0=1.0, alpha_c=1.0,
                    newton_tol=1e-6, tol=1e-6, verbose=False):
    """examples/02_signorini/signorini_dolfinx.py:317-358: it from 1; alpha (:324-329); SNES tolerances 10 tol on
    the first step, as atol and rtol (:331-332); line search none, max_it PETSc default 50; stop on the discrete
    ||u - u_prev||_2 <= tol (:337-342); u_prev <- u, psi_k <- psi (:343-344).
    ``problem``: oracle.forms.SignoriniOracle."""
    x = np.zeros(problem.num_rows)
    nu = problem.gd * problem.N
    u_prev = np.zeros(nu)
    iterations = []
    it = 0
    for it in range(1, max_iterations + 1):
        if alpha_scheme == "linear":
            problem.alpha = alpha_0 + alpha_c * it
        elif alpha_scheme == "doubling":
            problem.alpha = alpha_0 * 2**it
        else:
            problem.alpha = alpha_0
        stol = 10 * newton_tol if it < 2 else newton_tol
        xn, reason, n, _ = snes.newton_ls(problem.assemble_residual, problem.jacobian, x, "none",
                                          rtol=stol, atol=stol, max_it=50)
        if reason <= 0:
            raise RuntimeError(f"SNES did not converge: reason {reason}")  # snes_error_if_not_converged (:279)
        x = xn
        iterations.append(n)
        normed_diff = np.linalg.norm(x[:nu] - u_prev)
        if verbose:
            print(f"it {it}: alpha {problem.alpha} newton {n} increment {normed_diff:.3e}")
        if normed_diff <= tol:
            break
        u_prev = x[:nu].copy()
        problem.psi_k = x[nu:].copy()
    return x, {"it": it, "iterations": iterations}


def solve_obstacle_adaptive(problem, max_outer=100, alpha_0=1.0, r=2.0, nfail_max=50, alpha_max=None, tol_exit=1e-6,
                            snes_rtol=1e-6, snes_max_it=100, newton=None, verbose=False):
    """The failure-recovering outer loop of the reference, written for the
    obstacle problem.  Restates examples/03_fracture/fracture_dolfinx.py:215-283 (the same loop stands in
    examples/07_eigenvalue_constraints/eigenvalue_constraints_dolfinx.py:163-225 and
    examples/08_intersecting_constraints/intersecting_constraints_dolfinx.py:120-174):

    * alpha starts at ``alpha_0`` (= 1, :215), r = 2 (:218), failures are counted in ``nfail`` (:219);
    * a solve *fails* when SNES diverged (reason < 0, :239-240) or when it converged without doing a Newton step
      (:234-238: alpha has been reduced so far that the initial guess satisfies the equation);
    * on failure (:241-262): nfail += 1, alpha /= 2, the unknown goes back to the last accepted proximal iterate
      (z_prev on the first proximal step, z_iter afterwards -- both are ``xk`` here, :250-253), give up when
      nfail >= nfail_max (:255-260), else try again with the same k;
    * on success: stop on the increment (:265-273; here the H1 increment of obstacle_pg.py:203,222); alpha *= r when
      the solve took <= 4 Newton steps, alpha /= r when it took >= 10 (:276-279); z_iter <- z (:282); k += 1.

    ``alpha_max`` (None in the reference loops) clamps alpha like obstacle_pg.py:184.  ``newton(x0, xk, alpha) ->
    (x, reason, its)`` replaces the SNES solve (tests script failures with it).  Returns (x, history)."""
    x = np.zeros(problem.num_rows)
    xk = x.copy()
    alpha = float(alpha_0)
    hist = {k: [] for k in ("newton_steps", "alpha", "primal_increment", "reason", "attempts")}
    k, nfail = 1, 0
    gave_up = False

    def snes_solve(x0, xk_, a):
        xn, reason, n, _ = snes.newton_ls_none(lambda z: problem.assemble_residual(z, xk_, a), lambda z: problem.jacobian(z, a),
                                               x0, rtol=snes_rtol, max_it=snes_max_it)
        return xn, reason, n

    newton = newton or snes_solve
    while nfail <= nfail_max and k <= max_outer:
        xn, reason, n = newton(x, xk, alpha)
        hist["attempts"].append((k, alpha, int(n), int(reason)))
        if reason < 0 or (n == 0 and reason > 0):
            nfail += 1
            if verbose:
                print(f"failed to converge ({reason}), k={k} alpha={alpha}")
            alpha /= 2
            x = xk.copy()
            if nfail >= nfail_max:
                gave_up = True
                break
            continue
        x = xn
        obs = problem.observables(x, xk, alpha)
        increment = float(np.sqrt(obs[4]))
        hist["newton_steps"].append(int(n))
        hist["alpha"].append(alpha)
        hist["primal_increment"].append(increment)
        hist["reason"].append(int(reason))
        if verbose:
            print(f"solved k={k} newton {n} alpha={alpha} increment {increment:.3e}")
        if increment < tol_exit:
            break
        if n <= 4:
            alpha *= r
        elif n >= 10:
            alpha /= r
        if alpha_max is not None:
            alpha = min(alpha, alpha_max)
        xk = x.copy()
        k += 1
    hist["nfail"] = nfail
    hist["gave_up"] = gave_up
    return x, hist
