"""examples/04_multiphase/multiphase_dolfinx.py on the GPU: ``solve_problem`` keeps the reference's
signature and return value (Newton and LVPP iterations per time step).

Four-species Cahn-Hilliard, (u, z, psi) in (P1)^4 each on a crossed-diagonal unit square (:34-45),
softmax latent map (:81-90), no Dirichlet data, SNES newtonls with PETSc's default bt line search
(:128-143).  Mixed dof numbering: vertex n holds (u_0..3, z_0..3, psi_0..3) at 12 n + (0..11).
"""
import numpy as np

from . import _capi, fem, mesh as _mesh, quadrature
from .forms import FormNonlinearProblem, FormProblem, Integral

NUM_SPECIES = 4


def initial_condition(coords, cells, tol=1e-14):
    """u_prev of :92-125: species 0 everywhere, then species 1 / 2 / 3 interpolated on the cells that
    ``locate_entities`` finds for three rectangles (all vertices of the cell inside)."""
    x, y = coords[:, 0], coords[:, 1]
    markers = (
        (1, (0.2 - tol <= y) & (y <= 0.75 + tol) & (0.2 - tol <= x) & (x <= 0.8 + tol)),
        (2, (y <= 0.5 + tol) & (0.2 - tol <= y) & (0.2 - tol <= x) & (x <= 0.5 + tol)),
        (3, (y <= 0.5 + tol) & (0.2 <= y + tol) & (0.5 - tol <= x) & (x <= 0.8 + tol)),
    )
    u = np.zeros((coords.shape[0], NUM_SPECIES))
    u[:, 0] = 1.0
    for species, inside in markers:
        nodes = np.unique(cells[np.all(inside[cells], axis=1)])
        u[nodes, :] = 0.0
        u[nodes, species] = 1.0
    return u


def setup(N, M, tau0=1e-5, alpha_0=1.0, eps_0=1e-9, quadrature_degree=7, petsc_options=None, rule=None):
    msh = _mesh.create_rectangle(N, M, lo=(0.0, 0.0), hi=(1.0, 1.0), diagonal="crossed")
    qp, qw = rule or quadrature.make_quadrature("triangle", quadrature_degree)
    p1, _ = fem.tabulate_lagrange(1, qp)
    nv = msh.num_vertices
    dofs = (12 * msh.cells.astype(np.int64)[:, :, None] + np.arange(12)[None, None, :]).reshape(msh.num_cells, -1)
    itg = Integral(dofs, msh.cells, qw, p1)
    blocks = [12 * v + np.arange(12) for v in range(nv)]
    # params: alpha, tau, eps0, epsilon / circumradius (epsilon = 2 h, h = 2 Circumradius, :52-53)
    dev = FormProblem(_capi.FORM_MULTIPHASE, 2, 12 * nv, msh.coords, [itg], [alpha_0, tau0, eps_0, 4.0], blocks=blocks)
    opts = {  # :127-143; no snes_linesearch_type: PETSc's default "bt"
        "snes_type": "newtonls", "snes_atol": 1e-8, "snes_rtol": 1e-8, "snes_max_it": 25, "ksp_type": "preonly",
        "pc_type": "lu", "ksp_error_if_not_converged": True, "snes_error_if_not_converged": True,
    }
    opts.update(petsc_options or {})
    sol = np.zeros(12 * nv)
    return {"mesh": msh, "dev": dev, "sol": sol, "problem": FormNonlinearProblem(dev, sol, opts)}


def solve_problem(N, M, primal_degree=1, cell_type="triangle", alpha_scheme="constant", alpha_0=1.0, alpha_c=1.0,
                  alpha_max=50.0, max_iterations=20, stopping_tol=1e-5, result_dir=None, write_frequency=25,
                  tau0=1e-5, T=7e-3, petsc_options=None, verbose=False):
    """multiphase_dolfinx.py:16-238.  Returns (newton_iterations, lvpp_iterations) per time step."""
    if primal_degree != 1 or cell_type != "triangle":
        raise NotImplementedError("P1 on triangles (the reference's default)")
    s = setup(N, M, tau0, alpha_0, petsc_options=petsc_options)
    msh, dev, sol, problem = s["mesh"], s["dev"], s["sol"], s["problem"]
    nv, ns = msh.num_vertices, NUM_SPECIES
    S = sol.reshape(nv, 3, ns)
    u_prev = initial_condition(msh.coords, msh.cells)
    lvpp_old = np.zeros_like(sol)
    num_steps = int(np.ceil(T / tau0))
    newton_iterations = np.zeros(num_steps, dtype=np.int32)
    lvpp_iterations = np.zeros(num_steps, dtype=np.int32)
    X, X0 = dev.vector(), dev.vector()
    aux1 = np.zeros_like(sol)
    for j in range(1, num_steps + 1):
        psi0 = np.log(np.abs(S[:, 0, :]) + 1e-7) + 1.0  # :190-196
        S[:, 2, :] = psi0
        lvpp_old.reshape(nv, 3, ns)[:, 2, :] = psi0
        aux1.reshape(nv, 3, ns)[:, 0, :] = u_prev
        dev.set_aux(1, aux1)
        u_old = np.zeros_like(sol)  # only the u slots are compared
        i = 0
        for i in range(1, max_iterations + 1):
            alpha = alpha_0  # :200-205
            if alpha_scheme == "linear":
                alpha = min(alpha_0 + alpha_c * i, alpha_max)
            elif alpha_scheme == "doubling":
                alpha = min(alpha_0 * 2**i, alpha_max)
            dev.set_param(0, alpha)
            dev.set_aux(0, lvpp_old)
            problem.solve()
            newton_iterations[j - 1] += problem.solver.getIterationNumber()
            X.set(sol)
            X0.set(u_old)
            global_diff = np.sqrt(dev.increment_sq(X, X0))  # :210-211
            if verbose:
                print(f"Step {j} iteration {i}: converged={problem.solver.getConvergedReason()} alpha={alpha:.2e} "
                      f"num_iterations={problem.solver.getIterationNumber()} |delta u |= {global_diff}")
            u_old.reshape(nv, 3, ns)[:, 0, :] = S[:, 0, :]  # :222
            lvpp_old[:] = sol  # :223
            if global_diff < stopping_tol:
                break
        u_prev = S[:, 0, :].copy()  # :227
        lvpp_iterations[j - 1] += i
    solve_problem.last = s
    return newton_iterations, lvpp_iterations
