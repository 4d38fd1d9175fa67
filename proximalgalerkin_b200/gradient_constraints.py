"""examples/06_gradient_constraints/gradient_constraint_dolfinx.py on the GPU: ``solve_problem`` keeps
the reference's signature and return value (Newton iterations and L2 increments per proximal step).

u in P2, psi in (P1)^2 on the unit square (:36-45), quadrature degree 10 (:53), phi and f
interpolated into the primal space (:56-62), u = 0 on the boundary (:64-70,109-111), residual :100-107.
Mixed dof numbering: u at P2 node n -> n, psi at vertex v -> N2 + 2 v + component (a recorded
permutation of dolfinx's numbering).
"""
import numpy as np

from . import _capi, fem, mesh as _mesh, quadrature
from .forms import FormNonlinearProblem, FormProblem, Integral


def default_phi(x):
    return 0.1 + 0.2 * x[0] + 0.4 * x[1]  # :289-291


def default_f(x):
    return 15.0 * np.sin(np.pi * x[0]) * np.sin(np.pi * x[0])  # as coded at :296-297


def setup(N, M, phi_func=default_phi, f_func=default_f, quadrature_degree=10, alpha_0=1.0, petsc_options=None, rule=None):
    msh = _mesh.create_rectangle(N, M, lo=(0.0, 0.0), hi=(1.0, 1.0))
    rule = rule or quadrature.make_quadrature("triangle", quadrature_degree)
    qp, qw = rule
    V2 = fem.FunctionSpace(msh, 2, rule=rule)
    N2, N1 = V2.num_nodes, msh.num_vertices
    p2, d2 = fem.tabulate_lagrange(2, qp)
    p1, _ = fem.tabulate_lagrange(1, qp)
    cu = V2.cell_nodes.astype(np.int64)
    cp = (N2 + 2 * msh.cells.astype(np.int64)[:, :, None] + np.arange(2)[None, None, :]).reshape(msh.num_cells, -1)
    itg = Integral(np.concatenate([cu, cp], axis=1), msh.cells, qw, p2, d2, p1)
    n = N2 + 2 * N1
    coef0, coef1 = np.zeros(n), np.zeros(n)
    coef0[:N2] = phi_func(V2.node_coords.T)
    coef1[:N2] = f_func(V2.node_coords.T)
    # Jacobi blocks: (u, psi_x, psi_y) at a vertex, u alone at an edge node
    blocks = [np.array([v, N2 + 2 * v, N2 + 2 * v + 1]) for v in range(N1)] + [np.array([e]) for e in range(N1, N2)]
    dev = FormProblem(_capi.FORM_GRADIENT, 2, n, msh.coords, [itg], [alpha_0], bc_dofs=np.sort(V2.boundary_nodes.astype(np.int64)),
                      coef0=coef0, coef1=coef1, blocks=blocks)
    opts = {  # :118-131
        "snes_type": "newtonls", "ksp_type": "preonly", "pc_type": "lu", "snes_atol": 1e-9, "snes_rtol": 1e-9,
        "snes_stol": 1e-9, "snes_max_it": 20, "snes_error_if_not_converged": True, "snes_linesearch_type": "none",
    }
    opts.update(petsc_options or {})
    sol = np.zeros(n)
    return {"mesh": msh, "V2": V2, "dev": dev, "sol": sol, "problem": FormNonlinearProblem(dev, sol, opts), "N2": N2}


def solve_problem(N, M, primal_space="Lagrange", primal_degree=2, cell_type="triangle", alpha_scheme="doubling",
                  alpha_0=1.0, alpha_c=1.0, max_iterations=25, stopping_tol=1e-8, result_dir=None,
                  phi_func=default_phi, f_func=default_f, warm_start=False, petsc_options=None, verbose=False):
    """gradient_constraint_dolfinx.py:18-209.  Returns (newton_iterations, L2_diff) like the reference."""
    if primal_space not in ("Lagrange", "P", "CG") or primal_degree != 2 or cell_type != "triangle":
        raise NotImplementedError("P2 primal space on triangles (the reference's default)")
    if warm_start:
        raise NotImplementedError("warm_start")
    s = setup(N, M, phi_func, f_func, alpha_0=alpha_0, petsc_options=petsc_options)
    dev, sol, problem = s["dev"], s["sol"], s["problem"]
    w0 = np.zeros_like(sol)
    X, X0 = dev.vector(), dev.vector()
    newton_iterations = np.zeros(max_iterations, dtype=np.int32)
    L2_diff = np.zeros(max_iterations)
    i = 0
    for i in range(max_iterations):
        alpha = alpha_0  # :172-177
        if alpha_scheme == "linear":
            alpha = alpha_0 + alpha_c * i
        elif alpha_scheme == "doubling":
            alpha = alpha_0 * 2**i
        dev.set_param(0, alpha)
        dev.set_aux(0, w0)
        problem.solve()
        newton_iterations[i] = problem.solver.getIterationNumber()
        X.set(sol)
        X0.set(w0)
        L2_diff[i] = np.sqrt(dev.increment_sq(X, X0))  # :184-185
        if verbose:
            print(f"Iteration {i + 1}: converged={problem.solver.getConvergedReason()} num_newton_iterations={newton_iterations[i]}"
                  f" |delta u |= {L2_diff[i]}")
        if L2_diff[i] < stopping_tol:
            break
        w0[:] = sol  # :205
    s["history"] = {"alpha_last": alpha}
    solve_problem.last = s
    return newton_iterations[: i + 1], L2_diff[: i + 1]
