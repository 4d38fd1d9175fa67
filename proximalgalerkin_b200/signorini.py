"""examples/02_signorini/signorini_dolfinx.py on the GPU: ``solve_contact_problem`` keeps the
reference's call shape and return value ``(it, iterations)``.

Linear elasticity in (P1)^3 on tetrahedra with the latent contact variable psi in P1 on the submesh
of the contact facets (:199-225), residual :244-249, Dirichlet data u = (0, 0, disp) on the
displacement boundary (:256-267), SNES newtonls / line search none with tolerances set per proximal
step (:272-281,331-332).  Mixed dof numbering: u at vertex n -> 3 n + c; psi at contact vertex s
(ascending mesh vertex number) -> 3 N + s.
"""
import numpy as np

from . import _capi, fem, mesh as _mesh, quadrature
from .forms import BlockedForm, BlockFunction, FormProblem, Integral
from .problem import NonlinearProblem


def exterior_facets(cells):
    nv = cells.shape[1]
    f = np.concatenate([np.delete(cells, i, axis=1) for i in range(nv)], axis=0)
    f = np.sort(f, axis=1)
    uniq, counts = np.unique(f, axis=0, return_counts=True)
    return uniq[counts == 1]


def setup(msh, E=2.0e4, nu=0.3, gap=0.0, disp=-0.25, quadrature_degree=4, alpha_0=1.0, contact_coord=None,
          disp_coord=None, petsc_options=None, tol=1e-12):
    if msh.cell_name != "tetrahedron":
        raise NotImplementedError("tetrahedral meshes")
    N = msh.num_vertices
    last = msh.coords[:, 2]
    contact_coord = last.min() if contact_coord is None else contact_coord
    disp_coord = last.max() if disp_coord is None else disp_coord
    ef = exterior_facets(msh.cells.astype(np.int64))
    contact = ef[np.all(np.abs(last[ef] - contact_coord) < tol, axis=1)]  # facet_tag.find(contact), :199-203
    bc_vertices = np.unique(ef[np.all(np.abs(last[ef] - disp_coord) < tol, axis=1)])
    sub_vertices = np.unique(contact)  # create_submesh, :207
    sub_of = np.full(N, -1, dtype=np.int64)
    sub_of[sub_vertices] = np.arange(sub_vertices.size)
    dof_u = 3 * np.arange(N, dtype=np.int64)[:, None] + np.arange(3)[None, :]
    cd = dof_u[msh.cells].reshape(msh.num_cells, -1)
    fq, fw = quadrature.make_quadrature("triangle", quadrature_degree)
    fphi, _ = fem.tabulate_lagrange(1, fq)
    fd = np.concatenate([dof_u[contact].reshape(contact.shape[0], -1), 3 * N + sub_of[contact]], axis=1)
    mu = E / (2.0 * (1.0 + nu))  # :234-235
    lmbda = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    n = 3 * N + sub_vertices.size
    blocks = []
    for v in range(N):
        b = [3 * v, 3 * v + 1, 3 * v + 2]
        if sub_of[v] >= 0:
            b.append(3 * N + int(sub_of[v]))
        blocks.append(np.array(b))
    bc_dofs = dof_u[bc_vertices].ravel()
    bc_vals = np.tile(np.array([0.0, 0.0, disp]), bc_vertices.size)
    dev = FormProblem(_capi.FORM_SIGNORINI, 3, n, msh.coords, [Integral(cd, msh.cells), Integral(fd, contact, fw, fphi)],
                      [alpha_0, mu, lmbda, gap, 0.0, 0.0, -1.0], bc_dofs=bc_dofs, bc_values=bc_vals, blocks=blocks)
    opts = {  # :272-281
        "snes_type": "newtonls", "snes_linesearch_type": "none", "ksp_type": "preonly", "pc_type": "lu",
        "ksp_error_if_not_converged": True, "snes_error_if_not_converged": True,
    }
    opts.update(petsc_options or {})
    # unknowns and mutable inputs of the blocked form (signorini_dolfinx.py:227-232,240-252)
    u, psi, psi_k = BlockFunction(3 * N, "u"), BlockFunction(sub_vertices.size, "psi"), BlockFunction(sub_vertices.size, "psi_k")
    alpha = fem.Constant(msh, alpha_0)
    aux = np.zeros(n)

    def sync():  # what a dolfinx assembly reads at call time: alpha.value, psi_k.x.array
        dev.set_param(0, alpha.value)
        aux[3 * N:] = psi_k.x.array
        dev.set_aux(0, aux)

    F = BlockedForm(dev, [u, psi], sync)
    # signorini_dolfinx.py:283-291: NonlinearProblem(F, [u, psi], bcs=bcs, J=J, entity_maps=..., kind="mpi", ...)
    problem = NonlinearProblem(F, [u, psi], bcs=None, J=None, petsc_options=opts, petsc_options_prefix="signorini_",
                               entity_maps=None, kind="mpi")
    # "sol": the mixed host vector [u; psi] behind the two blocks (gathered before, scattered after every solve)
    return {"mesh": msh, "dev": dev, "F": F, "u": u, "psi": psi, "psi_k": psi_k, "alpha": alpha, "problem": problem,
            "sol": problem._blocked._x, "num_u": 3 * N, "bc_vertices": bc_vertices, "sub_vertices": sub_vertices}


def solve_contact_problem(mesh, facet_tag=None, boundary_conditions=None, degree=1, E=2.0e4, nu=0.3, gap=0.0,
                          disp=-0.25, newton_max_its=25, newton_tol=1e-6, max_iterations=25, alpha_scheme="doubling",
                          alpha_0=1.0, alpha_c=1.0, tol=1e-6, output=None, quadrature_degree=4, petsc_options=None,
                          verbose=False):
    """signorini_dolfinx.py:156-368.  ``facet_tag`` / ``boundary_conditions``: the contact boundary is the
    exterior boundary at the lowest last coordinate and the displacement boundary the one at the highest
    (the reference's built-in box, :373-395), unless ``boundary_conditions`` gives
    ``{"contact": z_c, "displacement": z_d}`` as coordinates.  ``newton_max_its`` is accepted and not applied,
    like the reference (:165,422)."""
    if degree != 1:
        raise NotImplementedError("degree 1")
    bcnd = boundary_conditions or {}
    s = setup(mesh, E, nu, gap, disp, quadrature_degree, alpha_0, bcnd.get("contact"), bcnd.get("displacement"), petsc_options)
    u, psi, psi_k, alpha_c_, problem = s["u"], s["psi"], s["psi_k"], s["alpha"], s["problem"]
    u_prev = np.zeros_like(u.x.array)
    iterations = []
    it = 0
    for it in range(1, max_iterations + 1):
        alpha = alpha_0  # :324-329
        if alpha_scheme == "linear":
            alpha = alpha_0 + alpha_c * it
        elif alpha_scheme == "doubling":
            alpha = alpha_0 * 2**it
        alpha_c_.value = alpha
        solver_tol = 10 * newton_tol if it < 2 else newton_tol  # :331-332
        problem.solver.setTolerances(atol=solver_tol, rtol=solver_tol)
        problem.solve()
        iterations.append(problem.solver.getIterationNumber())
        normed_diff = float(np.linalg.norm(u.x.array - u_prev))  # :337-339 (host copy of the device result)
        if verbose:
            print(f"it={it}/{max_iterations} alpha={alpha} newton={iterations[-1]} increment {normed_diff:.2e}")
        if normed_diff <= tol:
            break
        u_prev[:] = u.x.array
        psi_k.x.array[:] = psi.x.array  # :344
        if problem.solver.getConvergedReason() <= 0:
            break
    solve_contact_problem.last = s
    return it, iterations
