#!/usr/bin/env python
"""Golden files from the real reference stack ("Consequence").  NOT RUNNABLE IN THIS IMAGE: it needs
dolfinx >= 0.10, basix, ufl, petsc4py, mpi4py and the reference's own ``lvpp`` package (`pip install` of the
reference repository).  Written against the API calls the reference itself makes
(examples/01_obstacle_problem/obstacle_pg.py:68-142, src/lvpp/problem.py:54-77,105-124); untested until somebody
runs it on such a machine -- parity stays "unpinned" (oracle/__init__.py) until then.

  python tools/export_from_dolfinx.py --dim 2 --size 12 --out tests/golden/dolfinx_tri_p1_n12.npz

What it writes (one .npz, serial run), all in dolfinx's own numbering -- tests/test_dolfinx_golden.py does the
renumbering through the exported dofmaps, which is the point of feeding meshes and dofmaps as plain arrays:

  coords [nv, gdim], cells [nc, nv_per_cell]      mesh.geometry.x, mesh.geometry.dofmap (P1 geometry: = vertices)
  dof_u [nv], dof_psi [nv]                        mixed-space dof of u / psi at every geometry node
  qpts, qwts                                      basix.make_quadrature(cell, 6): the rule FFCx integrates with
  bc_dofs                                         the dofs of the DirichletBC (obstacle_pg.py:76-83)
  alpha, x, xk                                    the state the forms are evaluated at (nodal values of smooth fields)
  F [rows]                                        lvpp.SNESProblem.F(None, x, F)            (problem.py:54-67)
  J_indptr, J_indices, J_data                     lvpp.SNESProblem.J(None, x, A, A), A.getValuesCSR()  (:69-77)
  newton_steps, alphas, increments, u_final       the reference's own outer loop (obstacle_pg.py:154-227),
                                                  double-exponential alpha, alpha_max 1e2, tol_exit 1e-4 (compare_all.py:80-87)
"""
import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--size", type=int, default=12)
    ap.add_argument("--alpha", type=float, default=1.7)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    from mpi4py import MPI
    from petsc4py import PETSc

    import basix
    import basix.ufl
    import dolfinx
    import dolfinx.fem.petsc
    import ufl
    from dolfinx import default_scalar_type, fem, mesh
    from lvpp import SNESProblem

    comm = MPI.COMM_SELF
    n = args.size
    if args.dim == 2:
        msh = mesh.create_rectangle(comm, [np.array([-1.0, -1.0]), np.array([1.0, 1.0])], [n, n],
                                    cell_type=mesh.CellType.triangle, diagonal=mesh.DiagonalType.right)
    else:
        msh = mesh.create_box(comm, [np.array([-1.0, -1.0, -1.0]), np.array([1.0, 1.0, 1.0])], [n, n, n],
                              cell_type=mesh.CellType.tetrahedron)
    tdim = msh.topology.dim
    # spaces, BC, obstacle, forms: obstacle_pg.py:68-125
    P = basix.ufl.element("Lagrange", msh.basix_cell(), 1)
    V = fem.functionspace(msh, basix.ufl.mixed_element([P, P]))
    alpha = fem.Constant(msh, default_scalar_type(args.alpha))
    f = fem.Constant(msh, 0.0)
    msh.topology.create_connectivity(tdim - 1, tdim)
    facets = mesh.exterior_facet_indices(msh.topology)
    V0, _ = V.sub(0).collapse()
    bdofs = fem.locate_dofs_topological((V.sub(0), V0), entity_dim=tdim - 1, entities=facets)
    u_bc = fem.Function(V0)
    u_bc.x.array[:] = 0.0
    bcs = fem.dirichletbc(value=u_bc, dofs=bdofs, V=V.sub(0))
    sol, sol_k = fem.Function(V), fem.Function(V)
    u, psi = ufl.split(sol)
    _, psi_k = ufl.split(sol_k)

    def phi_set(x):  # obstacle_pg.py:92-104 with r over all gdim coordinates
        r = np.sqrt(sum(x[d] ** 2 for d in range(tdim)))
        r0, beta = 0.5, 0.9
        b = r0 * beta
        t = np.sqrt(r0**2 - b**2)
        return np.where(r > b, (t + b * b / t) - (b / t) * r, np.sqrt(np.maximum(r0**2 - r**2, 0.0)))

    qdeg = 6
    Vq = fem.functionspace(msh, basix.ufl.quadrature_element(msh.topology.cell_name(), degree=qdeg))
    phi = fem.Function(Vq)
    phi.interpolate(phi_set)
    v, w = ufl.TestFunctions(V)
    dx = ufl.Measure("dx", domain=msh, metadata={"quadrature_degree": qdeg})
    F = (alpha * ufl.inner(ufl.grad(u), ufl.grad(v)) * dx + psi * v * dx + u * w * dx - ufl.exp(psi) * w * dx
         - phi * w * dx - alpha * f * v * dx - psi_k * v * dx)
    J = ufl.derivative(F, sol)

    # dof <-> geometry node (P1 Lagrange and the affine coordinate element share the reference vertex order)
    gd = msh.geometry.dofmap
    nv = msh.geometry.x.shape[0]
    dof_u = np.full(nv, -1, dtype=np.int64)
    dof_psi = np.full(nv, -1, dtype=np.int64)
    dm_u, dm_p = V.sub(0).dofmap, V.sub(1).dofmap
    for c in range(gd.shape[0]):
        dof_u[gd[c]] = dm_u.cell_dofs(c)
        dof_psi[gd[c]] = dm_p.cell_dofs(c)
    assert dof_u.min() >= 0 and dof_psi.min() >= 0
    X = msh.geometry.x[:, :tdim]

    # the state: smooth nodal fields, so that the other side can evaluate them at its own nodes
    def u_field(x):
        return 0.3 * np.prod(np.sin(np.pi * x), axis=1)

    def psi_field(x):
        return 0.4 * np.cos(2.0 * x[:, 0]) - 0.2 * x[:, -1]

    x_state = np.zeros(V.dofmap.index_map.size_local * V.dofmap.index_map_bs)
    xk_state = np.zeros_like(x_state)
    x_state[dof_u], x_state[dof_psi] = u_field(X), psi_field(X)
    xk_state[dof_u], xk_state[dof_psi] = 0.5 * u_field(X), psi_field(X) - 0.3
    sol_k.x.array[:] = xk_state

    # residual and Jacobian through the reference's own callbacks (src/lvpp/problem.py:54-77)
    prob = SNESProblem(F, sol, J=fem.form(J), bcs=[bcs])
    A = dolfinx.fem.petsc.create_matrix(prob.a)
    bfun, xfun = fem.Function(V), fem.Function(V)
    xfun.x.array[:] = x_state
    prob.F(None, xfun.x.petsc_vec, bfun.x.petsc_vec)
    prob.J(None, xfun.x.petsc_vec, A, A)
    indptr, indices, data = A.getValuesCSR()

    # the reference's outer loop (obstacle_pg.py:154-227) with its own NonlinearProblem and options (:128-142)
    opts = {"ksp_type": "preonly", "pc_type": "lu", "pc_factor_mat_solver_type": "mumps", "ksp_error_if_not_converged": True,
            "snes_error_if_not_converged": True, "snes_linesearch_type": "none", "snes_rtol": 1e-6, "snes_max_it": 100}
    nlp = dolfinx.fem.petsc.NonlinearProblem(F, u=sol, bcs=[bcs], J=J, petsc_options=opts, petsc_options_prefix="obstacle_")
    u_k, _ = ufl.split(sol_k)
    h1 = fem.form(ufl.inner(ufl.grad(u - u_k), ufl.grad(u - u_k)) * dx + (u - u_k) ** 2 * dx)
    sol.x.array[:] = 0.0
    sol_k.x.array[:] = 0.0
    C, r, q, alpha_k, alpha_max = 1.0, 1.5, 1.5, 1, 1e2
    newton, alphas, incs = [], [], []
    for k in range(500):
        try:
            alpha.value = max(C * r ** (q**k) - alpha_k, C)
        except OverflowError:
            pass
        alpha_k = alpha.value
        alpha.value = min(alpha.value, alpha_max)
        nlp.solve()
        newton.append(nlp.solver.getIterationNumber())
        alphas.append(float(alpha.value))
        incs.append(float(np.sqrt(fem.assemble_scalar(h1))))
        if incs[-1] < 1e-4:
            break
        sol_k.x.array[:] = sol.x.array[:]

    pts, wts = basix.make_quadrature(msh.basix_cell(), qdeg)
    np.savez_compressed(
        args.out, coords=X, cells=np.asarray(gd, dtype=np.int32), dof_u=dof_u, dof_psi=dof_psi, qpts=pts, qwts=wts,
        bc_dofs=np.sort(np.asarray(bdofs[0], dtype=np.int64)), alpha=args.alpha, x=x_state, xk=xk_state,
        F=bfun.x.array.copy(), J_indptr=indptr, J_indices=indices, J_data=data, newton_steps=np.array(newton),
        alphas=np.array(alphas), increments=np.array(incs), u_final=sol.x.array[dof_u].copy(),
        versions=np.array([dolfinx.__version__, basix.__version__, ufl.__version__, PETSc.Sys.getVersion().__repr__()]))
    print(f"wrote {args.out}: {x_state.size} rows, {data.size} nnz, Newton steps {newton}")


if __name__ == "__main__":
    main()
