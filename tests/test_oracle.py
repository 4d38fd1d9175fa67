"""The CPU oracle against everything that can be pinned without the dolfinx/PETSc stack
(SURVEY.md section 8c "known answers") and against the committed golden fixtures."""
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import elements, lvpp_driver, mesh, obstacle, quadrature, snes

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_alpha_schedule_known_answers():
    """Pure arithmetic of obstacle_pg.py:175-186 with C = 1, r = 1.5, q = 1.5."""
    expect = [1.0, 1.0, 1.4900343193257237, 2.439200639104808, 5.349445965312164, 16.387223352344883,
              84.95478289516922, 935.2426972935323, 31650.405900023918]
    alpha_k, alpha, got = 1, 1.0, []
    for k in range(11):
        alpha, alpha_k = lvpp_driver.alpha_schedule("double_exponential", k, alpha_k, 1e5, alpha_current=alpha)
        got.append(alpha)
    assert np.allclose(got[:9], expect, rtol=1e-14)
    assert got[9] == 1e5 and got[10] == 1e5
    alpha_k, alpha, got = 1, 1.0, []
    for k in range(9):
        alpha, alpha_k = lvpp_driver.alpha_schedule("double_exponential", k, alpha_k, 1e2, alpha_current=alpha)
        got.append(alpha)
    assert got[7] == 100.0 and got[8] == 100.0 and np.isclose(got[6], 84.95478289516922)
    assert lvpp_driver.alpha_schedule("constant", 5, 1, 1e5)[0] == 1.0
    assert lvpp_driver.alpha_schedule("geometric", 3, 1, 1e5)[0] == 1.5**3


def test_obstacle_closed_form():
    f = obstacle.phi_set
    assert f(np.zeros((2, 1)))[0] == 0.5
    b = 0.45
    assert np.isclose(f(np.array([[b], [0.0]]))[0], 0.21794494717703367, rtol=1e-14)
    r_zero = 0.5 / 0.9
    assert abs(f(np.array([[r_zero], [0.0]]))[0]) < 1e-15
    # 3-D points use the Euclidean norm: (0.3, 0.4, 0) has r = 0.5, on the linear branch B + C r
    t = np.sqrt(0.25 - b * b)
    assert np.isclose(f(np.array([[0.3], [0.4], [0.0]]))[0], t + b * b / t - 0.5 * b / t, rtol=1e-13)
    assert np.isclose(f(np.array([[0.3], [0.4], [0.0]]))[0], f(np.array([[0.5], [0.0]]))[0], rtol=1e-15)


@pytest.mark.parametrize("cell,tdim", [("triangle", 2), ("tetrahedron", 3)])
def test_reference_element_matrices(cell, tdim):
    """Exact P1 mass matrix vol/((d+1)(d+2)) (1 + delta_ij); any rule exact to degree 2 reproduces it."""
    for deg in (2, 6):
        pts, wts = quadrature.make_quadrature(cell, deg)
        phi, dphi = elements.tabulate(1, pts)
        M = np.einsum("q,qa,qb->ab", wts, phi, phi)
        vol = 1.0 / np.prod(np.arange(1, tdim + 1))
        exact = vol / ((tdim + 1) * (tdim + 2)) * (np.ones((tdim + 1,) * 2) + np.eye(tdim + 1))
        assert np.allclose(M, exact, rtol=1e-14, atol=1e-17)
        assert np.isclose(wts.sum(), vol, rtol=1e-15)
        K = np.einsum("q,qad,qbd->ab", wts, dphi, dphi)
        g = np.vstack([-np.ones((1, tdim)), np.eye(tdim)])
        assert np.allclose(K, vol * g @ g.T, rtol=1e-14)


@pytest.mark.parametrize("cell,deg", [("triangle", 6), ("tetrahedron", 6), ("triangle", 9), ("tetrahedron", 8)])
def test_quadrature_exactness(cell, deg):
    from itertools import product
    from math import factorial

    pts, wts = quadrature.make_quadrature(cell, deg)
    tdim = pts.shape[1]
    assert np.all(wts > 0) and np.all(pts > 0) and np.all(pts.sum(axis=1) < 1)
    for e in product(range(deg + 1), repeat=tdim):
        if sum(e) > deg:
            continue
        exact = np.prod([factorial(k) for k in e]) / factorial(sum(e) + tdim)
        assert np.isclose(np.sum(wts * np.prod(pts ** np.array(e), axis=1)), exact, rtol=1e-13, atol=1e-17), e
    if deg == 6:
        assert wts.size == {"triangle": 12, "tetrahedron": 24}[cell]


@pytest.mark.parametrize("msh,degree", [(mesh.rectangle(7, 6), 1), (mesh.box_kuhn(3, 4, 3), 1), (mesh.rectangle(4, 4), 2)])
def test_structural_identities(msh, degree):
    orc = obstacle.ObstacleOracle(msh, degree=degree)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(orc.num_rows)
    J = orc.jacobian(x, 3.0)
    assert abs(J - J.T).max() < 1e-14
    bc = orc.bc_dofs
    assert np.array_equal(J[bc].toarray(), np.eye(orc.num_rows)[bc])
    assert np.array_equal(J[:, bc].toarray(), np.eye(orc.num_rows)[:, bc])
    F = orc.assemble_residual(x, 0 * x, 3.0)
    assert np.array_equal(F[bc], x[bc] - orc.bc_values[bc])
    # first residual at u = psi = 0: F_psi = -M 1 - int phi_obs phi_i
    z = np.zeros(orc.num_rows)
    F0 = orc.assemble_residual(z, z, 1.0)
    ws = orc.qwts[None, :] * orc.scale[:, None]
    load = np.einsum("cq,qa->ca", ws * (-1.0 - orc.phi_q), orc.phi_tab)
    expect = np.bincount(orc.dof_psi[orc.cell_nodes].ravel(), weights=load.ravel(), minlength=orc.num_rows)
    assert np.allclose(F0[1::2], expect[1::2], rtol=1e-13, atol=1e-16)
    # Jacobian is the derivative of the residual (finite differences on interior dofs)
    free = np.flatnonzero(~orc.is_bc)
    d = np.zeros(orc.num_rows)
    d[free] = rng.standard_normal(free.size)
    eps = 1e-6
    fd = (orc.assemble_residual(x + eps * d, z, 3.0) - orc.assemble_residual(x - eps * d, z, 3.0)) / (2 * eps)
    # lifting makes F independent of x at Dirichlet columns only through the BC rows
    assert np.allclose(fd[free], (J @ d)[free], rtol=1e-6, atol=1e-8)


def test_pattern_is_cellwise_tensor_product():
    msh = mesh.box_kuhn(2, 2, 2)
    orc = obstacle.ObstacleOracle(msh)
    A = sp.lil_matrix((orc.num_rows,) * 2)
    for cd in orc.cell_dofs:
        for i in cd:
            A[i, cd] = 1
    A = A.tocsr()
    A.sort_indices()
    assert np.array_equal(A.indptr, orc.indptr) and np.array_equal(A.indices, orc.indices)


def test_snes_reasons():
    # linear problem: converges in one step on the relative test
    A = sp.diags([2.0, 3.0, 4.0]).tocsr()
    b = np.array([1.0, 1.0, 1.0])
    x, reason, its, hist = snes.newton_ls_none(lambda z: A @ z - b, lambda z: A, np.zeros(3), rtol=1e-8)
    assert (reason, its) == (snes.CONVERGED_FNORM_RELATIVE, 1) or reason == snes.CONVERGED_FNORM_ABS
    x, reason, its, hist = snes.newton_ls_none(lambda z: z**3 - 1.0, lambda z: sp.diags(3 * z**2).tocsr(), np.full(3, 5.0), max_it=2)
    assert reason == snes.DIVERGED_MAX_IT and its == 2
    assert snes.converged_default(0, 0, 0, float("nan"), 0, 1, 1e-50, 1e-8, 1e4) == snes.DIVERGED_FNORM_NAN
    assert snes.converged_default(3, 1.0, 1e-10, 1.0, 1e-9, 1.0, 1e-50, 1e-8, 1e4) == snes.CONVERGED_SNORM_RELATIVE
    assert snes.converged_default(3, 1.0, 1.0, 1e5, 1e-9, 1.0, 1e-50, 1e-8, 1e4) == snes.DIVERGED_DTOL


@pytest.mark.parametrize("name,build,degree", [
    ("tri_p1_n12", lambda: mesh.rectangle(12, 12), 1),
    ("tet_p1_n5", lambda: mesh.box_kuhn(5, 5, 5), 1),
    ("tri_p2_n6", lambda: mesh.rectangle(6, 6), 2)])
def test_golden_fixtures(name, build, degree):
    g = np.load(GOLDEN / f"{name}.npz")
    orc = obstacle.ObstacleOracle(build(), degree=degree)
    assert np.array_equal(orc.indptr, g["indptr"]) and np.array_equal(orc.indices, g["indices"])
    alpha = float(g["alpha"])
    assert np.allclose(orc.assemble_residual(g["x"], g["xk"], alpha), g["F"], rtol=1e-13, atol=1e-15)
    assert np.allclose(orc.assemble_jacobian_values(g["x"], alpha), g["jac_values"], rtol=1e-13, atol=1e-16)
    assert np.allclose(orc.observables(g["x"], g["xk"], alpha), g["observables"], rtol=1e-13)
    xs, h = lvpp_driver.solve_obstacle(orc, max_outer=500, alpha_scheme="double_exponential", alpha_max=1e2, tol_exit=1e-4)
    assert h["newton_steps"] == g["newton_steps"].tolist()
    assert np.allclose(h["alpha"], g["alphas"], rtol=0, atol=0)
    assert np.linalg.norm(xs[0::2] - g["solution"][0::2]) <= 1e-11 * np.linalg.norm(g["solution"][0::2])
