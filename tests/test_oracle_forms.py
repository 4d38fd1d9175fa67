"""The oracle restatements of the gradient-constraint, multiphase and Signorini forms (oracle/forms.py):
hand-written Jacobians against finite differences of the residuals, structural identities, the
Dirichlet conventions, the alpha schedules of the three drivers and PETSc's bt line search on a
problem where it has to backtrack.  Parity is unpinned by the reference (no golden vectors)."""
import numpy as np
import pytest

from oracle import forms, lvpp_driver, mesh as omesh, snes


def _fd_error(orc, x):
    J = orc.jacobian(x)
    rng = np.random.default_rng(1)
    err = 0.0
    for _ in range(3):
        d = rng.standard_normal(orc.num_rows)
        d[orc.bc_dofs] = 0.0
        h = 1e-6
        fd = (orc.assemble_residual(x + h * d) - orc.assemble_residual(x - h * d)) / (2 * h)
        err = max(err, np.abs(fd - J @ d).max() / np.abs(J @ d).max())
    return err


def test_gradient_constraint_jacobian_and_structure():
    g = forms.GradientConstraintOracle(omesh.rectangle(4, 4, lo=0.0, hi=1.0))
    rng = np.random.default_rng(0)
    x = 0.3 * rng.standard_normal(g.num_rows)
    g.alpha, g.w0 = 1.7, 0.2 * rng.standard_normal(g.num_rows)
    assert _fd_error(g, x) < 1e-8
    J = g.jacobian(x).toarray()
    assert np.abs(J - J.T).max() < 1e-14 * np.abs(J).max()  # symmetric saddle point
    bc = g.bc_dofs
    assert np.all(J[bc][:, bc] == np.eye(bc.size)) and np.all(J[:, bc].sum(axis=0) == 1.0)
    assert np.allclose(g.assemble_residual(x)[bc], x[bc])  # set_bc: x - g with g = 0
    # phi and f are interpolated into P2 (gradient_constraint_dolfinx.py:56-62, 289-297)
    assert np.allclose(g.phi_h, 0.1 + 0.2 * g.node_coords[:, 0] + 0.4 * g.node_coords[:, 1])
    assert np.allclose(g.f_h, 15.0 * np.sin(np.pi * g.node_coords[:, 0]) ** 2)
    # 16 cells... P2 nodes 81, psi dofs 2 * 25
    assert g.num_rows == 81 + 50 and g.bc_nodes.size == 32


def test_multiphase_jacobian_and_structure():
    m = omesh.rectangle(3, 3, diagonal="crossed", lo=0.0, hi=1.0)
    mp = forms.MultiphaseOracle(m)
    rng = np.random.default_rng(0)
    x = 0.3 * rng.standard_normal(mp.num_rows)
    mp.alpha, mp.lvpp_old, mp.u_prev = 1.3, 0.2 * rng.standard_normal(mp.num_rows), rng.random((mp.N, 4))
    assert _fd_error(mp, x) < 1e-8
    # crossed mesh: right isosceles triangles with hypotenuse 1/3 -> circumradius 1/6, epsilon = 4 R
    assert np.allclose(mp.eps_cell, 4.0 / 6.0)
    assert mp.bc_dofs.size == 0 and mp.num_rows == 12 * 25
    # dolfinx stores the full 36 x 36 cell tensor product, structural zeros included
    assert mp.nnz == (mp.jacobian(x) != 0).sum() or mp.nnz > (mp.jacobian(x) != 0).sum()
    # softmax Jacobian rows sum to zero: the psi-psi block annihilates constant shifts up to the eps0 mass term
    J = mp.jacobian(x).toarray()
    P = np.zeros(mp.num_rows)
    P.reshape(mp.N, 3, 4)[:, 2, :] = 1.0
    r = (J @ P).reshape(mp.N, 3, 4)
    assert np.abs(r[:, 2, :]).max() < 1e-8  # only -eps0 * M * 1 is left


def test_multiphase_initial_condition():
    m = omesh.rectangle(10, 10, diagonal="crossed", lo=0.0, hi=1.0)
    u = lvpp_driver.multiphase_initial_condition(m.coords, m.cells)
    assert np.all(u.sum(axis=1) == 1.0)
    assert u[:, 1].sum() > 0 and u[:, 2].sum() > 0 and u[:, 3].sum() > 0


def test_signorini_jacobian_and_structure():
    s = forms.SignoriniOracle(omesh.box_kuhn(3, 3, 2, lo=(0, 0, 0), hi=(1, 1, 1)), disp=-0.1)
    rng = np.random.default_rng(0)
    x = 0.01 * rng.standard_normal(s.num_rows)
    s.alpha, s.psi_k = 0.7, 0.2 * rng.standard_normal(s.NS)
    assert _fd_error(s, x) < 1e-8
    assert s.NS == 16 and s.contact_facets.shape[0] == 18 and s.bc_dofs.size == 3 * 16
    # Lame parameters of signorini_dolfinx.py:234-235 with E = 2e4, nu = 0.3
    assert abs(s.mu - 2.0e4 / 2.6) < 1e-9 and abs(s.lmbda - 2.0e4 * 0.3 / (1.3 * 0.4)) < 1e-9
    # rigid translations are in the kernel of the elasticity block
    A = s._elasticity()
    t = np.tile(np.array([1.0, -2.0, 0.5]), 4)
    assert np.abs(A @ t).max() < 1e-9 * np.abs(A).max()
    # set_bc rows: x - g with g = (0, 0, disp)
    F = s.assemble_residual(x)
    assert np.allclose(F[s.bc_dofs], x[s.bc_dofs] - s.bc_values[s.bc_dofs])
    assert np.allclose(s.bc_values[s.bc_dofs].reshape(-1, 3), [0.0, 0.0, -0.1])


def test_driver_schedules_and_counts():
    g = forms.GradientConstraintOracle(omesh.rectangle(6, 6, lo=0.0, hi=1.0))
    _, h = lvpp_driver.solve_gradient_constraint(g, max_iterations=4)
    assert h["alpha"] == [1.0, 2.0, 4.0, 8.0]  # alpha_0 * 2**i, i from 0 (:177)
    s = forms.SignoriniOracle(omesh.box_kuhn(3, 3, 3, lo=(0, 0, 0), hi=(1, 1, 1)), disp=-0.1)
    x, h = lvpp_driver.solve_signorini(s, alpha_0=0.005)
    assert s.alpha == 0.005 * 2 ** h["it"]  # it from 1 (:329)
    nu = 3 * s.N
    # no penetration beyond quadrature accuracy: u.n_g <= g on the contact boundary nodes (g = z - gap = 0)
    assert np.all(-x[2:nu:3][s.sub_vertices] <= 1e-6)


def test_bt_linesearch_backtracks():
    # scalar problem where the full Newton step increases the residual: atan(x) from x0 = 2
    import scipy.sparse as sp

    F = lambda x: np.arctan(x)
    J = lambda x: sp.csr_matrix(np.diag(1.0 / (1.0 + x * x)))
    x, reason, its, hist = snes.newton_ls(F, J, np.array([2.0]), "bt", rtol=1e-12, atol=1e-12, max_it=50)
    assert reason > 0 and abs(x[0]) < 1e-10
    assert all(b < a for a, b in zip(hist, hist[1:]))  # monotone decrease (Armijo on 0.5 |F|^2)
    with np.errstate(all="ignore"):
        try:  # full steps diverge from x0 = 2 (|x| grows until the Jacobian underflows to a singular matrix)
            xb, reason_b, _, _ = snes.newton_ls(F, J, np.array([2.0]), "none", rtol=1e-12, atol=1e-12, max_it=50)
            assert reason_b < 0 or not np.isfinite(xb[0]) or abs(xb[0]) > 1.0
        except RuntimeError:
            pass
