"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): sparsity pattern bit-exact; assembled matrix and residual
entries within 1e-12 relative (measured against the largest entry of the array: entries that are
exact zeros in exact arithmetic carry rounding noise of that size in the oracle's quadrature sums);
identical Newton / proximal iteration counts; final u within 1e-10 relative L2.  The comparator is the
restatement in oracle/ (parity unpinned by the reference itself, see oracle/__init__.py).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu

TOL_ASSEMBLY = 1e-12
TOL_U = 1e-10


def _pair(kind, n, degree=1, f=0.0, obstacle="array"):
    """(device problem + objects, oracle) on the same mesh."""
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    if kind == "tri":
        msh, om = lvpp.mesh.create_rectangle(n, n), omesh.rectangle(n, n)
    else:
        msh, om = lvpp.mesh.create_box(n, n, n), omesh.box_kuhn(n, n, n)
    s = lvpp.obstacle_pg.setup(msh, degree, obstacle=lvpp.fem.phi_set if obstacle == "array" else "phi_set", f_value=f)
    orc = oobs.ObstacleOracle(om, degree=degree, f=f)
    return lvpp, s, s["problem"].device_problem, orc


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("kind,n,degree", [("tri", 9, 1), ("tet", 5, 1), ("tri", 6, 2), ("tet", 3, 2)])
def test_pattern_bit_exact(lib, kind, n, degree):
    lvpp, s, dev, orc = _pair(kind, n, degree)
    indptr, indices = dev.csr_pattern()
    assert indptr.dtype == np.int64 and indices.dtype == np.int32
    assert np.array_equal(indptr, orc.indptr)
    assert np.array_equal(indices, orc.indices)
    assert dev.stats()["nnz"] == orc.nnz


@pytest.mark.parametrize("kind,n,degree,f,obstacle", [
    ("tri", 12, 1, 0.0, "array"), ("tet", 6, 1, 0.0, "array"), ("tet", 6, 1, 0.7, "builtin"),
    ("tri", 12, 1, -1.3, "builtin"), ("tri", 7, 2, 0.4, "array"), ("tet", 3, 2, 0.0, "builtin")])
def test_residual_and_jacobian(lib, kind, n, degree, f, obstacle):
    lvpp, s, dev, orc = _pair(kind, n, degree, f, obstacle)
    rng = np.random.default_rng(42)
    for trial, alpha in enumerate((1.0, 37.5)):
        x = 0.5 * rng.standard_normal(orc.num_rows)  # violates the Dirichlet data: exercises lifting
        xk = 0.5 * rng.standard_normal(orc.num_rows)
        dev.set_alpha(alpha)
        dev.set_previous(xk)
        X, F = lvpp.DeviceVector(dev.n, dev.device), lvpp.DeviceVector(dev.n, dev.device)
        X.set(x)
        fnorm = dev.assemble_residual(X, F)
        Fo = orc.assemble_residual(x, xk, alpha)
        assert _rel(F.numpy(), Fo) < TOL_ASSEMBLY
        assert abs(fnorm - np.linalg.norm(Fo)) < 1e-12 * np.linalg.norm(Fo)
        vo = orc.assemble_jacobian_values(x, alpha)
        v = dev.jacobian_values().cpu().numpy()
        assert _rel(v, vo) < TOL_ASSEMBLY
        # structural identities
        J = sp.csr_matrix((v, orc.indices, orc.indptr), shape=(orc.num_rows,) * 2)
        assert abs(J - J.T).max() < 1e-13 * abs(J).max()
        bc = orc.bc_dofs
        assert np.all(J[bc].toarray() == np.eye(orc.num_rows)[bc])
        assert np.array_equal(F.numpy()[bc], x[bc] - orc.bc_values[bc])


def test_sm_jacobian_only_entry(lib):
    """SNESProblem.J without a preceding residual evaluation."""
    lvpp, s, dev, orc = _pair("tet", 4)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(orc.num_rows)
    X = lvpp.DeviceVector(dev.n, dev.device)
    X.set(x)
    s["alpha"].value = 3.0
    prob = s["problem"]._problem
    A = prob.create_matrix()
    prob.J(None, X, A, A)
    assert _rel(A.values().cpu().numpy(), orc.assemble_jacobian_values(x, 3.0)) < TOL_ASSEMBLY


@pytest.mark.parametrize("kind,n,degree", [("tri", 16, 1), ("tet", 6, 1), ("tri", 8, 2)])
def test_spmv_and_linear_solve(lib, kind, n, degree):
    lvpp, s, dev, orc = _pair(kind, n, degree)
    rng = np.random.default_rng(7)
    x = 0.2 * rng.standard_normal(orc.num_rows)
    x[orc.bc_dofs] = 0.0
    alpha = 2.5
    dev.set_alpha(alpha)
    dev.set_previous(np.zeros(orc.num_rows))
    X = lvpp.DeviceVector(dev.n, dev.device)
    X.set(x)
    dev.assemble_jacobian(X)
    J = orc.jacobian(x, alpha)
    v = rng.standard_normal(orc.num_rows)
    Vv, Y = lvpp.DeviceVector(dev.n, dev.device), lvpp.DeviceVector(dev.n, dev.device)
    Vv.set(v)
    dev.spmv(Vv, Y)
    assert _rel(Y.numpy(), J @ v) < 1e-13
    rhs = rng.standard_normal(orc.num_rows)
    R = lvpp.DeviceVector(dev.n, dev.device)
    R.set(rhs)
    opts = lvpp.newton_options({"ksp_rtol": 1e-13})
    its, reason, rnorm = dev.linear_solve(R, Y, opts)
    assert reason > 0 and its > 0
    ye = spla.splu(J.tocsc()).solve(rhs)
    assert np.linalg.norm(Y.numpy() - ye) / np.linalg.norm(ye) < 1e-9
    # bit-reproducible (deterministic reductions, no atomics)
    Y2 = lvpp.DeviceVector(dev.n, dev.device)
    its2, _, _ = dev.linear_solve(R, Y2, opts)
    assert its2 == its and np.array_equal(Y.numpy(), Y2.numpy())


@pytest.mark.parametrize("kind,n", [("tri", 24), ("tet", 8)])
def test_newton_solve_matches_lu_newton(lib, kind, n):
    from oracle import snes as osnes

    lvpp, s, dev, orc = _pair(kind, n)
    x0 = np.zeros(orc.num_rows)
    xo, reason_o, its_o, hist = osnes.newton_ls_none(
        lambda z: orc.assemble_residual(z, x0, 1.0), lambda z: orc.jacobian(z, 1.0), x0, rtol=1e-6, max_it=100)
    dev.set_alpha(1.0)
    dev.set_previous(x0)
    X = lvpp.DeviceVector(dev.n, dev.device)
    reason, its, fnorm, lin = dev.newton_solve(X, lvpp.newton_options(s["options"]))
    assert (reason, its) == (reason_o, its_o)
    assert abs(fnorm - hist[-1]) < 1e-6 * hist[0]
    u, uo = X.numpy()[0::2], xo[0::2]
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) < TOL_U


@pytest.mark.parametrize("kind,n,degree", [("tri", 20, 1), ("tet", 7, 1), ("tri", 8, 2)])
def test_full_lvpp_solve_matches_oracle(lib, kind, n, degree):
    """The CI configuration of the reference (compare_all.py:80-87): double-exponential alpha,
    alpha_max 1e2, tol 1e-4.  Same Newton count in every proximal step, same number of proximal
    steps, final u within 1e-10 relative (discrete and mass-weighted L2)."""
    import proximalgalerkin_b200 as lvpp
    from oracle import lvpp_driver

    lvpp_, s, dev, orc = _pair(kind, n, degree)
    xo, ho = lvpp_driver.solve_obstacle(orc, max_outer=500, alpha_scheme="double_exponential", alpha_max=1e2, tol_exit=1e-4)
    msh = s["V"].mesh
    sol, total, h = lvpp.obstacle_pg.solve_problem(msh, degree, 500, "double_exponential", 1e2, 1e-4, obstacle=lvpp.fem.phi_set)
    assert h["newton_steps"] == ho["newton_steps"]
    assert h["reason"] == ho["reason"]
    assert np.allclose(h["alpha"], ho["alpha"], rtol=0, atol=0)
    u, uo = sol.x.array[0::2], xo[0::2]
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) < TOL_U
    M = orc.jacobian(xo, 1.0)[1::2, 0::2]  # mass block (Dirichlet columns zeroed; u vanishes there)
    e = u - uo
    assert np.sqrt(e @ (M @ e)) / np.sqrt(uo @ (M @ uo)) < TOL_U
    for key in ("energy", "complementarity", "feasibility", "dual_feasibility", "primal_increment", "latent_increment"):
        assert np.allclose(h[key], ho[key], rtol=1e-7, atol=1e-12), key


def test_stepper_matches_driver(lib):
    """Device-resident one-Newton-step-at-a-time loop == host-buffer NonlinearProblem loop."""
    import proximalgalerkin_b200 as lvpp

    msh = lvpp.mesh.create_box(6, 6, 6)
    sol, total, h = lvpp.obstacle_pg.solve_problem(msh, 1, 500, "double_exponential", 1e2, 1e-4)
    st = lvpp.obstacle_pg.LvppStepper(msh, 1, "double_exponential", 1e2, 1e-4)
    while st.step():
        pass
    assert st.history["newton_steps"] == h["newton_steps"]
    assert st.total_newton == total
    assert np.array_equal(st.x.numpy(), sol.x.array)


def test_observables(lib):
    lvpp, s, dev, orc = _pair("tet", 5)
    rng = np.random.default_rng(3)
    x, xk = 0.4 * rng.standard_normal(orc.num_rows), 0.4 * rng.standard_normal(orc.num_rows)
    dev.set_alpha(2.0)
    dev.set_previous(xk)
    X = lvpp.DeviceVector(dev.n, dev.device)
    X.set(x)
    obs = dev.observables(X)
    assert np.allclose(obs, orc.observables(x, xk, 2.0), rtol=1e-12, atol=1e-14)


def test_solver_api_semantics(lib):
    """SNESSolver.solve returns (reason, its) and overwrites u only on convergence
    (src/lvpp/problem.py:120-124); error flags raise (obstacle_pg.py:132,135)."""
    import proximalgalerkin_b200 as lvpp

    msh = lvpp.mesh.create_rectangle(10, 10)
    s = lvpp.obstacle_pg.setup(msh, 1, petsc_options={"snes_max_it": 1, "snes_error_if_not_converged": False})
    prob = s["problem"]._problem
    solver = lvpp.SNESSolver(prob, {"snes_max_it": 1, "snes_rtol": 1e-6, "snes_linesearch_type": "none"})
    before = s["sol"].x.array.copy()
    reason, its = solver.solve()
    assert reason == -5 and its == 1
    assert np.array_equal(s["sol"].x.array, before)
    solver2 = lvpp.SNESSolver(prob, {"snes_max_it": 100, "snes_rtol": 1e-6, "snes_linesearch_type": "none"})
    reason, its = solver2.solve()
    assert reason == 3 and its > 1 and not np.array_equal(s["sol"].x.array, before)
    s2 = lvpp.obstacle_pg.setup(msh, 1, petsc_options={"snes_max_it": 1})
    with pytest.raises(RuntimeError):
        s2["problem"].solve()
    with pytest.raises(NotImplementedError):
        lvpp.newton_options({"snes_linesearch_type": "cp"})  # none, bt, l2 exist; PETSc's critical-point search does not


def test_error_codes(lib):
    import ctypes as C

    from proximalgalerkin_b200 import _capi

    d = _capi.ObstacleDesc()
    h = _capi.H()
    assert lib.lvpp_create(C.byref(d), C.byref(h)) == _capi.E_INVALID
    assert b"unsupported element" in lib.lvpp_last_error()
    assert lib.lvpp_set_alpha(None, 1.0) == _capi.E_INVALID


def test_periodic_obstacle_matches_oracle(lib):
    """The weak-scaling workload tiles the reference's obstacle along the last axis (one copy per GPU slab);
    the device closed form with ``obstacle_period`` must equal the oracle fed with the wrapped function."""
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    n, slabs = 4, 3
    lo, hi = (-1.0, -1.0, -float(slabs)), (1.0, 1.0, float(slabs))
    msh = lvpp.mesh.create_box(n, n, n * slabs, lo=lo, hi=hi)
    s = lvpp.obstacle_pg.setup(msh, 1, obstacle="phi_set", obstacle_period=2.0, obstacle_origin=-float(slabs))
    dev = s["problem"].device_problem

    def wrapped(x):
        x = np.array(x, dtype=np.float64)
        x[2] = np.mod(x[2] + slabs, 2.0) - 1.0
        return oobs.phi_set(x)

    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n * slabs, lo=lo, hi=hi), phi=wrapped)
    rng = np.random.default_rng(8)
    x, xk = 0.4 * rng.standard_normal(orc.num_rows), 0.4 * rng.standard_normal(orc.num_rows)
    dev.set_alpha(1.9)
    dev.set_previous(xk)
    X, F = lvpp.DeviceVector(dev.n, dev.device), lvpp.DeviceVector(dev.n, dev.device)
    X.set(x)
    dev.assemble_residual(X, F)
    assert _rel(F.numpy(), orc.assemble_residual(x, xk, 1.9)) < TOL_ASSEMBLY
    # every slab sees the same obstacle: the load vector is periodic, unlike with a single obstacle
    s1 = lvpp.obstacle_pg.setup(msh, 1, obstacle="phi_set")
    d1 = s1["problem"].device_problem
    d1.set_alpha(1.9)
    d1.set_previous(xk)
    F1 = lvpp.DeviceVector(d1.n, d1.device)
    d1.assemble_residual(X, F1)
    assert _rel(F1.numpy(), F.numpy()) > 1e-3


def test_dirichlet_values_and_forcing_are_read_at_solve_time(lib):
    """A dolfinx assembly reads Constants and the DirichletBC's Function at call time: mutating ``f.value`` or the bc
    values between solves changes the next residual (lvpp_set_forcing / lvpp_set_bc_values behind
    DeviceProblem.sync_coefficients), exactly like the oracle rebuilt with the new data."""
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    n = 6
    msh = lvpp.mesh.create_box(n, n, n)
    V = lvpp.fem.functionspace(msh, ("Lagrange", 1))
    alpha, f = lvpp.fem.Constant(msh, 1.3), lvpp.fem.Constant(msh, 0.0)
    g = lvpp.fem.Constant(msh, 0.0)
    dofs = lvpp.fem.locate_dofs_boundary(V.sub(0))
    bc = lvpp.fem.dirichletbc(value=g, dofs=dofs, V=V.sub(0))
    sol, sol_k = lvpp.fem.Function(V), lvpp.fem.Function(V)
    phi = lvpp.fem.QuadratureFunction(V, name="phi")
    phi.interpolate_phi_set(None, 0.0)
    F = lvpp.obstacle_residual(sol, sol_k, alpha, f, phi)
    prob = lvpp.SNESProblem(F, sol, bcs=[bc])
    dev = prob.device_problem
    rng = np.random.default_rng(9)
    x = 0.2 * rng.standard_normal(dev.n)
    sol_k.x.array[:] = 0.1 * rng.standard_normal(dev.n)
    X, R = lvpp.DeviceVector(dev.n, dev.device), lvpp.DeviceVector(dev.n, dev.device)
    X.set(x)
    for fval, gval in ((0.0, 0.0), (2.5, 0.0), (2.5, -0.3)):
        f.value, g.value = fval, gval
        prob.F(None, X, R)
        orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n), f=fval)
        orc.bc_values[orc.bc_dofs] = gval
        Fo = orc.assemble_residual(x, sol_k.x.array, alpha.value)
        assert np.abs(R.numpy() - Fo).max() <= 1e-12 * np.abs(Fo).max(), (fval, gval)
