"""The multigrid-preconditioned GMRES path (pc_type mg) against the oracle: linear solves against
sparse LU, and full LVPP solves with identical Newton / proximal iteration counts."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu
MG = {"ksp_type": "gmres", "pc_type": "mg"}


def _pair(kind, n, degree=1):
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    if kind == "tri":
        msh, om = lvpp.mesh.create_rectangle(n, n), omesh.rectangle(n, n)
    else:
        msh, om = lvpp.mesh.create_box(n, n, n), omesh.box_kuhn(n, n, n)
    s = lvpp.obstacle_pg.setup(msh, degree, obstacle="phi_set", petsc_options=MG)
    return lvpp, s, s["problem"].device_problem, oobs.ObstacleOracle(om, degree=degree)


@pytest.mark.parametrize("kind,n,degree", [("tri", 3, 1), ("tri", 40, 1), ("tet", 12, 1), ("tri", 12, 2), ("tet", 4, 2)])
def test_mg_gmres_linear_solve(lib, kind, n, degree):
    lvpp, s, dev, orc = _pair(kind, n, degree)
    rng = np.random.default_rng(11)
    x = 0.2 * rng.standard_normal(orc.num_rows)
    x[1::2] -= 3.0 * (rng.random(orc.num_rows // 2) > 0.6)  # strongly varying exp(psi)
    x[orc.bc_dofs] = 0.0
    alpha = 7.5
    dev.set_alpha(alpha)
    dev.set_previous(np.zeros(orc.num_rows))
    X, R, Y = (lvpp.DeviceVector(dev.n, dev.device) for _ in range(3))
    X.set(x)
    dev.assemble_jacobian(X)
    J = orc.jacobian(x, alpha)
    rhs = rng.standard_normal(orc.num_rows)
    R.set(rhs)
    opts = lvpp.newton_options(dict(MG, ksp_rtol=1e-12))
    its, reason, rnorm = dev.linear_solve(R, Y, opts)
    assert reason > 0, (its, reason, rnorm)
    y = Y.numpy()
    # the stopping test runs in the equilibrated norm (multigrid.cu: the psi rows are weighed by
    # w = alpha sum K_ii / sum M_ii over the free nodes), so that is the norm the promise ksp_rtol is checked in
    free = np.ones(orc.num_rows // 2, dtype=bool)
    free[np.asarray(orc.bc_dofs) // 2] = False
    Kd, Md = J[0::2][:, 0::2].diagonal()[free] / alpha, J[0::2][:, 1::2].diagonal()[free]
    w = alpha * Kd.sum() / Md.sum()
    wn = lambda v: np.sqrt(np.sum(v[0::2] ** 2) + w * np.sum(v[1::2] ** 2))  # noqa: E731
    assert wn(J @ y - rhs) <= 5e-12 * wn(rhs)
    ye = spla.splu(J.tocsc()).solve(rhs)
    assert np.linalg.norm(y - ye) / np.linalg.norm(ye) < 1e-8
    assert its < 80
    Y2 = lvpp.DeviceVector(dev.n, dev.device)
    its2, _, _ = dev.linear_solve(R, Y2, opts)
    assert its2 == its and np.array_equal(y, Y2.numpy())  # bit-reproducible


@pytest.mark.parametrize("kind,n,degree", [("tri", 20, 1), ("tet", 7, 1), ("tri", 8, 2)])
def test_full_lvpp_solve_mg_matches_oracle(lib, kind, n, degree):
    import proximalgalerkin_b200 as lvpp
    from oracle import lvpp_driver

    lvpp_, s, dev, orc = _pair(kind, n, degree)
    xo, ho = lvpp_driver.solve_obstacle(orc, max_outer=500, alpha_scheme="double_exponential", alpha_max=1e2, tol_exit=1e-4)
    sol, total, h = lvpp.obstacle_pg.solve_problem(s["V"].mesh, degree, 500, "double_exponential", 1e2, 1e-4, petsc_options=MG)
    assert h["newton_steps"] == ho["newton_steps"]
    assert h["reason"] == ho["reason"]
    u, uo = sol.x.array[0::2], xo[0::2]
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) < 1e-10
    for key in ("energy", "complementarity", "feasibility", "dual_feasibility", "primal_increment", "latent_increment"):
        assert np.allclose(h[key], ho[key], rtol=1e-7, atol=1e-12), key


@pytest.mark.parametrize("env", [{"LVPP_MG_CHEB": "0"}, {"LVPP_MG_FP32": "0"}, {"LVPP_MG_PACK": "fp32"},
                                 {"LVPP_MG_NPRE": "2", "LVPP_MG_NPOST": "2"}, {"LVPP_MG_NPRE": "1", "LVPP_MG_NPOST": "3"}])
def test_mg_variants_solve_the_same_system(lib, env, monkeypatch):
    """The tunables of the cycle (plain damping instead of Chebyshev roots, fp64 or single-precision records instead of
    the packed bf16 operator, other pre/post degrees) change the preconditioner, never the solution: read at
    lvpp_mg_setup time, i.e. per handle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    lvpp, s, dev, orc = _pair("tet", 9, 1)
    rng = np.random.default_rng(3)
    x = 0.2 * rng.standard_normal(orc.num_rows)
    x[orc.bc_dofs] = 0.0
    dev.set_alpha(3.0)
    dev.set_previous(np.zeros(orc.num_rows))
    X, R, Y = (lvpp.DeviceVector(dev.n, dev.device) for _ in range(3))
    X.set(x)
    dev.assemble_jacobian(X)
    rhs = rng.standard_normal(orc.num_rows)
    R.set(rhs)
    its, reason, _ = dev.linear_solve(R, Y, lvpp.newton_options(dict(MG, ksp_rtol=1e-12)))
    assert reason > 0 and its < 80, (its, reason)
    ye = spla.splu(orc.jacobian(x, 3.0).tocsc()).solve(rhs)
    assert np.linalg.norm(Y.numpy() - ye) / np.linalg.norm(ye) < 1e-8


@pytest.mark.parametrize("kind,n", [("tri", 128), ("tet", 32)])
def test_mid_solve_linear_solve_matches_host_lu_on_both_row_blocks(lib, kind, n):
    """After two proximal steps the contact set has formed (psi down to -20, D = int exp(psi) phi_i phi_j ~ 0 on it): the
    Krylov solution of the Newton system at that state agrees with a host sparse LU of the exported CSR on the u block AND
    on the psi block, and the true residual is small in both blocks (round-1 verdict: a stopping test that is blind to
    the psi rows would pass the first and fail the second)."""
    lvpp, s, dev, orc = _pair(kind, n, 1)
    st = lvpp.obstacle_pg.LvppStepper(s["V"].mesh, 1, "double_exponential", 1e2, 1e-4, petsc_options=MG, setup_objects=s)
    while st.k < 2 and st.step():
        pass
    assert st.k == 2 and st.history["newton_steps"][0] >= 4
    x = st.x.numpy()
    assert x[1::2].min() < -8.0  # a developed contact set
    rng = np.random.default_rng(5)
    X, R, Y, JY = (lvpp.DeviceVector(dev.n, dev.device) for _ in range(4))
    X.set(x)
    fn = dev.assemble_residual(X, R)  # the right-hand side of the next Newton step; leaves the Jacobian at x assembled
    assert fn > 0
    its, reason, _ = dev.linear_solve(R, Y, lvpp.newton_options(dict(MG, ksp_rtol=1e-12)))
    assert reason > 0 and its < 100, (its, reason)
    indptr, indices = dev.csr_pattern()
    J = sp.csr_matrix((dev.jacobian_values().cpu().numpy(), indices, indptr), shape=(dev.n, dev.n))
    rhs, y = R.numpy(), Y.numpy()
    ye = spla.splu(J.tocsc()).solve(rhs)
    for blk, name in ((slice(0, None, 2), "u"), (slice(1, None, 2), "psi")):
        err = np.linalg.norm(y[blk] - ye[blk]) / np.linalg.norm(ye[blk])
        assert err < 1e-8, (name, err)
    r = J @ y - rhs
    scale_u, scale_p = np.abs(J[0::2]).sum(axis=1).max(), np.abs(J[1::2]).sum(axis=1).max()  # row scales of the two blocks
    assert np.linalg.norm(r[0::2]) <= 1e-9 * scale_u * np.linalg.norm(y) / np.sqrt(y.size)
    assert np.linalg.norm(r[1::2]) <= 1e-9 * scale_p * np.linalg.norm(y) / np.sqrt(y.size)


def test_psi_increase_bound_is_inactive_when_large_and_converges_when_small(lib):
    """lvpp_psi_increase_max (not in the reference): a bound no step reaches leaves the whole LVPP solve
    bit-identical to the reference's full Newton step; a small bound changes the Newton path, not the solution."""
    import proximalgalerkin_b200 as lvpp

    def run(extra):
        msh = lvpp.mesh.create_box(10, 10, 10)
        sol, total, h = lvpp.obstacle_pg.solve_problem(msh, 1, 500, "double_exponential", 1e2, 1e-4, petsc_options=dict(MG, **extra))
        return sol.x.array.copy(), h

    x0, h0 = run({})
    x1, h1 = run({"lvpp_psi_increase_max": 1e3})
    assert h1["newton_steps"] == h0["newton_steps"] and np.array_equal(x0, x1)
    x2, h2 = run({"lvpp_psi_increase_max": 0.5})
    assert sum(h2["newton_steps"]) >= sum(h0["newton_steps"])
    u0, u2 = x0[0::2], x2[0::2]
    assert np.linalg.norm(u2 - u0) <= 2e-4 * np.linalg.norm(u0)  # both stop on an increment of 1e-4


def test_psi_bound_with_free_rise_converges(lib):
    """lvpp_psi_free_below: psi may rise freely up to the given value and by psi_increase_max per step beyond it."""
    import proximalgalerkin_b200 as lvpp

    def run(extra):
        msh = lvpp.mesh.create_box(10, 10, 10)
        sol, total, h = lvpp.obstacle_pg.solve_problem(msh, 1, 500, "double_exponential", 1e2, 1e-4, petsc_options=dict(MG, **extra))
        return sol.x.array.copy(), h

    x0, h0 = run({})
    x1, h1 = run({"lvpp_psi_increase_max": 1.0, "lvpp_psi_free_below": 0.0})
    assert np.linalg.norm(x1[0::2] - x0[0::2]) <= 2e-4 * np.linalg.norm(x0[0::2])
    assert sum(h1["newton_steps"]) <= 3 * sum(h0["newton_steps"])  # (a coarse mesh: psi legitimately rises past 1 early on)


def test_keeping_D_at_the_start_of_a_proximal_step_changes_nothing(lib):
    """lvpp_newton_begin_same_iterate keeps D(psi) when only alpha / the previous iterate changed (the start of every
    proximal step after the first): histories and the final iterate are bit-identical to re-assembling."""
    import proximalgalerkin_b200 as lvpp

    def run(reuse):
        msh = lvpp.mesh.create_box(9, 9, 9)
        st = lvpp.obstacle_pg.LvppStepper(msh, 1, "double_exponential", 1e2, 1e-4, petsc_options=MG)
        calls = {"same": 0}
        inner = st.dev.newton_begin

        def begin(x, same_iterate=False):
            calls["same"] += bool(same_iterate)
            return inner(x, same_iterate=same_iterate and reuse)

        st.dev.newton_begin = begin
        while st.step():
            pass
        return st.history, st.x.numpy().copy(), calls["same"]

    h1, x1, n1 = run(True)
    h0, x0, n0 = run(False)
    assert n1 == n0 == len(h1["newton_steps"]) - 1  # every proximal step but the first starts from an evaluated iterate
    assert h1["newton_steps"] == h0["newton_steps"] and h1["krylov_iterations"] == h0["krylov_iterations"]
    assert h1["primal_increment"] == h0["primal_increment"] and np.array_equal(x1, x0)
