"""GPU parity of the generic mixed-form engine (gradient constraint, multiphase, Signorini; SURVEY.md
section 8a rows a13-a18) against the numpy restatement in oracle/forms.py, through the C ABI:
pattern bit-exact, residual and Jacobian entries within 1e-12 of the largest entry, J*v, the Krylov
solve against sparse LU, and the full LVPP loops with identical Newton / proximal iteration counts."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _gradient(n):
    import proximalgalerkin_b200 as lvpp
    from oracle import forms as oforms, mesh as omesh

    s = lvpp.gradient_constraints.setup(n, n)
    orc = oforms.GradientConstraintOracle(omesh.rectangle(n, n, lo=0.0, hi=1.0))
    rng = np.random.default_rng(3)
    x = 0.3 * rng.standard_normal(orc.num_rows)
    aux = [0.2 * rng.standard_normal(orc.num_rows), None]
    orc.alpha, orc.w0 = 1.7, aux[0]
    return s["dev"], orc, x, aux, [(0, 1.7)]


def _multiphase(n):
    import proximalgalerkin_b200 as lvpp
    from oracle import forms as oforms, mesh as omesh

    s = lvpp.multiphase.setup(n, n)
    orc = oforms.MultiphaseOracle(omesh.rectangle(n, n, diagonal="crossed", lo=0.0, hi=1.0))
    rng = np.random.default_rng(4)
    x = 0.3 * rng.standard_normal(orc.num_rows)
    lv = 0.2 * rng.standard_normal(orc.num_rows)
    up = rng.random((orc.N, 4))
    a1 = np.zeros(orc.num_rows)
    a1.reshape(orc.N, 3, 4)[:, 0, :] = up
    orc.alpha, orc.lvpp_old, orc.u_prev = 1.3, lv, up
    return s["dev"], orc, x, [lv, a1], [(0, 1.3)]


def _signorini(n):
    import proximalgalerkin_b200 as lvpp
    from oracle import forms as oforms, mesh as omesh

    msh = lvpp.mesh.create_box(n, n, n - 1, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0))
    s = lvpp.signorini.setup(msh, disp=-0.1)
    orc = oforms.SignoriniOracle(omesh.box_kuhn(n, n, n - 1, lo=(0, 0, 0), hi=(1, 1, 1)), disp=-0.1)
    rng = np.random.default_rng(5)
    x = 0.01 * rng.standard_normal(orc.num_rows)
    pk = np.zeros(orc.num_rows)
    pk[3 * orc.N:] = 0.2 * rng.standard_normal(orc.NS)
    orc.alpha, orc.psi_k = 0.7, pk[3 * orc.N:].copy()
    return s["dev"], orc, x, [pk, None], [(0, 0.7)]


CASES = {"gradient": (_gradient, 5), "multiphase": (_multiphase, 4), "signorini": (_signorini, 4)}


@pytest.mark.parametrize("name", list(CASES))
def test_assembly_matches_oracle(lib, name):
    make, n = CASES[name]
    dev, orc, x, aux, params = make(n)
    assert np.array_equal(dev.indptr, orc.indptr) and np.array_equal(dev.indices, orc.indices)  # create_matrix
    for i, v in params:
        dev.set_param(i, v)
    for k, a in enumerate(aux):
        if a is not None:
            dev.set_aux(k, a)
    X, F, Y = dev.vector(x), dev.vector(), dev.vector()
    fnorm = dev.assemble_residual(X, F)
    Fo = orc.assemble_residual(x)
    assert np.abs(F.numpy() - Fo).max() <= TOL * np.abs(Fo).max()
    assert abs(fnorm - np.linalg.norm(Fo)) <= 1e-12 * np.linalg.norm(Fo)
    vals, vo = dev.jacobian_values(), orc.assemble_jacobian_values(x)
    assert np.abs(vals - vo).max() <= TOL * np.abs(vo).max()
    # bit-reproducible (fixed summation order, no atomics)
    F2 = dev.vector()
    dev.assemble_residual(X, F2)
    assert np.array_equal(F.numpy(), F2.numpy()) and np.array_equal(vals, dev.jacobian_values())
    # J*v
    J = sp.csr_matrix((vo, orc.indices, orc.indptr), shape=(orc.num_rows,) * 2)
    v = np.random.default_rng(9).standard_normal(orc.num_rows)
    dev.spmv(dev.vector(v), Y)
    assert np.abs(Y.numpy() - J @ v).max() <= 1e-13 * np.abs(J @ v).max()
    # Krylov solve against sparse LU
    import proximalgalerkin_b200 as lvpp

    opts = lvpp.newton_options({"ksp_rtol": 1e-13, "ksp_gmres_restart": 400}, generic=True)
    rhs = np.random.default_rng(10).standard_normal(orc.num_rows)
    its, reason, rnorm = dev.linear_solve(dev.vector(rhs), Y, opts)
    assert reason > 0, (its, reason, rnorm)
    ye = spla.splu(J.tocsc()).solve(rhs)
    assert np.linalg.norm(Y.numpy() - ye) <= 1e-8 * np.linalg.norm(ye), its


def test_gradient_constraint_lvpp_matches_oracle(lib):
    import proximalgalerkin_b200 as lvpp
    from oracle import forms as oforms, lvpp_driver, mesh as omesh

    n = 8
    orc = oforms.GradientConstraintOracle(omesh.rectangle(n, n, lo=0.0, hi=1.0))
    xo, ho = lvpp_driver.solve_gradient_constraint(orc)
    its, l2 = lvpp.gradient_constraints.solve_problem(n, n, petsc_options={"ksp_gmres_restart": 300})
    assert list(its) == ho["newton_steps"]
    assert np.allclose(l2, ho["l2_diff"], rtol=1e-6, atol=2e-11)  # increments reach 4e-9 of a solution of size 0.1
    sol = lvpp.gradient_constraints.solve_problem.last["sol"]
    N2 = orc.N2
    assert np.linalg.norm(sol[:N2] - xo[:N2]) <= 1e-9 * np.linalg.norm(xo[:N2])


def test_multiphase_time_steps_match_oracle(lib):
    import proximalgalerkin_b200 as lvpp
    from oracle import forms as oforms, lvpp_driver, mesh as omesh

    n = 6
    m = omesh.rectangle(n, n, diagonal="crossed", lo=0.0, hi=1.0)
    orc = oforms.MultiphaseOracle(m)
    orc.u_prev = lvpp_driver.multiphase_initial_condition(m.coords, m.cells)
    assert np.array_equal(orc.u_prev, lvpp.multiphase.initial_condition(m.coords, m.cells))
    xo, ho = lvpp_driver.solve_multiphase(orc, num_steps=2)
    newton, lvpp_its = lvpp.multiphase.solve_problem(n, n, T=2e-5, tau0=1e-5, petsc_options={"ksp_gmres_restart": 400})
    assert list(newton) == ho["newton_iterations"] and list(lvpp_its) == ho["lvpp_iterations"]
    sol = lvpp.multiphase.solve_problem.last["sol"]
    u, uo = sol.reshape(-1, 3, 4)[:, 0, :], xo.reshape(-1, 3, 4)[:, 0, :]
    assert np.linalg.norm(u - uo) <= 1e-8 * np.linalg.norm(uo)


@pytest.mark.parametrize("disp", [-0.1, -0.2])
def test_signorini_lvpp_matches_oracle(lib, disp):
    import proximalgalerkin_b200 as lvpp
    from oracle import forms as oforms, lvpp_driver, mesh as omesh

    n = 4
    orc = oforms.SignoriniOracle(omesh.box_kuhn(n, n, n, lo=(0, 0, 0), hi=(1, 1, 1)), disp=disp)
    xo, ho = lvpp_driver.solve_signorini(orc, alpha_0=0.005)  # the CI parameters, test_dolfinx.yml:39-41
    msh = lvpp.mesh.create_box(n, n, n, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0))
    it, iterations = lvpp.signorini.solve_contact_problem(msh, disp=disp, alpha_0=0.005, petsc_options={"ksp_gmres_restart": 400})
    assert it == ho["it"] and iterations == ho["iterations"]
    last = lvpp.signorini.solve_contact_problem.last
    sol = last["sol"]
    nu = 3 * orc.N
    assert np.linalg.norm(sol[:nu] - xo[:nu]) <= 1e-9 * np.linalg.norm(xo[:nu])
    # the blocked call shape of signorini_dolfinx.py:283-291: NonlinearProblem(F, [u, psi], ...) keeps the two blocks in
    # the caller's functions
    assert isinstance(last["problem"], lvpp.NonlinearProblem) and last["problem"].u == [last["u"], last["psi"]]
    assert np.array_equal(last["u"].x.array, sol[:nu]) and np.array_equal(last["psi"].x.array, sol[nu:])
    with pytest.raises(TypeError):
        lvpp.NonlinearProblem(object(), [last["u"], last["psi"]])
