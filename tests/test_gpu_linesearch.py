"""snes_linesearch_type bt on the obstacle engine (host loop of proximalgalerkin_b200/linesearch.py over the
library's assembly / Krylov / J*v entry points) against the oracle's restated PETSc loop."""
import numpy as np
import pytest

# Plain tests: all of them passed on a B200 in round 2 (profiles/r02_gpu_tests.txt); the blanket non-strict xfail of
# round 1 is gone.
pytestmark = [pytest.mark.gpu]


def test_obstacle_bt_matches_oracle(lib):
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs
    from oracle import snes as osnes
    from proximalgalerkin_b200 import linesearch as ls

    n, alpha = 6, 3.0
    msh = lvpp.mesh.create_box(n, n, n)
    s = lvpp.obstacle_pg.setup(msh, 1, petsc_options={"ksp_type": "gmres", "pc_type": "mg", "snes_linesearch_type": "bt"})
    dev = s["problem"].device_problem
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    rng = np.random.default_rng(2)
    xk = 0.3 * rng.standard_normal(orc.num_rows)
    x0 = np.zeros(orc.num_rows)
    x0[1::2] = -4.0  # the full Newton step overshoots from here: bt has to backtrack
    xo, reason_o, its_o, hist_o = osnes.newton_ls(lambda z: orc.assemble_residual(z, xk, alpha), lambda z: orc.jacobian(z, alpha),
                                                  x0, linesearch="bt", rtol=1e-10, max_it=60)
    dev.set_alpha(alpha)
    dev.set_previous(xk)
    X = lvpp.DeviceVector(dev.n, dev.device)
    X.set(x0)
    opts = lvpp.newton_options(s["options"])
    nb = ls.NewtonBT(ls.DeviceBackend(dev, opts), rtol=1e-10, max_it=60)
    hist, lams = [nb.begin(X)], []
    while not nb.reason:
        nb.step(X)
        hist.append(nb.fnorm)
        lams.append(nb.last_lambda)
    assert (nb.reason, nb.its) == (reason_o, its_o), (nb.reason, nb.its, reason_o, its_o)
    assert min(lams) < 1.0
    assert np.allclose(hist[: len(hist_o)], hist_o, rtol=1e-6, atol=1e-14)
    assert np.linalg.norm(X.numpy() - xo) <= 1e-8 * np.linalg.norm(xo)


def test_obstacle_l2_matches_oracle(lib):
    """snes_linesearch_type l2 (+ snes_linesearch_maxlambda 1, the setting of fracture_dolfinx.py:132-138) through
    NonlinearProblem.solve() against the restated PETSc loop."""
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs
    from oracle import snes as osnes

    n, alpha = 6, 3.0
    msh = lvpp.mesh.create_box(n, n, n)
    s = lvpp.obstacle_pg.setup(msh, 1, petsc_options={"ksp_type": "gmres", "pc_type": "mg", "ksp_rtol": 1e-12, "snes_rtol": 1e-10,
                                                       "snes_max_it": 80, "snes_linesearch_type": "l2", "snes_linesearch_maxlambda": 1})
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    rng = np.random.default_rng(2)
    xk = 0.3 * rng.standard_normal(orc.num_rows)
    x0 = np.zeros(orc.num_rows)
    x0[1::2] = -2.0
    xo, reason_o, its_o, _ = osnes.newton_ls(lambda z: orc.assemble_residual(z, xk, alpha), lambda z: orc.jacobian(z, alpha),
                                             x0, linesearch="l2", rtol=1e-10, max_it=80, maxstep=1.0)
    s["alpha"].value = alpha
    s["sol_k"].x.array[:] = xk
    s["sol"].x.array[:] = x0
    s["problem"].solve()
    assert (s["problem"].solver.getConvergedReason(), s["problem"].solver.getIterationNumber()) == (reason_o, its_o)
    assert np.linalg.norm(s["sol"].x.array - xo) <= 1e-8 * np.linalg.norm(xo)


def test_nonlinear_problem_bt_option(lib):
    """The option routes NonlinearProblem.solve() through the same loop; from the zero start of the LVPP iteration
    the full step is always accepted, so bt reproduces the line-search-free Newton counts."""
    import proximalgalerkin_b200 as lvpp

    msh = lvpp.mesh.create_rectangle(10, 10)
    base = {"ksp_type": "gmres", "pc_type": "mg"}
    _, tot_none, h_none = lvpp.obstacle_pg.solve_problem(msh, 1, 500, "double_exponential", 1e2, 1e-4, petsc_options=base)
    _, tot_bt, h_bt = lvpp.obstacle_pg.solve_problem(msh, 1, 500, "double_exponential", 1e2, 1e-4,
                                                     petsc_options=dict(base, snes_linesearch_type="bt"))
    assert h_bt["reason"] == h_none["reason"]
    assert h_bt["newton_steps"] == h_none["newton_steps"]
    assert np.allclose(h_bt["primal_increment"], h_none["primal_increment"], rtol=1e-6)


def test_residual_and_jacobian_at_extreme_latent_values(lib):
    """States of a late proximal step: psi between -60 (deep inside the contact set) and +1 next to each other, u of
    order 1e-2.  F and J against the oracle to the same 1e-12 as at the moderate states of tests/test_gpu_parity.py."""
    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    n = 8
    msh = lvpp.mesh.create_box(n, n, n)
    s = lvpp.obstacle_pg.setup(msh, 1)
    dev = s["problem"].device_problem
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    rng = np.random.default_rng(9)
    x = np.zeros(orc.num_rows)
    x[0::2] = 1e-2 * rng.standard_normal(orc.num_rows // 2)
    x[1::2] = np.where(rng.random(orc.num_rows // 2) > 0.5, -60.0 + 30.0 * rng.random(orc.num_rows // 2), rng.random(orc.num_rows // 2))
    xk = x.copy()
    xk[1::2] += 20.0 * rng.random(orc.num_rows // 2)
    alpha = 1.49
    dev.set_alpha(alpha)
    dev.set_previous(xk)
    X, F = lvpp.DeviceVector(dev.n, dev.device), lvpp.DeviceVector(dev.n, dev.device)
    X.set(x)
    dev.assemble_residual(X, F)
    Fo = orc.assemble_residual(x, xk, alpha)
    assert np.abs(F.numpy() - Fo).max() <= 1e-12 * np.abs(Fo).max()
    vals = dev.jacobian_values().cpu().numpy()
    vo = orc.assemble_jacobian_values(x, alpha)
    assert np.abs(vals - vo).max() <= 1e-12 * np.abs(vo).max()
    # entries of D span 26 orders of magnitude: compare them relative to themselves as well, where they are not
    # below the rounding level of the sums they sit in
    big = np.abs(vo) > 1e-14 * np.abs(vo).max()
    assert np.abs(vals[big] - vo[big]).max() <= 1e-9 * np.abs(vo[big]).max()


@pytest.mark.parametrize("env", [{}, {"LVPP_GMRES_WEIGHT": "off"}, {"LVPP_GMRES_RESTART": "40"},
                                 {"LVPP_GMRES_RESTART": "40", "LVPP_GMRES_WEIGHT": "off"}])
def test_gmres_switches_solve_the_same_system(lib, env, monkeypatch):
    """The residual norm (equilibrated by default, Euclidean with LVPP_GMRES_WEIGHT=off) and the
    restart length change the Krylov process, not the solution (restart 40: several cycles, the restart path of the
    device-resident recurrence)."""
    _solve_with_env(env, monkeypatch)


def _solve_with_env(env, monkeypatch):
    import scipy.sparse.linalg as spla

    import proximalgalerkin_b200 as lvpp
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n = 9
    msh = lvpp.mesh.create_box(n, n, n)
    opts_d = {"ksp_type": "gmres", "pc_type": "mg"}
    s = lvpp.obstacle_pg.setup(msh, 1, petsc_options=opts_d)
    dev = s["problem"].device_problem
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    rng = np.random.default_rng(3)
    x = 0.2 * rng.standard_normal(orc.num_rows)
    x[1::2] -= 30.0 * (rng.random(orc.num_rows // 2) > 0.7)  # a developed contact set: D ~ 0 on a third of the nodes
    x[orc.bc_dofs] = 0.0
    alpha = 5.0
    dev.set_alpha(alpha)
    dev.set_previous(np.zeros(orc.num_rows))
    X, R, Y = (lvpp.DeviceVector(dev.n, dev.device) for _ in range(3))
    X.set(x)
    dev.assemble_jacobian(X)
    rhs = rng.standard_normal(orc.num_rows)
    R.set(rhs)
    its, reason, _ = dev.linear_solve(R, Y, lvpp.newton_options(dict(opts_d, ksp_rtol=1e-12)))
    assert reason > 0 and its < (400 if "LVPP_GMRES_RESTART" in env else 120), (its, reason)
    J = orc.jacobian(x, alpha)
    ye = spla.splu(J.tocsc()).solve(rhs)
    assert np.linalg.norm(Y.numpy() - ye) / np.linalg.norm(ye) < 1e-7


@pytest.mark.parametrize("max_it", [100, 5])
def test_adaptive_alpha_with_recovery_matches_oracle(lib, max_it):
    """alpha_scheme="adaptive" (recovery.py): host-buffer driver and device-resident stepper against the
    restated loop of fracture_dolfinx.py:215-283.  snes_max_it = 5 makes the first proximal step fail at alpha = 1
    (reason -5), so the halve-restore-retry branch runs on real arithmetic."""
    import proximalgalerkin_b200 as lvpp
    from oracle import lvpp_driver
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    n = 8
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    xo, ho = lvpp_driver.solve_obstacle_adaptive(orc, max_outer=40, alpha_max=1e5, tol_exit=1e-4, nfail_max=12, snes_max_it=max_it)
    opts = {"ksp_type": "gmres", "pc_type": "mg", "ksp_rtol": 1e-12, "snes_max_it": max_it,
            "snes_error_if_not_converged": False, "ksp_error_if_not_converged": False}
    msh = lvpp.mesh.create_box(n, n, n)
    sol, total, h = lvpp.obstacle_pg.solve_problem(msh, 1, 40, "adaptive", 1e5, 1e-4, petsc_options=opts, adaptive=dict(nfail_max=12))
    assert [tuple(a) for a in h["attempts"]] == [tuple(a) for a in ho["attempts"]]
    assert h["alpha"] == ho["alpha"]
    u, uo = sol.x.array[0::2], xo[0::2]
    assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo)
    st = lvpp.obstacle_pg.LvppStepper(msh, 1, "adaptive", 1e5, 1e-4, max_outer=40, petsc_options=opts, adaptive=dict(nfail_max=12))
    while st.step():
        pass
    assert [tuple(a) for a in st.history["attempts"]] == [tuple(a) for a in ho["attempts"]]
    assert st.history["newton_steps"] == ho["newton_steps"]
    us = st.x.numpy()[0::2]
    assert np.linalg.norm(us - uo) <= 1e-10 * np.linalg.norm(uo)


def _export_cases():
    import glob
    from pathlib import Path

    return ["standin-2d", "standin-3d"] + sorted(glob.glob(str(Path(__file__).parent / "golden" / "dolfinx_*.npz")))


@pytest.mark.parametrize("case", _export_cases())
def test_device_path_against_reference_style_export(lib, case):
    """The device path fed with a mesh, dof maps and quadrature rule in somebody else's numbering (a
    tools/export_from_dolfinx.py file when one exists; otherwise the oracle's scrambled stand-in of
    tests/test_dolfinx_golden.py): pattern bit-exact, J and F to 1e-12, Newton counts, final u to 1e-10."""
    import scipy.sparse as sp

    import proximalgalerkin_b200 as lvpp
    from test_dolfinx_golden import standin_export

    g = standin_export(8, 2, seed=5) if case == "standin-2d" else standin_export(5, 3, seed=6) if case == "standin-3d" else dict(np.load(case))
    msh = lvpp.mesh.from_arrays(g["coords"], g["cells"])
    rule = (np.asarray(g["qpts"]), np.asarray(g["qwts"]))
    dof_u, dof_psi = np.asarray(g["dof_u"]), np.asarray(g["dof_psi"])
    nv = dof_u.size
    perm = np.empty(2 * nv, dtype=np.int64)   # device row (node-interleaved) -> exported dof
    perm[0::2], perm[1::2] = dof_u, dof_psi
    s = lvpp.obstacle_pg.setup(msh, 1, rule=rule)
    dev = s["problem"].device_problem
    alpha = float(g["alpha"])
    dev.set_alpha(alpha)
    dev.set_previous(np.asarray(g["xk"])[perm])
    X, F = lvpp.DeviceVector(dev.n, dev.device), lvpp.DeviceVector(dev.n, dev.device)
    X.set(np.asarray(g["x"])[perm])
    dev.assemble_residual(X, F)
    Fg = np.asarray(g["F"])[perm]
    assert np.abs(F.numpy() - Fg).max() <= 1e-12 * np.abs(Fg).max()
    vals = dev.jacobian_values().cpu().numpy()
    indptr, indices = dev.csr_pattern()
    Jd = sp.csr_matrix((vals, indices, indptr), shape=(2 * nv, 2 * nv))
    Jg = sp.csr_matrix((g["J_data"], g["J_indices"], g["J_indptr"]), shape=(2 * nv, 2 * nv))[perm][:, perm].tocsr()
    Jg.sort_indices()
    Jd.sort_indices()
    assert np.array_equal(Jg.indptr, Jd.indptr) and np.array_equal(Jg.indices, Jd.indices)
    assert np.abs(Jg.data - Jd.data).max() <= 1e-12 * np.abs(Jg.data).max()
    if "newton_steps" in g:
        s2 = lvpp.obstacle_pg.setup(msh, 1, rule=rule, petsc_options={"ksp_type": "gmres", "pc_type": "mg", "ksp_rtol": 1e-12})
        st = lvpp.obstacle_pg.LvppStepper(msh, 1, "double_exponential", 1e2, 1e-4, setup_objects=s2)
        while st.step():
            pass
        assert st.history["newton_steps"] == [int(v) for v in g["newton_steps"]]
        uf = st.x.numpy()[0::2]
        assert np.linalg.norm(uf - g["u_final"]) <= 1e-10 * np.linalg.norm(g["u_final"])


@pytest.mark.parametrize("env", [{"LVPP_MG_PACK": "bf16"}, {"LVPP_MG_PACK": "fp32"}, {"LVPP_MG_FP32": "0"}])
def test_cycle_record_formats_precondition_the_same_system(lib, env, monkeypatch):
    """The cycle's copy of the operator: bf16 pair records + single-precision block inverses (the default,
    block_op.cuh:k_packed2_op), single-precision records (LVPP_MG_PACK=fp32, k_packed_op) or the fp64 operator itself
    (LVPP_MG_FP32=0) -- different preconditioners, the same solution."""
    _solve_with_env(env, monkeypatch)
