"""The research tools keep running: tools/mg_prototype.py (numpy mirror of the multigrid GMRES) solves the Newton
systems of a small LVPP run to the same tolerance as sparse LU."""
import importlib.util
from pathlib import Path

import numpy as np
import scipy.sparse.linalg as spla

ROOT = Path(__file__).resolve().parents[1]


def _load(name):
    spec = importlib.util.spec_from_file_location(name, ROOT / "tools" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_mg_prototype_solves_newton_systems():
    mp = _load("mg_prototype")
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    orc = oobs.ObstacleOracle(omesh.box_kuhn(5, 5, 5))
    states = mp.newton_states(orc, 2)
    k, it, x, xk, alpha, F = states[-1]
    for slabs, cheb, gamma in ((1, 10.0, 1), (2, 4.0, 1), (1, 0.0, 2)):
        mg = mp.Multigrid(orc, x, alpha, slabs=slabs)
        mg.set_smoother(cheb=cheb)
        L0 = mg.levels[0]
        rhs = mp.to_blocked(orc, F)
        y, its = mp.gmres_right(L0.J, lambda v: mg.cycle(v, 0, gamma), rhs)
        assert its < 60
        assert np.linalg.norm(L0.J @ y - rhs) <= 1e-10 * np.linalg.norm(rhs)
        # the prototype's masked operator is the oracle's Jacobian in blocked ordering
        ye = spla.splu(orc.jacobian(x, alpha).tocsc()).solve(F)
        assert np.linalg.norm(y - mp.to_blocked(orc, ye)) <= 1e-7 * np.linalg.norm(ye)


def test_mg_precision_quantisers_and_pair_layout():
    """tools/mg_precision.py: bf16 rounding as k_pack_op_bf16 stores it (round to nearest even on the upper 16 bits of the
    single-precision pattern), and the address rule of the bf16 pair records (block_op.cuh: pair p of slice s at
    ((slice_ptr[s] + 32 s) >> 1) + 32 p + lane): slices never overlap and fit the allocation slots / 2 + 16 nslices."""
    mq = _load("mg_precision")
    a = np.array([1.0, 1.0 + 2.0**-8, 1.0 + 3 * 2.0**-8, 1.0 + 2.0**-7, -3.0e38, 1e-30, 0.0])
    q = mq.q_bf16(a)
    assert q[0] == 1.0 and q[1] == 1.0 and q[2] == 1.0 + 2.0**-6 and q[3] == 1.0 + 2.0**-7  # ties go to the even mantissa
    assert np.isfinite(q).all() and abs(q[4] / a[4] - 1) < 2.0**-8 and abs(q[5] / a[5] - 1) < 2.0**-8 and q[6] == 0.0
    assert mq.q_fp16(np.array([1e9]))[0] == 65504.0
    rng = np.random.default_rng(0)
    for _ in range(100):
        ns = int(rng.integers(1, 40))
        w = rng.integers(1, 40, size=ns)
        sp_ = np.concatenate([[0], np.cumsum(32 * w)])
        po = (sp_[:-1] + 32 * np.arange(ns)) >> 1
        end = po + 32 * ((w + 1) // 2)
        assert np.all(end[:-1] <= po[1:]) and end[-1] <= (sp_[-1] >> 1) + 16 * ns and np.all(po % 16 == 0)
    # the cycle with bf16 operator values still preconditions the Newton systems of a small run like the exact one
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    mp = mq.mp
    orc = oobs.ObstacleOracle(omesh.box_kuhn(6, 6, 6))
    k, it, x, xk, alpha, F = mp.newton_states(orc, 3)[-1]
    mg = mp.Multigrid(orc, x, alpha)
    mg.set_smoother(cheb=6.0)
    J0, rhs = mg.levels[0].J, mp.to_blocked(orc, F)
    _, its_exact = mp.gmres_right(J0, lambda v: mg.cycle(v, 0, 1), rhs)
    for L, (J, B) in zip(mg.levels, mq.quantised_cycle_operators(mg, alpha, mq.q_bf16, mq.q_fp32)):
        L.J, L.Binv = J, B
    y, its_bf16 = mp.gmres_right(J0, lambda v: mg.cycle(v, 0, 1), rhs)
    assert its_bf16 <= its_exact + 4  # (this late state moves by +-2 with any perturbation of the cycle: fp32 records give 16)
    # single-precision vectors inside the cycle need the flexible variant: plain right-preconditioned GMRES degrades
    prec32 = lambda v: mq.q_fp32(mq.cycle32(mg, mq.q_fp32(v)))  # noqa: E731
    yf, its_fg = mq.fgmres(J0, prec32, rhs)
    _, its_plain = mp.gmres_right(J0, prec32, rhs)
    assert its_fg <= its_exact + 5 and its_plain > its_fg
    assert np.linalg.norm(J0 @ yf - rhs) <= 1e-10 * np.linalg.norm(rhs)
    assert np.linalg.norm(J0 @ y - rhs) <= 1e-10 * np.linalg.norm(rhs)


def test_full_solve_cpu_tool_runs_a_whole_lvpp_solve(capsys, monkeypatch):
    """tools/full_solve_cpu.py (the whole LVPP solve with the prototype's multigrid GMRES as linear solver): Newton
    counts of the oracle's exact-LU solve on a small mesh, with the Euclidean and the equilibrated residual norm."""
    import sys

    from oracle import lvpp_driver
    from oracle import mesh as omesh
    from oracle import obstacle as oobs

    orc = oobs.ObstacleOracle(omesh.box_kuhn(6, 6, 6))
    _, h = lvpp_driver.solve_obstacle(orc, 500, "double_exponential", 1e2, 1e-4)
    for extra in ([], ["--equilibrate"]):
        fs = _load("full_solve_cpu")
        monkeypatch.setattr(sys, "argv", ["full_solve_cpu.py", "--size", "6"] + extra)
        fs.main()
        out = capsys.readouterr().out
        assert "converged" in out and "diverged" not in out
        counts = [int(ln.split(":")[1].split()[0]) for ln in out.splitlines() if ln.startswith("== outer")]
        assert counts == h["newton_steps"], (extra, counts, h["newton_steps"])
