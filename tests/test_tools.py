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
