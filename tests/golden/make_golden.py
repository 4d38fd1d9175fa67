"""Generates the golden fixtures of this directory from the CPU oracle (oracle/).

The reference holds no golden vectors for this path and its stack (dolfinx/PETSc) cannot be
imported here, so these fixtures pin the *restatement* (parity unpinned, see oracle/__init__.py):
they guard the oracle against regressions and give the GPU tests a committed target.

Run from the repository root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import lvpp_driver, mesh, obstacle  # noqa: E402

OUT = Path(__file__).resolve().parent


def case(name, msh, degree):
    orc = obstacle.ObstacleOracle(msh, degree=degree)
    rng = np.random.default_rng(2024)
    x = 0.3 * rng.standard_normal(orc.num_rows)
    xk = 0.3 * rng.standard_normal(orc.num_rows)
    alpha = 2.25
    F = orc.assemble_residual(x, xk, alpha)
    vals = orc.assemble_jacobian_values(x, alpha)
    obs = orc.observables(x, xk, alpha)
    xs, h = lvpp_driver.solve_obstacle(orc, max_outer=500, alpha_scheme="double_exponential", alpha_max=1e2, tol_exit=1e-4)
    np.savez_compressed(
        OUT / f"{name}.npz", x=x, xk=xk, alpha=alpha, F=F, jac_values=vals, indptr=orc.indptr, indices=orc.indices,
        observables=obs, solution=xs, newton_steps=np.array(h["newton_steps"]), alphas=np.array(h["alpha"]),
        primal_increment=np.array(h["primal_increment"]), energy=np.array(h["energy"]),
        fnorm0=np.array([f[0] for f in h["fnorms"]]), fnorm_last=np.array([f[-1] for f in h["fnorms"]]))
    print(name, orc.num_rows, "rows", h["newton_steps"])


if __name__ == "__main__":
    case("tri_p1_n12", mesh.rectangle(12, 12), 1)
    case("tet_p1_n5", mesh.box_kuhn(5, 5, 5), 1)
    case("tri_p2_n6", mesh.rectangle(6, 6), 2)
