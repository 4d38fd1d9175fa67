"""Generates the golden fixtures of the gradient-constraint, multiphase and Signorini forms from the CPU
oracle (oracle/forms.py, oracle/lvpp_driver.py).  Like make_golden.py these pin the *restatement*
(parity unpinned: the reference ships no vectors for these forms).

Run from the repository root:  python tests/golden/make_golden_forms.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import forms, lvpp_driver, mesh  # noqa: E402

OUT = Path(__file__).resolve().parent


def states(name):
    """(oracle, x, aux dict) at the seeded non-trivial state shared with tests/test_gpu_forms.py::test_golden."""
    rng = np.random.default_rng(777)
    if name == "gradient_n5":
        orc = forms.GradientConstraintOracle(mesh.rectangle(5, 5, lo=0.0, hi=1.0))
        x = 0.3 * rng.standard_normal(orc.num_rows)
        orc.alpha, orc.w0 = 2.25, 0.2 * rng.standard_normal(orc.num_rows)
        aux = {"aux0": orc.w0}
    elif name == "multiphase_n4":
        orc = forms.MultiphaseOracle(mesh.rectangle(4, 4, diagonal="crossed", lo=0.0, hi=1.0))
        x = 0.3 * rng.standard_normal(orc.num_rows)
        orc.alpha, orc.lvpp_old, orc.u_prev = 1.5, 0.2 * rng.standard_normal(orc.num_rows), rng.random((orc.N, 4))
        a1 = np.zeros(orc.num_rows)
        a1.reshape(orc.N, 3, 4)[:, 0, :] = orc.u_prev
        aux = {"aux0": orc.lvpp_old, "aux1": a1}
    else:
        orc = forms.SignoriniOracle(mesh.box_kuhn(4, 4, 3, lo=(0, 0, 0), hi=(1, 1, 1)), disp=-0.15)
        x = 0.01 * rng.standard_normal(orc.num_rows)
        pk = np.zeros(orc.num_rows)
        pk[3 * orc.N:] = 0.2 * rng.standard_normal(orc.NS)
        orc.alpha, orc.psi_k = 0.75, pk[3 * orc.N:].copy()
        aux = {"aux0": pk}
    return orc, x, aux


def main():
    for name in ("gradient_n5", "multiphase_n4", "signorini_n4"):
        orc, x, aux = states(name)
        np.savez_compressed(OUT / f"form_{name}.npz", x=x, alpha=orc.alpha, F=orc.assemble_residual(x),
                            jac_values=orc.assemble_jacobian_values(x), indptr=orc.indptr, indices=orc.indices, **aux)
        print(name, orc.num_rows, "rows", orc.nnz, "nnz")
    g = forms.GradientConstraintOracle(mesh.rectangle(6, 6, lo=0.0, hi=1.0))
    xg, hg = lvpp_driver.solve_gradient_constraint(g)
    m = mesh.rectangle(5, 5, diagonal="crossed", lo=0.0, hi=1.0)
    mp = forms.MultiphaseOracle(m)
    mp.u_prev = lvpp_driver.multiphase_initial_condition(m.coords, m.cells)
    xm, hm = lvpp_driver.solve_multiphase(mp, num_steps=2)
    s = forms.SignoriniOracle(mesh.box_kuhn(3, 3, 3, lo=(0, 0, 0), hi=(1, 1, 1)), disp=-0.15)
    xs, hs = lvpp_driver.solve_signorini(s, alpha_0=0.005)
    np.savez_compressed(OUT / "form_loops.npz", gradient_newton=np.array(hg["newton_steps"]), gradient_l2=np.array(hg["l2_diff"]),
                        gradient_solution=xg, multiphase_newton=np.array(hm["newton_iterations"]),
                        multiphase_lvpp=np.array(hm["lvpp_iterations"]), multiphase_solution=xm,
                        signorini_iterations=np.array(hs["iterations"]), signorini_it=hs["it"], signorini_solution=xs)
    print("loops", hg["newton_steps"], hm, hs)


if __name__ == "__main__":
    main()
