#!/usr/bin/env python
"""Whole LVPP obstacle solve on the CPU at sizes sparse LU cannot reach: the oracle's assembly and Newton loop
(line search none, the reference's options) with the prototype's multigrid-preconditioned GMRES (tools/mg_prototype.py)
as the linear solver, converged to 1e-12 like the device path.  A research tool: it answers whether the full Newton
step stays robust as the 3-D mesh is refined without a GPU.

  python tools/full_solve_cpu.py --size 48
"""
import argparse
import importlib.util
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import lvpp_driver, mesh as omesh, obstacle as oobs, snes  # noqa: E402

spec = importlib.util.spec_from_file_location("mg_prototype", ROOT / "tools" / "mg_prototype.py")
mp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--cheb", type=float, default=6.0)
    ap.add_argument("--outer", type=int, default=500)
    ap.add_argument("--restart", type=int, default=50)
    ap.add_argument("--equilibrate", action="store_true",
                    help="GMRES on S J S, S = diag(1, sqrt(a/m)) on (u, psi) with a, m the medians of alpha K_ii and M_ii: "
                         "the residual norm weighs the psi rows by a/m (LVPP_GMRES_WEIGHT=auto in the library)")
    ap.add_argument("--snes-rtol", dest="snes_rtol", type=float, default=1e-6)
    args = ap.parse_args()
    n = args.size
    t0 = time.time()
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    print(f"# n={n}: {orc.num_rows} rows, setup {time.time() - t0:.0f} s", flush=True)
    N = orc.num_nodes
    x = np.zeros(orc.num_rows)
    xk = x.copy()
    alpha_k, alpha = 1, 1.0
    for k in range(args.outer):
        alpha, alpha_k = lvpp_driver.alpha_schedule("double_exponential", k, alpha_k, 1e2, alpha_current=alpha)
        F = orc.assemble_residual(x, xk, alpha)
        f0 = np.linalg.norm(F)
        reason, it = 0, 0
        while not reason:
            mg = mp.Multigrid(orc, x, alpha)
            mg.set_smoother(cheb=args.cheb)
            L0 = mg.levels[0]
            rhs = mp.to_blocked(orc, F)
            if args.equilibrate:
                w = np.sqrt(np.median(alpha * L0.K.diagonal()) / np.median(L0.M.diagonal()))
                Sd = np.concatenate([np.ones(N), np.full(N, w)])
                S = sp.diags(Sd)
                yt, kits = mp.gmres_right((S @ L0.J @ S).tocsr(), lambda v: mg.cycle(v / Sd, 0, 1) / Sd, Sd * rhs, rtol=1e-12,
                                          restart=args.restart, maxit=600)
                yb = Sd * yt
            else:
                yb, kits = mp.gmres_right(L0.J, lambda v: mg.cycle(v, 0, 1), rhs, rtol=1e-12, restart=args.restart, maxit=600)
            y = np.empty_like(x)
            y[orc.dof_u], y[orc.dof_psi] = yb[:N], yb[N:]
            x = x - y
            with np.errstate(over="ignore", invalid="ignore"):
                F = orc.assemble_residual(x, xk, alpha)
            fn = float(np.linalg.norm(F))
            it += 1
            psi = x[orc.dof_psi]
            print(f"outer {k} alpha {alpha:.4g} newton {it}: |F| {fn:.3e} |y| {np.linalg.norm(y):.3e} krylov {kits} "
                  f"max +dpsi {(-y[orc.dof_psi]).max():.2f} psi [{psi.min():.1f}, {psi.max():.2f}] ({time.time() - t0:.0f} s)", flush=True)
            reason = snes.converged_default(it, np.linalg.norm(x), np.linalg.norm(y), fn, f0 * args.snes_rtol, f0, 1e-50, 1e-8, 1e4)
            if not reason and it >= 100:
                reason = snes.DIVERGED_MAX_IT
        if reason < 0:
            print(f"SNES diverged: reason {reason} in proximal step {k} (alpha {alpha:.4g})")
            return
        obs = orc.observables(x, xk, alpha)
        incr = float(np.sqrt(obs[4]))
        print(f"== outer {k} alpha {alpha:.4g}: {it} Newton steps, increment {incr:.3e}", flush=True)
        if incr < 1e-4:
            print("converged")
            return
        xk = x.copy()


if __name__ == "__main__":
    main()
