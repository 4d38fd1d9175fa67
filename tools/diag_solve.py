#!/usr/bin/env python
"""Diagnostic LVPP obstacle solve on one GPU: the Newton loop is driven from the host through the library's separate
entry points (residual + Jacobian assembly, Krylov solve, J*v) so that after EVERY linear solve the true residual
F - J y can be split into its u rows and psi rows, and the state of psi logged.  One line per Newton step on stderr,
one JSON line at the end.  Used to decide which residual norm the Krylov stopping test needs.

  python tools/diag_solve.py --size 215
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=215)
    ap.add_argument("--ksp-rtol", dest="ksp_rtol", type=float, default=1e-12)
    ap.add_argument("--ksp-max-it", dest="ksp_max_it", type=int, default=400)
    ap.add_argument("--snes-rtol", dest="snes_rtol", type=float, default=1e-6)
    ap.add_argument("--max-outer", dest="max_outer", type=int, default=30)
    ap.add_argument("--max-newton", dest="max_newton", type=int, default=30)
    ap.add_argument("--tol-exit", dest="tol_exit", type=float, default=1e-4)
    ap.add_argument("--alpha-max", dest="alpha_max", type=float, default=1e2)
    ap.add_argument("--alpha-scheme", dest="alpha_scheme", default="double_exponential")
    ap.add_argument("--lu", action="store_true", help="also solve every system with host sparse LU (small sizes only)")
    args = ap.parse_args()
    import numpy as np
    import torch

    import proximalgalerkin_b200 as lvpp
    from proximalgalerkin_b200.obstacle_pg import alpha_update

    n = args.size
    t0 = time.perf_counter()
    msh = lvpp.mesh.create_box(n, n, n)
    opts_d = {"ksp_rtol": args.ksp_rtol, "ksp_type": "gmres", "pc_type": "mg", "ksp_max_it": args.ksp_max_it,
              "snes_rtol": args.snes_rtol}
    s = lvpp.obstacle_pg.setup(msh, 1, petsc_options=opts_d)
    dev = s["problem"].device_problem
    opts = lvpp.newton_options(s["options"])
    N = dev.n
    mk = lambda: lvpp.DeviceVector(N, dev.device)  # noqa: E731
    x, xk, F, y, Jy = mk(), mk(), mk(), mk(), mk()
    nown = 2 * dev.stats()["num_owned"] if "num_owned" in dev.stats() else N
    torch.cuda.synchronize()
    print(f"setup {time.perf_counter() - t0:.1f} s, rows {N}", file=sys.stderr, flush=True)
    lu_pattern = dev.csr_pattern() if args.lu else None

    def split(v):
        t = v.tensor[:nown]
        return float(t[0::2].norm()), float(t[1::2].norm())

    alpha, alpha_k = 1.0, 1
    hist = []
    t1 = time.perf_counter()
    total_newton = total_krylov = 0
    for k in range(args.max_outer):
        alpha, alpha_k = alpha_update(args.alpha_scheme, k, alpha, alpha_k, args.alpha_max)
        dev.set_alpha(alpha)
        dev.set_previous(xk)
        fnorm0 = fnorm = dev.assemble_residual(x, F)
        its, reason = 0, 0
        while True:
            dev.assemble_jacobian(x)
            fu, fp = split(F)
            kits, kreason, krn = dev.linear_solve(F, y, opts)
            dev.spmv(y, Jy)
            Jy.tensor.sub_(F.tensor).neg_()  # r = F - J y
            ru, rp = split(Jy)
            yu, yp = split(y)
            extra = ""
            if args.lu:
                import scipy.sparse as sp
                import scipy.sparse.linalg as spl

                indptr, indices = lu_pattern
                A = sp.csr_matrix((dev.jacobian_values().cpu().numpy(), indices, indptr), shape=(N, N)).tocsc()
                ylu = spl.splu(A).solve(F.numpy())
                e = y.numpy() - ylu
                extra = (f" | vs LU: du {np.linalg.norm(e[0::2]) / max(np.linalg.norm(ylu[0::2]), 1e-300):.2e} "
                         f"dpsi {np.linalg.norm(e[1::2]) / max(np.linalg.norm(ylu[1::2]), 1e-300):.2e} "
                         f"max|dpsi| {np.abs(e[1::2]).max():.2e}")
            x.tensor.sub_(y.tensor)
            fnorm = dev.assemble_residual(x, F)
            its += 1
            total_newton += 1
            total_krylov += kits
            psi = x.tensor[1:nown:2]
            npos = int((psi > 1.0).sum())
            print(f"outer {k} alpha {alpha:.4g} newton {its}: |F| {fnorm:.3e} krylov {kits} ({kreason}) "
                  f"rhs u/psi {fu:.2e}/{fp:.2e} lin.res u/psi {ru:.2e}/{rp:.2e} |y| u/psi {yu:.2e}/{yp:.2e} "
                  f"psi [{float(psi.min()):.4g}, {float(psi.max()):.4g}] #psi>1: {npos}{extra}", file=sys.stderr, flush=True)
            if not np.isfinite(fnorm):
                reason = -4
            elif kreason < 0:
                reason = -3
            elif fnorm <= args.snes_rtol * fnorm0:
                reason = 3
            elif float(y.tensor[:nown].norm()) < 1e-8 * float(x.tensor[:nown].norm()):
                reason = 4
            elif fnorm > 1e4 * fnorm0:
                reason = -9
            elif its >= args.max_newton:
                reason = -5
            if reason:
                break
        obs = dev.observables(x)
        inc = float(np.sqrt(obs[4]))
        hist.append({"alpha": alpha, "newton": its, "reason": reason, "increment": inc})
        print(f"== outer {k}: alpha {alpha:.4g} newton {its} reason {reason} increment {inc:.3e}", file=sys.stderr, flush=True)
        if reason < 0 or inc < args.tol_exit:
            break
        xk.tensor.copy_(x.tensor)
    torch.cuda.synchronize()
    print(json.dumps({"size": n, "rows": N, "solve_s": time.perf_counter() - t1, "newton": total_newton,
                      "krylov": total_krylov, "history": hist,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("LVPP_")}}))


if __name__ == "__main__":
    main()
