#!/usr/bin/env python
"""Full LVPP obstacle solve (all proximal steps, the reference's CI parameters) at a given size on one
GPU or under torchrun; prints one JSON line with the iteration history, time and memory.

  python tools/full_solve.py --size 368            # 100.5 M rows on one B200
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
    ap.add_argument("--nz", type=int, default=None, help="cubes along z over all ranks (default: --size)")
    ap.add_argument("--tag", default="")
    ap.add_argument("--psi-free", dest="psi_free", type=float, default=None, help="lvpp_psi_free_below")
    ap.add_argument("--psi-cap", dest="psi_cap", type=float, default=None,
                    help="lvpp_psi_increase_max: bound on the growth of psi per Newton step (not in the reference)")
    ap.add_argument("--tol", type=float, default=1e-4, help="tol_exit of the outer loop (CI: 1e-4; script default 1e-6)")
    ap.add_argument("--alpha-max", dest="alpha_max", type=float, default=1e2)
    ap.add_argument("--pc", default="mg")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--snes-rtol", dest="snes_rtol", type=float, default=None,
                    help="override the reference's 1e-6 (the latent variable on the contact set is determined only to "
                         "||F_psi|| * O(h^-5): DESIGN.md 7a)")
    ap.add_argument("--linesearch", default="none", choices=["none", "bt"],
                    help="bt: PETSc's backtracking line search (the full step overshoots on fine 3-D meshes)")
    ap.add_argument("--alpha-scheme", dest="alpha_scheme", default="double_exponential",
                    choices=["double_exponential", "adaptive", "constant", "geometric"],
                    help="adaptive: alpha x2 / :2 by Newton count, halved and the step repeated when a Newton solve fails "
                         "(proximalgalerkin_b200/recovery.py; the reference's remedy in examples 03, 07, 08)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import proximalgalerkin_b200 as lvpp

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    n = args.size
    nz = args.nz if args.nz else world * max(2, round(n / world))
    t0 = time.perf_counter()
    msh = lvpp.mesh.create_box(n, n, nz, rank=rank, nranks=world)
    opts = {"ksp_rtol": 1e-12, "ksp_type": "gmres", "pc_type": "mg"} if args.pc == "mg" else {"ksp_rtol": 1e-12}
    opts["snes_linesearch_type"] = args.linesearch
    if args.psi_cap is not None:
        opts["lvpp_psi_increase_max"] = args.psi_cap
    if args.psi_free is not None:
        opts["lvpp_psi_free_below"] = args.psi_free
    if args.snes_rtol is not None:
        opts["snes_rtol"] = args.snes_rtol
    if args.alpha_scheme == "adaptive":  # a failed solve is reported as a reason, not raised
        opts.update({"snes_error_if_not_converged": False, "ksp_error_if_not_converged": False})
    st = lvpp.obstacle_pg.LvppStepper(msh, 1, args.alpha_scheme, args.alpha_max, args.tol, max_outer=100, petsc_options=opts)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    t1 = time.perf_counter()
    if args.verbose and rank == 0 and st.nb is not None:
        inner_bt = st.nb.step

        def logged_bt(x):
            r = inner_bt(x)
            print(f"outer {st.k} alpha {st.alpha_value:.4g} newton {st.nb.its}: |F| {st.nb.fnorm:.3e} lambda {st.nb.last_lambda:.3g} "
                  f"krylov {st.nb.linear_its} reason {r}", file=sys.stderr, flush=True)
            return r

        st.nb.step = logged_bt
    elif args.verbose and rank == 0:  # log every Newton step: norms, Krylov iterations and reason
        inner = st.dev.newton_step

        def logged(x, opts):
            out = inner(x, opts)
            (fnorm, ynorm, xnorm), kits, kreason = out
            print(f"outer {st.k} alpha {st.alpha_value:.4g} newton {st.newton_its + 1}: |F| {fnorm:.3e} |y| {ynorm:.3e} "
                  f"|x| {xnorm:.3e} krylov {kits} ({kreason})", file=sys.stderr, flush=True)
            return out

        st.dev.newton_step = logged
    failure = None
    try:
        while st.step():
            pass
    except lvpp.NotConvergedError as e:
        failure = {"outer": st.k, "alpha": st.alpha_value, "newton": st.newton_its, "reason": e.reason}
    torch.cuda.synchronize()
    t_solve = time.perf_counter() - t1
    s = st.dev.stats()
    if rank == 0:
        print(json.dumps({
            "workload": f"3-D P1 obstacle LVPP, {n}x{n}x{nz} cubes x 6 tets, full solve ({args.alpha_scheme} alpha, alpha_max 1e2, tol 1e-4)",
            "tag": args.tag, "psi_cap": args.psi_cap, "linesearch": args.linesearch, "snes_rtol": args.snes_rtol, "failure": failure,
            "n_gpus": world, "rows": s["num_rows"], "setup_s": t_setup, "solve_s": t_solve,
            "newton_steps": st.total_newton, "krylov_iterations": st.total_krylov, "outer_steps": len(st.history["newton_steps"]),
            "dofs_per_sec": s["num_rows"] * st.total_newton / t_solve, "history": st.history,
            "device_bytes": s["device_bytes"], "max_memory_allocated_torch": torch.cuda.max_memory_allocated(),
            "free_total_bytes": list(torch.cuda.mem_get_info()),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
