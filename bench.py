#!/usr/bin/env python
"""LVPP Newton-step throughput on B200 (BASELINE.json metric) -- see DESIGN.md section 7 "Measurement".

A "step" is one Newton step of the LVPP obstacle solve (Krylov solve of the saddle-point system,
fused update + norms, residual and Jacobian assembly at the new iterate; outer-loop work --
observables, alpha update, sol_k <- sol -- is inside the timed region when it falls between steps).
Workload at N GPUs: the 3-D P1 obstacle problem on [-1,1]^2 x [-N,N], n x n x (n N) cubes x 6 Kuhn
tetrahedra, one z-slab, one copy of the obstacle and one clamped box per GPU (weak scaling; n = 215 ->
20.2 M rows per GPU), started from the zero iterate with the reference script's default parameters
(obstacle_pg.py:291-321: constant alpha, tol 1e-6, full Newton step, SNES rtol 1e-6).  Warm-up steps are
the first W Newton steps of that solve; the K timed steps continue it and start a fresh solve when it ends,
so that exactly K steps are timed whatever K is.  Defaults: --steps 20 --warmup 5.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 215] [--impl b200|reference]
                  [--workload obstacle|obstacle2d|gradient|multiphase|signorini]

--impl reference times the CPU restatement of the reference algorithm (oracle/: numpy assembly +
SuperLU in place of dolfinx + MUMPS; the real stack is not installable here) on a bounded sample of the
same workload, and reports the C + OpenMP Krylov path beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# The LVPP parameters are the defaults of the reference script (obstacle_pg.py:291-321: --alpha-scheme constant,
# --alpha-max 1e5, --tol 1e-6, --max-iter 100).  The CI parameters of compare_all.py:80-87 (double_exponential, alpha_max
# 1e2, tol 1e-4) complete on the 3-D meshes up to ~128 cubes per axis; from 160 on the first FULL Newton step
# (snes_linesearch_type none, obstacle_pg.py:136) of the third proximal step (alpha 1 -> 1.49) overshoots psi next to the
# contact boundary by +20 and the exponential takes over -- with linear solves exact to 1e-15 in both row blocks and under
# two unrelated Krylov solvers (profiles/r02_newton_robustness.md).  Run them with --alpha-scheme double_exponential
# --alpha-max 1e2 --tol 1e-4.
SCHEDULE_NOTE = ("reference script defaults (obstacle_pg.py:291-321); the CI schedule of compare_all.py:80-87 diverges in exact "
                 "arithmetic on 3-D meshes of >= 160 cubes per axis with the full Newton step: profiles/r02_newton_robustness.md")
METRIC = "lvpp_newton_dofs_per_sec"
UNIT = "DOFs/s"  # rows of the mixed Newton system x Newton steps / second


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_newton_steps(n_cpu, steps, warmup, alpha_scheme="constant", alpha_max=1e5, tol_exit=1e-6, dim=3):
    """Times `steps` Newton steps of the oracle's LVPP solve (after `warmup`) on an n_cpu^3 mesh.
    Returns (DOFs/s, seconds, rows, steps_done)."""
    import numpy as np

    from oracle import lvpp_driver, mesh as omesh, obstacle as oobs, snes as osnes

    orc = oobs.ObstacleOracle(omesh.box_kuhn(n_cpu, n_cpu, n_cpu) if dim == 3 else omesh.rectangle(n_cpu, n_cpu))
    import scipy.sparse.linalg as spla

    x = np.zeros(orc.num_rows)
    xk = x.copy()
    alpha_k, alpha, k = 1, 1.0, 0
    done, t_timed, t0 = 0, 0.0, None
    total = steps + warmup
    while done < total:
        alpha, alpha_k = lvpp_driver.alpha_schedule(alpha_scheme, k, alpha_k, alpha_max, alpha_current=alpha)
        # one Newton step at a time so that exactly `steps` are timed
        F = orc.assemble_residual(x, xk, alpha)
        fnorm0 = np.linalg.norm(F)
        it = 0
        while done < total:
            if done == warmup and t0 is None:
                t0 = time.perf_counter()
            y = spla.splu(orc.jacobian(x, alpha).tocsc()).solve(F)
            x = x - y
            F = orc.assemble_residual(x, xk, alpha)
            fnorm = np.linalg.norm(F)
            it += 1
            done += 1
            if osnes.converged_default(it, np.linalg.norm(x), np.linalg.norm(y), fnorm, fnorm0 * 1e-6, fnorm0, 1e-50, 1e-8, 1e4):
                break
        obs = orc.observables(x, xk, alpha)
        if np.sqrt(obs[4]) < tol_exit or k + 1 >= 100:  # solve finished: the next step starts a fresh solve, as on the GPU arm
            x = np.zeros(orc.num_rows)
            xk = x.copy()
            alpha_k, alpha, k = 1, 1.0, 0
            continue
        xk = x.copy()
        k += 1
    t_timed = time.perf_counter() - t0 if t0 is not None else float("nan")
    nsteps = done - warmup
    return orc.num_rows * nsteps / t_timed, t_timed, orc.num_rows, nsteps


def blas_threads():
    """Threads of the BLAS behind scipy's SuperLU (its supernodal kernels are the only threaded part of the port)."""
    try:
        from threadpoolctl import threadpool_info

        return max([int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return 1


def workload_mesh(args, world):
    """(dim, n, nxy, nz, lo, hi, slabs) of the N-GPU workload -- shared by both arms."""
    if args.workload == "obstacle2d":
        # configs[0]: the 2-D obstacle problem of examples/01 on the unit square mapped to [-1,1]^2, N x N squares with
        # the right diagonal (N = 1000: 2 004 002 rows); N GPUs: N sqrt(world) squares per axis, y-slabs
        n = args.n2d
        nxy = int(round(n * world ** 0.5))
        return 2, n, nxy, world * max(1, int(round(nxy / world))), None, None, 1
    nxy, nz, lo, hi, slabs = weak_scaling_mesh(args.n, world, args.weak, args.slabs)
    return 3, args.n, nxy, nz, lo, hi, slabs


def workload_config(args, world):
    """The part of ``config`` that names the workload: identical in the b200 arm and in the reference arm (which times a
    bounded sample of it, described in its ``cpu_baseline.sample``)."""
    dim, n, nxy, nz, lo, hi, slabs = workload_mesh(args, world)
    rows = 2 * (nxy + 1) ** (dim - 1) * (nz + 1)
    return {
        "workload": (f"3-D P1 obstacle LVPP (configs[1]): {nxy}x{nxy}x{nz} cubes x 6 tets, " if dim == 3 else
                     f"2-D P1 obstacle LVPP (configs[0], examples/01): {nxy}x{nz} squares x 2 triangles on [-1,1]^2, ") + f"{rows} rows",
        "n": n, "rows": rows, "primal_dofs": rows // 2,  # the reference's CSV column "dofs" (obstacle_pg.py:237,255)
        "alpha_scheme": args.alpha_scheme, "alpha_max": args.alpha_max, "tol_exit": args.tol_exit, "max_outer": 100,
        "snes_linesearch_type": "none", "snes_rtol": 1e-6,
        "parallelism": f"slab{world}", "weak_scaling": (None if world == 1 else args.weak),
        "obstacle": "phi_set of obstacle_pg.py:92-104" + (", one copy per slab (period 2 in z)" if slabs > 1 else "") +
                    (", u = 0 on the planes between the slabs" if slabs > 1 and args.weak == "stack" else ""),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dim = 2 if args.workload == "obstacle2d" else 3
    n_cpu = args.n_cpu if dim == 3 else max(args.n_cpu, 150)
    val, secs, rows, nsteps = cpu_newton_steps(n_cpu, args.steps, args.warmup, args.alpha_scheme, args.alpha_max, args.tol_exit, dim)
    sample = (f"{nsteps} Newton steps of the same LVPP obstacle solve on a {n_cpu}^{dim} " + ("cube Kuhn" if dim == 3 else "square right-diagonal") + f" mesh ({rows} rows), "
              f"numpy assembly + scipy SuperLU (stand-in for dolfinx + MUMPS; sequential factorisation, {blas_threads()} BLAS threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": nsteps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(nsteps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the b200 arm's workload; this arm times a bounded sample of it (cpu_baseline.sample) because sparse LU of the
        # full size is out of reach of any host (3-D, 20 M rows)
        "config": dict(workload_config(args, int(os.environ.get("WORLD_SIZE", str(max(args.gpus, 1))))),
                       ksp="sparse LU (SuperLU), the stand-in for ksp_type preonly / pc_type lu / MUMPS (obstacle_pg.py:129-131)",
                       sample_rows=rows, sample_n=n_cpu),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": blas_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "newton_steps_per_sec": nsteps / secs,
    }
    if not args.no_aux:
        # beside the direct-solve port (the reference's algorithm, sequential factorisation): the Krylov path on all host
        # threads (C + OpenMP assembly + preconditioned MINRES, oracle/c) -- what a CPU user without MUMPS would run
        try:
            from oracle import cpu_kernels

            line["cpu_baseline"]["krylov_path"] = cpu_kernels.time_newton_steps_krylov(48 if dim == 3 else 24, 3)
        except Exception as e:  # optional infrastructure
            line["cpu_baseline"]["krylov_path"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line))


def weak_scaling_mesh(n, world, weak="refine", slabs_1gpu=1):
    """Cubes per axis and box of the N-GPU workload: (nxy, nz, lo, hi, slabs).  See the comment in run_b200."""
    slabs = max(1, slabs_1gpu) if world == 1 else (world if weak in ("stack", "stack-open") else 1)
    if world > 1 and weak == "refine":
        nxy = int(round(n * world ** (1.0 / 3.0)))
        # an even number of planes per rank: every rank aggregates from its own first plane, so an odd count leaves a
        # one-plane layer of thin aggregates at every slab boundary already on the first coarsening
        # (tools/mg_prototype.py --slabs: 16 -> 18 Krylov iterations with 6-7 planes per slab, 16 -> 16 with 8)
        nz = 2 * world * max(1, int(round(nxy / (2.0 * world))))
        return nxy, nz, (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), slabs
    return n, n * slabs, (-1.0, -1.0, -float(slabs)), (1.0, 1.0, float(slabs)), slabs


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import proximalgalerkin_b200 as lvpp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.n
    t_setup = time.perf_counter()
    # Weak scaling (N > 1).  "stack" (default, SURVEY 8d "weak-scaled by stacking slabs in z"): N copies of the n^3
    # problem stacked along z on [-1,1]^2 x [-N,N], one z-slab and one copy of the obstacle per GPU (a single obstacle
    # would leave the far slabs at phi = -16, where the first Newton step from psi = 0 overshoots past PETSc's divergence
    # tolerance).  The mesh width stays 2 / n, so the Newton iteration behaves as on one GPU -- the full Newton step
    # of the reference stops converging on finer meshes (profiles/r02_newton_robustness.md), which rules out
    # "refine": one obstacle on [-1,1]^3 on a mesh refined so that every GPU keeps ~n^3 cubes (nx = ny =
    # round(n N^(1/3)), nz = the multiple of 2N nearest to nx, N z-slabs).  --slabs S emulates the S-GPU "stack"
    # problem on one GPU (diagnostic).
    dim, n, nxy, nz, lo, hi, slabs = workload_mesh(args, world)
    if dim == 2:
        msh = lvpp.mesh.create_rectangle(nxy, nz, rank=rank, nranks=world)
    else:
        msh = lvpp.mesh.create_box(nxy, nxy, nz, lo=lo, hi=hi, rank=rank, nranks=world,
                                   clamp_every=n if (slabs > 1 and args.weak != "stack-open") else None)
    opts = {"ksp_rtol": args.ksp_rtol, "ksp_max_it": 200000}
    if args.pc == "mg":
        opts = {"ksp_rtol": args.ksp_rtol, "ksp_type": "gmres", "pc_type": "mg", "ksp_max_it": 400}
    if args.psi_cap is not None:  # the safeguard that is not in the reference; off for the headline line
        opts["lvpp_psi_increase_max"] = args.psi_cap
        opts["ksp_max_it"] = 4000
        if args.psi_free is not None:
            opts["lvpp_psi_free_below"] = args.psi_free
    st = lvpp.obstacle_pg.LvppStepper(msh, 1, args.alpha_scheme, args.alpha_max, args.tol_exit, max_outer=100, petsc_options=opts,
                                      obstacle_period=2.0 if slabs > 1 else None, obstacle_origin=-float(slabs))
    dev = st.dev
    stats0 = dev.stats()
    t_setup = time.perf_counter() - t_setup
    rows_global = stats0["num_rows"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # A step never raises out of the timed region and the region always holds exactly K steps: when the LVPP solve
    # finishes (about 20 Newton steps at the CI parameters) or a Newton solve fails, the next step starts a fresh solve
    # from the zero iterate on the same handle (LvppStepper.reset: zero x and x_k, alpha schedule from k = 0, residual +
    # Jacobian at the zero iterate -- all inside the timed region).
    solves, failures = [], []

    def advance():
        try:
            alive = st.step()
        except lvpp.NotConvergedError as e:
            failures.append({"solve": len(solves), "outer": st.k, "newton": st.newton_its, "reason": e.reason})
            alive = False
        if not alive:
            solves.append({"newton_steps": list(st.history["newton_steps"]), "krylov_iterations": list(st.history["krylov_iterations"]),
                           "alpha": list(st.history["alpha"]), "primal_increment": list(st.history["primal_increment"]),
                           "converged": bool(st.finished)})
            st.reset()

    # ---- warm-up: the first W Newton steps of the solve
    for _ in range(args.warmup):
        advance()
    # ---- timed region: exactly K Newton steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    s0 = dev.stats()
    barrier()
    dev.timer_start()
    t0 = time.perf_counter()
    done = 0
    while done < args.steps:
        advance()
        done += 1
    dev_ms = dev.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    s1 = dev.stats()
    t = torch.tensor([dev_ms / 1e3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs, wall = float(t[0]), float(t[1])
    value = rows_global * done / secs
    kry = s1["krylov_iterations"] - s0["krylov_iterations"]
    launches = s1["kernel_launches"] - s0["kernel_launches"]

    # ---- rooflines: launches sampled with CUDA events (on the library's stream) inside the timed solves
    V_own = stats0["local_rows"] // 2
    slots = stats0["sell_slots"]
    peak, peak_src = measured_peak()

    def traffic_of(fname):
        tf = ROOT / "profiles" / fname
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                if int(tj.get("n", -1)) == n:
                    return tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        return None

    # (1) fp64 J*v (k_block_op<0>): compulsory bytes in the stored format: per slot col(4) + K,M,D (24);
    # per node: gathered v (16; the own entry is one of them), y (16), bc flag (1), slice ptr (8/32)
    n_s = s1["spmv_samples"] - s0["spmv_samples"]
    spmv_ms = (s1["spmv_sampled_ms"] - s0["spmv_sampled_ms"]) / max(n_s, 1)
    spmv_bytes = 28 * slots + (16 + 16 + 1 + 0.25) * V_own
    fine_ops = s1["fine_op_launches"] - s0["fine_op_launches"]
    packed_ops = s1["packed_op_launches"] - s0["packed_op_launches"]
    roof_jv = {
        "bound": "hbm", "kernel": "k_block_op<0> (fp64 J*v and Krylov residual, 2x2-block sliced-ELL)",
        "achieved": spmv_bytes / (spmv_ms * 1e-3) / 1e9 if n_s else None, "peak": peak, "unit": "GB/s",
        "frac": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / peak if n_s else None, "traffic": traffic_of("spmv_traffic.json"),
        "peak_source": peak_src, "frac_of_nominal_8tbs": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / 8000.0 if n_s else None,
        "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms, "launches_sampled": n_s,
        "csr_equivalent_bytes": 12 * stats0["nnz"] + 16 * stats0["local_rows"] + 8 * (stats0["local_rows"] + 1),
        "launches_in_timed_region": fine_ops - packed_ops,
        "share_of_step": spmv_ms * (fine_ops - packed_ops) / (secs * 1e3) if secs > 0 else None,
    }
    # (2) the multigrid cycle's fine-level sweep (k_packed_op, the dominant kernel with pc mg): 16-byte record per slot;
    # per node: gathered v (16), b (16), node-block inverse (32), y (16), bc flag (1), slice ptr (8/32)
    n_p = s1["smooth_samples"] - s0["smooth_samples"]
    roofline = roof_jv
    roof_extra = None
    if n_p:
        sm_ms = (s1["smooth_sampled_ms"] - s0["smooth_sampled_ms"]) / n_p
        # (bf16 pair records, 20 bytes per pair of slots, kernel k_packed2_op; LVPP_MG_PACK=fp32: 16-byte records, k_packed_op)
        rec = 10 if os.environ.get("LVPP_MG_PACK", "bf16") != "fp32" and os.environ.get("LVPP_MG_FP32", "1") != "0" else 16
        sm_bytes = rec * slots + (16 + 16 + (16 if rec == 10 else 32) + 16 + 1 + 0.25) * V_own
        roofline = {
            "bound": "hbm", "kernel": ("k_packed2_op (multigrid smoother sweep on the fine level: bf16 pair records, 10 B / slot, " if rec == 10 else
                                       "k_packed_op (multigrid smoother sweep on the fine level: packed {col, alpha K, M, D} "
                                       "single-precision records, ") + "fp64 accumulation, fused node-block Jacobi update)",
            "achieved": sm_bytes / (sm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": sm_bytes / (sm_ms * 1e-3) / 1e9 / peak,
            "traffic": traffic_of("smooth_traffic.json" if rec == 16 else "smooth_traffic_bf16.json"), "peak_source": peak_src, "frac_of_nominal_8tbs": sm_bytes / (sm_ms * 1e-3) / 1e9 / 8000.0,
            "algorithmic_bytes_per_launch": sm_bytes,
            "ms_per_launch": sm_ms, "launches_sampled": n_p, "launches_in_timed_region": packed_ops,
            "share_of_step": sm_ms * packed_ops / (secs * 1e3) if secs > 0 else None,
        }
        roof_extra = roof_jv

    # ---- e2e: the same solve through NonlinearProblem.solve() on host buffers (H2D of sol and sol_k,
    # D2H of sol inside the timed region), whole outer iterations until >= K Newton steps
    e2e = None
    if not args.no_e2e:
        s = st.s
        sol, sol_k, alpha, problem = s["sol"], s["sol_k"], s["alpha"], s["problem"]
        sol.x.array[:] = 0.0
        sol_k.x.array[:] = 0.0
        alpha.value, alpha_k, k, steps_e2e = 1.0, 1, 0, 0
        n_solves = 0
        barrier()
        te = time.perf_counter()
        e2e_fail = []
        while steps_e2e < args.steps:
            alpha.value, alpha_k = lvpp.obstacle_pg.alpha_update(args.alpha_scheme, k, alpha.value, alpha_k, args.alpha_max)
            ok = True
            try:
                problem.solve()
            except lvpp.NotConvergedError as e:
                e2e_fail.append({"outer": k, "reason": e.reason})
                ok = False
            steps_e2e += max(problem.solver.getIterationNumber(), 1)
            n_solves += 1
            if ok:
                dev.x.set(sol.x.array)
                obs = dev.observables(dev.x)
            if not ok or np.sqrt(obs[4]) < args.tol_exit or k + 1 >= 100:  # solve finished (or failed): the next proximal step starts a fresh solve
                sol.x.array[:] = 0.0
                sol_k.x.array[:] = 0.0
                alpha.value, alpha_k, k = 1.0, 1, 0
                continue
            sol_k.x.array[:] = sol.x.array[:]
            k += 1
        barrier()
        te = time.perf_counter() - te
        t = torch.tensor([te], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t[0])
        vec_bytes = 8 * dev.n
        e2e = {
            "value": rows_global * steps_e2e / te, "unit": UNIT, "steps": steps_e2e, "seconds": te,
            "h2d_bytes_per_step": 3 * vec_bytes * n_solves / max(steps_e2e, 1),  # sol, sol_k (+ sol for observables)
            "d2h_bytes_per_step": vec_bytes * n_solves / max(steps_e2e, 1),
            "api": "NonlinearProblem.solve() per proximal step, host numpy buffers", "failures": e2e_fail,
        }

    # ---- assembly kernels (reported, not the headline)
    asm = None
    if rank == 0 and not args.no_aux:
        F = lvpp.DeviceVector(dev.n, dev.device)
        a, b, c = dev.time_assembly(dev.x, F, reps=3)
        asm = {"cell_exp_ms": a, "row_gather_ms": b, "residual_ms": c}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.n_cpu if dim == 3 else max(args.n_cpu, 150)
        val, secs_c, rows_c, ns = cpu_newton_steps(n_cpu, 5, 1, args.alpha_scheme, args.alpha_max, args.tol_exit, dim)
        cpu = {"value": val, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": f"{ns} Newton steps of the same LVPP solve on a {n_cpu}^{dim} structured mesh ({rows_c} rows), "
                         f"numpy assembly + SuperLU (oracle/; sequential factorisation, threaded BLAS), {secs_c:.1f} s; "
                         f"host has {os.cpu_count()} cores"}
        if not args.no_aux:
            # like-for-like kernels on all host threads: the C + OpenMP restatement of SNESProblem.J and MatMult on the
            # reference's monolithic CSR (GPU counterparts: `assembly` and `roofline_jv` of this line)
            try:
                from oracle import cpu_kernels

                cpu["kernels"] = cpu_kernels.time_kernels(40)
                # ... and the whole Krylov path: Newton steps with C + OpenMP assembly and
                # diagonally preconditioned MINRES on the monolithic CSR, all host threads (GPU counterpart: --pc jacobi)
                cpu["krylov_path"] = cpu_kernels.time_newton_steps_krylov(48 if dim == 3 else 24, 3)
            except Exception as e:  # the baseline library is optional infrastructure
                cpu["kernels"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": done, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(done, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(
                workload_config(args, world),
                nnz_per_gpu=stats0["nnz"],  # CSR-equivalent
                schedule_note=SCHEDULE_NOTE, psi_increase_max=args.psi_cap, psi_free_below=args.psi_free,
                ksp=("MINRES + block-Jacobi/Schur-diag" if args.pc == "jacobi" else
                     "GMRES(50) + monolithic aggregation multigrid V(2,3) (node-block Jacobi sweeps with Chebyshev-root "
                     "dampings, ratio 6; cycle operator from packed bf16 pair records with fp64 accumulation, fp64 Krylov operator)"),
                ksp_rtol=args.ksp_rtol,
                l2=("operator (>=4 GB) and vectors exceed the 126 MB L2; no flush needed" if (dim == 3 and n >= 100) else
                    ("operator (0.6 GB at N = 1000) and Krylov basis exceed the 126 MB L2; no flush" if dim == 2 and n >= 700 else
                     "inputs fit L2: kernel-level numbers are L2-warm"))),
            "newton_steps_per_sec": done / secs, "krylov_iterations": kry, "vcycles": s1["vcycles"] - s0["vcycles"],
            "mg_levels": s1["mg_levels"], "wall_s": wall, "setup_s": t_setup,
            "roofline": roofline, "roofline_jv": roof_extra, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "assembly": asm, "device_bytes": stats0["device_bytes"],
            "solves_completed": solves, "solve_in_progress": {k: list(v) for k, v in st.history.items()}, "failures": failures,
            "env": {k: v for k, v in sorted(os.environ.items()) if k.startswith("LVPP_")},  # experimental switches in force
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_forms(args):
    """Secondary measurement (not the driver's line): the mixed-form engine on one GPU -- element kernels +
    gathers (assembly), CSR SpMV and the first Newton solves of the example at a moderate size."""
    import numpy as np
    import torch

    import proximalgalerkin_b200 as lvpp

    torch.cuda.set_device(0)
    name, n = args.workload, args.n
    t0 = time.perf_counter()
    opts = {"ksp_gmres_restart": 100, "ksp_rtol": 1e-10, "snes_error_if_not_converged": False, "snes_max_it": 2,
            "ksp_max_it": 40000, "ksp_error_if_not_converged": False}
    if name == "gradient":
        s = lvpp.gradient_constraints.setup(n, n, petsc_options=opts)
        desc = f"examples/06 gradient constraint, unit square {n}x{n}, u in P2, psi in (P1)^2, quadrature degree 10"
    elif name == "multiphase":
        s = lvpp.multiphase.setup(n, n, petsc_options=opts)
        up = lvpp.multiphase.initial_condition(s["mesh"].coords, s["mesh"].cells)
        a1 = np.zeros_like(s["sol"])
        a1.reshape(-1, 3, 4)[:, 0, :] = up
        s["dev"].set_aux(1, a1)
        psi0 = np.log(1e-7) + 1.0
        s["sol"].reshape(-1, 3, 4)[:, 2, :] = psi0
        lo = np.zeros_like(s["sol"])
        lo.reshape(-1, 3, 4)[:, 2, :] = psi0
        s["dev"].set_aux(0, lo)
        desc = f"examples/04 multiphase, crossed unit square {n}x{n}, (u, z, psi) in (P1)^4 each"
    else:
        msh = lvpp.mesh.create_box(n, n, n, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0))
        s = lvpp.signorini.setup(msh, disp=-0.1, alpha_0=0.01, petsc_options=opts)
        desc = f"examples/02 Signorini, unit cube {n}^3 x 6 tets, u in (P1)^3, psi in P1 on the contact facets"
    setup_s = time.perf_counter() - t0
    dev, problem = s["dev"], s["problem"]
    X = dev.vector(s["sol"])
    F = dev.vector()
    dev.assemble_residual(X, F)
    ms_asm, ms_spmv = dev.time_kernels(X, reps=10)
    st0 = dev.stats()
    t1 = time.perf_counter()
    problem.solve()
    torch.cuda.synchronize()
    solve_s = time.perf_counter() - t1
    st1 = dev.stats()
    newton = st1["newton_steps"] - st0["newton_steps"]
    kry = st1["krylov_iterations"] - st0["krylov_iterations"]
    peak, peak_src = measured_peak()
    nnz, rows = st1["nnz"], st1["num_rows"]
    spmv_bytes = 12 * nnz + 16 * rows + 8 * (rows + 1)  # values + columns, x and y, row pointer
    line = {
        "metric": METRIC, "value": rows * newton / solve_s if newton else None, "unit": UNIT, "n_gpus": 1, "steps": newton,
        "dtype": "f64", "data": "synthetic", "higher_is_better": True,
        "config": {"workload": desc, "rows": rows, "nnz": nnz, "solver": "GMRES(100) + dof-block Jacobi, ksp_rtol 1e-10, 2 Newton steps",
                   "l2": "inputs fit L2: kernel-level numbers are L2-warm" if spmv_bytes < 120e6 else "operator exceeds L2"},
        "krylov_iterations": kry, "ms_per_step": 1e3 * solve_s / max(newton, 1), "setup_s": setup_s,
        "kernels": {"assembly_ms": ms_asm, "spmv_ms": ms_spmv},
        "roofline": {"bound": "hbm", "kernel": "k_csr_spmv (fp64 CSR)", "achieved": spmv_bytes / (ms_spmv * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": spmv_bytes / (ms_spmv * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": spmv_bytes},
        "gpu_launches": st1["kernel_launches"] - st0["kernel_launches"], "device_bytes": st1["device_bytes"],
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="timed Newton steps (default 20: with --warmup 5, steps 6-25 of a whole solve)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", dest="n", type=int, default=215, help="cubes per axis per GPU")
    ap.add_argument("--cpu-size", dest="n_cpu", type=int, default=20,
                    help="cubes per axis of the CPU sample (20: 18 522 rows, ~10 s for 6 Newton steps with sparse LU)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ksp-rtol", dest="ksp_rtol", type=float, default=1e-12)
    ap.add_argument("--pc", default="mg", choices=["jacobi", "mg"],
                    help="jacobi: block-diagonal MINRES; mg: multigrid-preconditioned GMRES")
    ap.add_argument("--workload", default="obstacle", choices=["obstacle", "obstacle2d", "gradient", "multiphase", "signorini"],
                    help="obstacle (the driver's line, configs[1]), obstacle2d (configs[0]: the 2-D problem of examples/01, "
                         "--size2d squares per axis) or one of the mixed-form examples (1 GPU, --size = N)")
    ap.add_argument("--size2d", dest="n2d", type=int, default=1000, help="obstacle2d: squares per axis (1000: 2 004 002 rows)")
    ap.add_argument("--alpha-scheme", dest="alpha_scheme", default="constant", choices=["constant", "double_exponential", "geometric"],
                    help="obstacle_pg.py --alpha-scheme (its default: constant)")
    ap.add_argument("--alpha-max", dest="alpha_max", type=float, default=1e5, help="obstacle_pg.py --alpha-max (default 1e5)")
    ap.add_argument("--tol", dest="tol_exit", type=float, default=1e-6, help="obstacle_pg.py --tol (default 1e-6)")
    ap.add_argument("--psi-cap", dest="psi_cap", type=float, default=None,
                    help="lvpp_psi_increase_max (NOT in the reference; off by default): bound on the growth of psi per Newton step")
    ap.add_argument("--psi-free", dest="psi_free", type=float, default=None, help="lvpp_psi_free_below (with --psi-cap)")
    ap.add_argument("--weak", default="stack", choices=["refine", "stack", "stack-open"],
                    help="N > 1: stack N copies of the n^3 problem along z (default; SURVEY 8d: 215 x 215 x 1720 cubes at 8 GPUs, "
                         "the mesh width -- and with it the behaviour of the full Newton step -- stays that of one GPU), every "
                         "copy clamped (u = 0) on all six faces like the single box; stack-open: no condition on the "
                         "planes between the copies; refine: the mesh of the one-obstacle problem refined")
    ap.add_argument("--slabs", type=int, default=1, help="1 GPU only: solve the global problem of an S-GPU run (diagnostic)")
    ap.add_argument("--skip-e2e", dest="no_e2e", action="store_true")
    ap.add_argument("--skip-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--skip-aux", dest="no_aux", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload not in ("obstacle", "obstacle2d"):
        run_forms(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
