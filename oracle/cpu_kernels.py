"""ctypes binding of oracle/c/lvpp_cpu.c (C + OpenMP restatement of the reference's Jacobian assembly and MatMult on
its own data layout) -- test / baseline infrastructure, see oracle/__init__.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liblvpp_cpu.so"
_lib = None


def build():
    subprocess.run(["make", "-C", str(HERE / "c")], check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        _lib = C.CDLL(str(LIB))
        _lib.lvpp_cpu_num_threads.restype = C.c_int
        _lib.lvpp_cpu_assemble_jacobian.restype = None
        _lib.lvpp_cpu_csr_spmv.restype = None
        _lib.lvpp_cpu_minres.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def num_threads():
    return int(load().lvpp_cpu_num_threads())


class JacobianAssembler:
    """Holds the arrays of an oracle.obstacle.ObstacleOracle in the layout the C kernel reads."""

    def __init__(self, orc):
        self.orc = orc
        c = np.ascontiguousarray
        self.cell_nodes = c(orc.cell_nodes, dtype=np.int64)
        self.dof_u, self.dof_psi = c(orc.dof_u, dtype=np.int64), c(orc.dof_psi, dtype=np.int64)
        self.phi = c(orc.phi_tab, dtype=np.float64)        # [nq, nld]
        self.gphi = c(orc.gphi, dtype=np.float64)          # [C, nq, nld, gdim]
        self.scale, self.w = c(orc.scale, dtype=np.float64), c(orc.qwts, dtype=np.float64)
        self.is_bc = c(orc.is_bc, dtype=np.uint8)
        self.map = c(orc.cell_to_nnz, dtype=np.int64)
        self.diag = c(orc._diag_positions(), dtype=np.int64)
        self.bc_dofs = c(orc.bc_dofs, dtype=np.int64)
        self.vals = np.zeros(orc.nnz)

    def assemble(self, x, alpha):
        o, lib = self.orc, load()
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib.lvpp_cpu_assemble_jacobian(
            C.c_int64(self.cell_nodes.shape[0]), C.c_int(o.nld), C.c_int(self.w.size), C.c_int(self.gphi.shape[3]),
            _p(self.cell_nodes, C.c_int64), _p(self.dof_u, C.c_int64), _p(self.dof_psi, C.c_int64), _p(self.phi, C.c_double),
            _p(self.gphi, C.c_double), _p(self.scale, C.c_double), _p(self.w, C.c_double), _p(self.is_bc, C.c_uint8),
            _p(self.map, C.c_int64), _p(self.diag, C.c_int64), _p(self.bc_dofs, C.c_int64), C.c_int64(self.bc_dofs.size),
            _p(x, C.c_double), C.c_double(alpha), C.c_int64(o.nnz), _p(self.vals, C.c_double))
        return self.vals


def csr_spmv(indptr, indices, vals, x, y=None):
    n = indptr.size - 1
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    if y is None:
        y = np.empty(n)
    load().lvpp_cpu_csr_spmv(C.c_int64(n), _p(indptr, C.c_int64), _p(indices, C.c_int32), _p(vals, C.c_double), _p(x, C.c_double),
                             _p(y, C.c_double))
    return y


def time_kernels(n=40, reps=5):
    """Bounded CPU sample for bench.py's cpu_baseline: Jacobian assembly and MatMult of the 3-D P1 obstacle problem
    on an n^3 Kuhn mesh, all host threads.  Returns a dict (milliseconds, cells/s, GB/s of the CSR-algorithmic bytes
    of SURVEY.md 8d: 12 B per nonzero + 16 B per row + the row pointer)."""
    import time

    from . import mesh as omesh
    from . import obstacle as oobs

    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    rng = np.random.default_rng(0)
    x = 0.3 * rng.standard_normal(orc.num_rows)
    ja = JacobianAssembler(orc)
    ja.assemble(x, 1.0)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    for _ in range(reps):
        vals = ja.assemble(x, 1.0)
    t_asm = (time.perf_counter() - t0) / reps
    y = np.empty(orc.num_rows)
    csr_spmv(orc.indptr, orc.indices, vals, x, y)
    t0 = time.perf_counter()
    for _ in range(4 * reps):
        csr_spmv(orc.indptr, orc.indices, vals, x, y)
    t_spmv = (time.perf_counter() - t0) / (4 * reps)
    spmv_bytes = 12 * orc.nnz + 16 * orc.num_rows + 8 * (orc.num_rows + 1)
    return {
        "threads": num_threads(), "n": n, "rows": int(orc.num_rows), "cells": int(orc.cell_nodes.shape[0]), "nnz": int(orc.nnz),
        "jacobian_assembly_ms": 1e3 * t_asm, "cells_per_s": orc.cell_nodes.shape[0] / t_asm,
        "spmv_ms": 1e3 * t_spmv, "spmv_gbs": spmv_bytes / t_spmv / 1e9,
        "what": "oracle/c/lvpp_cpu.c (C + OpenMP): SNESProblem.J (element tensors + MatSetValues(ADD) on the monolithic CSR) "
                "and MatMult, same data layout as the reference",
    }


def minres(indptr, indices, vals, pinv, rhs, rtol=1e-12, maxit=200000):
    """Diagonally preconditioned MINRES on a CSR matrix (oracle/c/lvpp_cpu.c: lvpp_cpu_minres).  Returns (y, its, rnorm)."""
    n = indptr.size - 1
    c = np.ascontiguousarray
    indptr, indices = c(indptr, dtype=np.int64), c(indices, dtype=np.int32)
    vals, pinv, rhs = c(vals, dtype=np.float64), c(pinv, dtype=np.float64), c(rhs, dtype=np.float64)
    y, work, rn = np.empty(n), np.empty(7 * n), C.c_double()
    its = load().lvpp_cpu_minres(C.c_int64(n), _p(indptr, C.c_int64), _p(indices, C.c_int32), _p(vals, C.c_double),
                                 _p(pinv, C.c_double), _p(rhs, C.c_double), _p(y, C.c_double), C.c_double(rtol),
                                 C.c_int64(maxit), _p(work, C.c_double), C.byref(rn))
    return y, int(its), rn.value


class BlockJacobiDiagonal:
    """The positive diagonal preconditioner of the product's pc_type jacobi path (csrc/krylov.cu:k_build_pinv; recipe of
    examples/09_eikonal/ex40.cpp:261-274): 1 / (alpha K_ii) on the u rows (1 on Dirichlet rows),
    1 / (D_ii + M_ii^2 / (alpha K_ii)) on the psi rows, read from the assembled CSR values through positions found once."""

    def __init__(self, orc):
        import scipy.sparse as sp

        self.orc = orc
        n = orc.num_rows
        pos = sp.csr_matrix((np.arange(1, orc.nnz + 1, dtype=np.float64), orc.indices, orc.indptr), shape=(n, n))
        self.p_uu = np.asarray(pos[orc.dof_u, orc.dof_u]).ravel().astype(np.int64) - 1
        self.p_pp = np.asarray(pos[orc.dof_psi, orc.dof_psi]).ravel().astype(np.int64) - 1
        self.p_up = np.asarray(pos[orc.dof_u, orc.dof_psi]).ravel().astype(np.int64) - 1  # M_ii sits in the u row (psi has no bc)

    def __call__(self, vals):
        o = self.orc
        a, d, m = vals[self.p_uu], -vals[self.p_pp], vals[self.p_up]
        pinv = np.empty(o.num_rows)
        pinv[o.dof_u] = 1.0 / a
        pinv[o.dof_psi] = 1.0 / (d + m * m / a)
        return pinv


def time_newton_steps_krylov(n=40, steps=3, rtol=1e-12):
    """Bounded CPU sample of the like-for-like path: Newton steps of the LVPP obstacle solve (first
    proximal step, alpha = 1, from the zero iterate) with the Jacobian assembled by the C + OpenMP kernel and the
    Newton system solved by diagonally preconditioned MINRES on all host threads.  Returns a dict."""
    import time

    from . import mesh as omesh
    from . import obstacle as oobs

    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n))
    ja = JacobianAssembler(orc)
    pc = BlockJacobiDiagonal(orc)
    x = np.zeros(orc.num_rows)
    xk = x.copy()
    F = orc.assemble_residual(x, xk, 1.0)
    ja.assemble(x, 1.0)
    its_all, t_asm, t_kry, t_res = [], 0.0, 0.0, 0.0
    for _ in range(steps):
        ta = time.perf_counter()
        vals = ja.assemble(x, 1.0)
        tb = time.perf_counter()
        pinv = pc(vals)
        y, its, _ = minres(orc.indptr, orc.indices, vals, pinv, F, rtol=rtol)
        x = x - y
        tc = time.perf_counter()
        F = orc.assemble_residual(x, xk, 1.0)
        td = time.perf_counter()
        its_all.append(its)
        t_asm += tb - ta
        t_kry += tc - tb
        t_res += td - tc
    secs = t_asm + t_kry  # the numpy residual assembly is a slow stand-in for a compiled cell loop: reported, not charged
    return {"threads": num_threads(), "n": n, "rows": int(orc.num_rows), "newton_steps": steps, "seconds": secs,
            "value": orc.num_rows * steps / secs, "unit": "DOFs/s", "minres_iterations": its_all,
            "jacobian_assembly_s": t_asm, "krylov_s": t_kry, "residual_numpy_s_not_charged": t_res,
            "residual_norm": float(np.linalg.norm(F)),
            "what": "oracle/c/lvpp_cpu.c: C + OpenMP Jacobian assembly + diagonally preconditioned MINRES (rtol 1e-12) on the "
                    "reference's monolithic CSR, all host threads -- the Krylov path the product's pc_type jacobi runs; the "
                    "(numpy) residual assembly is timed separately and not charged"}
