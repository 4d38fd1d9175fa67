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
