"""proximalgalerkin_b200 -- the LVPP (proximal Galerkin) Newton inner loop on B200.

Drop-in for the path the reference's FEniCSx examples hand to dolfinx/PETSc, behind the reference's
own API surface (src/lvpp/__init__.py:1-9 exports ``SNESProblem`` and ``SNESSolver``; the examples
use the ``NonlinearProblem(...).solve()`` call shape).  Host code is Python; all arithmetic runs in
hand-written sm_100a CUDA kernels reached through the C ABI of ``include/lvpp_b200.h``.
"""
from . import fem, forms, gradient_constraints, io, mesh, multiphase, obstacle_pg, quadrature, signorini
from .problem import (
    DeviceMatrix,
    DeviceProblem,
    DeviceVector,
    NonlinearProblem,
    NotConvergedError,
    SNESProblem,
    SNESSolver,
    derivative,
    newton_options,
    obstacle_residual,
)

__all__ = [
    "SNESProblem",
    "SNESSolver",
    "NonlinearProblem",
    "NotConvergedError",
    "DeviceProblem",
    "DeviceVector",
    "DeviceMatrix",
    "obstacle_residual",
    "derivative",
    "newton_options",
    "fem",
    "mesh",
    "io",
    "quadrature",
    "obstacle_pg",
    "forms",
    "gradient_constraints",
    "multiphase",
    "signorini",
]
