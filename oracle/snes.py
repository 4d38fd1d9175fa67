"""PETSc SNES ``newtonls`` control flow with a sparse-LU linear solve (oracle; see oracle/__init__.py).

Restates the published algorithm of PETSc's SNESSolve_NEWTONLS + SNESConvergedDefault as used via
the options at examples/01_obstacle_problem/obstacle_pg.py:128-139 (``ksp_type preonly``,
``pc_type lu``, MUMPS -> scipy ``splu`` here; ``snes_linesearch_type none`` -> full step
x <- x - y; ``snes_rtol``, ``snes_max_it``) and the SNESSolver wrapper at
src/lvpp/problem.py:114-124.  PETSc itself is not vendored in the reference nor installed here.
"""
import numpy as np
import scipy.sparse.linalg as spla

# SNESConvergedReason values (petsc/include/petscsnes.h)
CONVERGED_FNORM_ABS = 2
CONVERGED_FNORM_RELATIVE = 3
CONVERGED_SNORM_RELATIVE = 4
DIVERGED_FUNCTION_COUNT = -2
DIVERGED_LINEAR_SOLVE = -3
DIVERGED_FNORM_NAN = -4
DIVERGED_MAX_IT = -5
DIVERGED_DTOL = -9

DEFAULTS = dict(rtol=1e-8, atol=1e-50, stol=1e-8, max_it=50, divtol=1e4)


def converged_default(it, xnorm, snorm, fnorm, ttol, fnorm0, atol, stol, divtol):
    """SNESConvergedDefault."""
    if np.isnan(fnorm) or np.isinf(fnorm):
        return DIVERGED_FNORM_NAN
    if fnorm < atol:
        return CONVERGED_FNORM_ABS
    if it:
        if fnorm <= ttol:
            return CONVERGED_FNORM_RELATIVE
        if snorm < stol * xnorm:
            return CONVERGED_SNORM_RELATIVE
        if divtol > 0 and fnorm > divtol * fnorm0:
            return DIVERGED_DTOL
    return 0


def newton_ls_none(residual, jacobian, x0, rtol=1e-8, atol=1e-50, stol=1e-8, max_it=50, divtol=1e4,
                   linear_solve=None):
    """Returns (x, reason, its, fnorm_history).  ``residual(x) -> F``, ``jacobian(x) -> csr``."""
    x = x0.copy()
    F = residual(x)
    fnorm = np.linalg.norm(F)
    fnorm0 = fnorm
    ttol = fnorm * rtol
    hist = [fnorm]
    reason = converged_default(0, 0.0, 0.0, fnorm, ttol, fnorm0, atol, stol, divtol)
    if reason:
        return x, reason, 0, hist
    its = 0
    for i in range(max_it):
        J = jacobian(x)
        if linear_solve is None:
            y = spla.splu(J.tocsc()).solve(F)
        else:
            y = linear_solve(J, F)
        if not np.all(np.isfinite(y)):
            return x, DIVERGED_LINEAR_SOLVE, its, hist
        x = x - y  # line search "none" (basic): lambda = 1
        F = residual(x)
        fnorm = np.linalg.norm(F)
        ynorm = np.linalg.norm(y)
        xnorm = np.linalg.norm(x)
        its = i + 1
        hist.append(fnorm)
        reason = converged_default(its, xnorm, ynorm, fnorm, ttol, fnorm0, atol, stol, divtol)
        if reason:
            return x, reason, its, hist
    return x, DIVERGED_MAX_IT, its, hist
