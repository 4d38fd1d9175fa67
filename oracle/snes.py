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


DIVERGED_LINE_SEARCH = -6


def linesearch_bt(residual, x, F, y, fnorm, Jy, alpha=1e-4, maxstep=1e8, steptol=1e-12, max_it=40):
    """PETSc SNESLineSearchApply_BT, cubic order -- the default line search of ``newtonls`` when
    ``snes_linesearch_type`` is not set (examples/04_multiphase/multiphase_dolfinx.py:128-143).
    Restated from memory of petsc/src/snes/linesearch/impls/bt/linesearchbt.c [3P-mem]: sufficient
    decrease of 0.5 ||F||^2 along x - lambda y with slope F.(J y), quadratic then cubic backtracking.
    Returns (x_new, F_new, gnorm, lambda, ok)."""
    ynorm = np.linalg.norm(y)
    if ynorm == 0.0:
        return x.copy(), F.copy(), fnorm, 0.0, True
    if ynorm > maxstep:
        y = y * (maxstep / ynorm)
        ynorm = maxstep
    f = fnorm * fnorm
    initslope = float(np.dot(F, Jy))
    if initslope > 0.0:
        initslope = -initslope
    if initslope == 0.0:
        initslope = -1.0
    rellength = np.max(np.abs(y) / np.maximum(np.abs(x), 1.0))
    minlambda = steptol / rellength
    lam = 1.0
    w = x - lam * y
    G = residual(w)
    gnorm = np.linalg.norm(G)
    g = gnorm * gnorm
    if not np.isfinite(gnorm):
        return w, G, gnorm, lam, False
    if 0.5 * g <= 0.5 * f + lam * alpha * initslope:
        return w, G, gnorm, lam, True
    # quadratic fit
    lamtemp = -initslope / (g - f - 2.0 * lam * initslope)
    lamprev, gprev = lam, g
    if lamtemp > 0.5 * lam:
        lamtemp = 0.5 * lam
    lam = 0.1 * lam if lamtemp <= 0.1 * lam else lamtemp
    for _ in range(max_it):
        if lam <= minlambda:
            return w, G, gnorm, lam, False
        w = x - lam * y
        G = residual(w)
        gnorm = np.linalg.norm(G)
        g = gnorm * gnorm
        if 0.5 * g <= 0.5 * f + lam * alpha * initslope:
            return w, G, gnorm, lam, True
        t1 = 0.5 * (g - f) - lam * initslope
        t2 = 0.5 * (gprev - f) - lamprev * initslope
        a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
        b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
        d = b * b - 3.0 * a * initslope
        if d < 0.0:
            d = 0.0
        lamtemp = -initslope / (2.0 * b) if a == 0.0 else (-b + np.sqrt(d)) / (3.0 * a)
        lamprev, gprev = lam, g
        if lamtemp > 0.5 * lam:
            lamtemp = 0.5 * lam
        lam = 0.1 * lam if lamtemp <= 0.1 * lam else lamtemp
    return w, G, gnorm, lam, False


def newton_ls(residual, jacobian, x0, linesearch="none", rtol=1e-8, atol=1e-50, stol=1e-8, max_it=50, divtol=1e4,
              maxstep=1e8):
    """SNESSolve_NEWTONLS with line search "none" (basic, full step), "bt" or "l2".  Returns (x, reason, its, hist)."""
    if linesearch in ("none", "basic"):
        return newton_ls_none(residual, jacobian, x0, rtol=rtol, atol=atol, stol=stol, max_it=max_it, divtol=divtol)
    x = x0.copy()
    F = residual(x)
    fnorm = np.linalg.norm(F)
    fnorm0, ttol, hist = fnorm, fnorm * rtol, [fnorm]
    reason = converged_default(0, 0.0, 0.0, fnorm, ttol, fnorm0, atol, stol, divtol)
    if reason:
        return x, reason, 0, hist
    for i in range(max_it):
        J = jacobian(x)
        y = spla.splu(J.tocsc()).solve(F)
        if not np.all(np.isfinite(y)):
            return x, DIVERGED_LINEAR_SOLVE, i, hist
        if linesearch == "l2":
            xn, Fn, gnorm, lam, ok = linesearch_l2(residual, x, y, fnorm, maxstep=maxstep)
        else:
            xn, Fn, gnorm, lam, ok = linesearch_bt(residual, x, F, y, fnorm, J @ y, maxstep=maxstep)
        if not ok:
            return x, DIVERGED_LINE_SEARCH, i, hist
        snorm = np.linalg.norm(xn - x)
        x, F, fnorm = xn, Fn, gnorm
        hist.append(fnorm)
        reason = converged_default(i + 1, np.linalg.norm(x), snorm, fnorm, ttol, fnorm0, atol, stol, divtol)
        if reason:
            return x, reason, i + 1, hist
    return x, DIVERGED_MAX_IT, max_it, hist


def linesearch_l2(residual, x, y, fnorm, lam0=1.0, maxstep=1e8, steptol=1e-12, max_it=1):
    """PETSc's SNESLineSearchApply_L2 (secant search on ||F(x - lambda y)||^2; ``snes_linesearch_type l2`` of the
    reference's examples 03, 07-10, e.g. fracture_dolfinx.py:132-138 with ``snes_linesearch_maxlambda 1``), restated
    from memory of petsc/src/snes/linesearch/impls/l2/linesearchl2.c [3P-mem]: phi(lambda) = ||F(x - lambda y)||^2 is sampled at the ends and the midpoint of
    [lambda_old, lambda]; one-sided second-order differences give phi' at both ends, their difference phi'' ; the
    secant step lambda - phi'/|phi''| is taken (always downhill), reset to the midpoint when it falls below steptol,
    abandoned when it is not finite or exceeds maxstep.  ``max_it`` = 1 is PETSc's default for l2.  A non-finite norm
    halves the interval (and caps maxstep at 0.95 of the failed length).  Returns (x_new, F_new, gnorm, lambda, ok)."""
    lam, lam_old = float(lam0), 0.0
    phi_old = fnorm * fnorm
    lam_mid = 0.5 * (lam + lam_old)
    for _ in range(max_it):
        while True:
            with np.errstate(over="ignore", invalid="ignore"):
                phi_mid = float(np.linalg.norm(residual(x - lam_mid * y))) ** 2
                phi = float(np.linalg.norm(residual(x - lam * y))) ** 2
            if np.isfinite(phi) and np.isfinite(phi_mid):
                break
            if lam <= steptol:
                return x, None, np.nan, lam, False
            maxstep = 0.95 * lam
            lam = 0.5 * (lam + lam_old)
            lam_mid = 0.5 * (lam + lam_old)
        dl = lam - lam_old
        dphi = (3.0 * phi - 4.0 * phi_mid + phi_old) / dl
        dphi_old = (-3.0 * phi_old + 4.0 * phi_mid - phi) / dl
        d2phi = (dphi - dphi_old) / dl
        if d2phi > 0.0:
            lam_new = lam - dphi / d2phi
        elif d2phi < 0.0:
            lam_new = lam + dphi / d2phi
        else:
            break
        if lam_new < steptol:
            lam_new = 0.5 * (lam + lam_old)
        if not np.isfinite(lam_new) or lam_new > maxstep:
            break
        lam_old, lam, phi_old = lam, lam_new, phi
        lam_mid = 0.5 * (lam + lam_old)
    xn = x - lam * y
    with np.errstate(over="ignore", invalid="ignore"):
        Fn = residual(xn)
    gnorm = float(np.linalg.norm(Fn))
    return xn, Fn, gnorm, lam, bool(np.isfinite(gnorm))
