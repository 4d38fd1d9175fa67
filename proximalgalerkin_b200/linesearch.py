"""SNES ``newtonls`` with PETSc's ``bt`` (cubic backtracking) and ``l2`` (secant) line searches for the obstacle engine.

The obstacle driver sets ``snes_linesearch_type none`` (examples/01_obstacle_problem/obstacle_pg.py:136)
and that path runs entirely inside the library (``lvpp_newton_solve``).  The reference's other examples
switch the line search on where the full Newton step is not robust (``bt``:
examples/05_thermoforming/thermoforming_dolfinx.py:99-111, the PETSc default used by
examples/04_multiphase/multiphase_dolfinx.py:128-143) -- and the full step stops being
robust for the 3-D obstacle problem on fine meshes: the largest increase of psi in the first Newton step
of a proximal iteration depends on how the mesh meets the contact boundary (0.05, 2.8, 5.8, 7.2, 0.9 for
n = 8, 12, 16, 20, 24 cubes per axis at alpha = 1.49) and at n = 215 the exponential overshoots (||F|| jumps from 5e-3 to 0.23, then to
3e28).  With ``snes_linesearch_type bt`` the same Newton direction is damped until
0.5 ||F||^2 decreases sufficiently.

The loop is written against a small backend interface so that the control flow can be checked on the
CPU against the restated PETSc algorithm; on the device every operation is one of the library's entry
points (residual + Jacobian assembly, Krylov solve, J*v) or a torch vector operation on the owned
entries followed by an all-reduce.
"""
import math

# SNESConvergedReason (petscsnes.h)
CONVERGED_FNORM_ABS = 2
CONVERGED_FNORM_RELATIVE = 3
CONVERGED_SNORM_RELATIVE = 4
DIVERGED_LINEAR_SOLVE = -3
DIVERGED_FNORM_NAN = -4
DIVERGED_MAX_IT = -5
DIVERGED_LINE_SEARCH = -6
DIVERGED_DTOL = -9


def converged_default(it, xnorm, snorm, fnorm, ttol, fnorm0, atol, stol, divtol):
    """SNESConvergedDefault."""
    if math.isnan(fnorm) or math.isinf(fnorm):
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


class Backend:
    """What the Newton loop needs.  Vectors are opaque handles created by ``vector()``.

    residual(x, F) -> ||F(x)||      also leaves the Jacobian at x assembled
    solve(F, y) -> (its, reason)    y = J^-1 F with the Jacobian of the last ``residual`` call
    mult(y, Jy)                     Jy = J y, same Jacobian
    waxpy(w, a, y, x)               w = x + a y   (all local entries)
    copy(dst, src)
    dot(a, b), norm(a)              over the owned entries of all ranks
    rellength(y, x)                 max_i |y_i| / max(|x_i|, 1) over the owned entries of all ranks
    """


def linesearch_bt(be, x, F, y, Jy, w, G, fnorm, alpha=1e-4, maxstep=1e8, steptol=1e-12, max_it=40):
    """SNESLineSearchApply_BT, cubic order, on the step x - lambda y.  On return ``w`` holds the new
    iterate and ``G`` its residual (the Jacobian is assembled there).  Returns (gnorm, lambda, ok, ynorm)."""
    ynorm = be.norm(y)
    if ynorm == 0.0:
        be.copy(w, x)
        be.copy(G, F)
        return fnorm, 0.0, True, 0.0
    scale = 1.0
    if ynorm > maxstep:  # PETSc scales the direction itself
        scale = maxstep / ynorm
        ynorm = maxstep
    f = fnorm * fnorm
    be.mult(y, Jy)  # before any trial residual: the Jacobian is still the one at x
    initslope = scale * be.dot(F, Jy)
    if initslope > 0.0:
        initslope = -initslope
    if initslope == 0.0:
        initslope = -1.0
    minlambda = steptol / (scale * be.rellength(y, x))
    lam = 1.0
    be.waxpy(w, -lam * scale, y, x)
    gnorm = be.residual(w, G)
    g = gnorm * gnorm
    if not math.isfinite(gnorm):
        # exp(psi) overflowed at the full step: halve until the function is finite, then fit (an addition to the
        # restated PETSc control flow, which is otherwise followed as in oracle/snes.py)
        while not math.isfinite(gnorm):
            lam *= 0.5
            if lam <= minlambda:
                return gnorm, lam, False, ynorm
            be.waxpy(w, -lam * scale, y, x)
            gnorm = be.residual(w, G)
        g = gnorm * gnorm
    if 0.5 * g <= 0.5 * f + lam * alpha * initslope:
        return gnorm, lam, True, ynorm
    # quadratic fit through f, the slope at 0 and g
    lamtemp = -initslope / (g - f - 2.0 * lam * initslope)
    lamprev, gprev = lam, g
    if lamtemp > 0.5 * lam:
        lamtemp = 0.5 * lam
    lam = 0.1 * lam if lamtemp <= 0.1 * lam else lamtemp
    for _ in range(max_it):
        if lam <= minlambda:
            return gnorm, lam, False, ynorm
        be.waxpy(w, -lam * scale, y, x)
        gnorm = be.residual(w, G)
        g = gnorm * gnorm
        if math.isfinite(g) and 0.5 * g <= 0.5 * f + lam * alpha * initslope:
            return gnorm, lam, True, ynorm
        if not math.isfinite(g):
            lamprev, gprev = lam, g
            lam *= 0.5
            continue
        if not math.isfinite(gprev):
            # no usable second point: quadratic fit through the current one
            lamtemp = -initslope / (g - f - 2.0 * lam * initslope)
        else:
            t1 = 0.5 * (g - f) - lam * initslope
            t2 = 0.5 * (gprev - f) - lamprev * initslope
            a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
            b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
            d = b * b - 3.0 * a * initslope
            if d < 0.0:
                d = 0.0
            lamtemp = -initslope / (2.0 * b) if a == 0.0 else (-b + math.sqrt(d)) / (3.0 * a)
        lamprev, gprev = lam, g
        if lamtemp > 0.5 * lam:
            lamtemp = 0.5 * lam
        lam = 0.1 * lam if lamtemp <= 0.1 * lam else lamtemp
    return gnorm, lam, False, ynorm


def linesearch_l2(be, x, y, w, G, fnorm, lam0=1.0, maxstep=1e8, steptol=1e-12, max_it=1):
    """SNESLineSearchApply_L2 on the step x - lambda y (``snes_linesearch_type l2``: the reference's examples 03 and
    07-10, e.g. examples/03_fracture/fracture_dolfinx.py:132-138).  A secant iteration on
    phi(lambda) = ||F(x - lambda y)||^2 from three samples of the bracket [lambda_old, lambda] (both ends and the
    midpoint); PETSc's default is a single iteration.  On return ``w`` holds the new iterate and ``G`` its residual
    (Jacobian assembled there).  Returns (gnorm, lambda, ok, ynorm)."""
    ynorm = be.norm(y)

    def phi_at(lam):
        be.waxpy(w, -lam, y, x)
        g = be.residual(w, G)
        return g * g

    hi, lo = float(lam0), 0.0
    p_lo = fnorm * fnorm
    for _ in range(max_it):
        mid = 0.5 * (hi + lo)
        p_mid, p_hi = phi_at(mid), phi_at(hi)
        while not (math.isfinite(p_mid) and math.isfinite(p_hi)):
            if hi <= steptol:
                return float("nan"), hi, False, ynorm
            maxstep = 0.95 * hi  # never go back to a length where the function is not finite
            hi = 0.5 * (hi + lo)
            mid = 0.5 * (hi + lo)
            p_mid, p_hi = phi_at(mid), phi_at(hi)
        width = hi - lo
        slope_hi = (3.0 * p_hi - 4.0 * p_mid + p_lo) / width     # one-sided second-order differences
        slope_lo = (-3.0 * p_lo + 4.0 * p_mid - p_hi) / width
        curv = (slope_hi - slope_lo) / width
        if curv == 0.0:
            break
        # secant (Newton) step on phi'; PETSc's SNESLineSearchApply_L2 takes lambda - del/del2 for del2 > 0 and
        # lambda + del/del2 for del2 < 0 ("always go downhill"), i.e. the division by |del2| below
        cand = hi - slope_hi / abs(curv)
        if cand < steptol:
            cand = 0.5 * (hi + lo)
        if not math.isfinite(cand) or cand > maxstep:
            break
        lo, hi, p_lo = hi, cand, p_hi
    gnorm = math.sqrt(phi_at(hi))
    return gnorm, hi, math.isfinite(gnorm), ynorm


class NewtonBT:
    """SNESSolve_NEWTONLS with the bt (default) or l2 line search, one step at a time (``begin`` then ``step`` until
    the reason is non-zero) so that callers can time or log individual Newton steps."""

    def __init__(self, be, rtol=1e-8, atol=1e-50, stol=1e-8, max_it=50, divtol=1e4, linesearch="bt", maxstep=1e8):
        self.linesearch, self.maxstep = linesearch, maxstep
        self.be = be
        self.rtol, self.atol, self.stol, self.max_it, self.divtol = rtol, atol, stol, max_it, divtol
        self.F, self.y, self.Jy, self.w, self.G = (be.vector() for _ in range(5))
        self.its = self.linear_its = 0
        self.reason = 0
        self.fnorm = self.fnorm0 = self.ttol = float("nan")
        self.last_lambda = 1.0

    def begin(self, x):
        self.its = self.linear_its = 0
        self.fnorm = self.be.residual(x, self.F)
        self.fnorm0, self.ttol = self.fnorm, self.fnorm * self.rtol
        self.reason = converged_default(0, 0.0, 0.0, self.fnorm, self.ttol, self.fnorm0, self.atol, self.stol, self.divtol)
        return self.fnorm

    def step(self, x):
        """One Newton step on ``x`` (updated in place).  Returns the SNES reason (0 = keep going)."""
        be = self.be
        if self.reason:
            return self.reason
        if self.its >= self.max_it:
            self.reason = DIVERGED_MAX_IT
            return self.reason
        kits, kreason = be.solve(self.F, self.y)
        self.linear_its += kits
        if kreason < 0:
            self.reason = DIVERGED_LINEAR_SOLVE
            return self.reason
        if self.linesearch == "l2":
            gnorm, lam, ok, ynorm = linesearch_l2(be, x, self.y, self.w, self.G, self.fnorm, maxstep=self.maxstep)
        else:
            gnorm, lam, ok, ynorm = linesearch_bt(be, x, self.F, self.y, self.Jy, self.w, self.G, self.fnorm, maxstep=self.maxstep)
        self.last_lambda = lam
        if not ok:
            # the Jacobian now sits at the last trial point: restore it at x for a caller that carries on
            be.residual(x, self.F)
            self.reason = DIVERGED_LINE_SEARCH
            return self.reason
        be.copy(x, self.w)
        self.F, self.G = self.G, self.F
        self.fnorm = gnorm
        self.its += 1
        snorm = lam * min(ynorm, 1e8)
        self.reason = converged_default(self.its, be.norm(x), snorm, self.fnorm, self.ttol, self.fnorm0, self.atol,
                                        self.stol, self.divtol)
        return self.reason

    def solve(self, x):
        self.begin(x)
        while not self.reason:
            self.step(x)
        return self.reason, self.its


class DeviceBackend(Backend):
    """The obstacle engine's entry points behind the Backend interface (``dev``: problem.DeviceProblem)."""

    def __init__(self, dev, opts):
        import torch

        self.dev, self.opts, self.torch = dev, opts, torch
        self.nown2 = dev.n_owned        # owned entries of a mixed vector (ghosts last)
        self.nranks = dev.V.mesh.nranks

    def vector(self):
        from .problem import DeviceVector

        return DeviceVector(self.dev.n, self.dev.device)

    def residual(self, x, F):
        return self.dev.assemble_residual(x, F)

    def solve(self, F, y):
        its, reason, _ = self.dev.linear_solve(F, y, self.opts)
        return its, reason

    def mult(self, y, Jy):
        self.dev.spmv(y, Jy)

    def waxpy(self, w, a, y, x):
        self.torch.add(x.tensor, y.tensor, alpha=a, out=w.tensor)

    def copy(self, dst, src):
        dst.tensor.copy_(src.tensor)

    def _allreduce(self, t, op="sum"):
        if self.nranks > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
        return float(t)

    def dot(self, a, b):
        n = self.nown2
        return self._allreduce(self.torch.dot(a.tensor[:n], b.tensor[:n]))

    def norm(self, a):
        return math.sqrt(self.dot(a, a))

    def rellength(self, y, x):
        n = self.nown2
        t = (y.tensor[:n].abs() / x.tensor[:n].abs().clamp_min(1.0)).max()
        return self._allreduce(t, "max")
