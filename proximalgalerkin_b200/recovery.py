"""Failure-recovering proximal step control.

The reference's harder examples do not follow a fixed alpha schedule: they adapt alpha to the Newton count of
the last proximal step and, when a Newton solve fails, halve alpha, put the unknown back to the last accepted
proximal iterate and try the same step again
(examples/03_fracture/fracture_dolfinx.py:215-283,
examples/07_eigenvalue_constraints/eigenvalue_constraints_dolfinx.py:163-225,
examples/08_intersecting_constraints/intersecting_constraints_dolfinx.py:120-174 -- the same loop three times).

``AdaptiveAlpha`` is that rule as a small state machine, so that the host-buffer driver
(``obstacle_pg.solve_problem(alpha_scheme="adaptive")``) and the device-resident one
(``obstacle_pg.LvppStepper(alpha_scheme="adaptive")``) share it.  It holds no device state: the caller restores the
unknown when ``failed()`` says so.
"""


class GaveUp(RuntimeError):
    """nfail reached nfail_max (fracture_dolfinx.py:255-260 prints "Giving up" and leaves the loop)."""


class AdaptiveAlpha:
    """alpha control of fracture_dolfinx.py:215-283.

    alpha_0 = 1 (:215), r = 2 (:218).  After a converged solve with ``its`` Newton steps: alpha *= r if its <= 4,
    alpha /= r if its >= 10 (:276-279).  After a failed solve: alpha /= 2 (:249), nfail += 1 (:242).  ``alpha_max``
    (absent from the reference loops) clamps alpha like obstacle_pg.py:184 when given.
    """

    def __init__(self, alpha_0=1.0, r=2.0, nfail_max=50, alpha_max=None, fast=4, slow=10):
        self.alpha = float(alpha_0)
        self.r, self.nfail_max, self.alpha_max = float(r), int(nfail_max), alpha_max
        self.fast, self.slow = int(fast), int(slow)
        self.nfail = 0
        self.k = 1  # proximal step being attempted (the reference counts from 1, :216)
        self.attempts = []  # (k, alpha, Newton steps, reason) of every solve, failed ones included

    @staticmethod
    def is_failure(reason, its):
        """:234-240 -- diverged, or "converged" without a Newton step: alpha is so small that the initial guess
        already satisfies the equation and the proximal iteration would stall."""
        return reason < 0 or (its == 0 and reason > 0)

    def record(self, its, reason):
        self.attempts.append((self.k, self.alpha, int(its), int(reason)))

    def failed(self):
        """Bookkeeping of a failed solve.  The caller then restores the unknown to the last accepted proximal iterate
        (:250-253) and repeats the step; raises GaveUp when the failure budget is spent."""
        self.nfail += 1
        self.alpha /= 2
        if self.nfail >= self.nfail_max:
            raise GaveUp(f"LVPP: {self.nfail} failed Newton solves, alpha = {self.alpha:g}, proximal step {self.k}")

    def accepted(self, its):
        """Bookkeeping of a converged solve whose increment did not meet the stopping test (:276-283)."""
        if its <= self.fast:
            self.alpha *= self.r
        elif its >= self.slow:
            self.alpha /= self.r
        if self.alpha_max is not None:
            self.alpha = min(self.alpha, self.alpha_max)
        self.k += 1
